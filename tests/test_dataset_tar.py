"""SURVEY.md §8 row f1 to completion: the reference's dataset tar (memory_mapped_data.py:198-285), fold selection (:85-101) and
the chunked, shuffled iteration of ReadsDataset + DataLoader(collate_fn=Batch) (reads_dataset.py:109-209), against the
UNMODIFIED reference (oracle/_ref): a tar written by the reference's ``save_to_tarfile`` loads here, the batches of a seeded
pass are the reference's batches variant for variant, and a tar written here loads in the reference."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import reference

pytestmark = pytest.mark.skipif(not reference.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def _reference_dataset(tmp_path, n=307, seed=21):
    reference.load()
    import permutect.data.datum as rd
    import permutect.data.memory_mapped_data as rm
    from permutect_b200.synthetic import make_wgs_arrays
    ia, fa, reads = make_wgs_arrays(n, seed=seed)
    rng = np.random.default_rng(seed)
    ia[:, 4] = rng.integers(0, 2, n)                     # two sources
    ref_c, alt_c = ia[:, 0].astype(int), ia[:, 1].astype(int)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = ref_off[-1]
    data = [rd.Datum(ia[v], fa[v], np.vstack((reads[ref_off[v]:ref_off[v + 1]], reads[total_ref + alt_off[v]:total_ref + alt_off[v + 1]])),
                     compressed=True) for v in range(n)]
    mmd = rm.MemoryMappedData.from_generator(iter(data), estimated_num_data=n // 3, estimated_num_reads=100)   # forces regrowth
    tar = os.path.join(tmp_path, "dataset.tar")
    mmd.save_to_tarfile(tar)
    return tar, ia, fa, data, rm


def test_reference_tar_loads_here(tmp_path):
    from permutect_b200.data.memory_mapped_data import MemoryMappedData
    tar, ia, fa, data, rm = _reference_dataset(str(tmp_path))
    ref = rm.MemoryMappedData.load_from_tarfile(tar)
    mine = MemoryMappedData.load_from_tarfile(tar)
    assert (mine.num_data, mine.num_reads) == (ref.num_data, ref.num_reads) == (len(ia), sum(len(d.get_reads_array_re()) for d in data))
    # the reference maps the files' whole capacity (junk rows past num_data / num_reads included); here the used prefix
    np.testing.assert_array_equal(mine.int_mmap, ref.int_mmap[: ref.num_data])
    np.testing.assert_array_equal(mine.float_mmap, ref.float_mmap[: ref.num_data])
    np.testing.assert_array_equal(mine.reads_mmap, ref.reads_mmap[: ref.num_reads])
    np.testing.assert_array_equal(mine.read_end_indices, ref.read_end_indices)
    assert mine.read_end_indices.dtype == ref.read_end_indices.dtype
    # per-variant view and fold restriction against the reference's Datum generators
    for a, b in zip(mine.generate(num_folds=4, used_folds=[1, 3]), ref.generate(num_folds=4, used_folds=[1, 3])):
        np.testing.assert_array_equal(a.get_int_array(), b.get_int_array())
        np.testing.assert_array_equal(a.get_reads_array_re(), b.get_reads_array_re())
    sub_mine, sub_ref = mine.restrict_to_folds(4, [1, 3]), ref.restrict_to_folds(4, [1, 3])
    assert sub_mine.num_data == sub_ref.num_data and sub_mine.num_reads == sub_ref.num_reads
    np.testing.assert_array_equal(sub_mine.int_mmap[: sub_mine.num_data], sub_ref.int_mmap[: sub_ref.num_data])
    np.testing.assert_array_equal(sub_mine.reads_mmap[: sub_mine.num_reads], sub_ref.reads_mmap[: sub_ref.num_reads])
    # a tar written here loads in the reference
    out = os.path.join(str(tmp_path), "mine.tar")
    sub_mine.save_to_tarfile(out)
    back = rm.MemoryMappedData.load_from_tarfile(out)
    assert back.num_data == sub_ref.num_data
    np.testing.assert_array_equal(back.float_mmap[: back.num_data], sub_ref.float_mmap[: sub_ref.num_data])
    np.testing.assert_array_equal(back.reads_mmap[: back.num_reads], sub_ref.reads_mmap[: sub_ref.num_reads])


@pytest.mark.parametrize("n_chunks", [1, 3])
def test_seeded_pass_yields_the_reference_batches(tmp_path, monkeypatch, n_chunks):
    import psutil
    from permutect_b200.data.memory_mapped_data import MemoryMappedData
    from permutect_b200.data.reads_dataset import ReadsDataset
    tar, ia, fa, data, rm = _reference_dataset(str(tmp_path))
    import permutect.data.datum as rd
    import permutect.data.batch as rb
    import permutect.data.reads_dataset as rds
    ref_mmd, my_mmd = rm.MemoryMappedData.load_from_tarfile(tar), MemoryMappedData.load_from_tarfile(tar)
    ref_ds = rds.ReadsDataset(ref_mmd, num_folds=3, folds_to_use=[0, 2])
    my_ds = ReadsDataset(my_mmd, num_folds=3, folds_to_use=[0, 2])
    assert len(my_ds) == len(ref_ds)
    assert (my_ds.num_read_features(), my_ds.num_info_features(), my_ds.haplotypes_length()) == \
           (ref_ds.num_read_features(), ref_ds.num_info_features(), ref_ds.haplotypes_length())
    np.testing.assert_array_equal(my_ds.totals_slvra.as_subclass(torch.Tensor).numpy(), ref_ds.totals_slvra.as_subclass(torch.Tensor).numpy())
    assert my_ds.num_sources() == ref_ds.num_sources() == 2

    # the chunk count is 1 + 8 * bytes // available memory (reads_dataset.py:126-128): pick "available" accordingly
    # (the reference's restricted copy keeps its spare capacity, so its byte count is a little larger than the used prefix here)
    ref_bytes, my_bytes = ref_ds.memory_mapped_data.num_bytes(), my_ds.memory_mapped_data.num_bytes()
    assert my_bytes <= ref_bytes <= 1.25 * my_bytes

    class _Mem:
        available = 10 ** 15 if n_chunks == 1 else (8 * my_bytes) // (n_chunks - 1)
        percent = 0.0        # report_memory_usage reads it
    monkeypatch.setattr(psutil, "virtual_memory", lambda: _Mem)
    assert 1 + (8 * ref_bytes) // _Mem.available == 1 + (8 * my_bytes) // _Mem.available == n_chunks

    for batch_size in (16, 37):
        random.seed(5)
        ref_batches = list(ref_ds.make_data_loader(batch_size=batch_size))
        random.seed(5)
        my_batches = list(my_ds.make_data_loader(batch_size=batch_size))
        assert len(my_batches) == len(ref_batches) == (len(ref_ds) + batch_size - 1) // batch_size
        for mine, ref in zip(my_batches, ref_batches):
            assert torch.equal(mine.int_tensor.long(), ref.int_tensor)
            np.testing.assert_array_equal(mine.float_tensor.float().numpy(), ref.float_tensor.numpy())     # NaN == NaN
            # rows stay compressed here until the GPU decodes them: expand them with the reference's own decoder (batch.py:51-56)
            rows = mine.reads.numpy()
            decoded = np.hstack((np.unpackbits(rows[:, :rd.NUMBER_OF_BYTES_IN_PACKED_READ], axis=1).astype(np.float32),
                                 rb.convert_uint8_to_quantile_normalized(rows[:, rd.NUMBER_OF_BYTES_IN_PACKED_READ:])))
            np.testing.assert_array_equal(decoded, ref.reads_re.numpy())
    # per-variant iteration in the same order
    random.seed(9)
    ref_order = [d.get_int_array().copy() for d in ref_ds]
    random.seed(9)
    my_order = [d.get_int_array().copy() for d in my_ds]
    np.testing.assert_array_equal(np.vstack(my_order), np.vstack(ref_order))
