"""Batch assembly from the dataset memory maps: dataset-order slices + device gather indices must give
exactly the batches (and logits) of the reference-style per-Datum collate."""
import numpy as np
import pytest
import torch

from permutect_b200.data.batch import Batch
from permutect_b200.data.datum import Datum
from permutect_b200.data.reads_dataset import MemoryMappedBatches
from permutect_b200.synthetic import make_wgs_arrays


def _dataset(n, seed):
    """A MemoryMappedData-style triple: reads in DATASET order (per variant: ref rows, then alt rows)."""
    ia, fa, reads_batch_order = make_wgs_arrays(n, seed=seed)
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = int(ref_off[-1])
    rows = []
    for v in range(n):
        rows.append(reads_batch_order[ref_off[v]:ref_off[v + 1]])
        rows.append(reads_batch_order[total_ref + alt_off[v]:total_ref + alt_off[v + 1]])
    return ia, fa, np.concatenate(rows), ref_c, alt_c


def _expected_indices(ref_c, alt_c):
    """numpy restatement of batch.py:45-47 applied to a dataset-order slice."""
    start = np.concatenate(([0], np.cumsum(ref_c + alt_c)))[:-1]
    ref_idx = [start[v] + np.arange(ref_c[v]) for v in range(len(ref_c))]
    alt_idx = [start[v] + ref_c[v] + np.arange(alt_c[v]) for v in range(len(ref_c))]
    return np.concatenate(ref_idx + alt_idx).astype(np.int64)


def test_memory_mapped_batches_slices_cover_the_dataset():
    ia, fa, reads_ds, ref_c, alt_c = _dataset(1000, seed=3)
    loader = MemoryMappedBatches(ia, fa, reads_ds, batch_size=128, pin_memory=False)
    assert len(loader) == 8
    seen, rows = 0, 0
    for b in loader:
        assert b._dataset_order and b.read_indices is None
        n = b.size()
        np.testing.assert_array_equal(b.int_tensor.numpy(), ia[seen:seen + n])
        k = int((ref_c[seen:seen + n] + alt_c[seen:seen + n]).sum())
        np.testing.assert_array_equal(b.reads.numpy(), reads_ds[rows:rows + k])
        seen, rows = seen + n, rows + k
    assert seen == 1000 and rows == len(reads_ds)
    shard = MemoryMappedBatches(ia, fa, reads_ds, batch_size=100, start=250, stop=777, pin_memory=False)
    assert sum(b.size() for b in shard) == 527


@pytest.mark.gpu
def test_dataset_order_batches_equal_the_per_datum_collate():
    from golden_utils import load
    from helpers import model_from_golden
    from permutect_b200.data.prefetch_generator import prefetch_generator
    from permutect_b200.utils.enums import Epoch
    dev = torch.device("cuda:0")
    g = load("v040_seed0_b64")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    ia, fa, reads_ds, ref_c, alt_c = _dataset(700, seed=9)
    starts = np.concatenate(([0], np.cumsum(ref_c + alt_c)))
    loader = MemoryMappedBatches(ia, fa, reads_ds, batch_size=256)
    v0 = 0
    with torch.inference_mode():
        for b in prefetch_generator(loader, dev):
            n = b.size()
            want_idx = _expected_indices(ref_c[v0:v0 + n], alt_c[v0:v0 + n])
            np.testing.assert_array_equal(b.read_indices.cpu().numpy(), want_idx)                   # bit-exact index work
            # the reference's collate: one Datum per variant (reads = its ref rows then alt rows), Batch(list)
            data = [Datum(ia[v], fa[v], reads_ds[starts[v]:starts[v + 1]], compressed=True) for v in range(v0, v0 + n)]
            ref_batch = Batch(data).copy_to(dev)
            assert torch.equal(b.get_reads_re(), ref_batch.get_reads_re())
            assert torch.equal(model.compute_batch_output(b).logits_b, model.compute_batch_output(ref_batch).logits_b)
            v0 += n
    assert v0 == 700


@pytest.mark.gpu
def test_downsampling_a_dataset_order_batch_matches_downsampling_the_collated_batch():
    from permutect_b200.data.batch import DownsampledBatch
    dev = torch.device("cuda:0")
    ia, fa, reads_ds, ref_c, alt_c = _dataset(300, seed=21)
    starts = np.concatenate(([0], np.cumsum(ref_c + alt_c)))
    a = Batch.from_dataset_slice(ia, fa, reads_ds).copy_to(dev)
    b = Batch([Datum(ia[v], fa[v], reads_ds[starts[v]:starts[v + 1]], compressed=True) for v in range(300)]).copy_to(dev)
    fr = torch.full((300,), 0.6, device=dev)
    for offset in (False, True):           # reference behaviour (quirk Q1) and the corrected offset
        import random
        random.seed(17)                      # the forced-alt-read draw (batch.py:418) comes from Python's generator
        da = DownsampledBatch(a, fr, fr, seed=5, offset_alt_rows=offset)
        random.seed(17)
        db = DownsampledBatch(b, fr, fr, seed=5, offset_alt_rows=offset)
        assert torch.equal(da.ref_counts, db.ref_counts) and torch.equal(da.alt_counts, db.alt_counts)
        n = int(da.ref_counts.sum() + da.alt_counts.sum())
        assert da.get_reads_re().shape[0] == n and torch.equal(da.get_reads_re(), db.get_reads_re())


def test_parallel_copy_and_plain_slices_of_the_dataset_loader():
    """Host logic of the ingest path without a GPU: the row-range copy the staging threads run is exact for every split, and
    the unpinned loader yields the maps' own slices in dataset order (variant count, row count, contents)."""
    from concurrent.futures import ThreadPoolExecutor
    from permutect_b200.data.reads_dataset import MemoryMappedBatches, _parallel_copy
    from permutect_b200.synthetic import make_wgs_arrays
    rng = np.random.default_rng(3)
    src = rng.integers(0, 255, (10_001, 12), dtype=np.uint8)
    with ThreadPoolExecutor(4) as pool:
        for parts in (1, 3, 4, 7):
            dst = np.zeros_like(src)
            for job in _parallel_copy(pool, dst, src, parts):
                job.result()
            assert np.array_equal(dst, src), parts
    ia, fa, reads = make_wgs_arrays(1000, seed=8)
    n = ia[:, 0].astype(np.int64) + ia[:, 1].astype(np.int64)
    rows = np.arange(int(n.sum()), dtype=np.int64)           # stand-in for dataset-order rows: only the slicing is under test
    reads_ds = reads[rows]
    loader = MemoryMappedBatches(ia, fa, reads_ds, 300, pin_memory=False)
    assert len(loader) == 4
    seen_v = seen_r = 0
    for b in loader:
        nv = b.size()
        assert np.array_equal(b.int_tensor.numpy(), ia[seen_v:seen_v + nv])
        nr = int(n[seen_v:seen_v + nv].sum())
        assert b.reads.shape[0] == nr and np.array_equal(b.reads.numpy(), reads_ds[seen_r:seen_r + nr])
        seen_v += nv
        seen_r += nr
    assert seen_v == 1000 and seen_r == int(n.sum())
