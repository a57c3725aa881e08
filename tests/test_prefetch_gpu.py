"""prefetch_generator: side-stream H2D through the ring of device buffers must deliver every batch intact, in order,
for ragged batch sizes (ring buffers grow) and for more batches than ring slots."""
import numpy as np
import pytest
import torch

from golden_utils import load
from helpers import model_from_golden
from permutect_b200.data.batch import Batch
from permutect_b200.data.prefetch_generator import prefetch_generator
from permutect_b200.synthetic import make_wgs_arrays
from permutect_b200.utils.enums import Epoch

GPU = pytest.mark.gpu


@GPU
def test_prefetched_batches_give_the_same_logits_as_direct_copies():
    g = load("v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    sizes = [700, 64, 1500, 1, 333, 2048, 900]
    host = [Batch.from_arrays(*make_wgs_arrays(n, seed=50 + i)).pin_memory() for i, n in enumerate(sizes)]
    with torch.inference_mode():
        want = [model.compute_batch_output(b.copy_to(dev)).logits_b.cpu() for b in host]
        for depth in (1, 2, 3):
            got = []
            for b in prefetch_generator(host, dev, depth=depth):
                assert b.reads.device == dev and b.size() == sizes[len(got)]
                got.append(model.compute_batch_output(b).logits_b.cpu())
            assert len(got) == len(want)
            for a, w in zip(got, want):
                assert torch.equal(a, w)


def test_prefetch_generator_has_no_cpu_path():
    import pytest as _pytest
    host = [Batch.from_arrays(*make_wgs_arrays(10, seed=1))]
    with _pytest.raises(RuntimeError, match="no CPU path"):
        list(prefetch_generator(host, torch.device("cpu")))


def test_make_optimizer_has_no_cpu_path():
    import pytest as _pytest
    from permutect_b200.training.step import make_optimizer
    with _pytest.raises(RuntimeError, match="no CPU optimiser path"):
        make_optimizer(torch.nn.Linear(3, 2))


@GPU
def test_ring_is_shared_between_passes_and_survives_early_exit_and_nesting():
    g = load("v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    sizes = [300, 1200, 64, 800, 5]
    host = [Batch.from_arrays(*make_wgs_arrays(n, seed=70 + i)).pin_memory() for i, n in enumerate(sizes)]
    with torch.inference_mode():
        want = [model.compute_batch_output(b.copy_to(dev)).logits_b.cpu() for b in host]
        for _ in range(3):                                   # consecutive passes reuse the ring
            got = [model.compute_batch_output(b).logits_b.cpu() for b in prefetch_generator(host, dev)]
            assert all(torch.equal(a, w) for a, w in zip(got, want))
        for b in prefetch_generator(host, dev):              # the consumer stops early
            break
        got = [model.compute_batch_output(b).logits_b.cpu() for b in prefetch_generator(host, dev)]
        assert all(torch.equal(a, w) for a, w in zip(got, want))
        outer_got, inner_got = [], []
        for i, b in enumerate(prefetch_generator(host, dev)):    # a second generator while the first is alive
            if i == 1:
                inner_got = [model.compute_batch_output(c).logits_b.cpu() for c in prefetch_generator(host, dev)]
            outer_got.append(model.compute_batch_output(b).logits_b.cpu())
        assert all(torch.equal(a, w) for a, w in zip(outer_got, want))
        assert all(torch.equal(a, w) for a, w in zip(inner_got, want))
    # a ring filled under inference_mode must be refillable by a training pass (its buffers are ordinary tensors)
    got = [model.compute_batch_output(b).logits_b.detach().cpu() for b in prefetch_generator(host, dev)]
    assert all(torch.equal(a, w) for a, w in zip(got, want))


@GPU
def test_staged_dataset_batches_equal_the_plain_slices_and_feed_the_model():
    """MemoryMappedBatches(staging_threads > 0): a producer thread stages every batch into a ring of pinned buffers with
    copy threads.  Consumed through prefetch_generator (which leaves the H2D event the ring waits for before it refills a
    slot) over more batches than the ring holds, twice, the device batches equal the plain loader's and the logits too."""
    import numpy as np
    import bench
    from permutect_b200.data.prefetch_generator import prefetch_generator
    from permutect_b200.data.reads_dataset import MemoryMappedBatches
    from permutect_b200.engine import library as L
    from permutect_b200.synthetic import make_wgs_arrays
    dev = torch.device("cuda:0")
    ia, fa, reads = make_wgs_arrays(9000, seed=17)
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    n = ref_c + alt_c
    v = np.repeat(np.arange(len(n)), n)
    k = np.arange(int(n.sum())) - np.repeat(np.concatenate(([0], np.cumsum(n)))[:-1], n)
    reads_ds = reads[np.where(k < ref_c[v], ref_off[v] + k, ref_off[-1] + alt_off[v] + (k - ref_c[v]))]
    plain = MemoryMappedBatches(ia, fa, reads_ds, 700, pin_memory=True)
    staged = MemoryMappedBatches(ia, fa, reads_ds, 700, pin_memory=True, staging_threads=3, prefetch=2)
    assert len(plain) == len(staged) == 13                      # more batches than the ring's five slots
    model = bench.make_model(dev)
    L.set_precision("tf32x3")
    try:
        with torch.inference_mode():
            want = [(b.reads.clone(), b.int_tensor.clone(), model.compute_batch_output(b).logits_b.clone())
                    for b in prefetch_generator(plain, dev)]
            for _ in range(2):
                got = 0
                for b, (reads_w, int_w, logits_w) in zip(prefetch_generator(staged, dev), want):
                    assert b.reads.is_cuda and torch.equal(b.reads, reads_w) and torch.equal(b.int_tensor, int_w)
                    assert torch.equal(model.compute_batch_output(b).logits_b, logits_w)
                    got += 1
                assert got == 13
        first = next(iter(staged))
        assert first.reads.is_pinned() and first.int_tensor.is_pinned() and first.float_tensor.is_pinned()
    finally:
        L.set_precision("fp32")


@GPU
def test_dataset_arrays_registered_in_place_give_zero_copy_pinned_batches():
    """MemoryMappedBatches(pin_memory="register"): the arrays are page-locked in place, batches are views of them (no copy),
    DMA-able, and equal to the plain loader's."""
    import numpy as np
    from permutect_b200.data.reads_dataset import MemoryMappedBatches
    from permutect_b200.synthetic import make_wgs_arrays
    ia, fa, reads = make_wgs_arrays(5000, seed=23)
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    n = ref_c + alt_c
    v = np.repeat(np.arange(len(n)), n)
    k = np.arange(int(n.sum())) - np.repeat(np.concatenate(([0], np.cumsum(n)))[:-1], n)
    reads_ds = np.ascontiguousarray(reads[np.where(k < ref_c[v], ref_off[v] + k, ref_off[-1] + alt_off[v] + (k - ref_c[v]))])
    ia, fa = np.ascontiguousarray(ia), np.ascontiguousarray(fa)
    plain = list(MemoryMappedBatches(ia, fa, reads_ds, 1200, pin_memory=False))
    loader = MemoryMappedBatches(ia, fa, reads_ds, 1200, pin_memory="register")
    assert loader.registered
    got = list(loader)
    assert len(got) == len(plain) == 5
    for a, b in zip(got, plain):
        assert a.reads.is_pinned() and a.int_tensor.is_pinned() and a.float_tensor.is_pinned()
        assert torch.equal(a.reads, b.reads) and torch.equal(a.int_tensor, b.int_tensor)
        assert torch.equal(a.float_tensor.view(torch.int16), b.float_tensor.view(torch.int16))      # bit patterns: the scalar block holds NaNs
    assert got[1].int_tensor.data_ptr() == ia[1200:].ctypes.data          # a view of the dataset's own memory
    dev = torch.device("cuda:0")
    assert torch.equal(got[2].copy_to(dev).reads.cpu(), plain[2].reads)
    # the page lock ends with the loader (a freed range that still counted as pinned would poison later host tensors)
    torch.cuda.synchronize()
    del got, loader
    import gc
    gc.collect()
    assert not torch.from_numpy(ia[:1]).is_pinned() and not torch.from_numpy(reads_ds[:1]).is_pinned()


@GPU
def test_dataset_loaded_into_pinned_memory_gives_pinned_views():
    import numpy as np
    from permutect_b200.data.reads_dataset import MemoryMappedBatches, load_into_pinned_memory
    ia, fa, reads = make_wgs_arrays(800, seed=29)
    pinned = load_into_pinned_memory(ia)
    assert pinned.dtype == ia.dtype and np.array_equal(pinned, ia) and torch.from_numpy(pinned[100:200]).is_pinned()
    n = ia[:, 0].astype(np.int64) + ia[:, 1].astype(np.int64)
    loader = MemoryMappedBatches(pinned, load_into_pinned_memory(fa), load_into_pinned_memory(reads[: int(n.sum())]), 300,
                                 pin_memory="register")
    assert loader.registered and all(b.int_tensor.is_pinned() and b.reads.is_pinned() for b in loader)
