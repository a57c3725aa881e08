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
