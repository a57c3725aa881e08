"""Inference caller tail: the oracle restatement against the reference-generated golden (CPU), and
pmt_pack_posterior through the C-ABI against both (GPU).  Integer / fp16-derived outputs are compared BIT-EXACTLY."""
import numpy as np
import pytest
import torch

import os

from oracle import posterior_tail_oracle as pto

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posterior_tail.npz")


def _golden():
    return dict(np.load(GOLDEN))


def _same_bits(a: np.ndarray, b: np.ndarray):
    assert a.dtype == b.dtype and a.shape == b.shape
    np.testing.assert_array_equal(a.view(np.uint8 if a.dtype.itemsize == 1 else f"u{a.dtype.itemsize}"),
                                  b.view(np.uint8 if b.dtype.itemsize == 1 else f"u{b.dtype.itemsize}"))


def test_oracle_matches_the_reference_generated_records():
    g = _golden()
    int_out, float_out = pto.posterior_arrays(g["int_array"], g["float_array"], g["logits"], g["embeddings"])
    _same_bits(int_out, g["int_out"])
    _same_bits(float_out, g["float_out"])          # NaN slots included: bit patterns must match
    assert float_out.dtype == np.float32 and (int_out[:, :2] == 0).all()


def test_oracle_on_empty_and_single_variant():
    for B in (0, 1):
        ia, fa = np.zeros((B, 58), np.int16), np.zeros((B, 77), np.float16)
        io, fo = pto.posterior_arrays(ia, fa, np.zeros(B, np.float32), np.ones((B, 10), np.float32))
        assert io.shape == (B, 58) and fo.shape == (B, 16)


@pytest.mark.gpu
def test_pack_posterior_kernel_is_bit_exact():
    from permutect_b200.data.batch import Batch
    from permutect_b200.tools.filter_variants import posterior_arrays_on_device
    g = _golden()
    dev = torch.device("cuda:0")
    B = len(g["logits"])
    reads = np.zeros((int(g["int_array"][:, 0].sum() + g["int_array"][:, 1].sum()), 12), np.uint8)
    batch = Batch.from_arrays(g["int_array"], g["float_array"], reads).copy_to(dev)
    int_out, float_out = posterior_arrays_on_device(batch, torch.from_numpy(g["logits"]).to(dev), torch.from_numpy(g["embeddings"]).to(dev))
    _same_bits(int_out.cpu().numpy(), g["int_out"])
    _same_bits(float_out.cpu().numpy(), g["float_out"])
    # a large random case against the oracle
    rng = np.random.default_rng(5)
    B = 100_003
    ia = rng.integers(-32768, 32767, (B, 58)).astype(np.int16)
    ia[:, 0], ia[:, 1] = 0, 1
    fa = rng.normal(0, 50, (B, 77)).astype(np.float16)
    logits = rng.normal(0, 10, B).astype(np.float32)
    emb = rng.normal(0, 3, (B, 10)).astype(np.float32)
    batch = Batch.from_arrays(ia, fa, np.zeros((B, 12), np.uint8)).copy_to(dev)
    io, fo = posterior_arrays_on_device(batch, torch.from_numpy(logits).to(dev), torch.from_numpy(emb).to(dev))
    want_i, want_f = pto.posterior_arrays(ia, fa, logits, emb)
    _same_bits(io.cpu().numpy(), want_i)
    _same_bits(fo.cpu().numpy(), want_f)


@pytest.mark.gpu
def test_generate_posterior_data_yields_reference_shaped_records():
    from golden_utils import load
    from helpers import model_from_golden
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    from permutect_b200.tools.filter_variants import generate_posterior_data
    from permutect_b200.utils.enums import Epoch
    g = load("v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    host = [Batch.from_arrays(*make_wgs_arrays(n, seed=70 + i)).pin_memory() for i, n in enumerate((100, 37))]
    data = list(generate_posterior_data(host, model))
    assert len(data) == 137
    with torch.inference_mode():
        out = model.compute_batch_output(host[0].copy_to(dev))
    d0 = data[0]
    assert d0.get_int_array()[0] == 0 and d0.get_int_array()[1] == 0 and d0.get_float_array().dtype == np.float32
    assert d0.get_float_array()[5] == np.float32(np.float16(out.logits_b[0].item()))
    np.testing.assert_array_equal(d0.get_float_array()[6:], out.features_be[0].cpu().numpy())
    assert len(d0.get_reads_array_re()) == 0
