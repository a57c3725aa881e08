"""Parity of the CUDA forward (through the C-ABI) against the oracle and the reference-generated
golden fixtures.  Tolerances: artifact logits 1e-3 absolute in fp32 (BASELINE.json north_star);
log-likelihood sums are O(100) so they carry a matching relative term."""
import numpy as np
import pytest
import torch

from golden_utils import CASES, load
from helpers import golden_batch, model_from_golden
from oracle import artifact_oracle as orc
from permutect_b200.utils.enums import Epoch

pytestmark = pytest.mark.gpu
LOGIT_ATOL = 1e-3


def _run(case):
    g = load(case)
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    batch = golden_batch(g, dev)
    with torch.inference_mode():
        out = model.compute_batch_output(batch)
    return g, model, batch, out


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference_golden(case):
    g, model, batch, out = _run(case)
    got = {k: getattr(out, k).cpu().numpy() for k in g.out}
    np.testing.assert_allclose(got["logits_b"], g.out["logits_b"], rtol=0, atol=LOGIT_ATOL)
    np.testing.assert_allclose(got["logits_bk"], g.out["logits_bk"], rtol=2e-5, atol=LOGIT_ATOL)
    np.testing.assert_allclose(got["outlier_binary_logits"], g.out["outlier_binary_logits"], rtol=2e-5, atol=LOGIT_ATOL)
    np.testing.assert_allclose(got["features_be"], g.out["features_be"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["ref_features_be"], g.out["ref_features_be"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["artifact_probs_b"], g.out["artifact_probs_b"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_oracle_and_filter_decisions(case):
    g, model, batch, out = _run(case)
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, g.raw())
    logits = out.logits_b.cpu()
    torch.testing.assert_close(logits, want["logits_b"], rtol=0, atol=LOGIT_ATOL)
    # what filter_variants persists and gates on (quirk Q6): sign and fp16 rounding of the logit
    far_from_boundary = (want["logits_b"].abs() > 2 * LOGIT_ATOL)
    assert torch.equal(torch.sign(logits)[far_from_boundary], torch.sign(want["logits_b"])[far_from_boundary])
    # measured 1.5 % at 1.25 M variants (quirk Q6); the fixtures hold 64 variants, so at most 3 % or two of them
    mismatch = int((logits.half() != want["logits_b"].half()).sum())
    assert mismatch <= max(2, 0.03 * len(logits)), f"{mismatch} of {len(logits)} fp16-rounded logits differ"


@pytest.mark.parametrize("case", ["v040_seed0_b64", "small_hp"])
def test_calculate_features_rows(case):
    g = load(case)
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    batch = golden_batch(g, dev)
    ref, alt, seq = model.calculate_features(batch)
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, g.raw())
    torch.testing.assert_close(ref.flattened_tensor_nf.cpu(), want["final_ref_re"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(alt.flattened_tensor_nf.cpu(), want["final_alt_re"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(seq.cpu(), want["ref_seq_emb"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(alt.means_over_sets().cpu(), want["features_be"], rtol=1e-4, atol=1e-4)


def test_decode_kernel_is_bit_exact():
    rng = np.random.default_rng(7)
    reads = rng.integers(0, 256, (5000, 12), dtype=np.uint8)
    reads[:6, 7] = [0, 96, 127, 128, 160, 255]          # quirk Q2 probe values
    from helpers import batch_from_raw
    raw = dict(reads_u8=reads, ref_counts=np.array([0]), alt_counts=np.array([5000]), labels=np.array([0]),
               info=np.zeros((1, 71), np.float32), haplotypes=np.zeros((1, 42), np.int16))
    batch = batch_from_raw(raw, torch.device("cuda:0"))
    got = batch.get_reads_re().cpu().numpy()
    np.testing.assert_array_equal(got, orc.decode_reads(reads))
    np.testing.assert_array_equal(got[:6, 56], [4.0, 7.0, 7.96875, 0.0, 1.0, 3.96875])


def test_large_ragged_batch_against_oracle():
    """A few thousand variants with the WGS-like count distribution: many tiles, many claims."""
    from permutect_b200.synthetic import make_wgs_arrays
    g = load("v040_perturbed_edge")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    ia, fa, reads = make_wgs_arrays(3000, seed=5)
    from permutect_b200.data.batch import Batch
    batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    with torch.inference_mode():
        out = model.compute_batch_output(batch)
    raw = dict(reads_u8=reads, read_indices=None, ref_counts=ia[:, 0], alt_counts=ia[:, 1], info=fa[:, 6:].astype(np.float32),
               haplotypes=ia[:, 16:], labels=ia[:, 2], sources=ia[:, 4])
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, raw)
    torch.testing.assert_close(out.logits_b.cpu(), want["logits_b"], rtol=0, atol=LOGIT_ATOL)
    torch.testing.assert_close(out.features_be.cpu(), want["features_be"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out.ref_features_be.cpu(), want["ref_features_be"], rtol=1e-4, atol=1e-4)


def test_prepared_forward_reuses_images_only_while_the_weights_are_unchanged():
    """Repeated inference goes through pmt_forward_prepared (packed weight images kept in the workspace); any parameter
    update, precision switch, other model or intervening backward must rebuild them."""
    from permutect_b200.data.batch import Batch
    from permutect_b200.engine import function as engine
    from permutect_b200.engine import library as L
    from permutect_b200.synthetic import make_wgs_arrays
    g = load("v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    other = model_from_golden(load("v040_perturbed_edge"), dev)
    other.set_epoch_type(Epoch.VALID)
    small = Batch.from_arrays(*make_wgs_arrays(300, seed=1)).copy_to(dev)
    big = Batch.from_arrays(*make_wgs_arrays(5000, seed=2)).copy_to(dev)
    fresh = model_from_golden(g, dev)
    fresh.set_epoch_type(Epoch.VALID)
    with torch.no_grad():
        fresh.reducer._model[-1].bias.add_(0.05)
    try:
        for mode in ("tf32x3", "fp32"):
            L.set_precision(mode)
            with torch.inference_mode():
                a1 = model.compute_batch_output(small).logits_b.clone()
                assert engine._PREPARED[dev] is not None
                b1 = model.compute_batch_output(big).logits_b.clone()          # prepared call, different batch size
                a2 = model.compute_batch_output(small).logits_b.clone()
                assert torch.equal(a1, a2)
                o1 = other.compute_batch_output(small).logits_b.clone()        # another model: images rebuilt
                a3 = model.compute_batch_output(small).logits_b.clone()
                assert torch.equal(a1, a3) and not torch.equal(a1, o1)
                with torch.no_grad():
                    model.reducer._model[-1].bias.add_(0.05)                   # in-place update bumps the version
                c1 = model.compute_batch_output(small).logits_b.clone()
                assert not torch.equal(a1, c1)
                assert torch.equal(c1, fresh.compute_batch_output(small).logits_b)
                assert not torch.equal(model.compute_batch_output(big).logits_b, b1)
                with torch.no_grad():
                    model.reducer._model[-1].bias.sub_(0.05)
    finally:
        L.set_precision("fp32")
