"""The tcgen05 forward modes against the reference-generated fixtures and the oracle.
tf32x3 (split-precision TF32 on the tensor cores) must meet the fp32 contract (logits within 1e-3);
plain tf32 is a separately stated, looser mode."""
import numpy as np
import pytest
import torch

from golden_utils import load
from helpers import golden_batch, model_from_golden
from permutect_b200.engine import library as L
from permutect_b200.utils.enums import Epoch

pytestmark = pytest.mark.gpu
CASES = ["v040_seed0_b64", "v040_seed0_downsampled", "v040_perturbed_edge", "v040_two_sources"]


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    L.set_precision("fp32")


def _forward(case, mode):
    g = load(case)
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    batch = golden_batch(g, dev)
    L.set_precision(mode)
    with torch.inference_mode():
        out = model.compute_batch_output(batch)
    torch.cuda.synchronize()
    return g, out


@pytest.mark.parametrize("case", CASES)
def test_tf32x3_meets_the_fp32_contract(case):
    g, out = _forward(case, "tf32x3")
    np.testing.assert_allclose(out.logits_b.cpu().numpy(), g.out["logits_b"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(out.logits_bk.cpu().numpy(), g.out["logits_bk"], rtol=5e-5, atol=2e-3)
    np.testing.assert_allclose(out.features_be.cpu().numpy(), g.out["features_be"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(out.ref_features_be.cpu().numpy(), g.out["ref_features_be"], rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("case", CASES)
def test_tf32_mode_stated_tolerance(case):
    """Plain TF32 (10-bit mantissa) through ~45 layers: logits within 0.25 absolute, same sign away from zero."""
    g, out = _forward(case, "tf32")
    got, want = out.logits_b.cpu().numpy(), g.out["logits_b"]
    np.testing.assert_allclose(got, want, rtol=0, atol=0.25)
    far = np.abs(want) > 0.5
    assert np.array_equal(np.sign(got[far]), np.sign(want[far]))


def test_tensor_core_modes_on_a_large_ragged_batch():
    from oracle import artifact_oracle as orc
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    g = load("v040_perturbed_edge")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    ia, fa, reads = make_wgs_arrays(3000, seed=5)
    batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    raw = dict(reads_u8=reads, read_indices=None, ref_counts=ia[:, 0], alt_counts=ia[:, 1], info=fa[:, 6:].astype(np.float32),
               haplotypes=ia[:, 16:], labels=ia[:, 2], sources=ia[:, 4])
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, raw)
    L.set_precision("tf32x3")
    with torch.inference_mode():
        out = model.compute_batch_output(batch)
    torch.testing.assert_close(out.logits_b.cpu(), want["logits_b"], rtol=0, atol=1e-3)
    torch.testing.assert_close(out.features_be.cpu(), want["features_be"], rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("n_variants", [1, 15, 16, 17, 333, 5000])
def test_tensor_core_haplotype_cnn_matches_the_oracle(n_variants):
    """hap_cnn_tc_kernel (shifted-window MMAs, 16 variants per group): partial groups, single variants, many CTAs.
    Indel codes (4) included; the reference sequence embedding is info_seq's last d_seq columns."""
    from oracle import artifact_oracle as orc
    from permutect_b200.data.batch import Batch
    from permutect_b200.engine import function as engine
    from permutect_b200.synthetic import make_wgs_arrays
    g = load("v040_perturbed_edge")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    ia, fa, reads = make_wgs_arrays(n_variants, seed=11 + n_variants)
    rng = np.random.default_rng(n_variants)
    ia[:, 16:] = rng.integers(0, 5, ia[:, 16:].shape)          # every code incl. the deletion marker, at every position
    batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    raw = dict(reads_u8=reads, read_indices=None, ref_counts=ia[:, 0], alt_counts=ia[:, 1], info=fa[:, 6:].astype(np.float32),
               haplotypes=ia[:, 16:], labels=ia[:, 2], sources=ia[:, 4])
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, raw)
    d_info = model.descriptor().d_info
    for mode, tol in (("tf32x3", 2e-5), ("tf32", 5e-3)):
        L.set_precision(mode)
        with torch.inference_mode():
            out = engine.forward_call(model.descriptor(), model.flat_weights(), batch)
        torch.testing.assert_close(out["info_seq"][:, d_info:].cpu(), want["ref_seq_emb"], rtol=0, atol=tol, msg=lambda m: f"{mode}: {m}")
    L.set_precision("tf32x3")
    with torch.inference_mode():
        logits = model.compute_batch_output(batch).logits_b
    torch.testing.assert_close(logits.cpu(), want["logits_b"], rtol=0, atol=1e-3)


def test_cnn_outside_the_tensor_core_envelope_runs_the_simt_kernel():
    """small_hp has a 64-channel CNN with flatten > 1: the tensor-core modes must still give reference results."""
    g = load("small_hp")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    batch = golden_batch(g, dev)
    from permutect_b200.engine import function as engine
    outs = {}
    for mode in ("fp32", "tf32x3"):
        L.set_precision(mode)
        try:
            with torch.inference_mode():
                outs[mode] = engine.forward_call(model.descriptor(), model.flat_weights(), batch)["info_seq"].clone()
        except RuntimeError as e:       # the READ kernel may be outside its own envelope for this shape: must fail loudly, not silently
            assert "envelope" in str(e)
            return
    torch.testing.assert_close(outs["tf32x3"], outs["fp32"], rtol=1e-5, atol=1e-5)


def test_sharded_evaluation_is_bitwise_identical_to_one_batch():
    """Size-independent property (variants are independent): evaluating contiguous variant shards separately
    gives bit-identical logits to one big batch, however the tile planner packs them."""
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    g = load("v040_perturbed_edge")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    n = 40_000
    ia, fa, reads = make_wgs_arrays(n, seed=99)
    ref_off = np.concatenate(([0], np.cumsum(ia[:, 0].astype(np.int64))))
    alt_off = np.concatenate(([0], np.cumsum(ia[:, 1].astype(np.int64))))
    total_ref = int(ref_off[-1])
    L.set_precision("tf32x3")
    with torch.inference_mode():
        whole = model.compute_batch_output(Batch.from_arrays(ia, fa, reads).copy_to(dev))
        parts = []
        for v0, v1 in ((0, 1), (1, 12_345), (12_345, 12_409), (12_409, n)):
            sub = np.concatenate((reads[ref_off[v0]:ref_off[v1]], reads[total_ref + alt_off[v0]:total_ref + alt_off[v1]]))
            parts.append(model.compute_batch_output(Batch.from_arrays(ia[v0:v1], fa[v0:v1], sub).copy_to(dev)))
    for name in ("logits_b", "features_be", "ref_features_be", "logits_bk"):
        assert torch.equal(torch.cat([getattr(p, name) for p in parts]), getattr(whole, name)), name


# ---- sets longer than a tile on the tensor cores (BASELINE config 5; pmt_tc.cuh: LongTile) ----
LONG_SHAPES = [(7, 3), (125, 3), (700, 300), (10, 15), (0, 200), (130, 1), (3, 1), (257, 255), (9, 2), (2900, 1500), (128, 128), (129, 1)]


def _long_forward(raw, mode, simt=False):
    import os
    from helpers import batch_from_raw
    g = load("v040_perturbed_edge")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    batch = batch_from_raw(raw, dev)
    L.set_precision(mode)
    os.environ["PMT_LONG_SIMT"] = "1" if simt else "0"
    try:
        with torch.inference_mode():
            out = model.compute_batch_output(batch)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("PMT_LONG_SIMT", None)
    return g, out


@pytest.mark.parametrize("mode,atol", [("tf32x3", 1e-3), ("tf32", 0.25)])
def test_long_sets_on_the_tensor_cores_match_the_oracle(mode, atol):
    """Sets of up to 4 400 reads cut into single-side tiles that meet through global memory in every gated block
    (gated_mlp.py:236-248) and for the set sums (ragged_sets.py:144-158), interleaved with tile-sized sets, sets without ref
    reads and sets that only just pass a tile (125 + 3 padded, 128 + 128, 129 + 1)."""
    from oracle import artifact_oracle as orc
    from test_backward_gpu import _long_set_raw
    raw = _long_set_raw(11, LONG_SHAPES)
    g, out = _long_forward(raw, mode)
    with torch.no_grad():
        want = orc.forward(g.sd, g.hp, raw)
    torch.testing.assert_close(out.logits_b.cpu(), want["logits_b"], rtol=0, atol=atol)
    if mode == "tf32x3":
        torch.testing.assert_close(out.logits_bk.cpu(), want["logits_bk"], rtol=5e-5, atol=2e-3)
        torch.testing.assert_close(out.features_be.cpu(), want["features_be"], rtol=1e-4, atol=2e-4)
        torch.testing.assert_close(out.ref_features_be.cpu(), want["ref_features_be"], rtol=1e-4, atol=2e-4)
        # and against the FP32 long-set kernel on the same batch
        _, simt = _long_forward(raw, mode, simt=True)
        torch.testing.assert_close(out.logits_b, simt.logits_b, rtol=0, atol=1e-3)


def test_long_sets_on_the_tensor_cores_are_bitwise_reproducible():
    """Partials are summed in tile order, not in arrival order: two runs (and two different tile placements -- the list of
    long sets is built with an atomic append) give identical bits."""
    from test_backward_gpu import _long_set_raw
    rng = np.random.default_rng(3)
    shapes = [(int(rng.integers(0, 900)), int(rng.integers(1, 700))) for _ in range(60)]
    raw = _long_set_raw(13, shapes)
    _, a = _long_forward(raw, "tf32x3")
    _, b = _long_forward(raw, "tf32x3")
    for name in ("logits_b", "logits_bk", "features_be", "ref_features_be"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_packed_tile_planner_changes_nothing_but_the_tile_count(monkeypatch):
    """Inference batches of >= 2 M reads are tiled by plan_tiles_packed_kernel (best fit through a permutation of the
    variants) instead of in batch order.  There is no arithmetic across the read sets of a tile, so every output must be
    BITWISE what the sequential planner gives (PMT_TC_PACKED=0), including variants longer than a tile left to the long-set
    path and a last claim that is not full."""
    import bench
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    dev = torch.device("cuda:0")
    model = bench.make_model(dev)
    L.set_precision("tf32x3")
    ia, fa, reads = make_wgs_arrays(140_123, seed=77)
    batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    assert batch.reads.shape[0] >= 2_000_000
    with torch.inference_mode():
        monkeypatch.setenv("PMT_TC_PACKED", "0")
        want = model.compute_batch_output(batch)
        want = {k: getattr(want, k).clone() for k in ("logits_b", "logits_bk", "features_be", "ref_features_be", "outlier_binary_logits")}
        monkeypatch.delenv("PMT_TC_PACKED")
        got = model.compute_batch_output(batch)
    for k, w in want.items():
        assert torch.equal(getattr(got, k), w), k
    L.set_precision("fp32")
