"""Gradient parity of the TENSOR-CORE backward (pmt_backward with a tensor-core precision mode selected: recompute in the
split-precision mode, then data- and weight-gradient MMAs on TF32 operands rounded to nearest, fp32 accumulation) against
the gradients the unmodified reference left in param.grad (tests/golden) and against the FP32 kernels.

Stated tolerance of this mode: every tensor with eight or more entries within 4e-3 of its largest entry
(measured: <= 1.7e-3); the one-entry tensors (gate scalars, DenseSkipBlock.alpha, reg_weight) are sums in which the
TF32 rounding of the chain does not average out against the tensor's own size, so they are held to 4e-3 of the largest
small-tensor gradient of the same top-level module.  The FP32 mode keeps the 2e-3 contract (test_backward_gpu.py)."""
import numpy as np
import pytest
import torch

from golden_utils import load
from helpers import golden_batch, model_from_golden
from permutect_b200.engine import library as L
from permutect_b200.utils.enums import Epoch

pytestmark = pytest.mark.gpu
CASES = ["v040_seed0_b64", "v040_seed0_downsampled", "v040_perturbed_edge", "v040_two_sources"]
BIG, SMALL = 4e-3, 4e-3


@pytest.fixture(autouse=True)
def _tensor_core_mode():
    L.set_precision("tf32x3")
    yield
    L.set_precision("fp32")


def _grads(case):
    g = load(case)
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.TRAIN)
    batch = golden_batch(g, dev)
    losses = model.compute_batch_losses(model.compute_batch_output(batch), batch)
    losses.total_loss.backward()
    return g, {n: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
               for n, p in model.named_parameters()}


def _violations(got, want):
    family = {}
    for name, w in want.items():
        if w.size < 8:
            top = name.split(".")[0]
            family[top] = max(family.get(top, 1e-3), float(np.abs(w).max()))
    bad = []
    for name, w in want.items():
        err = float(np.abs(got[name] - w).max())
        if w.size >= 8:
            rel, tol = err / max(float(np.abs(w).max()), 1e-3), BIG
        else:
            rel, tol = err / family[name.split(".")[0]], SMALL
        if rel > tol:
            bad.append((rel, name))
    return sorted(bad, reverse=True)


@pytest.mark.parametrize("case", CASES)
def test_tensor_core_gradients_match_reference(case):
    g, got = _grads(case)
    bad = _violations(got, g.grad)
    assert not bad, "gradient error too large:\n" + "\n".join(f"{e:.3e} {n}" for e, n in bad[:25])


def test_tensor_core_gradients_are_bitwise_reproducible():
    _, a = _grads("v040_perturbed_edge")
    _, b = _grads("v040_perturbed_edge")
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_large_batch_several_recompute_ranges_against_fp32_kernels(monkeypatch):
    """40 000 WGS-shaped variants: more tiles than one recompute range holds (the backward walks the tile list in bounded
    ranges -- 8 192 tiles by default, cut to 1 024 here so that this batch takes six of them), every slot of every CTA busy;
    compared with the FP32 kernels, which the reference pins."""
    monkeypatch.setenv("PMT_BWD_CHUNK_TILES", "1024")
    import bench
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    dev = torch.device("cuda:0")
    model = bench.make_model(dev)
    model.set_epoch_type(Epoch.TRAIN)
    batch = Batch.from_arrays(*make_wgs_arrays(40000, seed=3000)).copy_to(dev)

    def grads(mode):
        L.set_precision(mode)
        for p in model.parameters():
            p.grad = None
        model.compute_batch_losses(model.compute_batch_output(batch), batch).total_loss.backward()
        return {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}

    tc, fp32, tc2 = grads("tf32x3"), grads("fp32"), grads("tf32x3")
    bad = _violations(tc, fp32)
    assert not bad, "gradient error too large:\n" + "\n".join(f"{e:.3e} {n}" for e, n in bad[:25])
    for k in tc:
        np.testing.assert_array_equal(tc[k], tc2[k], err_msg=k)
    flat = lambda d: np.concatenate([d[k].ravel() for k in sorted(d)])
    rel = np.linalg.norm(flat(tc) - flat(fp32)) / np.linalg.norm(flat(fp32))
    assert rel < 1e-3, f"whole-gradient relative error {rel:.3e}"


def test_mixed_long_and_tile_sized_sets():
    """Sets longer than a tile keep the FP32 long-set kernels; the tile-sized sets around them take the tensor cores."""
    from helpers import batch_from_raw
    from oracle import artifact_oracle as orc
    from test_backward_gpu import _long_set_raw
    g = load("v040_perturbed_edge")
    raw = _long_set_raw(11, [(7, 3), (125, 3), (300, 100), (10, 15), (0, 20), (130, 1), (3, 1), (9, 2), (0, 1), (40, 60)])
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.TRAIN)
    batch = batch_from_raw(raw, dev)
    model.compute_batch_losses(model.compute_batch_output(batch), batch).total_loss.backward()
    names = [n for n, _ in model.named_parameters()]
    _, _, want = orc.loss_and_grads(g.sd, g.hp, raw, names)
    got = {n: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
           for n, p in model.named_parameters()}
    bad = _violations(got, {k: v.numpy() for k, v in want.items()})
    assert not bad, "gradient error too large:\n" + "\n".join(f"{e:.3e} {n}" for e, n in bad[:25])


@pytest.mark.parametrize("layers", [
    # 16 channels, kernel sizes 3 / 5 / 3 / 3, the flatten keeps two positions (the Linear is a convolution of length 2)
    ['convolution/kernel_size=3/out_channels=16', 'selu', 'pool/kernel_size=2/stride=1',
     'convolution/kernel_size=5/out_channels=16', 'selu', 'pool/kernel_size=1',
     'convolution/kernel_size=3/out_channels=24', 'selu', 'pool/kernel_size=2',
     'convolution/kernel_size=3/out_channels=8', 'selu', 'pool/kernel_size=2',
     'flatten', 'linear/out_features=10'],
    # no pooling after the first convolution, one pooled layer, a SELU-free last convolution
    ['convolution/kernel_size=4/out_channels=32', 'selu',
     'convolution/kernel_size=3/out_channels=32', 'selu', 'pool/kernel_size=2',
     'convolution/kernel_size=2/out_channels=12',
     'flatten', 'linear/out_features=10'],
])
def test_haplotype_cnn_backward_other_shapes_against_the_oracle(layers):
    """cnn_backward_mma_kernel (pmt_cnn_bwd.cu) is driven by the layer program, not by the v0.4.0 shapes: other channel
    counts, kernel sizes, pool placements and a flatten that keeps more than one position, against the oracle's autograd."""
    import bench
    from helpers import batch_from_raw, params_from_hp
    from oracle import artifact_oracle as orc
    from permutect_b200.architecture.artifact_model import ArtifactModel
    from permutect_b200.synthetic import make_wgs_arrays
    hp = dict(bench.V040, ref_seq_layer_strings=layers, source_adversarial_strength=0.0)
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = ArtifactModel(params_from_hp(hp), 61, 71, 42, device=dev)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() == 0:
                p.copy_(0.05 + 0.1 * torch.rand((), generator=gen))
    model.set_epoch_type(Epoch.TRAIN)
    assert L.backward_kernels(model.descriptor()) == {"reads_tc": True, "cnn_tc": True}     # not the SIMT fall-back
    ia, fa, reads = make_wgs_arrays(203, seed=31)
    raw = bench.oracle_inputs(ia, fa, reads)
    batch = batch_from_raw(raw, dev)
    model.compute_batch_losses(model.compute_batch_output(batch), batch).total_loss.backward()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    _, _, want = orc.loss_and_grads(sd, hp, raw, names)
    got = {n: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
           for n, p in model.named_parameters()}
    bad = _violations(got, {k: v.numpy() for k, v in want.items()})
    assert not bad, "gradient error too large:\n" + "\n".join(f"{e:.3e} {n}" for e, n in bad[:25])
    cnn = [n for n in names if n.startswith("haplotypes_cnn")]
    assert cnn and all(np.abs(got[n]).max() > 0 for n in cnn)


def test_saved_forward_and_recompute_give_the_same_outputs_and_gradients(monkeypatch):
    """pmt_forward_train (the training forward keeps the tile list, every tile's operand panels and the CNN's activations;
    pmt_backward skips its recompute passes) against the recompute path (PERMUTECT_B200_TRAIN_SAVED=0): same kernels over
    the same deterministic tile list, so outputs and every gradient are BITWISE equal."""
    import bench
    from permutect_b200.data.batch import Batch, DownsampledBatch
    from permutect_b200.synthetic import make_wgs_arrays
    dev = torch.device("cuda:0")
    model = bench.make_model(dev)
    model.set_epoch_type(Epoch.TRAIN)
    parent = Batch.from_arrays(*make_wgs_arrays(3000, seed=41)).copy_to(dev)
    frac = torch.full((3000,), 0.7, device=dev)
    batch = DownsampledBatch(parent, frac, frac, seed=9)
    desc = model.descriptor()
    assert L.load().pmt_train_saved_bytes(__import__("ctypes").byref(desc), __import__("ctypes").byref(batch.pmt_batch())) > 0

    def run():
        for p in model.parameters():
            p.grad = None
        out = model.compute_batch_output(batch)
        model.compute_batch_losses(out, batch).total_loss.backward()
        return (out.logits_b.detach().clone(), out.features_be.detach().clone(),
                {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})

    logits_s, feat_s, grads_s = run()
    monkeypatch.setenv("PERMUTECT_B200_TRAIN_SAVED", "0")
    logits_r, feat_r, grads_r = run()
    assert torch.equal(logits_s, logits_r) and torch.equal(feat_s, feat_r)
    assert grads_s.keys() == grads_r.keys()
    for k in grads_s:
        assert torch.equal(grads_s[k], grads_r[k]), k
    # the FP32 mode has no saved-forward path: the library says so and the call falls back
    L.set_precision("fp32")
    assert L.load().pmt_train_saved_bytes(__import__("ctypes").byref(desc), __import__("ctypes").byref(batch.pmt_batch())) == 0
