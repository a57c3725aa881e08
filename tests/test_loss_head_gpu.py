"""The fused loss head (pmt_losses_forward / pmt_losses_backward through the C-ABI) against the oracle's
restatement of artifact_model.py:267-325 on random inputs: non-unit balancer weights, upstream gradients on
every loss component, one and three sources."""
import numpy as np
import pytest
import torch

from golden_utils import load
from helpers import batch_from_raw, model_from_golden
from oracle import artifact_oracle as orc
from permutect_b200.engine import function as engine
from permutect_b200.utils.enums import Epoch

pytestmark = pytest.mark.gpu


def _case(num_sources, B=300, seed=0):
    g = load("v040_two_sources" if num_sources > 1 else "v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    if num_sources > 1 and model.num_sources != num_sources:
        torch.manual_seed(3)
        model.reset_source_predictor(num_sources)
        model.source_predictor.set_adversarial_strength(0.3)
    model.set_epoch_type(Epoch.TRAIN)
    rng = np.random.default_rng(seed)
    E = model.reducer.output_dimension()
    raw = dict(ref_counts=np.ones(B, np.int64), alt_counts=rng.integers(1, 16, B), labels=rng.integers(0, 3, B),
               sources=rng.integers(0, num_sources, B), haplotypes=np.zeros((B, 42), np.int64), info=np.zeros((B, 71), np.float32),
               reads_u8=np.zeros((int(B + 0), 12), np.uint8))
    raw["reads_u8"] = np.zeros((int(raw["ref_counts"].sum() + raw["alt_counts"].sum()), 12), np.uint8)
    batch = batch_from_raw(raw, dev)
    t = lambda a: torch.from_numpy(a.astype(np.float32))
    out = dict(logits_b=t(rng.normal(0, 6, B)), outlier_binary_logits=t(rng.normal(6, 5, B)), features_be=t(rng.normal(0, 1.5, (B, E))))
    w, sw = t(rng.uniform(0.2, 3, B)), t(rng.uniform(0.2, 3, B))
    ups = [t(rng.normal(0, 1, B)) for _ in range(5)]
    return g, model, batch, raw, out, w, sw, ups


@pytest.mark.parametrize("num_sources", [1, 2, 3])
def test_loss_head_forward_and_backward_match_the_oracle(num_sources):
    g, model, batch, raw, out, w, sw, ups = _case(num_sources)
    dev = model._device
    sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    o_cpu = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    strength = float(model.source_predictor.gradient_reversal.alpha)
    want = orc.losses(sd, g.hp, raw, o_cpu, weights=w, source_weights=sw, num_sources=num_sources, source_adversarial_strength=strength)
    names = ["supervised_losses_b", "unsupervised_losses_b", "alt_count_losses_b", "source_prediction_losses_b", "total_losses_b"]
    scalar = sum((want[n] * u).sum() for n, u in zip(names, ups))
    head = [k for k in sd if k.startswith(("alt_count_predictor", "source_predictor")) and sd[k].requires_grad]
    grads = torch.autograd.grad(scalar, [o_cpu["logits_b"], o_cpu["outlier_binary_logits"], o_cpu["features_be"]] + [sd[k] for k in head],
                                allow_unused=True)

    flat = model.flat_weights()
    ins = [out["logits_b"].to(dev).requires_grad_(True), out["outlier_binary_logits"].to(dev).requires_grad_(True),
           out["features_be"].to(dev).requires_grad_(True)]
    got = engine.FusedLossFunction.apply(flat, ins[0], ins[1], ins[2], w.to(dev), sw.to(dev), model.loss_descriptor(), batch)
    for n, a in zip(names, got):
        np.testing.assert_allclose(a.detach().cpu().numpy(), want[n].detach().numpy(), rtol=2e-5, atol=2e-6, err_msg=n)
    sum((a * u.to(dev)).sum() for a, u in zip(got, ups)).backward()
    for name, a, b in zip(["d logits", "d outlier", "d features"], ins, grads[:3]):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.numpy(), rtol=2e-4, atol=2e-6, err_msg=name)
    params = dict(model.named_parameters())
    for k, b in zip(head, grads[3:]):
        if b is None:
            continue
        scale = max(float(b.abs().max()), 1e-6)
        err = float((params[k].grad.cpu() - b).abs().max()) / scale
        assert err < 1e-4, (k, err)


def test_loss_head_gradients_are_bitwise_reproducible():
    res = []
    for _ in range(2):
        g, model, batch, raw, out, w, sw, ups = _case(3, B=1000, seed=4)
        dev = model._device
        flat = model.flat_weights()
        got = engine.FusedLossFunction.apply(flat, out["logits_b"].to(dev), out["outlier_binary_logits"].to(dev),
                                             out["features_be"].to(dev), w.to(dev), sw.to(dev), model.loss_descriptor(), batch)
        got[4].sum().backward()
        res.append({n: p.grad.cpu().numpy().copy() for n, p in model.named_parameters() if p.grad is not None})
    for k in res[0]:
        np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)
