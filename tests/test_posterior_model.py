"""PosteriorModel inference: the oracle against the reference-generated golden on the CPU, the kernel
(pmt_posterior_log_posteriors through the C-ABI, behind the reference's PosteriorModel surface) against both on the GPU.
Tolerance: every table entry is a sum of ~7 fp32 lgamma terms as large as lgamma(depth + 1) (279 000 at depth 30 000,
one ulp = 0.03) that cancel to O(10); the reference's own fp32 result carries that rounding, so rows are held to
2e-5 relative + 2e-4 + 8 ulp(lgamma(depth + 2)) absolute (3e-4 at depth 100, 6e-3 at 1 000, 0.27 at 30 000).
Calls (argmax) identical wherever the top two posteriors differ by more than that."""
import os

import numpy as np
import pytest
import torch

from oracle import posterior_model_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posterior_model.npz")
CASES = ["default", "no_context", "het_beta", "no_germline"]
TABLES = ["log_priors_bc", "spectra_log_lks_bc", "normal_log_lks_bc", "log_posteriors_bc"]


def _row_atol(int_array):
    from scipy.special import gammaln
    depth = np.maximum(int_array[:, 5], int_array[:, 7]).astype(np.float64)
    return (2e-4 + 8 * np.finfo(np.float32).eps * gammaln(depth + 2.0))[:, None]


def _assert_tables_close(got, want, int_array, msg):
    excess = np.abs(got - want) - (_row_atol(int_array) + 2e-5 * np.abs(want))
    worst = np.unravel_index(np.argmax(excess), excess.shape)
    assert excess.max() <= 0, f"{msg}: row {worst[0]} call {worst[1]}: got {got[worst]} want {want[worst]} (depth {int_array[worst[0], 5]})"


def _case(z, tag):
    sd = {k[len(tag) + 4:]: z[k] for k in z.files if k.startswith(tag + "/sd/")}
    cfg = z[tag + "/cfg"]
    return sd, bool(cfg[0]), (None if cfg[1] < 0 else float(cfg[1])), bool(cfg[2])


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_golden(tag):
    z = np.load(GOLDEN)
    sd, no_germline, het_beta, context = _case(z, tag)
    out = orc.log_posterior_and_ingredients(sd, z["int_array"], z["float_array"], no_germline, het_beta, context)
    for k in TABLES + ["posterior_probabilities_bc", "error_probabilities_b"]:
        np.testing.assert_allclose(out[k].numpy(), z[f"{tag}/out/{k}"], rtol=2e-6, atol=1e-5, err_msg=k)


def test_state_dict_keys_and_initial_values_match_the_reference():
    from permutect_b200.architecture.posterior_model import PosteriorModel
    z = np.load(GOLDEN)
    model = PosteriorModel(-10.0, -10.0, device="cpu")
    sd = model.state_dict()
    want = {k[len("default/sd/"):]: z[k] for k in z.files if k.startswith("default/sd/")}
    assert set(sd.keys()) == set(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k].shape), k
    # un-perturbed constants of the reference's constructor (somatic_spectrum.py:62-69)
    np.testing.assert_allclose(sd["spectra.somatic_spectrum.log_background_weight"].numpy(), np.log(np.float32(0.0001)), rtol=1e-6)
    assert float(sd["spectra.somatic_spectrum.background_alpha"][0]) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_kernel_matches_reference_golden_and_oracle(tag):
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    z = np.load(GOLDEN)
    sd, no_germline, het_beta, context = _case(z, tag)
    dev = torch.device("cuda:0")
    model = PosteriorModel(-3.0, -4.0, no_germline_mode=no_germline, device=dev, het_beta=het_beta)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    if not context:
        model.priors.disable_context_dependent_snv_priors()
    batch = PosteriorBatch(z["int_array"], z["float_array"], dev)
    got = dict(zip(TABLES, model.log_posterior_and_ingredients(batch)))
    for k in TABLES:
        _assert_tables_close(got[k].cpu().numpy(), z[f"{tag}/out/{k}"], z["int_array"], k)
    probs = model.posterior_probabilities_bc(batch).cpu().numpy()
    shallow = z["int_array"][:, 5] <= 400            # probabilities: rows whose log-table tolerance is below 3e-3
    np.testing.assert_allclose(probs[shallow], z[f"{tag}/out/posterior_probabilities_bc"][shallow], rtol=0, atol=3e-3)
    np.testing.assert_allclose(model.error_probabilities_b(batch).cpu().numpy()[shallow],
                               z[f"{tag}/out/error_probabilities_b"][shallow], rtol=0, atol=3e-3)
    want_post = z[f"{tag}/out/log_posteriors_bc"]
    top2 = np.sort(want_post, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 2 * _row_atol(z["int_array"])[:, 0] + 1e-3
    assert decided.mean() > 0.9
    assert np.array_equal(probs.argmax(1)[decided], want_post.argmax(1)[decided])       # identical calls
    # fp16 float block (an ArtifactModel Batch keeps its side arrays in fp16): same tables, the inputs are fp16-exact
    half = PosteriorBatch(z["int_array"], z["float_array"].astype(np.float16), dev)
    _assert_tables_close(model.log_relative_posteriors_bc(half).cpu().numpy(), want_post, z["int_array"], "fp16 float block")


@pytest.mark.gpu
def test_kernel_matches_oracle_on_a_large_random_batch_and_rejects_cpu():
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    z = np.load(GOLDEN)
    sd, _, _, _ = _case(z, "default")
    rng = np.random.default_rng(3)
    reps = 40
    ia = np.tile(z["int_array"], (reps, 1))
    fa = np.tile(z["float_array"], (reps, 1))
    depth = rng.integers(1, 2000, len(ia))
    ia[:, 5] = depth
    ia[:, 6] = np.maximum(1, rng.binomial(depth, rng.uniform(0.001, 0.9, len(ia))))
    ia[:, 7] = rng.integers(0, 500, len(ia))
    ia[:, 8] = rng.binomial(ia[:, 7], 0.03)
    fa[:, 5] = np.float16(rng.normal(0, 6, len(ia)))
    dev = torch.device("cuda:0")
    model = PosteriorModel(-3.0, -4.0, device=dev)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    got = model.log_relative_posteriors_bc(PosteriorBatch(ia, fa, dev)).cpu().numpy()
    want = orc.log_posterior_and_ingredients(sd, ia, fa)["log_posteriors_bc"].numpy()
    _assert_tables_close(got, want, ia, "random batch")
    with pytest.raises(RuntimeError, match="CUDA"):
        model.log_relative_posteriors_bc(PosteriorBatch(ia[:4], fa[:4]))


def test_gradient_oracle_matches_reference_autograd():
    """The loss of learn_priors_and_spectra and what its backward leaves in .grad of the spectra parameters
    (tests/golden/make_posterior_model_golden.py ran them through the unmodified reference)."""
    z = np.load(GOLDEN)
    for tag in ("default", "het_beta", "no_germline"):
        sd, no_germline, het_beta, context = _case(z, tag)
        loss, _, raw = orc.negative_log_evidence_and_grads(sd, z["int_array"], z["float_array"], no_germline, het_beta, context)
        np.testing.assert_allclose(float(loss), float(z[f"{tag}/loss"]), rtol=1e-6)
        for name, g in raw.items():
            np.testing.assert_allclose(g.numpy(), z[f"{tag}/grad/{name}"], rtol=1e-5, atol=1e-7, err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["default", "het_beta", "no_germline"])
def test_fit_step_matches_reference_loss_gradients_and_totals(tag):
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    z = np.load(GOLDEN)
    sd, no_germline, het_beta, context = _case(z, tag)
    keep = z["int_array"][:, 5] <= 400          # rows whose fp32 lgamma rounding stays below 3e-3 (module docstring)
    ia, fa = z["int_array"][keep], z["float_array"][keep]
    dev = torch.device("cuda:0")
    model = PosteriorModel(-3.0, -4.0, no_germline_mode=no_germline, device=dev, het_beta=het_beta)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    totals = torch.zeros((5, 5), device=dev)
    snv = torch.zeros((5, 5, 5, 5), device=dev)
    ctx = torch.zeros((5, 5, 5, 5), device=dev)
    loss = model.negative_log_evidence(PosteriorBatch(ia, fa, dev), totals, snv, ctx)
    loss.backward()
    want_loss, _, want_raw = orc.negative_log_evidence_and_grads(sd, ia, fa, no_germline, het_beta, context)
    np.testing.assert_allclose(float(loss.detach()), float(want_loss), rtol=2e-5, atol=2e-4)
    got = {k: v for k, v in model.named_parameters()}
    for name, w in want_raw.items():
        g = got[name].grad.cpu().numpy()
        scale = max(float(np.abs(w.numpy()).max()), 1e-4)
        assert np.abs(g - w.numpy()).max() <= 2e-3 * scale + 1e-6, (name, np.abs(g - w.numpy()).max(), scale)
    post = orc.log_posterior_and_ingredients(sd, ia, fa, no_germline, het_beta, context)["posterior_probabilities_bc"].numpy()
    want_totals = np.zeros((5, 5))
    np.add.at(want_totals, ia[:, 3].astype(int), post)
    np.testing.assert_allclose(totals.cpu().numpy(), want_totals, rtol=1e-3, atol=2e-3)
    is_snv = ia[:, 3] == 0
    assert abs(float(ctx.sum()) - is_snv.sum()) < 1e-3 and abs(float(snv.sum()) - post[is_snv, 0].sum()) < 2e-3
    # reproducible: the fixed-order sums give the same loss and gradients again
    model.zero_grad()
    loss2 = model.negative_log_evidence(PosteriorBatch(ia, fa, dev))
    loss2.backward()
    assert float(loss2.detach()) == float(loss.detach())
    for name in want_raw:
        assert torch.equal(got[name].grad, dict(model.named_parameters())[name].grad)


@pytest.mark.gpu
def test_learn_priors_and_spectra_lowers_the_negative_log_evidence():
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    z = np.load(GOLDEN)
    keep = z["int_array"][:, 5] <= 400
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = PosteriorModel(-10.0, -10.0, device=dev)
    batches = [PosteriorBatch(z["int_array"][keep][i::3], z["float_array"][keep][i::3], dev) for i in range(3)]
    history = model.learn_priors_and_spectra(batches, num_iterations=15, ignored_to_non_ignored_ratio=10.0, learning_rate=0.05)
    assert len(history) == 15 and np.all(np.isfinite(history)) and history[-1] < history[0]
    pri = model.priors.log_priors_vc.detach().cpu().numpy()
    assert np.all(pri[:, 2] == 0) and np.all(pri[:, 3] == 0) and np.all(pri[:, [0, 1, 4]] < 0)
    # context-dependent SNV priors are on from the second half (posterior_model.py:125-127): the SNV entries hold the fitted
    # log mutation rates (snv_context_priors.py), the deletion / ref == alt entries what the first half's M steps filled in
    assert model.priors.use_context_dependent_snv_priors
    rrra = model.priors.somatic_snv_log_priors_rrra.detach().cpu()
    from permutect_b200.architecture.snv_context_priors import convert_rrra_tensor_to_sc
    sc = convert_rrra_tensor_to_sc(rrra)
    assert torch.all(torch.isfinite(sc)) and torch.all(sc < 0) and float(sc.max() - sc.min()) > 0
    assert float(rrra[4, 0, 0, 0]) == float(rrra[0, 1, 0, 1]) and float(rrra[4, 0, 0, 0]) < 0
    # and the kernel consumes them: posteriors stay normalised and finite
    probs = model.posterior_probabilities_bc(batches[0])
    assert torch.all(torch.isfinite(probs)) and torch.allclose(probs.sum(1), torch.ones_like(probs[:, 0]), atol=1e-4)


@pytest.mark.gpu
def test_artifact_model_records_flow_into_the_posterior_model():
    """filter_variants' chain (tools/filter_variants.py:292-320 then posterior_model.py:58-67): ArtifactModel logits ->
    posterior records (pmt_pack_posterior) -> error probabilities (pmt_posterior_log_posteriors), against the oracles fed
    with the same records."""
    import bench
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    from permutect_b200.tools.filter_variants import generate_posterior_arrays
    from permutect_b200.utils.enums import Epoch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    ia, fa, reads = make_wgs_arrays(500, seed=21)
    depth = rng.integers(10, 300, len(ia))
    ia[:, 5], ia[:, 6] = depth, np.maximum(1, rng.binomial(depth, 0.2))
    ia[:, 7] = rng.integers(0, 200, len(ia))
    ia[:, 8] = rng.binomial(ia[:, 7], 0.02)
    fa[:, 0], fa[:, 1] = -rng.uniform(1, 40, len(ia)), -rng.uniform(0, 5, len(ia))
    fa[:, 2], fa[:, 3], fa[:, 4] = 10.0 ** rng.uniform(-4, -0.5, len(ia)), rng.uniform(0.1, 0.5, len(ia)), rng.uniform(0.1, 0.5, len(ia))
    model = bench.make_model(dev)
    model.set_epoch_type(Epoch.VALID)
    loader = [Batch.from_arrays(ia[:300], fa[:300], np.concatenate((reads[:ia[:300, 0].sum()],
                                reads[ia[:, 0].sum():ia[:, 0].sum() + ia[:300, 1].sum()])))]
    (int_out, float_out), = list(generate_posterior_arrays(loader, model, dev))
    assert int_out.shape[0] == 300 and float_out.dtype == np.float32
    pm = PosteriorModel(-10.0, -10.0, device=dev)
    err = pm.error_probabilities_b(PosteriorBatch(int_out, float_out, dev)).cpu().numpy()
    sd = {k: v.detach().cpu() for k, v in pm.state_dict().items()}
    want = orc.log_posterior_and_ingredients(sd, int_out, float_out)["error_probabilities_b"].numpy()
    np.testing.assert_allclose(err, want, rtol=0, atol=3e-3)
    assert np.all((err >= 0) & (err <= 1)) and 0.01 < err.mean() < 0.999


def _reference_roc_loop(artifact_probs, recall_weight=1.0):
    """get_theoretical_roc_data's loop (metrics/plotting.py:153-190), second output only."""
    beta_sqr = recall_weight ** 2
    artifact_probs = sorted(artifact_probs)
    total_artifact = sum(artifact_probs) + 0.0001
    total_non_artifact = len(artifact_probs) - total_artifact + 0.0002
    art_found, non_art_found = total_artifact, 0
    best_threshold, best_hm = (0, 1, 0), 0
    for prob in artifact_probs:
        art_found -= prob
        non_art_found += 1 - prob
        tp, fp = non_art_found, total_artifact - art_found
        sensitivity, precision = tp / total_non_artifact, tp / (tp + fp)
        hm = (1 + beta_sqr) * sensitivity * precision / (sensitivity + (beta_sqr * precision) + 0.0001)
        if hm > best_hm:
            best_hm, best_threshold = hm, (prob, precision, sensitivity)
    return best_threshold


def test_probability_threshold_search_matches_the_reference_loop():
    from permutect_b200.architecture.posterior_model import theoretical_roc_best_threshold
    rng = np.random.default_rng(8)
    for n, recall_weight in ((1, 1.0), (7, 1.0), (500, 1.0), (5000, 2.0), (5000, 0.5)):
        probs = np.clip(rng.beta(0.3, 0.3, n), 0, 1).astype(np.float32)
        want = _reference_roc_loop([float(x) for x in probs], recall_weight)
        got = theoretical_roc_best_threshold(torch.from_numpy(probs), recall_weight)
        assert got[0] == want[0], (n, got, want)
        np.testing.assert_allclose(got[1:], want[1:], rtol=1e-9)
    assert theoretical_roc_best_threshold(torch.zeros(0)) == (0, 1, 0)


@pytest.mark.gpu
def test_calculate_probability_thresholds_per_variant_type():
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    from permutect_b200.utils.enums import Variation
    z = np.load(GOLDEN)
    sd, _, _, _ = _case(z, "default")
    dev = torch.device("cuda:0")
    model = PosteriorModel(-3.0, -4.0, device=dev)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    batches = [PosteriorBatch(z["int_array"][i::2], z["float_array"][i::2], dev) for i in range(2)]
    got = model.calculate_probability_thresholds(batches)
    err = np.concatenate([model.error_probabilities_b(b).cpu().numpy() for b in batches])
    types = np.concatenate([z["int_array"][i::2, 3] for i in range(2)])
    assert set(got.keys()) == set(Variation)
    for var_type in Variation:
        mine = sorted(float(x) for x in err[types == int(var_type)])
        want = _reference_roc_loop(mine)
        if got[var_type] != want[0]:
            # a parallel prefix sum rounds differently from the reference's sequential one: only an exact tie of the
            # F score may then pick another element; the F score at the chosen threshold must still be the maximum
            k = mine.index(got[var_type])
            total_art = sum(mine) + 0.0001
            tp = sum(1 - x for x in mine[:k + 1])
            fp = sum(mine[:k + 1])
            sens, prec = tp / (len(mine) - total_art + 0.0002), tp / (tp + fp)
            best = 2 * want[2] * want[1] / (want[2] + want[1] + 0.0001)
            assert abs(2 * sens * prec / (sens + prec + 0.0001) - best) < 1e-9, var_type
