"""The context-dependent SNV-prior M step (permutect_b200/architecture/snv_context_priors.py) against the reference's index
conventions (posterior_model_priors.py:39-60, 210-222) and against the limits of the model the reference fits with PyMC
(:161-190).  The fit itself has no reference golden (pymc is not installable here; the reference's ADVI is stochastic)."""
import itertools

import numpy as np
import pytest
import torch

from permutect_b200.architecture import snv_context_priors as S


def test_rrra_to_sc_matches_the_reference():
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref not installed")
    reference.load()
    from permutect.architecture.posterior_model_priors import convert_rrra_tensor_to_sc
    x = torch.arange(625, dtype=torch.float32).reshape(5, 5, 5, 5)
    assert torch.equal(S.convert_rrra_tensor_to_sc(x), convert_rrra_tensor_to_sc(x))


def test_write_back_follows_the_reference_loop():
    values_sc = torch.arange(12 * 16, dtype=torch.float64).reshape(12, 16) + 1
    got = torch.full((5, 5, 5, 5), -7.0)
    S.scatter_sc_to_rrra(values_sc, got)
    want = torch.full((5, 5, 5, 5), -7.0)
    for lf, rf in itertools.product(range(4), range(4)):          # posterior_model_priors.py:210-222
        for ref, alt in itertools.product(range(4), range(4)):
            if ref != alt:
                with_trivial = ref * 4 + alt
                want[lf, ref, rf, alt] = values_sc[with_trivial - (with_trivial // 5) - 1, lf * 4 + rf]
    assert torch.equal(got, want)
    # and the two index maps are inverse to each other on the SNV entries
    assert torch.equal(S.convert_rrra_tensor_to_sc(got.double()), values_sc)


def _counts(total_per_cell, seed=0):
    g = torch.Generator().manual_seed(seed)
    true = 3e-6 * torch.exp(0.7 * torch.randn(12, 16, dtype=torch.float64, generator=g))
    total = torch.full((12, 16), float(total_per_cell), dtype=torch.float64)
    snv = torch.round(total * true + torch.sqrt(total * true) * torch.randn(12, 16, dtype=torch.float64, generator=g)).clamp(min=0)
    return true, total, snv


def test_large_counts_give_the_empirical_rates():
    _, total, snv = _counts(5e9)
    rates = S.fit_mutation_rates_sc(total, snv)
    assert rates.shape == (12, 16)
    np.testing.assert_allclose(rates.numpy(), (snv / total).numpy(), rtol=0.02)


def test_fit_is_deterministic_and_shrinks_empty_cells():
    true, total, snv = _counts(2e7)
    total[3, 5], snv[3, 5] = 1000.0, 0.0              # a context that was hardly ever seen
    a, b = S.fit_mutation_rates_sc(total, snv), S.fit_mutation_rates_sc(total, snv)
    assert torch.equal(a, b)
    assert torch.all(a > 0) and torch.all(torch.isfinite(torch.log(a)))
    row_mean = float((snv[3].sum() / total[3].sum()))
    assert 0.2 * row_mean < float(a[3, 5]) < 5 * row_mean      # pulled to its substitution's rate, not to 0 / 1000
    # sparse data: the overall rate survives, every cell stays within the prior's reach of it
    _, total_s, snv_s = _counts(2e5, seed=1)
    r = S.fit_mutation_rates_sc(total_s, snv_s)
    assert 0.5 < float(r.mean() / true.mean()) < 2.0


def test_context_m_step_rounds_and_writes_logs():
    _, total, snv = _counts(5e8, seed=2)
    totals_rrra, snv_rrra = torch.zeros(5, 5, 5, 5), torch.zeros(5, 5, 5, 5)
    S.scatter_sc_to_rrra(total - 640.0, totals_rrra)
    S.scatter_sc_to_rrra(snv + 0.3, snv_rrra)                    # posterior sums are not integers; the reference rounds them
    totals_rrra[4, 1, 2, 3] = 99.0                               # deletion contexts are ignored by the fit
    log_priors = torch.full((5, 5, 5, 5), -10.0)
    rates = S.context_m_step(log_priors, snv_rrra, totals_rrra, total_ignored_per_context=640.0)
    np.testing.assert_allclose(rates.numpy(), S.fit_mutation_rates_sc(total, snv).numpy(), rtol=1e-12)
    np.testing.assert_allclose(S.convert_rrra_tensor_to_sc(log_priors.double()).numpy(), torch.log(rates).numpy(), rtol=1e-6)
    assert float(log_priors[4, 1, 2, 3]) == -10.0 and float(log_priors[1, 2, 3, 2]) == -10.0
