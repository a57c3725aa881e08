"""The context-dependent somatic-SNV prior M step of the reference's PosteriorModelPriors
(permutect/architecture/posterior_model_priors.py:157-222) without PyMC.

The reference fits, with PyMC's mean-field ADVI, the hierarchical model

    overall_rate ~ Beta(1, 1e6)
    substitution_concentration, context_concentration ~ Gamma(4, 5)          (shape, rate)
    theta_s  ~ Dirichlet(substitution_concentration * 1_12)                    12 substitutions
    theta_sc ~ Dirichlet(context_concentration * 1_16), independently per s   16 flanking-base pairs
    rate_sc  = overall_rate * 12 theta_s * 16 theta_sc
    snv_sc   ~ Binomial(total_sc, rate_sc)

to the rounded E-step totals and writes ``log E_q[rate_sc]`` (the mean of 1000 draws from the fitted Gaussian) into
``somatic_snv_log_priors_rrra``.  ADVI is a stochastic optimiser, so the reference's own result is not reproducible run to
run; here the same variational family (a diagonal Gaussian over PyMC's unconstrained coordinates: log-odds for the Beta,
log for the Gammas, PyMC's ``SimplexTransform`` for the Dirichlets, each with its log-Jacobian) is fitted to the same
ELBO with a FIXED set of quasi-random standard-normal draws (sample-average approximation), which makes the objective
deterministic and lets L-BFGS take it to its optimum; the mean of ``rate_sc`` is then taken over a second fixed set of draws.
PARITY UNPINNED for this step: pymc is absent from the build image, so no golden from the reference's M step exists; the
tests hold it to the model's own limits (large counts: rate -> snv / total; empty cells shrink to their substitution's
mean) and to the reference's index conventions (``convert_rrra_tensor_to_sc``, :39-60; the write-back loop :210-222).

Host side (float64, 194 variational means + 194 scales), like the PyMC fit it replaces: the E step that produces the two
5x5x5x5 totals is the CUDA kernel ``pmt_posterior_fit_step``.
"""
import math
from typing import Tuple

import torch
from torch import Tensor

NUM_SUBSTITUTIONS, NUM_CONTEXTS = 12, 16
_NONTRIVIAL_S = (1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14)          # ref * 4 + alt with ref != alt (:53)
_ELBO_DRAWS, _MEAN_DRAWS = 64, 1000                                 # the reference averages 1000 draws (:203)


def convert_rrra_tensor_to_sc(input_rrra: Tensor) -> Tensor:
    """[left flank, ref, right flank, alt] (5 values each, index 4 = deletion) -> [12 substitutions, 16 flank pairs]
    (posterior_model_priors.py:39-60)."""
    t = input_rrra[0:4, 0:4, 0:4, 0:4].permute(0, 2, 1, 3)           # lf, rf, ref, alt
    t = t.reshape(NUM_CONTEXTS, 16)[:, list(_NONTRIVIAL_S)]
    return t.transpose(0, 1).contiguous()


def scatter_sc_to_rrra(values_sc: Tensor, out_rrra: Tensor) -> None:
    """The write-back loop of :210-222: out[lf, ref, rf, alt] = values[substitution, lf * 4 + rf] for ref != alt; every
    other entry of ``out_rrra`` (deletion indices, ref == alt) is left as it is."""
    v = values_sc.to(out_rrra.dtype).to(out_rrra.device)
    for s, flat in enumerate(_NONTRIVIAL_S):
        ref, alt = divmod(flat, 4)
        out_rrra[0:4, ref, 0:4, alt] = v[s].view(4, 4)


def _sobol_normal(n: int, dim: int, seed: int) -> Tensor:
    u = torch.quasirandom.SobolEngine(dim, scramble=True, seed=seed).draw(n, dtype=torch.float64)
    u = u.clamp(1e-9, 1 - 1e-9)
    return math.sqrt(2.0) * torch.erfinv(2 * u - 1)


def _simplex_backward(z: Tensor) -> Tuple[Tensor, Tensor]:
    """PyMC's SimplexTransform: x = softmax([z, -sum z]); returns (log x, log|det J|) with log|det J| = log N + sum log x."""
    y = torch.cat([z, -z.sum(-1, keepdim=True)], dim=-1)
    log_x = torch.log_softmax(y, dim=-1)
    return log_x, math.log(y.shape[-1]) + log_x.sum(-1)


def _log_joint(z: Tensor, total_sc: Tensor, snv_sc: Tensor) -> Tuple[Tensor, Tensor]:
    """log p(snv | z) + log p(T(z)) + log|det J_T(z)| for a batch of unconstrained points z [n, 194]; also rate_sc [n, 12, 16]."""
    S, Cx = NUM_SUBSTITUTIONS, NUM_CONTEXTS
    z_rate, z_sub, z_ctx = z[:, 0], z[:, 1], z[:, 2]
    z_s = z[:, 3:3 + (S - 1)]
    z_sc = z[:, 3 + (S - 1):].reshape(-1, S, Cx - 1)
    # overall_rate ~ Beta(1, 1e6) through log-odds
    log_r, log_1mr = torch.nn.functional.logsigmoid(z_rate), torch.nn.functional.logsigmoid(-z_rate)
    lp = math.log(1e6) + (1e6 - 1) * log_1mr + log_r + log_1mr
    # concentrations ~ Gamma(4, rate 5) through log
    for zc in (z_sub, z_ctx):
        lp = lp + 4 * math.log(5.0) - math.lgamma(4.0) + 3 * zc - 5 * torch.exp(zc) + zc
    a_s, a_c = torch.exp(z_sub), torch.exp(z_ctx)
    log_ts, jac_s = _simplex_backward(z_s)
    lp = lp + torch.lgamma(S * a_s) - S * torch.lgamma(a_s) + (a_s - 1) * log_ts.sum(-1) + jac_s
    log_tsc, jac_sc = _simplex_backward(z_sc)
    lp = lp + S * (torch.lgamma(Cx * a_c) - Cx * torch.lgamma(a_c)) + (a_c - 1) * log_tsc.sum((-1, -2)) + jac_sc.sum(-1)
    log_rate = log_r[:, None, None] + math.log(S * Cx) + log_ts[:, :, None] + log_tsc
    log_rate = log_rate.clamp(max=-1e-12)
    log_1m = torch.log1p(-torch.exp(log_rate))
    lp = lp + (snv_sc * log_rate + (total_sc - snv_sc) * log_1m).sum((-1, -2))      # binomial coefficient is constant
    return lp, torch.exp(log_rate)


def fit_mutation_rates_sc(total_sc: Tensor, snv_sc: Tensor, max_iter: int = 300) -> Tensor:
    """E_q[rate_sc] [12, 16] (float64, CPU) of the mean-field fit described in the module docstring.  ``total_sc`` and
    ``snv_sc`` are the integer-rounded totals the reference hands to its Binomial (:158-159)."""
    total = total_sc.detach().double().cpu().reshape(NUM_SUBSTITUTIONS, NUM_CONTEXTS)
    snv = torch.minimum(snv_sc.detach().double().cpu().reshape(NUM_SUBSTITUTIONS, NUM_CONTEXTS), total)
    dim = 3 + (NUM_SUBSTITUTIONS - 1) + NUM_SUBSTITUTIONS * (NUM_CONTEXTS - 1)
    # start at the smoothed maximum-likelihood point
    overall = float((snv.sum() + 1.0) / (total.sum() + 1e6))
    ts = (snv.sum(1) + 1.0) / (snv.sum() + NUM_SUBSTITUTIONS)
    tsc = (snv + 1.0) / (snv.sum(1, keepdim=True) + NUM_CONTEXTS)
    simplex_fwd = lambda x: (torch.log(x) - torch.log(x).mean(-1, keepdim=True))[..., :-1]
    mu = torch.cat([torch.tensor([math.log(overall) - math.log1p(-overall), 0.0, 0.0], dtype=torch.float64),
                    simplex_fwd(ts), simplex_fwd(tsc).reshape(-1)]).requires_grad_(True)
    log_sigma = torch.full((dim,), -2.0, dtype=torch.float64, requires_grad=True)
    eps = _sobol_normal(_ELBO_DRAWS, dim, seed=1234)
    opt = torch.optim.LBFGS([mu, log_sigma], lr=1.0, max_iter=max_iter, tolerance_grad=1e-9, tolerance_change=1e-12,
                            history_size=100, line_search_fn="strong_wolfe")

    def closure():
        opt.zero_grad()
        z = mu + torch.exp(log_sigma) * eps
        lp, _ = _log_joint(z, total, snv)
        loss = -(lp.mean() + log_sigma.sum())                     # entropy of the diagonal Gaussian up to a constant
        loss.backward()
        return loss

    opt.step(closure)
    with torch.no_grad():
        z = mu + torch.exp(log_sigma) * _sobol_normal(_MEAN_DRAWS, dim, seed=4321)
        _, rate = _log_joint(z, total, snv)
        return rate.mean(0)


def context_m_step(somatic_snv_log_priors_rrra: Tensor, somatic_snv_totals_rrra: Tensor, snv_context_totals_rrra: Tensor,
                   total_ignored_per_context: float) -> Tensor:
    """posterior_model_priors.py:157-222: rounds the totals as the reference does, fits the rates and writes their logs
    into ``somatic_snv_log_priors_rrra`` in place.  Returns the fitted rates [12, 16]."""
    total_sc = torch.round(convert_rrra_tensor_to_sc(snv_context_totals_rrra.detach().double().cpu()) + total_ignored_per_context)
    snv_sc = torch.round(convert_rrra_tensor_to_sc(somatic_snv_totals_rrra.detach().double().cpu()))
    rates_sc = fit_mutation_rates_sc(total_sc, snv_sc)
    with torch.no_grad():
        scatter_sc_to_rrra(torch.log(rates_sc), somatic_snv_log_priors_rrra)
    return rates_sc
