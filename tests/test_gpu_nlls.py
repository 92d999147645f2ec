"""GPU: the batched bounded least-squares driver (`pk_local_nlls_batch`) against SciPy's
`least_squares(method='trf', x_scale='jac', bounds=...)` — what `curve_fit` runs at paramest/normest.py:278-290 —
on the SAME residual model (normest.py:403-423), evaluated on the CPU with the oracle's exact solution of the
linear models.  The optimiser policy differs (projected Levenberg-Marquardt vs TRF; outside the parity contract,
SURVEY.md §8(c)); what must agree is the minimum: cost within 1e-5 relative, parameters within 1e-3 relative when
both start in the same basin.
"""
import os
import sys

import numpy as np
import pytest
from scipy.optimize import least_squares

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import local_models as om  # noqa: E402
import loss as ol  # noqa: E402
from phoskintime_b200 import paramest  # noqa: E402
from phoskintime_b200.steady import initial_condition  # noqa: E402

pytestmark = pytest.mark.gpu
T = om.TIME_POINTS


def cpu_residual(model, ns, y0, target, sigma, lam):
    P = om.n_params(model, ns)

    def f(theta):
        p = np.exp(theta) if model == "randmod" else theta
        flat = om.flat_from_sol(model, om.exact_linear(model, p, y0, ns, T), ns)
        return (np.concatenate([flat, lam / P * theta ** 2]) - np.concatenate([target, np.zeros(P)])) / sigma
    return f


@pytest.mark.parametrize("model,ns,lam", [("distmod", 3, 0.0), ("succmod", 4, 0.1), ("distmod", 2, 0.05)])
def test_minima_agree_with_scipy_trf(engine, model, ns, lam):
    rng = np.random.default_rng(5 + ns)
    n, P = om.n_states(model, ns), om.n_params(model, ns)
    y0 = np.asarray(initial_condition(ns, model))
    th_true = rng.uniform(0.3, 2.0, P)
    clean = om.flat_from_sol(model, om.exact_linear(model, th_true, y0, ns, T), ns)
    target = clean * (1.0 + 0.05 * rng.standard_normal(clean.size))          # noisy: the minimum is not zero
    L = target.size
    sigma = np.concatenate([rng.uniform(0.5, 1.5, L), np.ones(P)])
    lb, ub = np.full(P, 1e-2), np.full(P, 20.0)
    starts = np.clip(th_true * np.exp(0.15 * rng.standard_normal((6, P))), lb, ub)
    res = engine.nlls_local_batch(model, starts, y0, ns, T, target, lb, ub, sigma=sigma, lam=lam, max_iter=200)
    assert (res["status"] > 0).all(), res["status"]
    f = cpu_residual(model, ns, y0, target, sigma, lam)
    for b in range(starts.shape[0]):
        ref = least_squares(f, starts[b], bounds=(lb, ub), method="trf", x_scale="jac", xtol=1e-12, ftol=1e-12,
                            gtol=1e-12, max_nfev=4000)
        c_cpu_at_gpu = 0.5 * np.sum(f(res["theta"][b]) ** 2)
        # the fused cost equals the CPU residual's cost at the same point (integrator error only)
        assert np.isclose(res["cost"][b], c_cpu_at_gpu, rtol=1e-5, atol=1e-12), (b, res["cost"][b], c_cpu_at_gpu)
        assert res["cost"][b] <= ref.cost * (1.0 + 1e-5) + 1e-12, (b, res["cost"][b], ref.cost, res["status"][b])
        if np.isclose(res["cost"][b], ref.cost, rtol=1e-5):
            free = (ref.x > lb * 1.001) & (ref.x < ub * 0.999)
            assert np.allclose(res["theta"][b][free], ref.x[free], rtol=5e-3, atol=1e-4), (b, res["theta"][b], ref.x)
        # score_fit at the optimum, as normest.py:293-306 ranks the starts
        flat = om.flat_from_sol(model, om.exact_linear(model, res["theta"][b], y0, ns, T), ns)
        assert np.isclose(res["score"][b], ol.score_fit(res["theta"][b], target, flat), rtol=1e-5)


def test_exact_data_is_recovered_and_bounds_hold(engine):
    """Noise-free data from a hidden parameter set: the cost falls to integrator-noise level; a box that excludes
    the true value of one parameter pins it to the bound and the others re-fit."""
    model, ns = "distmod", 3
    rng = np.random.default_rng(2)
    P = om.n_params(model, ns)
    y0 = np.asarray(initial_condition(ns, model))
    th_true = rng.uniform(0.3, 2.0, P)
    target = om.flat_from_sol(model, om.exact_linear(model, th_true, y0, ns, T), ns)
    lb, ub = np.full(P, 1e-2), np.full(P, 20.0)
    starts = np.clip(th_true * np.exp(0.3 * rng.standard_normal((32, P))), lb, ub)
    res = engine.nlls_local_batch(model, starts, y0, ns, T, target, lb, ub, max_iter=200)
    assert (res["status"] > 0).all()
    c0 = np.array([0.5 * np.sum((om.flat_from_sol(model, om.exact_linear(model, s, y0, ns, T), ns) - target) ** 2) for s in starts])
    assert (res["cost"] <= 1e-6 * c0).mean() >= 0.9 and res["cost"].min() < 1e-12
    # per-problem initial conditions travel with their problem through the compacted batches
    res_y = engine.nlls_local_batch(model, starts, np.repeat(y0[None, :], starts.shape[0], axis=0), ns, T, target, lb, ub,
                                    max_iter=200)
    assert np.allclose(res_y["theta"], res["theta"], rtol=1e-9, atol=1e-12) and np.array_equal(res_y["status"], res["status"])
    ub2 = ub.copy()
    ub2[2] = 0.5 * th_true[2]
    res2 = engine.nlls_local_batch(model, starts, y0, ns, T, target, lb, ub2, max_iter=200)
    assert (res2["theta"] >= lb - 1e-15).all() and (res2["theta"] <= ub2 + 1e-15).all()
    best = int(np.argmin(res2["cost"]))
    assert np.isclose(res2["theta"][best, 2], ub2[2], rtol=1e-9) and res2["cost"][best] > res["cost"].min()


def test_fit_multistart_many_proteins(engine):
    """`_curve_fit_multistart` for several proteins in one batch (BASELINE configs[3] shape, reduced): per-protein
    targets, normest's start sampling, best start by score_fit; randmod goes through log-parameters."""
    model, ns, G = "distmod", 4, 12
    rng = np.random.default_rng(4)
    P = om.n_params(model, ns)
    y0 = np.asarray(initial_condition(ns, model))
    truth = rng.uniform(0.2, 3.0, (G, P))
    targets = np.array([om.flat_from_sol(model, om.exact_linear(model, th, y0, ns, T), ns) for th in truth])
    targets *= 1.0 + 0.05 * rng.standard_normal(targets.shape)
    lb, ub = np.full(P, 1e-2), np.full(P, 20.0)
    fit = paramest.fit_multistart(model, np.ones(P), lb, ub, y0, ns, T, targets, genes=[f"G{p}" for p in range(G)],
                                  n_starts=24, engine=engine, max_iter=150)
    assert fit["popt"].shape == (G, P) and (fit["n_ok"] > 0).all()
    per = fit["theta"].shape[0] // G
    for p in range(G):
        sl = slice(p * per, (p + 1) * per)
        assert fit["best_score"][p] == np.nanmin(np.where(fit["status"][sl] > 0, fit["score"][sl], np.inf))
        flat = om.flat_from_sol(model, om.exact_linear(model, fit["popt"][p], y0, ns, T), ns)
        assert np.isclose(fit["best_score"][p], ol.score_fit(fit["popt"][p], targets[p], flat), rtol=1e-5)
    # the best-cost start of (nearly) every protein is at least as good as the parameters that generated the data
    truth_cost = np.array([0.5 * np.sum((om.flat_from_sol(model, om.exact_linear(model, truth[p], y0, ns, T), ns) - targets[p]) ** 2)
                           for p in range(G)])
    best_cost = np.array([fit["cost"][p * per:(p + 1) * per][fit["status"][p * per:(p + 1) * per] > 0].min() for p in range(G)])
    assert (best_cost <= 1.02 * truth_cost).mean() >= 0.75, (best_cost / truth_cost)
    # random model through log-parameters (normest.py:54): theta = log(params)
    ns_r = 3
    Pr = om.n_params("randmod", ns_r)
    y0r = np.asarray(initial_condition(ns_r, "randmod"))
    th = np.log(rng.uniform(0.3, 2.0, Pr))
    tgt = om.flat_from_sol("randmod", om.exact_linear("randmod", np.exp(th), y0r, ns_r, T), ns_r)
    st = th + 0.1 * rng.standard_normal((4, Pr))
    r = engine.nlls_local_batch("randmod", st, y0r, ns_r, T, tgt, np.full(Pr, -5.0), np.full(Pr, 3.0), log_params=True,
                                max_iter=200)
    assert (r["status"] > 0).all() and r["cost"].min() < 1e-10
