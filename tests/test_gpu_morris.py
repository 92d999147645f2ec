"""GPU: Morris elementary effects and the config-1 sensitivity flow against the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import local_models as om  # noqa: E402
import loss as ol  # noqa: E402
import morris as omor  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scaled", [False, True])
def test_morris_ee_kernel_matches_oracle(engine, scaled):
    rng = np.random.default_rng(0)
    bounds = [omor.compute_bound(v) for v in rng.uniform(0.1, 3.0, 7)]
    X = omor.sample(bounds, N=200, num_levels=400, seed=1)
    Y = np.sin(X).sum(axis=1) + X[:, 0] * X[:, 3]
    ref = omor.analyze(X, Y, 7, 400, scaled=scaled)
    got = engine.morris_ee(X, Y, 400, scaled=scaled, want_ee=True)
    for k in ("mu", "mu_star", "sigma"):
        assert np.allclose(got[k], ref[k], rtol=1e-10, atol=1e-13), k
    assert np.allclose(got["ee"], ref["ee"], rtol=1e-10, atol=1e-13)


def test_config1_morris_ranking_identical_to_reference_path(engine):
    """BASELINE config 1 shape (distributive, 3 psites, D=10) on a reduced trajectory count so the
    CPU oracle finishes in seconds: the SAME X goes through the stock-reference restatement and
    through the GPU; mu*/sigma rankings must be identical (north star)."""
    from phoskintime_b200 import sensitivity
    from phoskintime_b200.steady import initial_condition
    ns, N, levels = 3, 120, 400
    theta = np.random.default_rng(1).uniform(0.05, 3.0, 10)
    y0 = np.asarray(initial_condition(ns, "distmod"))
    prob = sensitivity.define_sensitivity_problem_ds(ns, theta)
    X = sensitivity.morris_sample(prob, N, levels, seed=42)
    Si, _ = sensitivity.sensitivity_analysis(theta, om.TIME_POINTS, ns, y0, "distmod", N=N, num_levels=levels,
                                             X=X, engine=engine)
    Y_ref = np.array([ol.compute_Y(om.solve_ode("distmod", x, y0, ns, om.TIME_POINTS)[0], ns) for x in X])
    assert np.all(np.abs(Si["Y"] - Y_ref) <= 1e-6 * np.abs(Y_ref) + 1e-6)
    ref = omor.analyze(X, Y_ref, 10, levels, scaled=True)
    for k in ("mu_star", "sigma"):
        assert np.array_equal(np.argsort(Si[k]), np.argsort(ref[k])), k
        assert np.allclose(Si[k], ref[k], rtol=1e-4, atol=1e-9)
    assert np.allclose(Si["mu"], ref["mu"], rtol=1e-4, atol=1e-8)


def test_config1_morris_full_size_n1000(engine):
    """BASELINE configs[0] at FULL size: distributive, 3 psites, Morris N = 1000 trajectories x (D + 1) = 11 000 rows at the
    14 experimental time points.  The same X goes through the reference path on the CPU (oracle port of `solve_ode`:
    LSODA at its defaults + `_compute_Y`) and through ONE launch on the GPU; identical mu* / sigma ranking, Y within
    the parity bound."""
    from phoskintime_b200 import sensitivity
    from phoskintime_b200.steady import initial_condition
    ns, N, levels = 3, 1000, 400
    theta = np.random.default_rng(1).uniform(0.05, 3.0, 10)
    y0 = np.asarray(initial_condition(ns, "distmod"))
    prob = sensitivity.define_sensitivity_problem_ds(ns, theta)
    X = sensitivity.morris_sample(prob, N, levels, seed=42)
    assert X.shape == (11000, 10)
    Si, _ = sensitivity.sensitivity_analysis(theta, om.TIME_POINTS, ns, y0, "distmod", N=N, num_levels=levels,
                                             X=X, engine=engine)
    Y_ref = np.array([ol.compute_Y(om.solve_ode("distmod", x, y0, ns, om.TIME_POINTS)[0], ns) for x in X])
    assert np.all(np.abs(Si["Y"] - Y_ref) <= 1e-6 * np.abs(Y_ref) + 1e-6)
    ref = omor.analyze(X, Y_ref, 10, levels, scaled=True)
    for k in ("mu_star", "sigma"):
        assert np.array_equal(np.argsort(Si[k]), np.argsort(ref[k])), k
        assert np.allclose(Si[k], ref[k], rtol=1e-4, atol=1e-9)
    assert np.allclose(Si["mu"], ref["mu"], rtol=1e-4, atol=1e-8)


def test_sensitivity_top_k_selection(engine):
    from phoskintime_b200 import sensitivity
    ns, N, levels = 3, 40, 400
    theta = np.random.default_rng(2).uniform(0.2, 2.0, 10)
    y0 = np.array([1.0, 0.4, 0.2, 0.2, 0.2])
    sol, _ = om.solve_ode("distmod", theta, y0, ns, om.TIME_POINTS)
    pr, p, rna = sol[:, 1], sol[:, 2:].T, sol[-9:, 0]
    Si, best = sensitivity.sensitivity_analysis(theta, om.TIME_POINTS, ns, y0, "distmod", pr_data=pr, p_data=p,
                                                rna_data=rna, N=N, num_levels=levels, seed=3, engine=engine)
    assert len(best) == int(np.ceil(N * 10 / levels)) and best[0]["rmse"] <= best[-1]["rmse"]
    assert best[0]["solution"].shape == (14, 5)
    # the fused RMSE (kernel epilogue, weighted residual) equals the reference formula on full trajectories
    # (sensitivity/analysis.py:277-284) and picks the same rows
    full = engine.solve_local_batch("distmod", Si["X"], y0, ns, om.TIME_POINTS, want=("sol",))["sol"]
    host = sensitivity.select_closest(full, Si["X"], pr, p, rna, ns, N, levels)
    assert [tuple(b["params"]) for b in best] == [tuple(b["params"]) for b in host]
    assert np.allclose([b["rmse"] for b in best], [b["rmse"] for b in host], rtol=1e-9, atol=1e-15)
    assert np.array_equal(best[0]["solution"], host[0]["solution"])
    # data that do not line up with the flat layout fall back to the trajectory path
    Si2, best2 = sensitivity.sensitivity_analysis(theta, om.TIME_POINTS, ns, y0, "distmod", pr_data=pr, p_data=p,
                                                  rna_data=sol[-8:, 0], N=4, num_levels=levels, seed=3, engine=engine)
    assert len(best2) == 1 and "rmse" not in Si2


def test_multistart_loss_batch_config4_shape(engine):
    """normest shape (config 4, reduced): G proteins x S starts, per-protein targets, one launch."""
    from phoskintime_b200 import paramest
    from phoskintime_b200.steady import initial_condition
    ns, G, S = 4, 6, 32
    rng = np.random.default_rng(4)
    y0 = np.asarray(initial_condition(ns, "distmod"))
    hidden = rng.uniform(0.05, 3.0, (G, 12))
    targets = np.array([om.solve_ode("distmod", h, y0, ns, om.TIME_POINTS)[1] for h in hidden])
    starts = np.concatenate([paramest.multistart_points(hidden[g], np.zeros(12), np.full(12, 20.0), S, seed=42 + g)
                             for g in range(G)])
    group = np.repeat(np.arange(G, dtype=np.int32), S)
    lam = 0.1
    r = paramest.evaluate_starts("distmod", starts, y0, ns, om.TIME_POINTS, targets, group=group, lam=lam,
                                 engine=engine, want=("ssr", "score", "flat"))
    for b in range(0, G * S, 7):
        g = group[b]
        assert abs(r["ssr"][b] - ol.weighted_ssr(starts[b], r["flat"][b], targets[g], None, lam)) <= 1e-12 * max(1, r["ssr"][b])
        sc = ol.score_fit(starts[b], targets[g], r["flat"][b])
        assert abs(r["score"][b] - sc) <= 1e-12 * max(1.0, abs(sc))
    best = paramest.best_per_group(r["ssr"], group, G)
    assert list(best) == [g * S for g in range(G)]      # the hidden truth (start 0 of each protein) wins
