"""CPU: the oracle (oracle/*.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py ran /root/reference's own solve_ode / score_fit / _compute_Y /
initial_condition).  This pins the checker before any GPU result is compared with it."""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import local_models as om  # noqa: E402
import loss as ol  # noqa: E402
import morris as omor  # noqa: E402

FILES = sorted(glob.glob(os.path.join(GOLDEN, "local_*.npz")))


def _case(path):
    name = os.path.basename(path)[6:-4]
    model, ns = name.split("_ns")
    return model, int(ns), np.load(path)


def test_goldens_present():
    assert len(FILES) == 10
    assert os.path.exists(os.path.join(GOLDEN, "steady.npz"))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[6:-4] for f in FILES])
def test_solve_ode_matches_reference(path):
    """Restated RHS + scipy LSODA == reference solve_ode (same SciPy build: bit-level; a different
    SciPy may differ at LSODA's tolerance, hence the loose fallback bound)."""
    model, ns, g = _case(path)
    import scipy
    same_scipy = f"scipy={scipy.__version__}" in list(g["versions"])
    tol = 1e-12 if same_scipy else 5e-6
    for b in range(0, g["params"].shape[0], 3):
        sol, flat = om.solve_ode(model, g["params"][b], g["y0"][b], ns, g["t"])
        assert sol.shape == g["sol"][b].shape and flat.shape == g["flat"][b].shape
        assert np.abs(sol - g["sol"][b]).max() <= tol
        assert np.abs(flat - g["flat"][b]).max() <= tol
        assert (sol >= 0).all()


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[6:-4] for f in FILES])
def test_exact_solution_matches_tight_reference(path):
    """O3 (matrix exponential of the restated linear system) vs O2 (reference RHS through LSODA at
    rtol=atol=1e-12): independent routes to the same trajectory."""
    model, ns, g = _case(path)
    for b in range(0, g["params"].shape[0], 4):
        ex = om.exact_linear(model, g["params"][b], g["y0"][b], ns, g["t"])
        ref = g["sol_tight"][b]
        assert np.all(np.abs(ex - ref) <= 2e-8 * np.abs(ref) + 1e-10)


@pytest.mark.parametrize("path", FILES[:4], ids=[os.path.basename(f)[6:-4] for f in FILES[:4]])
def test_score_and_Y_match_reference(path):
    model, ns, g = _case(path)
    for b in range(g["params"].shape[0]):
        s = ol.score_fit(g["params"][b], g["target"], g["flat"][b])
        assert abs(s - g["score"][b]) <= 1e-13 * max(1.0, abs(g["score"][b]))
        for m in ol.Y_METRICS:
            y = ol.compute_Y(g["sol"][b], ns, m)
            assert abs(y - g[f"Y_{m}"][b]) <= 1e-12 * max(1.0, abs(g[f"Y_{m}"][b])), m


def test_flat_layout():
    model, ns, g = _case(FILES[1])
    sol = g["sol"][0]
    flat = om.flat_from_sol(model, sol, ns)
    T = sol.shape[0]
    assert np.array_equal(flat[:T - 5], sol[5:, 0])
    assert np.array_equal(flat[T - 5:2 * T - 5], sol[:, 1])
    assert np.array_equal(flat[2 * T - 5:], sol[:, 2:].T.ravel())


def test_randmod_flat_uses_bitmask_columns():
    """randmod's site block is columns 2..2+ns-1 (bitmask states 1..ns), randmod.py:297-302."""
    model, ns, g = _case([f for f in FILES if "randmod_ns3" in f][0])
    sol, flat = om.solve_ode(model, g["params"][0], g["y0"][0], ns, g["t"])
    T = sol.shape[0]
    assert np.array_equal(flat[2 * T - 5:], sol[:, 2:2 + ns].T.ravel())


def test_steady_matches_reference():
    g = np.load(os.path.join(GOLDEN, "steady.npz"))
    for key in g.files:
        if key == "versions":
            continue
        model, ns = key.split("_ns")
        assert np.abs(np.array(om.initial_condition(model, int(ns))) - g[key]).max() < 1e-12
    # the closed form is a steady state of the all-ones distributive system
    y = np.array(om.initial_condition("distmod", 3))
    assert np.abs(om.rhs("distmod", y, np.ones(10), 3)).max() < 1e-15


def test_weighted_ssr_definition():
    rng = np.random.default_rng(0)
    flat, tgt, sig, p = rng.random(20), rng.random(20), rng.random(26) + 0.5, rng.random(6)
    want = np.sum(((flat - tgt) / sig[:20]) ** 2) + np.sum(((0.7 / 6 * p ** 2) / sig[20:]) ** 2)
    assert abs(ol.weighted_ssr(p, flat, tgt, sig, lam=0.7) - want) < 1e-14
    assert abs(ol.weighted_ssr(p, flat, tgt) - np.sum((flat - tgt) ** 2)) < 1e-14


def test_morris_sample_structure_and_known_answer():
    bounds = [omor.compute_bound(v) for v in (1.0, 2.0, 0.0, 4.0)]
    assert bounds[2] == [0.0, 0.1] and bounds[0] == [0.5, 1.5]
    X = omor.sample(bounds, N=64, num_levels=400, seed=5)
    D = 4
    assert X.shape == (64 * (D + 1), D)
    b = np.asarray(bounds)
    assert (X >= b[:, 0] - 1e-12).all() and (X <= b[:, 1] + 1e-12).all()
    d = np.diff(X.reshape(64, D + 1, D), axis=1)
    assert ((np.abs(d) > 0).sum(axis=2) == 1).all()          # one coordinate per move
    assert ((np.abs(d) > 0).sum(axis=1) == 1).all()          # every coordinate moves once
    delta = omor.delta_of(400)
    assert np.allclose(np.abs(d).max(axis=1) / (b[:, 1] - b[:, 0]), delta)
    # linear model: EE_i = a_i * range_i exactly, sigma = 0
    a = np.array([3.0, -2.0, 5.0, 0.25])
    Y = X @ a
    res = omor.analyze(X, Y, D, 400)
    assert np.allclose(res["mu"], a * (b[:, 1] - b[:, 0]), rtol=1e-9)
    assert np.allclose(res["mu_star"], np.abs(res["mu"]))
    assert np.all(res["sigma"] < 1e-9)
