"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the committed
goldens of the unmodified reference.

Tolerances:
  * vs O2 (reference RHS through LSODA at rtol=atol=1e-12) and O3 (exact matrix exponential):
    |ours - ref| <= 1e-6*|ref| + 1e-9          (north star: max rel err <= 1e-6; the 1e-9 floor is
                                                where LSODA-tight itself stops resolving states)
  * vs O1 (stock reference, LSODA default tolerances ~1.5e-8, itself up to ~1e-6..1e-5 off O2):
    |ours - stock| <= 1e-5*|stock| + 1e-6, and at least 99 % of entries within 1e-6*|stock| + 1e-7
  * fused scalars (flat/Y/ssr/score) vs the oracle formulas applied to OUR trajectories: 1e-12 rel.
"""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import local_models as om  # noqa: E402
import loss as ol  # noqa: E402

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN, "local_*.npz")))
T14 = om.TIME_POINTS


def _case(path):
    name = os.path.basename(path)[6:-4]
    model, ns = name.split("_ns")
    return model, int(ns), np.load(path)


def _close(a, ref, rtol, atol):
    return np.abs(a - ref) <= rtol * np.abs(ref) + atol


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[6:-4] for f in FILES])
def test_parity_with_reference_goldens(engine, path):
    model, ns, g = _case(path)
    r = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol", "flat"))
    assert (r["status"] == 0).all()
    sol = r["sol"]
    assert sol.shape == g["sol"].shape and (sol >= 0).all()
    assert _close(sol, g["sol_tight"], 1e-6, 1e-9).all(), \
        float((np.abs(sol - g["sol_tight"]) / (1e-6 * np.abs(g["sol_tight"]) + 1e-9)).max())
    assert _close(sol, g["sol"], 1e-5, 1e-6).all()
    assert _close(sol, g["sol"], 1e-6, 1e-7).mean() >= 0.99
    # margin at the library defaults: within a quarter of the parity bound (the stock reference itself is up to
    # 14x the bound away from the tight solution)
    assert float((np.abs(sol - g["sol_tight"]) / (1e-6 * np.abs(g["sol_tight"]) + 1e-9)).max()) <= 0.25
    assert _close(r["flat"], g["flat"], 1e-5, 1e-6).all()


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[6:-4] for f in FILES])
def test_parity_with_exact_solution(engine, path):
    model, ns, g = _case(path)
    r = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol",))
    for b in range(0, g["params"].shape[0], 3):
        ex = om.exact_linear(model, g["params"][b], g["y0"][b], ns, g["t"])
        assert _close(r["sol"][b], ex, 1e-6, 1e-9).all(), (model, ns, b)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[6:-4] for f in FILES])
def test_fused_epilogue_matches_oracle_formulas(engine, path):
    model, ns, g = _case(path)
    B, P = g["params"].shape
    rng = np.random.default_rng(3)
    L = g["flat"].shape[1]
    sigma = rng.uniform(0.5, 2.0, L + P)
    lam = 0.37
    r = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol", "flat", "ssr", "score"),
                                 target=g["target"], sigma=sigma, lam=lam)
    for b in range(B):
        flat = om.flat_from_sol(model, r["sol"][b], ns)
        assert np.array_equal(flat, r["flat"][b])                         # layout is exact
        s = ol.score_fit(g["params"][b], g["target"], flat)
        assert abs(r["score"][b] - s) <= 1e-12 * max(1.0, abs(s))
        w = ol.weighted_ssr(g["params"][b], flat, g["target"], sigma, lam)
        assert abs(r["ssr"][b] - w) <= 1e-12 * max(1.0, abs(w))
    # and against the reference's own score on ITS trajectories (differs only by LSODA's error)
    assert np.all(np.abs(r["score"] - g["score"]) <= 1e-6 * np.abs(g["score"]) + 1e-7)
    # sigma of length L (no regularisation rows) and no sigma
    r2 = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("ssr", "flat"),
                                  target=g["target"], sigma=sigma[:L])
    r3 = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("ssr",), target=g["target"])
    for b in range(0, B, 5):
        assert abs(r2["ssr"][b] - ol.weighted_ssr(g["params"][b], r2["flat"][b], g["target"], sigma[:L])) \
            <= 1e-12 * max(1.0, r2["ssr"][b])
        assert abs(r3["ssr"][b] - np.sum((r2["flat"][b] - g["target"]) ** 2)) <= 1e-12 * max(1.0, r3["ssr"][b])


@pytest.mark.parametrize("metric", ol.Y_METRICS)
@pytest.mark.parametrize("path", [FILES[1], FILES[9], FILES[5]], ids=["dist3", "succ5", "rand4"])
def test_morris_Y_metrics(engine, path, metric):
    model, ns, g = _case(path)
    r = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol", "Y"), y_metric=metric)
    for b in range(g["params"].shape[0]):
        y = ol.compute_Y(r["sol"][b], ns, metric)
        assert abs(r["Y"][b] - y) <= 1e-11 * max(1.0, abs(y)), (metric, b)
    ref = g[f"Y_{metric}"]
    assert np.all(np.abs(r["Y"] - ref) <= 1e-5 * np.abs(ref) + 1e-6)


def test_reference_signature_single_solve(engine):
    """solve_ode(params, init_cond, num_psites, t) -> (sol[T,n], flat[L]) with tuple params, as
    sensitivity/analysis.py:188-193 calls it."""
    from phoskintime_b200 import models
    from phoskintime_b200.steady import initial_condition
    for name, ns in (("distmod", 3), ("succmod", 5), ("randmod", 3)):
        mod = models.set_model(name)
        n, P, L = __import__("phoskintime_b200").local_dims(name, ns, 14)
        p = tuple(np.random.default_rng(ns).uniform(0.1, 2.5, P))
        y0 = initial_condition(ns)
        sol, flat = models.solve_ode(p, y0, ns, T14)
        assert sol.shape == (14, n) and flat.shape == (L,) and sol.flags["C_CONTIGUOUS"]
        ref_sol, ref_flat = om.solve_ode(name, p, y0, ns, T14)
        assert _close(sol, ref_sol, 1e-5, 1e-6).all() and _close(flat, ref_flat, 1e-5, 1e-6).all()
        assert np.array_equal(sol[0], np.asarray(y0))
    models.set_model("randmod")


def test_edge_cases_time_grid_and_inputs(engine):
    rng = np.random.default_rng(0)
    p = rng.uniform(0.1, 2.0, (5, 10))
    y0 = np.array([1.0, 0.4, 0.2, 0.2, 0.2])
    # T = 1: only the initial state; T < 5: flat has an empty RNA block
    r = engine.solve_local_batch("distmod", p, y0, 3, [0.0], want=("sol", "flat"))
    assert r["sol"].shape == (5, 1, 5) and np.array_equal(r["sol"][:, 0], np.tile(y0, (5, 1)))
    assert r["flat"].shape == (5, 4)
    t4 = np.array([0.0, 1.0, 2.0, 4.0])
    r = engine.solve_local_batch("distmod", p, y0, 3, t4, want=("sol", "flat"))
    assert r["flat"].shape == (5, 16)
    for b in range(5):
        assert np.array_equal(r["flat"][b], om.flat_from_sol("distmod", r["sol"][b], 3))
        assert _close(r["sol"][b], om.exact_linear("distmod", p[b], y0, 3, t4), 1e-6, 1e-9).all()
    # repeated output time and a start time != 0 (autonomous system: only differences matter)
    t_rep = np.array([10.0, 11.0, 11.0, 14.0])
    r = engine.solve_local_batch("succmod", p, y0, 3, t_rep, want=("sol",))
    assert np.array_equal(r["sol"][:, 1], r["sol"][:, 2])
    ex = om.exact_linear("succmod", p[0], y0, 3, t_rep - 10.0)
    assert _close(r["sol"][0], ex, 1e-6, 1e-9).all()
    # empty batch
    r = engine.solve_local_batch("distmod", np.empty((0, 10)), y0, 3, T14, want=("sol",))
    assert r["sol"].shape == (0, 14, 5)
    # wrong shapes are API errors, not crashes
    with pytest.raises(ValueError):
        engine.solve_local_batch("distmod", p[:, :9], y0, 3, T14)
    with pytest.raises(ValueError):
        engine.solve_local_batch("distmod", p, y0[:4], 3, T14)
    with pytest.raises(ValueError):
        engine.solve_local_batch("distmod", p, y0, 3, T14, want=("score",))


def test_knockouts_zero_parameters_and_steady_state(engine):
    """Zeroed rates (knockout/helper.py) and an exact steady state (f(y0) = 0) stay finite."""
    y0 = np.array([1.0, 0.4, 0.2, 0.2, 0.2])
    p = np.ones((4, 10))
    p[1, 4:7] = 0.0            # no phosphorylation
    p[2, 0] = 0.0              # no transcription
    p[3, :] = 0.0              # everything off: only unit dephosphorylation remains
    r = engine.solve_local_batch("distmod", p, y0, 3, T14, want=("sol",))
    assert (r["status"] == 0).all() and np.isfinite(r["sol"]).all()
    assert np.abs(r["sol"][0] - y0).max() < 1e-12          # all-ones parameters: y0 is the steady state
    for b in range(1, 4):
        assert _close(r["sol"][b], om.exact_linear("distmod", p[b], y0, 3, T14), 1e-6, 1e-9).all()


def test_failed_systems_are_flagged_not_fatal(engine):
    y0 = np.array([1.0, 0.4, 0.2, 0.2, 0.2])
    p = np.random.default_rng(1).uniform(0.1, 2.0, (6, 10))
    p[2, 3] = np.nan
    r = engine.solve_local_batch("distmod", p, y0, 3, T14, want=("sol", "Y", "score"), target=np.ones(65))
    assert r["status"][2] == 3 and np.isnan(r["sol"][2, 1:]).all() and np.isnan(r["score"][2])
    ok = np.array([0, 1, 3, 4, 5])
    assert (r["status"][ok] == 0).all() and np.isfinite(r["sol"][ok]).all()
    r = engine.solve_local_batch("distmod", p[:2], y0, 3, T14, want=("sol",), max_steps=20)
    assert (r["status"] == 1).all() and np.isnan(r["sol"][:, -1]).all() and (r["nsteps"] + r["nrej"] == 20).all()
    # dense (warp) kernel takes the same exits
    pr = np.random.default_rng(2).uniform(0.1, 2.0, (3, 14))
    pr[1, 0] = np.inf
    rr = engine.solve_local_batch("randmod", pr, np.full(9, 0.1), 3, T14, want=("sol",))
    assert rr["status"][1] == 3 and (rr["status"][[0, 2]] == 0).all()


def test_per_system_initial_conditions_normalize_and_log_params(engine):
    rng = np.random.default_rng(4)
    for model, ns in (("succmod", 4), ("randmod", 2)):
        n, P, L = __import__("phoskintime_b200").local_dims(model, ns, 14)
        p = rng.uniform(0.1, 2.0, (7, P))
        y0 = rng.uniform(0.2, 1.5, (7, n))
        r = engine.solve_local_batch(model, p, y0, ns, T14, want=("sol",))
        rn = engine.solve_local_batch(model, p, y0, ns, T14, want=("sol", "flat"), normalize=True)
        rl = engine.solve_local_batch(model, np.log(p), y0, ns, T14, want=("sol",), log_params=True)
        for b in range(7):
            assert _close(r["sol"][b], om.exact_linear(model, p[b], y0[b], ns, T14), 1e-6, 1e-9).all()
            assert np.allclose(rn["sol"][b], r["sol"][b] * (1.0 / y0[b])[None, :], rtol=1e-15, atol=0)
        assert np.allclose(rl["sol"], r["sol"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("method,rtol,atol", [("rodas4", 1e-8, 1e-11), ("ros5l", 1e-7, 1e-10), ("ros5l", 1e-9, 1e-12)])
def test_both_integrators_meet_the_parity_bound(engine, method, rtol, atol):
    """RODAS4 (order 4(3)) and ROS5L (order 5(4), the default) against O2 on every golden case; the
    tighter ROS5L run must also be closer to O2 than the default one (the error is tolerance driven)."""
    worst = 0.0
    for path in FILES:
        model, ns, g = _case(path)
        r = engine.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol",), method=method,
                                     rtol=rtol, atol=atol)
        assert (r["status"] == 0).all()
        e = (np.abs(r["sol"] - g["sol_tight"]) / (1e-6 * np.abs(g["sol_tight"]) + 1e-9)).max()
        worst = max(worst, float(e))
    assert worst < (0.05 if rtol > 5e-9 else 0.005), worst


def test_host_and_device_paths_agree_and_are_deterministic(engine):
    import torch
    rng = np.random.default_rng(5)
    p = rng.uniform(0.05, 3.0, (3000, 14))
    y0 = np.asarray(om.initial_condition("succmod", 5))
    target = rng.random(93)
    a = engine.solve_local_batch("succmod", p, y0, 5, T14, want=("flat", "score", "ssr"), target=target)
    b = engine.solve_local_batch("succmod", torch.from_numpy(p).cuda(), torch.from_numpy(y0).cuda(), 5,
                                 torch.from_numpy(T14).cuda(), want=("flat", "score", "ssr"),
                                 target=torch.from_numpy(target).cuda())
    for k in ("flat", "score", "ssr", "status", "nsteps"):
        assert np.array_equal(a[k], b[k].cpu().numpy()), k
    # results do not depend on which lane/warp picked a system up: permute the batch
    perm = rng.permutation(3000)
    c = engine.solve_local_batch("succmod", p[perm], y0, 5, T14, want=("flat", "score"), target=target)
    assert np.array_equal(c["flat"], a["flat"][perm]) and np.array_equal(c["score"], a["score"][perm])


def test_pipelined_host_path_equals_device_path(engine):
    """Host batches >= 400k systems are staged in 5 chunks of growing size (H2D / kernel / D2H overlapped on three
    streams); every output slice, per-system y0 and the group index must land where the one-launch
    device path puts them."""
    import torch
    B = 420_001
    rng = np.random.default_rng(8)
    p = rng.uniform(0.05, 3.0, (B, 10))
    y0 = rng.uniform(0.1, 1.0, (B, 5))
    tg = rng.random((3, 65))
    grp = rng.integers(0, 3, B).astype(np.int32)
    a = engine.solve_local_batch("distmod", p, y0, 3, T14, want=("flat", "ssr", "score", "Y"), target=tg, group=grp)
    b = engine.solve_local_batch("distmod", torch.from_numpy(p).cuda(), torch.from_numpy(y0).cuda(), 3,
                                 torch.from_numpy(T14).cuda(), want=("flat", "ssr", "score", "Y"),
                                 target=torch.from_numpy(tg).cuda(), group=torch.from_numpy(grp).cuda())
    for k in ("flat", "ssr", "score", "Y", "status", "nsteps", "nrej"):
        assert np.array_equal(a[k], b[k].cpu().numpy()), k
    assert engine.last_launch_info()[0] == 1


@pytest.mark.parametrize("ns", [3, 4, 6])
def test_dense_kernel_ros6l_option_matches_goldens(engine, ns):
    """The seven-solve family on the dense kernel (method='ros6l': run-time ROS6L(gamma') coefficients for steps that
    reuse an inverse) against the reference goldens of the random model; fewer steps than the default ROS5L."""
    g = np.load(os.path.join(GOLDEN, f"local_randmod_ns{ns}.npz"))
    a = engine.solve_local_batch("randmod", g["params"], g["y0"], ns, g["t"], want=("sol",), method="ros6l")
    b = engine.solve_local_batch("randmod", g["params"], g["y0"], ns, g["t"], want=("sol",))
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    tight = g["sol_tight"]
    assert float((np.abs(a["sol"] - tight) / (1e-6 * np.abs(tight) + 1e-9)).max()) <= 0.5
    assert a["nsteps"].mean() < 0.85 * b["nsteps"].mean()


def test_states_beyond_fp32_range_are_flagged(engine):
    """The ROS6L error scale is evaluated in FP32: a state beyond 3.4e38 must end as status 3 with NaN outputs, never as
    an accepted step with "zero" error; its neighbours in the batch are unaffected."""
    rng = np.random.default_rng(4)
    p = rng.uniform(0.05, 3.0, (8, 10))
    y0 = np.repeat(rng.uniform(0.1, 1.0, (1, 5)), 8, axis=0)
    y0[3, 1] = 1e39
    r = engine.solve_local_batch("distmod", p, y0, 3, T14, want=("sol",))
    assert r["status"][3] == 3 and np.isnan(r["sol"][3, 1:]).all()
    ok = np.delete(np.arange(8), 3)
    assert (r["status"][ok] == 0).all() and np.isfinite(r["sol"][ok]).all()


def test_fused_gather_single_rank(engine):
    """pk_local_solve_allgather at world 1: the batch is integrated in pieces, every output equals the one-launch
    result bit for bit and the gathered buffer equals the requested per-sample output."""
    import torch
    B = 50_003
    rng = np.random.default_rng(9)
    p = torch.from_numpy(rng.uniform(0.05, 3.0, (B, 14))).cuda()
    y0 = torch.from_numpy(rng.uniform(0.1, 1.0, (B, 7))).cuda()
    t = torch.from_numpy(T14).cuda()
    tg = torch.from_numpy(rng.random(93)).cuda()
    a = engine.solve_local_batch("succmod", p, y0, 5, t, want=("flat", "ssr", "score", "Y"), target=tg)
    assert engine.last_launch_info()[0] == 1
    for key in ("score", "ssr", "Y"):
        recv = torch.full((B,), -1.0, dtype=torch.float64, device="cuda")
        b = engine.solve_local_batch("succmod", p, y0, 5, t, want=("flat", "ssr", "score", "Y"), target=tg,
                                     gather=(key, recv, 3))
        assert engine.last_launch_info()[0] == 3
        for k in ("flat", "ssr", "score", "Y", "status", "nsteps", "nrej"):
            assert torch.equal(a[k], b[k]), k
        assert torch.equal(b["gathered"], a[key])


def test_peer_memory_gather_single_rank(engine):
    """pk_local_solve_gather_p2p at world 1 (pk_sym_alloc / pk_sym_buffer, the kernel's finish stage stores every system's
    scalar into the symmetric buffer): row 0 of the buffer equals the requested per-sample output bit for bit, every
    other output equals the plain call's.  (Two and eight ranks: tools/multi_gpu_check.py, profiles/r2_bench_*gpu_p2p.json.)"""
    import torch
    B = 40_007
    rng = np.random.default_rng(19)
    p = torch.from_numpy(rng.uniform(0.05, 3.0, (B, 14))).cuda()
    y0 = torch.from_numpy(rng.uniform(0.1, 1.0, (B, 7))).cuda()
    t = torch.from_numpy(T14).cuda()
    tg = torch.from_numpy(rng.random(93)).cuda()
    a = engine.solve_local_batch("succmod", p, y0, 5, t, want=("ssr", "score", "Y"), target=tg)
    view = engine.sym_setup(B + 5, lambda h: [h])
    assert tuple(view.shape) == (1, B + 5)
    try:
        for key in ("score", "ssr", "Y"):
            view.fill_(-1.0)
            b = engine.solve_local_batch("succmod", p, y0, 5, t, want=("ssr", "score", "Y"), target=tg, gather_p2p=key)
            assert torch.equal(b["gathered"][0, :B], a[key]), key
            assert bool((b["gathered"][0, B:] == -1.0).all())
            for k in ("ssr", "score", "Y", "status", "nsteps", "nrej"):
                assert torch.equal(a[k], b[k]), k
    finally:
        engine.lib.pk_sym_free(engine._h)
        engine._sym = None


def test_large_batch_properties(engine):
    """Size-independent checks at bench scale (2^18 systems): steady states stay put, the flow is
    a semigroup (solve to t1 then on to t2 == solve to t2), and the map is affine in (A, y0)."""
    import torch
    B = 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(11)
    p = torch.rand((B, 14), generator=gen, device="cuda", dtype=torch.float64) * 2.95 + 0.05
    t = torch.from_numpy(T14).cuda()
    y0 = torch.rand((B, 7), generator=gen, device="cuda", dtype=torch.float64) + 0.1
    r = engine.solve_local_batch("succmod", p, y0, 5, t, want=("sol",))
    assert int((r["status"] != 0).sum()) == 0
    sol = r["sol"]
    # semigroup: restart from the state at t=16 and integrate the remaining grid
    k = 7
    t2 = t[k:].clone()
    r2 = engine.solve_local_batch("succmod", p, sol[:, k].contiguous(), 5, t2, want=("sol",))
    d = (r2["sol"] - sol[:, k:]).abs() / (1e-6 * sol[:, k:].abs() + 1e-9)
    assert float(d.max()) < 1.0
    # affine: y(t; A, y0) with A and y0 doubled equals 2*y(t; A, y0)  (b and y0 scale together)
    p2 = p.clone()
    p2[:, 0] *= 2.0
    r3 = engine.solve_local_batch("succmod", p2, 2.0 * y0, 5, t, want=("sol",))
    d = (r3["sol"] - 2.0 * sol).abs() / (1e-6 * (2.0 * sol).abs() + 1e-9)
    assert float(d.max()) < 1.0
    # steady state of each system (solve M y + b = 0 on the GPU with torch as the checker)
    ys = torch.zeros_like(y0)
    ys[:, 0] = p[:, 0] / p[:, 1]
    # build the tridiagonal system for (P, sites) and solve it densely
    M = torch.zeros((B, 6, 6), device="cuda", dtype=torch.float64)
    S, Dr = p[:, 4:9], p[:, 9:14]
    M[:, 0, 0] = -(p[:, 3] + S[:, 0])
    M[:, 0, 1] = 1.0
    for i in range(5):
        M[:, 1 + i, i] = S[:, i]
        M[:, 1 + i, 1 + i] = -(1.0 + Dr[:, i] + (S[:, i + 1] if i < 4 else 0.0))
        if i < 4:
            M[:, 1 + i, 2 + i] = 1.0
    rhs = torch.zeros((B, 6), device="cuda", dtype=torch.float64)
    rhs[:, 0] = -p[:, 2] * ys[:, 0]
    ys[:, 1:] = torch.linalg.solve(M, rhs)
    r4 = engine.solve_local_batch("succmod", p, ys, 5, t, want=("sol",))
    drift = (r4["sol"] - ys[:, None, :]).abs() / (1e-6 * ys[:, None, :].abs() + 1e-9)
    assert float(drift.max()) < 1.0


@pytest.mark.parametrize("model,ns", [("succmod", 5), ("distmod", 3)])
def test_million_adversarial_draws_stay_inside_the_parity_bound(engine, model, ns):
    """2^20 parameter sets drawn log-uniformly over the reference's fit box [1e-2, 20] (config.toml:189-195) — a quarter
    with random per-system initial states, 5 % with zeroed rates (knockouts), 5 % with all rates equal to a few ulps
    (confluent eigenvalues) — at the library defaults against a tight run of a DIFFERENT integrator (RODAS4 at 1e-10 /
    1e-14), which is itself anchored to the oracle's exact solution (matrix exponential) on a sample.  Every state at
    every output time within the parity bound 1e-6*|ref| + 1e-9."""
    import torch
    from phoskintime_b200.steady import initial_condition
    B = 1 << 20
    n, P, L = __import__("phoskintime_b200").local_dims(model, ns, 14)
    gen = torch.Generator(device="cuda").manual_seed(2024 + ns)
    lo, hi = np.log(1e-2), np.log(20.0)
    p = torch.exp(torch.rand((B, P), generator=gen, device="cuda", dtype=torch.float64) * (hi - lo) + lo)
    q = B // 20
    zero = torch.rand((q, P), generator=gen, device="cuda", dtype=torch.float64) < 0.25
    zero[:, 1] = False                                       # (B = 0: mRNA grows without bound; the reference never fits that)
    p[:q] = torch.where(zero, torch.zeros((), device="cuda", dtype=torch.float64), p[:q])
    p[q:2 * q] = p[q:2 * q, :1] * (1.0 + 4e-16 * torch.randint(-3, 4, (q, P), generator=gen, device="cuda").double())
    y0 = torch.from_numpy(np.asarray(initial_condition(ns, model))).cuda().repeat(B, 1)
    y0[-B // 4:] = torch.rand((B // 4, n), generator=gen, device="cuda", dtype=torch.float64) * 1.4 + 0.1
    t = torch.from_numpy(T14).cuda()
    r = engine.solve_local_batch(model, p, y0, ns, t, want=("sol",))
    assert int((r["status"] != 0).sum()) == 0
    ref = engine.solve_local_batch(model, p, y0, ns, t, want=("sol",), method="rodas4", rtol=1e-10, atol=1e-14)
    assert int((ref["status"] != 0).sum()) == 0
    ratio = ((r["sol"] - ref["sol"]).abs() / (1e-6 * ref["sol"].abs() + 1e-9)).amax(dim=(1, 2))
    worst = float(ratio.max())
    assert worst < 1.0, (worst, int(ratio.argmax()))
    # anchor: the tight run against the oracle's exact solution on the 64 worst and 192 random systems
    idx = torch.cat([ratio.topk(64).indices, torch.randint(0, B, (192,), generator=gen, device="cuda")]).cpu().numpy()
    pc, yc, rc, dc = p[idx].cpu().numpy(), y0[idx].cpu().numpy(), ref["sol"][idx].cpu().numpy(), r["sol"][idx].cpu().numpy()
    for j in range(len(idx)):
        ex = om.exact_linear(model, pc[j], yc[j], ns, T14)
        assert _close(rc[j], ex, 1e-7, 1e-10).all(), j
        assert _close(dc[j], ex, 1e-6, 1e-9).all(), j


def test_knockout_sweep_is_one_batched_launch(engine):
    """paramest/core.py:144-187: every knockout setting solved in one launch equals the oracle's solve of the
    same modified parameter vector."""
    from phoskintime_b200 import knockout
    rng = np.random.default_rng(5)
    ns = 3
    p = rng.uniform(0.2, 2.0, 4 + 2 * ns)
    y0 = np.asarray(om.initial_condition("distmod", ns))
    res = knockout.simulate_knockouts(p, y0, ns, T14, model="distmod")
    assert len(res) == 4 * (ns + 2) and "WT" in res and "Transcription KO_Translation KO_Phospho KO" in res
    for name, r in res.items():
        assert r["status"] == 0
        ex = om.exact_linear("distmod", knockout.apply_knockout(p, r["knockout_setting"], ns), y0, ns, T14)
        assert _close(r["sol_ko"], np.clip(ex, 0, None), 1e-6, 1e-9).all(), name
    assert _close(res["Phospho KO"]["sol_ko"][:, 2:], y0[2:] * np.exp(-np.outer(T14, 1.0 + p[4 + ns:])), 1e-6, 1e-9).all()
