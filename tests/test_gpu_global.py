"""GPU parity tests of the global-network path: `global_net_kernel`
called through the C ABI against the committed goldens of the UNMODIFIED reference
(tests/golden/global_*.npz: `simulate_odeint` stock at rtol=atol=1e-8 = O1, reference RHS bucket by
bucket at 1e-12 = O2, `LOSS_FN` in all 8 modes) and against the oracle restatement.

Tolerances:
  * vs O2:  |ours - ref| <= 1e-6*|ref| + 1e-9   at the library defaults (rtol 2e-6, atol 2e-9)
  * vs O1:  |ours - ref| <= 1e-5*|ref| + 1e-6 everywhere, >= 99 % within 1e-6*|ref| + 1e-7
            (the stock reference integrates THROUGH the kinase-bucket jumps, SURVEY.md quirk 8, and is
             itself up to 0.6x the O2 bound away from O2)
  * loss / objectives / Morris scalar vs the oracle formulas on OUR trajectories: 1e-11 relative
  * LOSS_FN on the REFERENCE's trajectories vs the reference's own loss values: 1e-11 relative
"""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import global_models as og  # noqa: E402
from phoskintime_b200.global_model import (LOSS_FN, GlobalODE_MOO, init_raw_params, metric_time_indices,  # noqa: E402
                                           run_sensitivity_analysis, simulate_batch, simulate_odeint, solve_custom, fold_change_tables,
                                           simulate_and_measure,
                                           synthetic_loss_data, synthetic_system, unpack_params)
import morris as omorris  # noqa: E402

pytestmark = pytest.mark.gpu
# the seven round-1 cases first (tests below address some of them by position), the capacity cases of round 2 last
_LATE = ("global_m0_N300.npz", "global_m2_N7.npz")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "global_*.npz")), key=lambda f: (os.path.basename(f) in _LATE, f))
IDS = [os.path.basename(f)[7:-4] for f in FILES]
T_PROT = [0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0]
T_RNA = [4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0]


def load_case(path):
    g = np.load(path)
    s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]),
                         model=int(g["model"]))
    ld = {k[3:]: g[k] for k in g.files if k.startswith("ld_")}
    for k in ("prot_base_idx", "rna_base_idx", "pho_base_idx"):
        ld[k] = int(ld[k])
    return g, s, ld


def _ratio(a, ref, rtol, atol):
    return float((np.abs(a - ref) / (rtol * np.abs(ref) + atol)).max())


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_trajectories_match_reference(engine, path):
    g, s, _ = load_case(path)
    r = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], engine=engine)
    assert (r["status"] == 0).all(), r["status"]
    Y = r["Y"]
    assert Y.shape == g["Y"].shape
    assert _ratio(Y, g["Y_tight"], 1e-6, 1e-9) <= 1.0
    assert _ratio(Y, g["Y"], 1e-5, 1e-6) <= 1.0
    assert (np.abs(Y - g["Y"]) <= 1e-6 * np.abs(g["Y"]) + 1e-7).mean() >= 0.99
    # margin at the library defaults: within half the parity bound (the stock reference, run at 1e-8, is itself
    # 0.2-0.6 of the bound away from the tight solution); the 1371-state capacity case measures 0.53 of the bound
    assert _ratio(Y, g["Y_tight"], 1e-6, 1e-9) <= (0.5 if int(g["N"]) <= 120 else 0.75)
    assert (r["nsteps"] > 0).all() and (r["nrej"] >= 0).all()


@pytest.mark.parametrize("path", FILES[:2], ids=IDS[:2])
def test_tighter_tolerance_converges(engine, path):
    g, s, _ = load_case(path)
    r6 = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], engine=engine)
    r8 = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], rtol=1e-8, atol=1e-11, engine=engine)
    e6, e8 = _ratio(r6["Y"], g["Y_tight"], 1e-6, 1e-9), _ratio(r8["Y"], g["Y_tight"], 1e-6, 1e-9)
    assert e8 < 0.1 * e6 + 1e-3 and e8 < 0.02
    assert (r8["nsteps"] > r6["nsteps"]).all()


def test_reference_signature_single_system(engine):
    g, s, _ = load_case(FILES[0])
    s.update(**s.unpack_params(g["params"][1]))
    Y = simulate_odeint(s, g["t"], 1e-6, 1e-9, 200000)
    assert Y.shape == g["Y"][1].shape and Y.flags["C_CONTIGUOUS"] and Y.dtype == np.float64
    assert _ratio(Y, g["Y_tight"][1], 1e-6, 1e-9) <= 1.0


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_fused_loss_objectives_metric(engine, path):
    g, s, ld = load_case(path)
    net = s.as_dict()
    mt = metric_time_indices(g["t"], T_PROT, T_RNA, T_PROT)
    lambdas, lam_prior = (1.0, 0.7, 1.3), 0.1
    for mode in (0, 1, 3, 4, 5, 6, 7):
        for metric in (("total_signal",) if mode else ("total_signal", "mean", "variance", "l2_norm")):
            r = simulate_batch(s, g["params"], g["t"], ("Y", "loss", "F", "metric"), y0=g["y0"], loss_data=ld,
                               loss_mode=mode, metric=metric, metric_times=mt, lambdas=lambdas, lambda_prior=lam_prior,
                               engine=engine)
            for b in range(g["params"].shape[0]):
                ref = np.array(og.loss(s.model, r["Y"][b], ld, mode if mode < 7 else -1))
                assert np.allclose(r["loss"][b], ref, rtol=1e-11, atol=1e-13), (mode, b)
                F = og.objectives(ref, ld, og.unpack_params(g["params"][b], net), net["defaults"], lambdas, lam_prior)
                assert np.allclose(r["F"][b], F, rtol=1e-11, atol=1e-13), (mode, b)
                m = og.scalar_metric(r["Y"][b], net, mt, metric)
                assert np.isclose(r["metric"][b], m, rtol=1e-10, atol=1e-12), (metric, b)


def test_outputs_without_trajectory_match(engine):
    """loss/metric computed from the per-CTA trajectory slot (no Y output) equal the ones computed with Y."""
    g, s, ld = load_case(FILES[1])
    mt = metric_time_indices(g["t"], T_PROT, T_RNA, T_PROT)
    P = np.repeat(g["params"], 40, axis=0)            # more systems than a few CTAs hold at once
    a = simulate_batch(s, P, g["t"], ("Y", "loss", "metric"), y0=g["y0"], loss_data=ld, metric_times=mt, engine=engine)
    b = simulate_batch(s, P, g["t"], ("loss", "metric"), y0=g["y0"], loss_data=ld, metric_times=mt, engine=engine)
    assert np.array_equal(a["loss"], b["loss"]) and np.array_equal(a["metric"], b["metric"])
    # repeated parameter rows are reproduced bit for bit whichever CTA integrates them
    assert np.array_equal(a["Y"][0], a["Y"][1]) and np.array_equal(a["loss"][0], a["loss"][39])


@pytest.mark.parametrize("path", [f for f in FILES if "loss_mode0" in np.load(f).files],
                         ids=[i for f, i in zip(FILES, IDS) if "loss_mode0" in np.load(f).files])
def test_loss_fn_on_reference_trajectories(engine, path):
    """LOSS_FN with the reference's signature on the reference's own Y vs the reference's own loss values."""
    g, s, ld = load_case(path)
    for mode in (0, 1, 2, 3, 4, 5, 6, -1):
        ref = g[f"loss_mode{mode}"]
        for b in range(ref.shape[0]):
            mine = np.array(LOSS_FN(g["Y"][b], ld["p_prot"], ld["t_prot"], ld["obs_prot"], ld["w_prot"], ld["p_rna"],
                                    ld["t_rna"], ld["obs_rna"], ld["w_rna"], ld["p_pho"], ld["s_pho"], ld["t_pho"],
                                    ld["obs_pho"], ld["w_pho"], ld["prot_map"], ld["prot_base_idx"], ld["rna_base_idx"],
                                    ld["pho_base_idx"], loss_mode=mode, model=int(g["model"]), engine=engine))
            both_nan = np.isnan(mine) & np.isnan(ref[b])        # mode 2 is NaN by design
            assert np.all(both_nan | (np.abs(mine - ref[b]) <= 1e-11 * np.abs(ref[b]) + 1e-13)), (mode, b, mine, ref[b])


def test_theta_mode_and_device_buffers(engine):
    """raw theta -> softplus on the device (params.py:106-132); torch CUDA tensors are used in place."""
    import torch
    g, s, ld = load_case(FILES[0])
    phys = g["params"]
    theta = np.log(np.expm1(phys))                     # inverse softplus
    a = simulate_batch(s, phys, g["t"], ("Y",), y0=g["y0"], engine=engine)
    b = simulate_batch(s, torch.tensor(theta, device="cuda:0"), g["t"], ("Y",), y0=torch.tensor(g["y0"], device="cuda:0"),
                       theta_mode=True, engine=engine)
    assert isinstance(b["Y"], torch.Tensor) and b["Y"].is_cuda
    assert np.allclose(b["Y"].cpu().numpy(), a["Y"], rtol=1e-9, atol=1e-12)


def test_per_system_initial_conditions_and_failures(engine):
    g, s, _ = load_case(FILES[0])
    B = 6
    P = np.repeat(g["params"][:1], B, axis=0)
    y0 = np.repeat(g["y0"][None, :], B, axis=0)
    y0[1] *= 1.5
    r = simulate_batch(s, P, g["t"], ("Y",), y0=y0, engine=engine)
    assert np.array_equal(r["Y"][0], r["Y"][2]) and not np.allclose(r["Y"][0], r["Y"][1])
    assert np.allclose(r["Y"][1][0], y0[1])
    # a step budget that cannot be met -> status 1 and NaN rows, the batch itself succeeds
    r = simulate_batch(s, P, g["t"], ("Y",), y0=y0, mxstep=5, engine=engine)
    assert (r["status"] == 1).all() and np.isnan(r["Y"]).all()
    # non-finite parameters -> status 3
    Pbad = P.copy()
    Pbad[3, 0] = np.nan
    r = simulate_batch(s, Pbad, g["t"], ("Y",), y0=y0, engine=engine)
    assert r["status"][3] != 0 and np.isnan(r["Y"][3]).all() and (np.delete(r["status"], 3) == 0).all()


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_generic_schur_path_agrees(engine, path):
    """The shared-memory LU fallback (used beyond 128 regulators) against the register-resident
    Gauss-Jordan path and the tight reference."""
    g, s, _ = load_case(path)
    topo = engine.global_upload(s, force_generic=True)
    try:
        r = engine.global_solve_batch(topo, g["params"], g["y0"], g["t"], ("Y",))
    finally:
        engine.global_release(topo)
    assert (r["status"] == 0).all()
    assert _ratio(r["Y"], g["Y_tight"], 1e-6, 1e-9) <= 1.0
    fast = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], engine=engine)
    assert np.allclose(r["Y"], fast["Y"], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("name", ["m0_N36", "m1_N14", "m4_N14", "m2_N10", "m2_N7"])
def test_capacity_fallback_layout_is_bit_identical(engine, name):
    """Capacity fallback (networks beyond one CTA's shared memory): with `force_generic=2` the Schur matrix, the stage
    vectors, the block factors, the state and the parameters of every system live in the per-CTA global scratch instead
    of shared memory.  Same kernel arithmetic, different address space -> bit-identical trajectories, step counts and
    fused outputs as the all-shared generic layout (`force_generic=1`)."""
    g, s, ld = load_case(os.path.join(GOLDEN, f"global_{name}.npz"))
    out = {}
    for level in (1, 2):
        topo = engine.global_upload(s, force_generic=level)
        try:
            engine.global_set_loss_data(topo, ld)
            out[level] = engine.global_solve_batch(topo, g["params"], g["y0"], g["t"], ("Y", "loss"))
            dims = engine.global_dims(topo)
        finally:
            engine.global_release(topo)
    assert dims["smem_bytes"] < 227 * 1024
    assert (out[1]["status"] == 0).all() and (out[2]["status"] == 0).all()
    assert np.array_equal(out[1]["Y"], out[2]["Y"]) and np.array_equal(out[1]["loss"], out[2]["loss"])
    assert np.array_equal(out[1]["nsteps"], out[2]["nsteps"]) and np.array_equal(out[1]["nrej"], out[2]["nrej"])
    assert _ratio(out[2]["Y"], g["Y_tight"], 1e-6, 1e-9) <= 0.5


@pytest.mark.parametrize("name", ["m0_N36", "m0_N120", "m1_N14", "m4_N14", "m2_N10", "m2_N24", "m0_N300"])
def test_iterative_schur_solve_equals_exact_inversion(engine, name, monkeypatch):
    """Steps whose Schur system has ||K||_inf < 0.5 solve it by fixed-point sweeps to a bound of 1e-12 instead of the
    register Gauss-Jordan inverse (csrc/global_net.cuh: schur_neumann).  PHOSKIN_SCHUR_ITER=0 forces the exact inversion in
    every step: same accepted / rejected step counts and trajectories equal to 1e-9 relative (+1e-12) - two decades below the
    step's own tolerance, three below the parity bound."""
    g, s, _ = load_case(os.path.join(GOLDEN, f"global_{name}.npz"))
    P = g["params"][:2]
    monkeypatch.setenv("PHOSKIN_SCHUR_ITER", "0")
    exact = simulate_batch(s, P, g["t"], ("Y",), y0=g["y0"], engine=engine)
    monkeypatch.delenv("PHOSKIN_SCHUR_ITER")
    it = simulate_batch(s, P, g["t"], ("Y",), y0=g["y0"], engine=engine)
    assert (exact["status"] == 0).all() and (it["status"] == 0).all()
    assert np.array_equal(exact["nsteps"], it["nsteps"]) and np.array_equal(exact["nrej"], it["nrej"])
    assert np.all(np.abs(it["Y"] - exact["Y"]) <= 1e-9 * np.abs(exact["Y"]) + 1e-12), _ratio(it["Y"], exact["Y"], 1e-9, 1e-12)
    assert _ratio(it["Y"], g["Y_tight"][:2], 1e-6, 1e-9) <= 1.0


def test_network_beyond_one_cta_runs(engine):
    """N = 300 proteins (1371 states, 2332 parameters, 254 regulators): the Schur block alone (254 x 255 doubles = 518 KB)
    exceeds a CTA's 227 KB, so the upload must choose the overflow layout by itself instead of refusing; the trajectories are
    held to the golden of the unmodified reference by the parametrised tests above - here: layout facts and a fused batch."""
    g, s, ld = load_case(os.path.join(GOLDEN, "global_m0_N300.npz"))
    topo = engine.global_upload(s)
    try:
        dims = engine.global_dims(topo)
        assert dims["state_dim"] == 1371 and dims["n_reg"] > 128 and dims["smem_bytes"] <= 227 * 1024
        engine.global_set_loss_data(topo, ld)
        rng = np.random.default_rng(3)
        P = g["params"][:1] * np.exp(0.05 * rng.standard_normal((6, g["params"].shape[1])))
        r = engine.global_solve_batch(topo, P, g["y0"], g["t"], ("Y", "loss"))
        assert (r["status"] == 0).all() and np.isfinite(r["loss"]).all()
        for b in range(0, 6, 2):
            assert np.allclose(r["loss"][b], og.loss_noncomb(r["Y"][b], ld, 0), rtol=1e-11, atol=1e-13)
    finally:
        engine.global_release(topo)


def test_time_grid_subsets(engine):
    """t_eval need not contain the kinase-grid points: steps still land on them (same values at shared times)."""
    g, s, _ = load_case(FILES[2])
    full = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], engine=engine)["Y"]
    sub_t = g["t"][[0, 3, 7, 11, 14]]
    sub = simulate_batch(s, g["params"], sub_t, ("Y",), y0=g["y0"], engine=engine)["Y"]
    assert _ratio(sub, g["Y_tight"][:, [0, 3, 7, 11, 14]], 1e-6, 1e-9) <= 1.0
    assert np.allclose(sub, full[:, [0, 3, 7, 11, 14]], rtol=2e-6, atol=1e-9)


def test_unsupported_model_and_bad_inputs(engine):
    from phoskintime_b200 import PhoskinError
    s = synthetic_system(seed=1, N=6, K=3, max_sites=2, model=0)
    s.model = 3
    with pytest.raises(PhoskinError):
        engine.global_upload(s)
    s.model = 0
    with pytest.raises(ValueError):
        simulate_batch(s, np.ones((2, 5)), [0.0, 1.0], engine=engine)
    with pytest.raises(PhoskinError):
        simulate_batch(s, s.pack_params()[None], [0.0, 1.0, 1.0], engine=engine)      # not increasing
    with pytest.raises(PhoskinError):
        simulate_batch(s, s.pack_params()[None], [0.0, 1.0], ("loss",), engine=engine)  # no loss tables


def test_full_size_network_sanity(engine):
    """BASELINE configs[4] shape (N=120, state_dim ~500): every system integrates, results are positive,
    finite, and a tighter tolerance agrees to 1e-6 (size-independent self-consistency)."""
    s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
    rng = np.random.default_rng(0)
    base = s.pack_params()
    P = base[None, :] * np.exp(0.05 * rng.standard_normal((32, base.size)))
    t = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
    a = simulate_batch(s, P, t, ("Y",), engine=engine)
    b = simulate_batch(s, P, t, ("Y",), rtol=1e-8, atol=1e-11, engine=engine)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.isfinite(a["Y"]).all() and (a["Y"] > -1e-9).all()
    assert _ratio(a["Y"], b["Y"], 1e-6, 1e-9) <= 1.0


def test_combinatorial_full_size_and_uncoupled(engine):
    """MODEL 2 at the BASELINE configs[4] network size (N = 120, ~1000 pattern states): every system integrates,
    a tighter tolerance agrees to 1e-6, the shared-memory LU fallback agrees; and a network without TF edges (no
    Schur block) matches the oracle's tight solution block by block."""
    s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=2)
    assert s.idx.state_dim == int((1 + (1 << s.idx.n_sites)).sum())
    rng = np.random.default_rng(0)
    base = s.pack_params()
    P = base[None, :] * np.exp(0.05 * rng.standard_normal((8, base.size)))
    t = np.array([0.0, 0.5, 1.0, 4.0, 15.0, 16.0, 60.0, 240.0, 960.0])
    a = simulate_batch(s, P, t, ("Y",), engine=engine)
    b = simulate_batch(s, P, t, ("Y",), rtol=1e-8, atol=1e-11, engine=engine)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.isfinite(a["Y"]).all() and (a["Y"] > -1e-9).all()
    assert _ratio(a["Y"], b["Y"], 1e-6, 1e-9) <= 1.0
    u = synthetic_system(seed=8, N=6, K=3, max_sites=4, tf_density=0.0, model=2)
    tu = np.array([0.0, 0.5, 1.0, 4.0, 16.0, 60.0, 960.0])
    r = simulate_batch(u, u.pack_params()[None, :], tu, ("Y",), engine=engine)
    assert r["status"][0] == 0 and engine.global_dims(u._topo_id[engine.token])["n_reg"] == 0
    assert _ratio(r["Y"][0], og.simulate_exact_buckets(2, u.as_dict(), tu), 1e-6, 1e-9) <= 1.0


def test_simulate_until_steady_long_log_grid(engine):
    """analysis.py:29-67: 1000 log-spaced outputs over 24 h (beyond the last kinase-grid point): every output is hit
    exactly; a sample of rows is compared with the oracle's tight solution, the convergence rate with its definition."""
    from phoskintime_b200.global_model import final_rate_of_change, simulate_until_steady, steady_check_batch
    g, s, _ = load_case(FILES[0])
    s.update(**s.unpack_params(g["params"][1]))
    t, Y = simulate_until_steady(s, t_max=1440.0, n_points=1000)
    assert t.shape == (1000,) and t[0] == 0.0 and np.isclose(t[1], 1e-3) and np.isclose(t[-1], 1440.0)
    assert Y.shape == (1000, s.idx.state_dim) and np.isfinite(Y).all()
    rows = np.array([0, 1, 150, 400, 555, 700, 850, 930, 999])
    ref = og.simulate_exact_buckets(int(g["model"]), s.as_dict(), t[rows], params=s.unpack_params(g["params"][1]))
    assert _ratio(Y[rows], ref, 1e-6, 1e-9) <= 1.0
    rate = final_rate_of_change(t, Y)
    assert np.isclose(rate, np.linalg.norm(Y[-1] - Y[-2]) / (t[-1] - t[-2])) and rate < 1e-3
    tb, Yb, rb, st = steady_check_batch(s, g["params"], engine=engine)
    assert (st == 0).all() and np.array_equal(Yb[1], Y) and np.isclose(rb[1], rate)


@pytest.mark.parametrize("path", [FILES[0], FILES[4]], ids=[IDS[0], IDS[4]])
def test_fold_change_tables_match_oracle(engine, path):
    """simulate_and_measure's three tables (simulate.py:105-182) from the kernel epilogue vs the oracle's tabulation of
    the same trajectories; the reference-signature wrapper returns the same numbers as DataFrames."""
    g, s, _ = load_case(path)
    net = s.as_dict()
    tab = fold_change_tables(s, g["params"], T_PROT, T_RNA, T_PROT, engine=engine)
    assert (tab["status"] == 0).all() and np.array_equal(tab["times"], np.unique(np.concatenate([T_PROT, T_RNA])))
    mt = metric_time_indices(tab["times"], T_PROT, T_RNA, T_PROT)
    Y = simulate_batch(s, g["params"], tab["times"], ("Y",), rtol=1e-5, atol=1e-7, mxstep=5000, engine=engine)["Y"]
    for b in range(g["params"].shape[0]):
        P, R, PH = og.fc_tables(Y[b], net, mt)
        assert np.allclose(tab["fc_prot"][b], P, rtol=1e-12, atol=0) and np.allclose(tab["fc_rna"][b], R, rtol=1e-12, atol=0)
        assert np.allclose(tab["fc_pho"][b], PH, rtol=1e-12, atol=0)
    s.update(**s.unpack_params(g["params"][1]))
    df_p, df_r, df_ph = simulate_and_measure(s, None, T_PROT, T_RNA, T_PROT)
    assert list(df_p.columns) == ["protein", "time", "pred_fc"] and list(df_ph.columns) == ["protein", "psite", "time", "pred_fc"]
    assert len(df_p) == s.idx.N * len(T_PROT) and len(df_r) == s.idx.N * len(T_RNA) and len(df_ph) == s.idx.total_sites * len(T_PROT)
    assert np.array_equal(df_p["pred_fc"].to_numpy(), tab["fc_prot"][1].reshape(-1))
    assert np.array_equal(df_r["time"].to_numpy()[:len(T_RNA)], np.asarray(T_RNA))
    assert (df_p["pred_fc"].to_numpy()[::len(T_PROT)] == 1.0).all()          # every protein's t = 0 row is its own baseline


def test_solve_custom_signature(engine):
    g, s, _ = load_case(FILES[2])
    s.update(**s.unpack_params(g["params"][2]))
    Y = solve_custom(s, g["y0"] * 1.0, g["t"], 1e-6, 1e-9, method="rosenbrock")
    assert _ratio(Y, g["Y_tight"][2], 1e-6, 1e-9) <= 1.0
    # the default reproduces the reference's DOPRI5: its Hermite outputs are ~1e-5 off the tight solution by design
    Yd = solve_custom(s, g["y0"] * 1.0, g["t"], 1e-5, 1e-7)
    assert Yd.shape == Y.shape and _ratio(Yd, g["Y_tight"][2], 1e-3, 1e-6) <= 1.0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "globalrhs_*.npz"))),
                         ids=[os.path.basename(f)[10:-4] for f in sorted(glob.glob(os.path.join(GOLDEN, "globalrhs_*.npz")))])
def test_custom_dopri5_matches_reference_solve_custom(engine, path):
    """`pk_global_solve_custom` against the UNMODIFIED reference `solve_custom` (jacspeedup.py:31-67 ->
    adaptive_rk45_model01 / _model2, solvers.py:292-758) on the same systems and tolerances: the same step sequence up
    to rounding, so the outputs agree far below the solver's own error (bound: 1e-7 relative + 1e-10; the reference's
    DOPRI5 itself is ~1e-5 off the tight solution)."""
    from phoskintime_b200.global_model import solve_custom_batch
    r = np.load(path)
    g = np.load(os.path.join(GOLDEN, f"global_m{int(r['model'])}_N{int(r['N'])}.npz"))
    s = synthetic_system(seed=int(r["seed"]), N=int(r["N"]), K=int(r["K"]), max_sites=int(r["max_sites"]), model=int(r["model"]))
    rows = r["custom_rows"]
    for tag in ("a", "b"):
        rt, at = r[f"custom_tol_{tag}"]
        out = solve_custom_batch(s, g["params"][rows], s.y0(), r["custom_t"], rt, at, engine=engine)
        assert (out["status"] == 0).all() and (out["nsteps"] >= 960).all()          # dt <= 1 over 960 minutes
        ref = r[f"custom_Y_{tag}"]
        assert np.all(np.abs(out["Y"] - ref) <= 1e-7 * np.abs(ref) + 1e-10), _ratio(out["Y"], ref, 1e-7, 1e-10)


def test_population_objectives_match_oracle(engine):
    """GlobalODE_MOO.evaluate_batch: raw thetas -> softplus, prior, solve, loss, normalised objectives on the
    device vs the oracle's objectives (optproblem.py:87-160) evaluated on the device trajectories."""
    g, s, ld = load_case(FILES[0])
    net = s.as_dict()
    bounds = {k: (1e-4, 50.0) for k in ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i", "tf_scale")}
    theta0, slices, xl, xu = init_raw_params(s.defaults, bounds)
    lam = {"protein": 1.0, "rna": 0.5, "phospho": 2.0, "prior": 0.1}
    prob = GlobalODE_MOO(s, slices, ld, s.defaults, lam, g["t"], xl, xu, engine=engine)
    rng = np.random.default_rng(3)
    X = theta0[None, :] + 0.3 * rng.standard_normal((6, theta0.size))
    X[5, 0] = np.nan                                             # a broken individual
    F = prob.evaluate_batch(X)
    assert F.shape == (6, 3) and (F[5] == 1e12).all()
    for b in range(5):
        p = unpack_params(X[b], slices)
        Y = simulate_batch(s, s.pack_params(p)[None], g["t"], ("Y",), engine=engine)["Y"][0]
        ref = og.objectives(og.loss_noncomb(Y, ld, 0), ld, p, net["defaults"], (1.0, 0.5, 2.0), 0.1)
        assert np.allclose(F[b], ref, rtol=1e-9, atol=1e-12), (b, F[b], ref)
    out = {}
    prob._evaluate(X[1], out)
    assert np.array_equal(out["F"], F[1]) and np.allclose(s.pack_params(), s.pack_params(unpack_params(X[1], slices)))


def test_global_morris_ranking_matches_oracle(engine):
    """Morris mu*/sigma of the fused scalar metric vs the oracle pipeline on the SAME sample X: Y_ref of EVERY row comes
    from the oracle port (reference RHS + finite-difference Jacobian through LSODA, then the reference's fold-change
    scalar), goes through the restated SALib analysis, and the ranking of the parameters must be identical."""
    g, s, _ = load_case(FILES[0])
    net = s.as_dict()
    s.update(**s.unpack_params(g["params"][0]))
    res = run_sensitivity_analysis(s, metric="total_signal", N=3, num_levels=4, seed=11, engine=engine)
    X, D = res["X"], len(res["names"])
    assert X.shape == (3 * (D + 1), D) and (res["status"] == 0).all()
    times = np.unique(np.concatenate([T_PROT, T_RNA]))
    mt = metric_time_indices(times, T_PROT, T_RNA, T_PROT)
    Y_ref = np.empty(X.shape[0])
    for r in range(X.shape[0]):
        Y = og.simulate_odeint(0, net, times, 1e-8, 1e-8, 200000, params=og.unpack_params(X[r], net))
        Y_ref[r] = og.scalar_metric(Y, net, mt, "total_signal")
    assert np.allclose(res["Y"], Y_ref, rtol=2e-5, atol=0), float(np.max(np.abs(res["Y"] - Y_ref) / np.abs(Y_ref)))
    Si = omorris.analyze(X, Y_ref, D, num_levels=4, scaled=False)
    ref_order = np.argsort(-Si["mu_star"], kind="stable")
    # identical ranking wherever the reference separates neighbours by more than the integration noise of either side
    assert np.array_equal(res["order"][:10], ref_order[:10])
    gap_ok = np.abs(np.diff(Si["mu_star"][ref_order])) > 1e-3 * Si["mu_star"][ref_order][:-1]
    mine = np.asarray(res["order"])
    assert all(mine[i] == ref_order[i] for i in range(D - 1) if gap_ok[i] and (i == 0 or gap_ok[i - 1]))
    assert np.allclose(res["mu_star"], Si["mu_star"], rtol=5e-3, atol=1e-9 * np.max(Si["mu_star"]))
    assert np.array_equal(np.argsort(-res["sigma"], kind="stable")[:5], np.argsort(-Si["sigma"], kind="stable")[:5])
    # and the device reduction itself is the restated analysis to rounding
    Sg = omorris.analyze(X, res["Y"], D, num_levels=4, scaled=False)
    assert np.allclose(res["mu_star"], Sg["mu_star"], rtol=1e-9, atol=1e-12) and np.allclose(res["sigma"], Sg["sigma"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("N,K,expect_tile_range", [(70, 20, (33, 64)), (150, 40, (97, 128))], ids=["tile4", "tile8"])
def test_other_register_tile_sizes(engine, N, K, expect_tile_range):
    """Regulator sets of 33-64 and 97-128 proteins select the 4x4 / 8x8 register tiles: the default run agrees with
    a tight-tolerance run, and (where shared memory allows) with the shared-memory LU fallback."""
    s = synthetic_system(seed=21, N=N, K=K, max_sites=3, model=0)
    topo = engine.global_upload(s)
    dims = engine.global_dims(topo)
    assert expect_tile_range[0] <= dims["n_reg"] <= expect_tile_range[1], dims
    rng = np.random.default_rng(1)
    base = s.pack_params()
    P = base[None, :] * np.exp(0.1 * rng.standard_normal((6, base.size)))
    t = np.array([0.0, 0.5, 1.0, 4.0, 15.0, 16.0, 60.0, 240.0, 960.0])
    a = engine.global_solve_batch(topo, P, s.y0(), t, ("Y",))
    b = engine.global_solve_batch(topo, P, s.y0(), t, ("Y",), rtol=1e-8, atol=1e-11)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert _ratio(a["Y"], b["Y"], 1e-6, 1e-9) <= 1.0
    if dims["n_reg"] <= 64:
        tg = engine.global_upload(s, force_generic=True)
        c = engine.global_solve_batch(tg, P, s.y0(), t, ("Y",), rtol=1e-8, atol=1e-11)
        engine.global_release(tg)
        assert np.allclose(c["Y"], b["Y"], rtol=1e-7, atol=1e-10)
    engine.global_release(topo)


def test_network_without_transcriptional_coupling(engine):
    """No TF edges: empty regulator set (no Schur block at all); every protein block integrates on its own and
    must match the oracle's tight solution."""
    s = synthetic_system(seed=8, N=6, K=3, max_sites=3, tf_density=0.0, model=1)
    assert s.TF_indptr[-1] == 0
    net = s.as_dict()
    t = np.array([0.0, 0.5, 1.0, 4.0, 16.0, 60.0, 960.0])
    r = simulate_batch(s, s.pack_params()[None, :], t, ("Y",), engine=engine)
    assert r["status"][0] == 0 and engine.global_dims(s._topo_id[engine.token])["n_reg"] == 0
    ref = og.simulate_exact_buckets(1, net, t)
    assert _ratio(r["Y"][0], ref, 1e-6, 1e-9) <= 1.0


def test_oversized_network_is_rejected_with_a_message(engine):
    """N = 400 (1800 states, > 380 regulators) runs through the overflow layout; a network whose staged topology alone
    (TF CSR of 3.6e5 entries) exceeds a CTA's shared memory is refused with a message instead of a launch failure."""
    from phoskintime_b200 import PhoskinError
    s = synthetic_system(seed=2, N=400, K=40, max_sites=4, model=0)
    topo = engine.global_upload(s)
    try:
        r = engine.global_solve_batch(topo, s.pack_params()[None, :], s.y0(), np.array([0.0, 1.0, 16.0, 960.0]), ("Y",))
        assert (r["status"] == 0).all() and np.isfinite(r["Y"]).all()
    finally:
        engine.global_release(topo)
    huge = synthetic_system(seed=2, N=6000, K=40, max_sites=2, tf_density=0.01, model=0)
    with pytest.raises(PhoskinError, match="shared memory"):
        engine.global_upload(huge)


# ------------------------------------------------------------------------------------------------------------------
# The objective / observable layer against the UNMODIFIED reference (tests/golden/globalobj_*.npz, written by
# oracle/gen_golden_objectives.py: the reference's own GlobalODE_MOO._evaluate, simulate_and_measure and
# _compute_scalar_metric).  The reference integrates these with LSODA (rtol = atol = 1e-8 for the objectives, through the
# kinase-bucket jumps; rtol 1e-5 / atol 1e-7 for the observables, simulate.py:109), so the bounds below are the
# reference's OWN integration error, measured on B200: F <= 8.3e-8, tables <= 6.9e-5, scalars <= 5.6e-6 relative.
OBJ_FILES = sorted(glob.glob(os.path.join(GOLDEN, "globalobj_*.npz")))
OBJ_IDS = [os.path.basename(f)[10:-4] for f in OBJ_FILES]
OBJ_KEYS = ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i", "tf_scale")


def load_obj_case(path):
    g, s, ld = load_case(path)
    defaults = {k: (g[f"def_{k}"] if k != "tf_scale" else float(g[f"def_{k}"])) for k in OBJ_KEYS}
    slices = {k: slice(int(a), int(b)) for k, (a, b) in zip(OBJ_KEYS, g["slices"])}
    return g, s, ld, defaults, slices


@pytest.mark.parametrize("path", OBJ_FILES, ids=OBJ_IDS)
def test_objectives_match_reference_evaluate(engine, path):
    """GlobalODE_MOO.evaluate_batch (raw thetas in, F[B,3] out, ONE launch) vs the reference's _evaluate per vector."""
    g, s, ld, defaults, slices = load_obj_case(path)
    lam = dict(zip(("protein", "rna", "phospho", "prior"), (float(v) for v in g["lambdas"])))
    prob = GlobalODE_MOO(s, slices, ld, defaults, lam, g["t_grid"], engine=engine)
    F = prob.evaluate_batch(g["theta"])
    assert F.shape == g["F"].shape
    assert np.all(np.abs(F - g["F"]) <= 1e-6 * np.abs(g["F"])), np.max(np.abs(F - g["F"]) / np.abs(g["F"]))


@pytest.mark.parametrize("path", OBJ_FILES, ids=OBJ_IDS)
def test_fold_change_tables_and_metrics_match_reference(engine, path):
    """fold_change_tables / the fused Morris scalar vs the reference's simulate_and_measure / _compute_scalar_metric."""
    g, s, _, _, _ = load_obj_case(path)
    tab = fold_change_tables(s, g["phys"], g["t_prot"], g["t_rna"], g["t_prot"], engine=engine)
    assert (tab["status"] == 0).all() and np.array_equal(tab["times"], g["t_grid"])
    for k in ("fc_prot", "fc_rna", "fc_pho"):
        mine = tab[k].reshape(g[k].shape)
        assert np.all(np.abs(mine - g[k]) <= 3e-4 * np.abs(g[k])), (k, np.max(np.abs(mine - g[k]) / np.abs(g[k])))
    mt = metric_time_indices(g["t_grid"], g["t_prot"], g["t_rna"], g["t_prot"])
    for m, name in enumerate(g["metric_names"]):
        r = simulate_batch(s, g["phys"], g["t_grid"], ("metric",), rtol=1e-5, atol=1e-7, mxstep=5000, metric=str(name),
                           metric_times=mt, engine=engine)
        assert np.all(np.abs(r["metric"] - g["metrics"][:, m]) <= 3e-5 * np.abs(g["metrics"][:, m])), name


# ------------------------------------------------------------------------------------------------------------------
# RHS / Jacobian export (pk_global_rhs_batch) and the model_ivp closures against the UNMODIFIED reference
# (tests/golden/globalrhs_*.npz, oracle/gen_golden_rhs.py: rhs_odeint, fd_jacobian_odeint, make_solve_ivp_fun_*).
RHS_FILES = sorted(glob.glob(os.path.join(GOLDEN, "globalrhs_*.npz")))
RHS_IDS = [os.path.basename(f)[10:-4] for f in RHS_FILES]


@pytest.mark.parametrize("path", RHS_FILES, ids=RHS_IDS)
def test_rhs_and_jacobian_match_reference(engine, path):
    """f = the integrator's own eval_rhs vs the reference's rhs_odeint (1e-12: the reference kernels are numba fastmath);
    analytic J vs the reference's forward-difference Jacobian (step 1e-8*max(1,|y|): 1e-6 relative to the row scale),
    and vs a central difference of the device RHS itself (1e-8)."""
    from phoskintime_b200.global_model import make_solve_ivp_fun
    g = np.load(path)
    s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]), model=int(g["model"]))
    fun = make_solve_ivp_fun(s, engine=engine)
    B, n = g["Y"].shape
    f, J = engine.global_rhs_batch(fun_topology(s, engine), g["params"], g["Y"], g["t"], want_jac=True)
    assert f.shape == (B, n) and J.shape == (B, n, n)
    assert np.all(np.abs(f - g["f"]) <= 1e-12 * np.abs(g["f"]) + 1e-13), float(np.max(np.abs(f - g["f"])))
    scale = np.maximum(np.abs(g["J_fd"]).max(axis=2, keepdims=True), 1.0)
    assert np.all(np.abs(J - g["J_fd"]) <= 2e-6 * scale), float(np.max(np.abs(J - g["J_fd"]) / scale))
    assert np.array_equal(J != 0, np.abs(g["J_fd"]) > 1e-7 * scale) or np.mean((J != 0) == (np.abs(g["J_fd"]) > 1e-7 * scale)) > 0.995
    # central difference of the device RHS: rounding-level agreement with the analytic Jacobian
    b = 3
    y0, h = g["Y"][b], 1e-6 * np.maximum(1.0, np.abs(g["Y"][b]))
    Yp = np.concatenate([y0[None, :] + np.diag(h), y0[None, :] - np.diag(h)])
    fp = engine.global_rhs_batch(fun_topology(s, engine), g["params"][b], Yp, float(g["t"][b]))
    Jc = ((fp[:n] - fp[n:]) / (2 * h)[:, None]).T
    assert np.all(np.abs(J[b] - Jc) <= 1e-7 * scale[b]), float(np.max(np.abs(J[b] - Jc) / scale[b]))
    # the closure form: current parameters of the system, one state
    s.update(**s.unpack_params(g["params"][0]))
    assert np.array_equal(fun(float(g["t"][0]), g["Y"][0]), f[0]) and np.array_equal(fun.jac(float(g["t"][0]), g["Y"][0]), J[0])


def fun_topology(s, engine):
    from phoskintime_b200.global_model.simulate import _topology
    return _topology(s, engine)


@pytest.mark.parametrize("path", RHS_FILES, ids=RHS_IDS)
def test_model_ivp_closures_match_reference(engine, path):
    """make_solve_ivp_fun_{distributive,sequential,combinatorial,saturating} (model_ivp.py:49-277): same keyword
    signature, caller-supplied TF inputs and S_all, dy within 1e-12 of the reference closure."""
    from phoskintime_b200.global_model import (make_solve_ivp_fun_combinatorial, make_solve_ivp_fun_distributive,
                                               make_solve_ivp_fun_saturating, make_solve_ivp_fun_sequential)
    g = np.load(path)
    model = int(g["model"])
    s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]), model=model)
    p = s.unpack_params(g["ivp_params"])
    kw = dict(A_i=p["A_i"], B_i=p["B_i"], C_i=p["C_i"], D_i=p["D_i"], Dp_i=p["Dp_i"], E_i=p["E_i"], tf_scale=p["tf_scale"],
              tf_input=g["ivp_tf"], offset_y=s.idx.offset_y, offset_s=s.idx.offset_s, n_sites=s.idx.n_sites, engine=engine)
    if model == 2:
        fun = make_solve_ivp_fun_combinatorial(S_cache=g["ivp_S_cache"], jb=int(g["ivp_jb"]), n_states=s.idx.n_states, **kw)
    else:
        fun = {0: make_solve_ivp_fun_distributive, 1: make_solve_ivp_fun_sequential, 4: make_solve_ivp_fun_saturating}[model](
            S_all=g["ivp_S_all"], **kw)
    for y, ref in zip(g["ivp_Y"], g["ivp_f"]):
        dy = fun(1.0, y)
        assert np.all(np.abs(dy - ref) <= 1e-12 * np.abs(ref) + 1e-13), float(np.max(np.abs(dy - ref)))
    dyb = fun.batch(1.0, g["ivp_Y"])
    assert np.allclose(dyb, g["ivp_f"], rtol=1e-12, atol=1e-13)
    # a callable TF input (time and state dependent) is honoured per state
    fun2_kw = dict(kw, tf_input=lambda t, y=None: g["ivp_tf"] * (1.0 + 0.0 * t))
    if model != 2:
        fun2 = {0: make_solve_ivp_fun_distributive, 1: make_solve_ivp_fun_sequential, 4: make_solve_ivp_fun_saturating}[model](
            S_all=g["ivp_S_all"], **fun2_kw)
        assert np.array_equal(fun2(1.0, g["ivp_Y"][0]), fun(1.0, g["ivp_Y"][0]))
    Jb = fun.jac(1.0, g["ivp_Y"][1])
    assert Jb.shape == (s.idx.state_dim, s.idx.state_dim) and np.isfinite(Jb).all()
