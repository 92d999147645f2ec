"""CPU: host-side mirror of the reference's global_model interface (no GPU, no compute calls)."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

import phoskintime_b200 as pk
from phoskintime_b200 import _lib
from phoskintime_b200.global_model import (GlobalODE_MOO, compute_bounds, init_raw_params, metric_time_indices,
                                           synthetic_loss_data, synthetic_system, unpack_params)
from phoskintime_b200.global_model.network import PARAM_KEYS
from phoskintime_b200.global_model.optproblem import inv_softplus, softplus

BOUNDS = {k: (1e-4, 50.0) for k in PARAM_KEYS + ("tf_scale",)}
T15 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])


def test_global_struct_layouts_match_header():
    lib = _lib.load()
    assert lib.pk_sizeof_global_job() == ctypes.sizeof(_lib.PkGlobalJob)
    job = _lib.PkGlobalJob()
    lib.pk_global_job_init(ctypes.byref(job))
    assert job.metric == -1 and list(job.lambdas) == [1.0, 1.0, 1.0] and job.lambda_prior == 0.0
    # field order of the topology mirror = declaration order in include/phoskin_b200.h
    text = open(os.path.join(ROOT, "include", "phoskin_b200.h")).read()
    body = text[text.index("typedef struct pk_global_topology"):text.index("} pk_global_topology;")]
    order = [f for f, _ in _lib.PkGlobalTopology._fields_]
    pos = [body.index(" " + f) if (" " + f) in body else body.index("*" + f) for f in order]
    assert pos == sorted(pos)


def test_system_layout_and_argument_packing():
    s = synthetic_system(seed=3, N=12, K=5, max_sites=3, model=1)
    idx = s.idx
    assert idx.state_dim == int((2 + idx.n_sites).sum()) and idx.offset_y[0] == 0
    assert np.array_equal(np.diff(idx.offset_y), 2 + idx.n_sites[:-1])
    args = s.odeint_args()
    assert len(args) == 23                                    # network.py:508-526
    assert args[7] == s.tf_scale and args[13] == idx.total_sites and args[17] == idx.N
    assert args[10].dtype == np.int32 and args[12].dtype == np.float64
    y0 = s.y0()
    assert y0[idx.offset_y[0]] == 1.0 and y0[idx.offset_y[0] + 1] == 1.0
    v = s.pack_params()
    assert v.size == s.n_params == s.K + 5 * idx.N + idx.total_sites + 1
    p = s.unpack_params(v * 2.0)
    s.update(**p)
    assert np.array_equal(s.pack_params(), v * 2.0)           # write-through (network.py:293-302)
    with pytest.raises(KeyError):
        synthetic_system(seed=1, N=6, K=3, model="michaelis")


def test_combinatorial_layout_and_27_tuple():
    """MODEL 2 mirror: block [mRNA, 2^ns patterns] (network.py:131-149), y0 (network.py:431-436), the 27-tuple of
    network.py:471-505 and the hypercube edge lists (models.py:435-485, restated independently in the oracle)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import global_models as og
    s = synthetic_system(seed=4, N=9, K=4, max_sites=3, model="combinatorial")
    idx = s.idx
    assert s.model == 2 and np.array_equal(idx.n_states, 1 << idx.n_sites)
    assert idx.state_dim == int((1 + idx.n_states).sum())
    assert np.array_equal(np.diff(idx.offset_y), 1 + idx.n_states[:-1])
    y0 = s.y0()
    for i in range(idx.N):
        st = idx.offset_y[i]
        assert y0[st] == 1.0 and y0[st + 1] == 1.0 and np.all(y0[st + 2:st + 1 + idx.n_states[i]] == 0.01)
    with pytest.raises(ValueError):
        s.odeint_args()
    args = s.odeint_args(s.build_S_cache())
    assert len(args) == 27 and args[9].shape == (idx.total_sites, s.kin_grid.size)
    assert np.allclose(args[9], og.s_cache(s.as_dict(), s.c_k), rtol=1e-14, atol=0.0)
    for mine, ref in zip(args[18:23], og.comb_tables(idx.n_sites)):
        assert mine.dtype == np.int32 and np.array_equal(mine, ref)
    # every pattern has one forward edge per unset bit
    assert args[22].sum() == sum(ns * (1 << ns) // 2 for ns in idx.n_sites)
    ld = synthetic_loss_data(s, T15, seed=1)
    assert np.array_equal(ld["prot_map"][:, 1], idx.n_states)
    # blocks of 32..256 patterns are supported (one warp per block on the device); 2^9 patterns per protein are refused
    big = synthetic_system(seed=27, N=7, K=4, max_sites=6, model=2)
    assert big.idx.n_sites.max() == 6 and big.idx.state_dim == 7 + int((1 << big.idx.n_sites).sum())
    with pytest.raises(ValueError):
        synthetic_system(seed=4, N=9, K=4, max_sites=12, model=2)


def test_raw_parameter_transform_roundtrip():
    s = synthetic_system(seed=4, N=9, K=4, max_sites=2)
    theta0, slices, xl, xu = init_raw_params(s.defaults, BOUNDS)
    assert theta0.size == s.n_params and (xl < theta0).all() and (theta0 < xu).all()
    assert list(slices) == list(PARAM_KEYS) + ["tf_scale"]
    p = unpack_params(theta0, slices)
    assert np.allclose(s.pack_params(p), s.pack_params(s.defaults), rtol=1e-12)
    x = np.array([-30.0, -1.0, 0.0, 5.0, 25.0])
    assert np.allclose(inv_softplus(softplus(x)[1:]), x[1:]) and softplus(x)[4] == 25.0
    # the batch evaluator insists on the packed order it hands to the device
    ld = synthetic_loss_data(s, T15, seed=1)
    lam = {"protein": 1.0, "rna": 1.0, "phospho": 1.0, "prior": 0.1}
    prob = GlobalODE_MOO(s, slices, ld, s.defaults, lam, T15)
    assert prob.n_var == s.n_params and prob.n_obj == 3
    assert np.isclose(prob.norm_p, 1.0 / ld["w_prot"].sum())
    bad = dict(slices)
    bad["A_i"], bad["B_i"] = slices["B_i"], slices["A_i"]
    with pytest.raises(ValueError):
        GlobalODE_MOO(s, bad, ld, s.defaults, lam, T15)


def test_sensitivity_problem_and_metric_rows():
    s = synthetic_system(seed=5, N=7, K=3, max_sites=2)
    params = {**{k: getattr(s, k) for k in PARAM_KEYS}, "tf_scale": s.tf_scale}
    prob = compute_bounds(params)
    assert prob["num_vars"] == s.n_params and prob["names"][0] == "c_k_0" and prob["names"][-1] == "tf_scale"
    b = np.asarray(prob["bounds"])
    v = s.pack_params()
    assert np.allclose(b[:, 0], 0.95 * v) and np.allclose(b[:, 1], 1.05 * v)
    params["A_i"] = params["A_i"].copy()
    params["A_i"][0] = 0.0
    assert compute_bounds(params)["bounds"][s.K] == [0.0, 0.01]          # sensitivity.py:65-66
    mt = metric_time_indices(T15, [0.0, 1.0, 960.0], [4.0, 15.0], [0.0, 0.5])
    assert list(mt["t_prot"]) == [0, 3, 14] and list(mt["t_rna"]) == [5, 7] and list(mt["t_pho"]) == [0, 1]
    assert (mt["prot_b"], mt["rna_b"], mt["pho_b"]) == (0, 5, 0)


def test_loss_tables_are_consistent():
    s = synthetic_system(seed=6, N=8, K=3, max_sites=3)
    ld = synthetic_loss_data(s, T15, seed=2)
    assert ld["prot_base_idx"] == 0 and ld["rna_base_idx"] == 5 and ld["pho_base_idx"] == 0
    assert (T15[ld["t_rna"]] >= 4.0).all()
    assert (ld["s_pho"] < s.idx.n_sites[ld["p_pho"]]).all()
    assert np.array_equal(ld["prot_map"][:, 0], s.idx.offset_y) and np.array_equal(ld["prot_map"][:, 1], s.idx.n_sites)


def test_no_gpu_global_calls_fail_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phoskintime_b200.global_model import simulate_odeint
    s = synthetic_system(seed=1, N=6, K=3, max_sites=2)
    with pytest.raises(pk.PhoskinError):
        simulate_odeint(s, T15, 1e-6, 1e-9, 1000)


def test_fixed_point_schur_bound_and_sweep_count():
    """The algorithm behind csrc/global_net.cuh::schur_neumann, restated in numpy: for K = diag(f) G_QQ with
    ||K||_inf < 0.5 the sweeps z <- z0 + K z started from z0 reach the solution of (I - K) z = z0 to 1e-12 relative within the
    sweep count the kernel derives from the norm bound ||K||^k / (1 - ||K||)."""
    s = synthetic_system(seed=5, N=120, K=40, max_sites=4)
    N = s.idx.N
    G = np.zeros((N, N))
    for i in range(N):
        for q in range(s.TF_indptr[i], s.TF_indptr[i + 1]):
            G[i, s.TF_indices[q]] += s.TF_data[q]
    used = (np.abs(G).sum(axis=0) > 0) & (np.asarray(s.driver_map) < 0)        # regulator set Q: non-driven regulators
    Q = np.flatnonzero(used)
    Gq = G[np.ix_(Q, Q)]
    rabs = np.abs(Gq).sum(axis=1)
    rng = np.random.default_rng(3)
    for target in (0.05, 0.2, 0.49):
        f = rng.uniform(-1.0, 1.0, Q.size)
        f *= target / np.max(np.abs(f) * rabs)                                   # scale the row factors to the wanted norm
        Kmat = f[:, None] * Gq
        kn = float(np.max(np.abs(f) * rabs)) * 1.0001
        assert abs(np.abs(Kmat).sum(axis=1).max() * 1.0001 - kn) < 1e-12
        k_it = int(np.ceil(np.log(1e-12 * (1.0 - kn)) / np.log(kn)))
        k_it = max(2, (k_it + 1) & ~1)
        z0 = rng.standard_normal(Q.size)
        z = z0.copy()
        for _ in range(k_it):
            z = z0 + Kmat @ z
        exact = np.linalg.solve(np.eye(Q.size) - Kmat, z0)
        assert np.max(np.abs(z - exact)) <= 1e-12 * np.max(np.abs(z0)) * 1.01, (target, k_it)
        assert k_it <= 42
