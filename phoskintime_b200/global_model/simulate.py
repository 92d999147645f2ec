"""Drop-in surface of the reference's `global_model/simulate.py`, `lossfn.py` dispatch and the batched
forms the GPU path adds.

`simulate_odeint(sys, t_eval, rtol, atol, mxstep)` keeps the reference's signature
(global_model/simulate.py:34-80) and return value (`Y[T, state_dim]`, C-contiguous float64); it is a
batch of one through `simulate_batch`.  All arithmetic runs in libphoskin_b200.so (kernel
`global_net_kernel`, csrc/global_net.cuh); there is no CPU path — without the library or a GPU every
call raises `PhoskinError`.

Differences a caller can observe (documented, not hidden):
  * the integrator is a Rosenbrock method with the analytic Jacobian that lands on the kinase-grid
    points where the RHS jumps; the reference's LSODA integrates through them with a finite-difference
    Jacobian.  Results agree with the reference's own tight-tolerance solution to <=1e-6 relative
    (tests/test_gpu_global.py); `mxstep` bounds the steps of the whole solve, not of one output interval;
  * failed systems come back as NaN rows with a non-zero status instead of SciPy warnings.
"""
import numpy as np

from ..engine import get_engine

METRIC_BASE_TIMES = (0.0, 4.0, 0.0)       # simulate.py:116-118: protein, RNA (t=4), phospho baselines


def _topology(sys_, engine):
    key = engine.token
    if key not in sys_._topo_id:
        sys_._topo_id[key] = engine.global_upload(sys_)
    return sys_._topo_id[key]


def _loss_fingerprint(loss_data):
    """Digest of the loss tables' contents (a few KB: cheap next to one solve)."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for k in sorted(loss_data):
        a = np.ascontiguousarray(loss_data[k])
        h.update(k.encode()); h.update(str(a.dtype).encode()); h.update(a.tobytes())
    return h.digest()


def simulate_batch(sys_, params, t_eval, want=("Y",), *, y0=None, rtol=None, atol=None, mxstep=0, theta_mode=False,
                   loss_data=None, loss_mode=0, metric="total_signal", metric_times=None, lambdas=(1.0, 1.0, 1.0),
                   lambda_prior=0.0, engine=None, out=None):
    """B parameter vectors of one network -> any of Y[B,T,n], loss[B,3], F[B,3], metric[B] (+status, nsteps, nrej).

    params: [B,P] in the order `GlobalSystem.pack_params` (c_k|A|B|C|D|Dp|E|tf_scale); numpy (host path, the
    library stages through HBM) or torch CUDA tensors (used in place, outputs are torch tensors).
    """
    eng = engine or get_engine()
    topo = _topology(sys_, eng)
    if loss_data is not None:
        fp = _loss_fingerprint(loss_data)               # content, not identity: in-place edits of the dict are seen
        if sys_._loss_key.get(eng.token) != fp:
            eng.global_set_loss_data(topo, loss_data)
            sys_._loss_key[eng.token] = fp
    if "F" in want:
        eng.global_set_prior(topo, sys_.pack_params(sys_.defaults) if lambda_prior else None)
    if y0 is None:
        y0 = sys_.y0()
    return eng.global_solve_batch(topo, params, y0, t_eval, want, rtol=rtol, atol=atol, max_steps=mxstep,
                                  theta_mode=theta_mode, loss_mode=loss_mode, metric=metric, metric_times=metric_times,
                                  lambdas=lambdas, lambda_prior=lambda_prior, out=out)


def simulate_odeint(sys, t_eval, rtol, atol, mxstep):
    """Reference signature (global_model/simulate.py:34): current parameters of `sys` -> Y[T, state_dim]."""
    t_eval = np.asarray(t_eval, dtype=np.float64)
    res = simulate_batch(sys, sys.pack_params()[None, :], t_eval, ("Y",), rtol=rtol, atol=atol, mxstep=int(mxstep))
    return np.ascontiguousarray(res["Y"][0], dtype=np.float64)


def solve_custom(sys, y0, t_eval, rtol, atol, *, method="dopri5", engine=None):
    """Reference signature of the `USE_CUSTOM_SOLVER` branch (global_model/jacspeedup.py:31-67, used at
    simulate.py:55-58): explicit y0, current parameters of `sys` -> Y[T, state_dim].

    method="dopri5" (default) reproduces the reference function: its Numba DOPRI5(4) with PI control, dt <= 1,
    bucket landing and cubic-Hermite output (solvers.py:292-758), step for step on the device
    (`pk_global_solve_custom`) — including that solver's own output error (third-order interpolant: ~1e-5 relative
    at its default tolerances).  method="rosenbrock" integrates the same problem with the implicit kernel of
    `simulate_odeint` (lands on every output time; ~1e-7 of the tight solution) for callers that want the branch's
    signature but not its error."""
    if method == "rosenbrock":
        res = simulate_batch(sys, sys.pack_params()[None, :], np.asarray(t_eval, dtype=np.float64), ("Y",),
                             y0=np.asarray(y0, dtype=np.float64), rtol=rtol, atol=atol, engine=engine)
        return np.ascontiguousarray(res["Y"][0], dtype=np.float64)
    if method != "dopri5":
        raise ValueError("method must be 'dopri5' or 'rosenbrock'")
    return solve_custom_batch(sys, sys.pack_params()[None, :], y0, t_eval, rtol, atol, engine=engine)["Y"][0]


def solve_custom_batch(sys_, params, y0, t_eval, rtol=None, atol=None, *, max_steps=0, theta_mode=False, engine=None):
    """`solve_custom` for B parameter vectors in one launch: dict(Y[B,T,n], status, nsteps, nrej)."""
    eng = engine or get_engine()
    return eng.global_solve_custom(_topology(sys_, eng), params, np.asarray(y0, dtype=np.float64) if not hasattr(y0, "device") else y0,
                                   t_eval, rtol=rtol, atol=atol, max_steps=max_steps, theta_mode=theta_mode)


def metric_time_indices(times, t_points_p, t_points_r, t_points_pho):
    """Which rows of the union grid `times` simulate_and_measure keeps per modality (simulate.py:107-124,
    :189-200) and the three baseline rows."""
    times = np.asarray(times, float)
    pick = lambda tp: np.flatnonzero(np.isin(times, np.asarray(tp, float))).astype(np.int32)
    bidx = lambda t0: int(np.argmin(np.abs(times - t0)))
    return {"t_prot": pick(t_points_p), "t_rna": pick(t_points_r), "t_pho": pick(t_points_pho),
            "prot_b": bidx(METRIC_BASE_TIMES[0]), "rna_b": bidx(METRIC_BASE_TIMES[1]), "pho_b": bidx(METRIC_BASE_TIMES[2])}


def fold_change_tables(sys_, params, t_points_p, t_points_r, t_points_pho, *, rtol=1e-5, atol=1e-7, mxstep=5000, engine=None):
    """The three fold-change tables of `simulate_and_measure` (simulate.py:105-182) for B parameter vectors, computed
    in the solve kernel's epilogue: returns dict(times, fc_prot[B,N,len(tp)], fc_rna[B,N,len(tr)], fc_pho[B,total_sites,
    len(tph)], status) — rows in the reference's order (protein-major, sites in block order, times ascending), floors
    1e-12, baselines t=0 / t=4 / t=0, tolerances of simulate.py:109."""
    times = np.unique(np.concatenate([t_points_p, t_points_r, t_points_pho]).astype(np.float64))
    mt = metric_time_indices(times, t_points_p, t_points_r, t_points_pho)
    params = np.atleast_2d(params) if isinstance(params, np.ndarray) else params
    r = simulate_batch(sys_, params, times, ("fc",), rtol=rtol, atol=atol, mxstep=mxstep, metric_times=mt, engine=engine)
    fc = r["fc"]
    B, N, S = fc.shape[0], sys_.idx.N, sys_.idx.total_sites
    np_, nr_, nph = mt["t_prot"].size, mt["t_rna"].size, mt["t_pho"].size
    return {"times": times, "t_prot": times[mt["t_prot"]], "t_rna": times[mt["t_rna"]], "t_pho": times[mt["t_pho"]],
            "fc_prot": fc[:, :N * np_].reshape(B, N, np_), "fc_rna": fc[:, N * np_:N * (np_ + nr_)].reshape(B, N, nr_),
            "fc_pho": fc[:, N * (np_ + nr_):].reshape(B, S, nph), "status": r["status"]}


def simulate_and_measure(sys, idx, t_points_p, t_points_r, t_points_pho):
    """Reference signature (global_model/simulate.py:83): current parameters of `sys` -> (df_prot, df_rna, df_phos)
    with columns [protein, time, pred_fc] / [protein, psite, time, pred_fc].  `idx` supplies the names
    (`idx.proteins`, `idx.sites`); pass None for integer labels.  Needs pandas, like the reference."""
    import pandas as pd
    tab = fold_change_tables(sys, sys.pack_params()[None, :], t_points_p, t_points_r, t_points_pho)
    N = sys.idx.N
    names = list(idx.proteins) if idx is not None and hasattr(idx, "proteins") else list(range(N))
    sites = (idx.sites if idx is not None and hasattr(idx, "sites") else
             [list(range(int(k))) for k in sys.idx.n_sites])
    mk = lambda fc, t: pd.DataFrame({"protein": np.repeat(np.asarray(names, dtype=object), t.size),
                                     "time": np.tile(t, N), "pred_fc": fc.reshape(-1)})
    df_p, df_r = mk(tab["fc_prot"][0], tab["t_prot"]), mk(tab["fc_rna"][0], tab["t_rna"])
    prot_col = [names[i] for i in range(N) for _ in sites[i]]
    site_col = [s for i in range(N) for s in sites[i]]
    t = tab["t_pho"]
    df_ph = pd.DataFrame({"protein": np.repeat(np.asarray(prot_col, dtype=object), t.size),
                          "psite": np.repeat(np.asarray(site_col, dtype=object), t.size),
                          "time": np.tile(t, len(site_col)), "pred_fc": tab["fc_pho"][0].reshape(-1)})
    return df_p, df_r, df_ph


def LOSS_FN(Y, p_prot, t_prot, obs_prot, w_prot, p_rna, t_rna, obs_rna, w_rna, p_pho, s_pho, t_pho, obs_pho, w_pho,
            prot_map, prot_base_idx, rna_base_idx, pho_base_idx, *, loss_mode=0, model=0, engine=None):
    """Reference signature of `LOSS_FN` (global_model/lossfn.py:113-121, :386): one trajectory
    Y[T, state_dim] -> (loss_p, loss_r, loss_ph).  The reference binds `loss_function_comb` for MODEL 2 and
    `loss_function_noncomb` otherwise at import time (lossfn.py:386); here `model` selects.
    `prot_map[i] = (offset_y, n_sites)` — `(offset_y, n_states)` for the combinatorial model (cache.py:138-145) —
    is all the topology the loss needs, so a loss-only topology is uploaded per distinct prot_map and cached."""
    eng = engine or get_engine()
    prot_map = np.ascontiguousarray(prot_map, dtype=np.int32)
    comb = int(model) == 2
    key = (eng.token, comb, prot_map.tobytes())
    topo = _LOSS_TOPOS.get(key)
    if topo is None:
        from .network import GlobalSystem
        if comb:
            n_states = prot_map[:, 1]
            n_sites = np.round(np.log2(np.maximum(n_states, 1))).astype(np.int32)
            if not np.array_equal(1 << n_sites, n_states):
                raise ValueError("combinatorial prot_map: second column must be 2**n_sites (cache.py:145)")
            block = 1 + n_states
        else:
            n_sites = prot_map[:, 1]
            block = 2 + n_sites
        N, S = n_sites.size, int(n_sites.sum())
        if not np.array_equal(prot_map[:, 0], np.concatenate([[0], np.cumsum(block)[:-1]])):
            raise ValueError("prot_map offsets must be the packed per-protein block layout (network.py:28-167)")
        z = np.zeros(N)
        shell = GlobalSystem(n_sites=n_sites, W_indptr=np.zeros(S + 1, np.int32), W_indices=[], W_data=[],
                             TF_indptr=np.zeros(N + 1, np.int32), TF_indices=[], TF_data=[], kin_grid=[0.0],
                             kin_Kmat=np.ones((1, 1)), tf_deg=np.ones(N), driver_map=np.full(N, -1),
                             defaults={"c_k": [1.0], "A_i": z, "B_i": z, "C_i": z, "D_i": z, "Dp_i": np.zeros(S), "E_i": z,
                                       "tf_scale": 0.0}, model=2 if comb else 0)
        topo = _LOSS_TOPOS[key] = eng.global_upload(shell)
    eng.global_set_loss_data(topo, dict(p_prot=p_prot, t_prot=t_prot, obs_prot=obs_prot, w_prot=w_prot, p_rna=p_rna,
                                        t_rna=t_rna, obs_rna=obs_rna, w_rna=w_rna, p_pho=p_pho, s_pho=s_pho, t_pho=t_pho,
                                        obs_pho=obs_pho, w_pho=w_pho, prot_base_idx=prot_base_idx,
                                        rna_base_idx=rna_base_idx, pho_base_idx=pho_base_idx))
    out = eng.global_loss_batch(topo, np.asarray(Y, dtype=np.float64), loss_mode)
    return float(out[0, 0]), float(out[0, 1]), float(out[0, 2])


_LOSS_TOPOS = {}
