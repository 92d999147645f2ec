"""Global Morris sensitivity on the batched engine — mirrors the reference's
`global_model/sensitivity.py:33-297`.

Reference flow: ±5 % box around every fitted parameter (`compute_bounds` :33-79), SALib Morris sample,
one `simulate_and_measure` + `_compute_scalar_metric` per row in a process pool (:81-140, :219-246),
SALib `analyze` (:266).  Here the N·(D+1) rows are ONE launch of `global_net_kernel` with the scalar
metric fused into its epilogue (only 8 bytes per system leave HBM) and the elementary-effect statistics
are `pk_morris_ee`.  SALib is not available in this image: sampling follows the published method
(see phoskintime_b200/sensitivity.py and oracle/morris.py); `local_optimization=True` trajectory
selection of SALib is not reproduced (it only picks a spread-out subset of the sampled trajectories).
"""
import numpy as np

from ..engine import get_engine
from ..sensitivity import morris_sample
from .network import PARAM_KEYS
from .simulate import metric_time_indices, simulate_batch

SENSITIVITY_PERTURBATION = 0.05       # config.toml:351
TIME_POINTS_PROTEIN = [0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0]
TIME_POINTS_RNA = [4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0]
TIME_POINTS_PHOSPHO = TIME_POINTS_PROTEIN


def compute_bounds(params, perturbation=SENSITIVITY_PERTURBATION):
    """SALib problem dict over every scalar of the parameter dict (sensitivity.py:33-79): arrays are
    flattened to `key_i`, near-zero values get [0, 0.01]."""
    names, bounds = [], []
    for key in PARAM_KEYS + ("tf_scale",):
        value = params[key]
        vals = np.atleast_1d(np.asarray(value, dtype=np.float64))
        scalar = np.ndim(value) == 0
        for i, v in enumerate(vals):
            lb, ub = v * (1 - perturbation), v * (1 + perturbation)
            if abs(v) < 1e-6:
                lb, ub = 0.0, 0.01
            bounds.append([max(0.0, lb), ub])
            names.append(key if scalar else f"{key}_{i}")
    return {"num_vars": len(names), "names": names, "bounds": bounds}


def run_sensitivity_analysis(sys, fitted_params=None, metric="total_signal", *, N=10, num_levels=4, seed=None, X=None,
                             rtol=1e-5, atol=1e-7, mxstep=5000, engine=None, sharded=None,
                             t_points=(TIME_POINTS_PROTEIN, TIME_POINTS_RNA, TIME_POINTS_PHOSPHO)):
    """Morris indices of the scalar `metric` over all parameters of `sys`.

    Returns dict(names, mu, mu_star, sigma, Y, X, status).  `rtol/atol/mxstep` default to the values
    the reference hard-codes in `simulate_and_measure` (simulate.py:109).  With `sharded`
    (a `phoskintime_b200.parallel.ShardedRun`) every rank integrates a block of whole trajectories and
    Y is all-gathered."""
    eng = engine or get_engine()
    params = fitted_params or {**{k: getattr(sys, k) for k in PARAM_KEYS}, "tf_scale": sys.tf_scale}
    problem = compute_bounds(params)
    D = problem["num_vars"]
    if X is None:
        X = morris_sample(problem, N, num_levels=num_levels, seed=seed)
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.shape[1] != D or X.shape[0] % (D + 1):
        raise ValueError("X must hold N*(D+1) rows of D parameters")
    times = np.unique(np.concatenate([np.asarray(t, dtype=np.float64) for t in t_points]))
    mt = metric_time_indices(times, *t_points)
    lo, hi = (0, X.shape[0]) if sharded is None else sharded.bounds(X.shape[0], align=D + 1)
    res = simulate_batch(sys, X[lo:hi], times, ("metric",), rtol=rtol, atol=atol, mxstep=mxstep, metric=metric,
                         metric_times=mt, engine=eng)
    Y = np.where(np.asarray(res["status"]) == 0, np.asarray(res["metric"]), 0.0)      # failed run -> 0.0 (sensitivity.py:102-104)
    status = np.asarray(res["status"])
    if sharded is not None and sharded.world > 1:
        Y = sharded.allgather(Y, X.shape[0], align=D + 1)
        status = sharded.allgather(status.astype(np.float64), X.shape[0], align=D + 1).astype(np.int32)
    Si = eng.morris_ee(X, Y, num_levels=num_levels, scaled=False)
    order = np.argsort(-Si["mu_star"], kind="stable")
    return {"names": problem["names"], "mu": Si["mu"], "mu_star": Si["mu_star"], "sigma": Si["sigma"], "order": order,
            "Y": Y, "X": X, "status": status}
