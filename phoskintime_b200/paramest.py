"""Batched loss evaluation for multistart estimation — the data-parallel core of the reference's
`paramest/normest.py`.

The reference runs its 48 starts sequentially (normest.py:274-316); each start is a SciPy
`curve_fit` whose every residual evaluation is one `solve_ode` (normest.py:403-423).  The part
of that loop that is data-parallel — "given many parameter vectors for many proteins, return the
regularised weighted residual cost and the score_fit of each" — is one kernel launch here
(`evaluate_starts`).  `fit_multistart` runs the whole multistart fit on the device: every start of every
protein is one row of `pk_local_nlls_batch` (bounded Levenberg-Marquardt, forward-difference Jacobians
as B*(P+1) solves per launch), and the best start per protein is picked by `score_fit` exactly as
normest.py:293-306 does.  SciPy's TRF itself is optimiser policy outside the parity contract
(SURVEY.md §8(c)); tests compare the minima found with SciPy's on the same residual.
"""
import numpy as np

from .engine import get_engine


def multistart_points(base_p0, lb, ub, n_starts=48, jitter_frac=0.10, seed=42, gene=""):
    """Start points exactly as `_curve_fit_multistart` draws them (normest.py:224-264):
    base, n_starts//3 Gaussian jitters of 10 % of the span, then stratified uniform fills."""
    lb = np.asarray(lb, dtype=float)
    ub = np.asarray(ub, dtype=float)
    if not (np.all(np.isfinite(lb)) and np.all(np.isfinite(ub))):
        raise ValueError("free_bounds must be finite for multistart sampling.")
    gene_hash = sum(ord(c) for c in str(gene)) % 1000003
    rng = np.random.default_rng(int(seed + gene_hash))
    base = np.clip(np.asarray(base_p0, dtype=float), lb, ub)
    pts = [base]
    span = ub - lb
    span[span <= 0] = 1.0
    for _ in range(max(0, n_starts // 3)):
        pts.append(np.clip(base + (jitter_frac * span) * rng.normal(0.0, 1.0, size=base.shape[0]), lb, ub))
    remaining = max(0, n_starts - len(pts))
    if remaining > 0:
        U = np.empty((remaining, base.shape[0]))
        for j in range(base.shape[0]):
            u = (np.arange(remaining) + rng.random(remaining)) / float(remaining)
            rng.shuffle(u)
            U[:, j] = u
        pts.extend(lb + U * (ub - lb))
    return np.asarray(pts)


def evaluate_starts(model, starts, init_cond, num_psites, time_points, targets, *, group=None,
                    sigma=None, lam=0.0, engine=None, want=("ssr", "score"), **kw):
    """Cost of every start: starts[B,P] (log-parameters for randmod, as normest passes them),
    targets[G,L] (one row per protein), group[B] protein index of each start.
    Returns dict(ssr[B], score[B], status[B], ...)."""
    eng = engine or get_engine()
    return eng.solve_local_batch(model, starts, init_cond, num_psites, time_points, want=want,
                                 target=targets, sigma=sigma, group=group, lam=lam,
                                 log_params=(model == "randmod"), **kw)


def fit_multistart(model, base_p0, lb, ub, init_cond, num_psites, time_points, targets, *, genes=None, sigma=None,
                   lam=0.0, n_starts=48, jitter_frac=0.10, seed=42, engine=None, **nlls_kw):
    """`_curve_fit_multistart` (normest.py:167-326) for G proteins at once.

    targets[G,L] one row per protein; base_p0 [P] or [G,P]; genes: names used for the per-gene seed
    (normest.py:222-223).  Returns dict(popt[G,P], best_score[G], best_start[G], n_ok[G], n_fail[G]) plus the
    per-start arrays (`theta`, `score`, `cost`, `status`, `group`)."""
    eng = engine or get_engine()
    targets = np.atleast_2d(np.asarray(targets, dtype=float))
    G = targets.shape[0]
    lb = np.asarray(lb, dtype=float)
    ub = np.asarray(ub, dtype=float)
    base = np.broadcast_to(np.asarray(base_p0, dtype=float), (G, lb.size))
    genes = [""] * G if genes is None else list(genes)
    starts = np.concatenate([multistart_points(base[p], lb, ub, n_starts, jitter_frac, seed, genes[p]) for p in range(G)])
    per = starts.shape[0] // G
    group = np.repeat(np.arange(G, dtype=np.int32), per)
    res = eng.nlls_local_batch(model, starts, init_cond, num_psites, time_points, targets, lb, ub, sigma=sigma,
                               group=group, lam=lam, log_params=(model == "randmod"), **nlls_kw)
    ok = (res["status"] > 0) & np.isfinite(res["score"])            # a failed start is skipped (normest.py:312-316)
    score = np.where(ok, res["score"], np.inf)
    best = best_per_group(score, group, G)
    n_ok = np.bincount(group[ok], minlength=G)
    if (n_ok == 0).any():
        raise RuntimeError(f"multistart fit: all starts failed for proteins {np.flatnonzero(n_ok == 0).tolist()}")
    return {"popt": res["theta"][best], "best_score": score[best], "best_start": best - np.arange(G) * per,
            "n_ok": n_ok, "n_fail": per - n_ok, "theta": res["theta"], "score": res["score"], "cost": res["cost"],
            "status": res["status"], "iters": res["iters"], "nfev": res["nfev"], "group": group}


def best_per_group(values, group, n_groups):
    """argmin of `values` within each group (host reduction over a few thousand numbers)."""
    values = np.asarray(values)
    group = np.zeros(values.shape[0], dtype=np.int64) if group is None else np.asarray(group)
    best = np.full(n_groups, -1, dtype=np.int64)
    order = np.lexsort((values, group))
    first = np.r_[True, group[order][1:] != group[order][:-1]]
    best[group[order][first]] = order[first]
    return best
