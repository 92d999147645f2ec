"""Batched loss evaluation for multistart estimation — the data-parallel core of the reference's
`paramest/normest.py`.

The reference runs its 48 starts sequentially (normest.py:274-316); each start is a SciPy
`curve_fit` whose every residual evaluation is one `solve_ode` (normest.py:403-423).  The part
of that loop that is data-parallel — "given many parameter vectors for many proteins, return the
regularised weighted residual cost and the score_fit of each" — is one kernel launch here.  The
optimiser policy (TRF) stays with the caller (SURVEY.md §8(f) row 1 is the follow-up).
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


def best_per_group(values, group, n_groups):
    """argmin of `values` within each group (host reduction over a few thousand numbers)."""
    values = np.asarray(values)
    group = np.zeros(values.shape[0], dtype=np.int64) if group is None else np.asarray(group)
    best = np.full(n_groups, -1, dtype=np.int64)
    order = np.lexsort((values, group))
    first = np.r_[True, group[order][1:] != group[order][:-1]]
    best[group[order][first]] = order[first]
    return best
