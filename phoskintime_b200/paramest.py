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

from .engine import get_engine, local_dims
from .models.weights import early_emphasis, get_weight_options


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


def find_best_lambda(gene, target, p0, time_points, free_bounds, init_cond, num_psites, p_data, pr_data,
                     lambdas=np.logspace(-2, 0, 10), *, model=None, ms_gauss_weights=None, use_custom_weights=None,
                     engine=None, return_table=False, **nlls_kw):
    """`find_best_lambda` / `worker_find_lambda` (paramest/normest.py:22-165): for every lambda of the scan and every
    sigma option of `get_weight_options` one bounded fit from `p0`, ranked by `score_fit` at the optimum; returns
    (best_lambda, best_weight_key).

    The reference farms the 10 lambdas to a process pool and loops over the weight options inside each worker —
    10 x W sequential `curve_fit`s.  Here every (lambda, weight option) pair is one GROUP of a single
    `pk_local_nlls_batch` call: its own sigma row [L+P] and its own lambda (`lam_group`), all fits advancing together
    on the device.  `target` may be [L] (one protein) or [G,L] with `p_data` / `pr_data` / `ms_gauss_weights` lists of
    the same length: then the proteins are batched too and lists are returned.
    `ms_gauss_weights`: the measurement uncertainties the reference reads from its CSV files (`get_protein_weights`,
    models/weights.py:80-146); None -> ones."""
    from .models import ODE_MODEL
    model = model or ODE_MODEL
    eng = engine or get_engine()
    t = np.asarray(time_points, dtype=np.float64)
    targets = np.atleast_2d(np.asarray(target, dtype=np.float64))
    single = np.ndim(target) == 1
    G = targets.shape[0]
    listify = lambda v: [v] if single else list(v)
    p_list, pr_list = listify(p_data), listify(pr_data)
    ms_list = [None] * G if ms_gauss_weights is None else listify(ms_gauss_weights)
    lb, ub = (np.asarray(b, dtype=np.float64) for b in free_bounds)
    p0 = np.broadcast_to(np.asarray(p0, dtype=np.float64), (G, lb.size))
    lambdas = np.asarray(lambdas, dtype=np.float64)
    L = targets.shape[1]
    sig_rows, tg_rows, lam_rows, starts, keys = [], [], [], [], []
    for g in range(G):
        early = early_emphasis(pr_list[g], p_list[g], t, num_psites)
        ms = np.ones(L - 9) if ms_list[g] is None else np.asarray(ms_list[g], dtype=np.float64)
        opts = get_weight_options(targets[g], t, num_psites, True, lb.size, early, ms, use_custom_weights)
        for lam in lambdas:
            for key, sigma in opts.items():
                if np.size(sigma) != L + lb.size:
                    # the time-index based options of the reference are built over num_psites*n_times entries
                    # (models/weights.py:186), one block short of the protein + site layout: curve_fit rejects such a
                    # sigma ("sigma has incorrect shape") and the reference's worker dies on it — they are skipped
                    continue
                sig_rows.append(sigma); tg_rows.append(targets[g]); lam_rows.append(lam); starts.append(p0[g])
                keys.append((g, float(lam), key))
    n = len(keys)
    res = eng.nlls_local_batch(model, np.asarray(starts), init_cond, num_psites, t, np.asarray(tg_rows), lb, ub,
                               sigma=np.asarray(sig_rows), group=np.arange(n, dtype=np.int32), lam=np.asarray(lam_rows),
                               log_params=(model == "randmod"), **nlls_kw)
    score = np.where((res["status"] > 0) & np.isfinite(res["score"]), res["score"], np.inf)
    out = []
    for g in range(G):
        idx = [i for i, k in enumerate(keys) if k[0] == g]
        # the reference keeps the first strictly smaller score, walking lambdas in completion order (unordered); ties
        # are broken here by the scan order (lambda ascending, options in dict order)
        best = idx[int(np.argmin(score[idx]))]
        out.append((keys[best][1], keys[best][2]) if np.isfinite(score[best]) else (None, None))
    table = {"keys": keys, "score": score, "theta": res["theta"], "status": res["status"], "cost": res["cost"]}
    result = out[0] if single else out
    return (result, table) if return_table else result


def bootstrap_refit(model, popt_best, lb, ub, init_cond, num_psites, time_points, target_fit, *, n_boot, sigma=None, lam=0.0,
                    noise_sd=0.05, rng=None, engine=None, **nlls_kw):
    """The bootstrap loop of `normest` (paramest/normest.py:488-523): `n_boot` refits from `popt_best` against
    `target_fit * (1 + N(0, noise_sd))`, all in ONE `pk_local_nlls_batch` call (every bootstrap replicate is a group
    with its own noisy target row).  `target_fit` = [target | zeros(P)] as the reference builds it (the regularisation
    targets stay zero under multiplicative noise).  Returns dict(popt_mean[P], estimates[n_boot,P], ok[n_boot])."""
    eng = engine or get_engine()
    rng = rng or np.random.default_rng()
    lb = np.asarray(lb, dtype=np.float64)
    P = lb.size
    target_fit = np.asarray(target_fit, dtype=np.float64)
    L = local_dims(model, num_psites, len(time_points))[2]          # target_fit = [target (L) | zeros (P)] or just target
    if target_fit.size not in (L, L + P):
        raise ValueError(f"target_fit must have {L} or {L + P} entries")
    noisy = target_fit[None, :] * (1.0 + rng.normal(0.0, noise_sd, size=(n_boot, target_fit.size)))
    sg = None if sigma is None else np.broadcast_to(np.asarray(sigma, dtype=np.float64), (n_boot, np.size(sigma)))
    res = eng.nlls_local_batch(model, np.tile(np.asarray(popt_best, dtype=np.float64), (n_boot, 1)), init_cond, num_psites,
                               time_points, noisy[:, :L], lb, ub, sigma=sg, group=np.arange(n_boot, dtype=np.int32),
                               lam=lam, log_params=(model == "randmod"), **nlls_kw)
    ok = res["status"] > 0
    est = np.where(ok[:, None], res["theta"], np.asarray(popt_best, dtype=np.float64)[None, :])   # a failed refit keeps popt_best
    return {"popt_mean": est.mean(axis=0), "estimates": est, "ok": ok, "score": res["score"], "cost": res["cost"]}
