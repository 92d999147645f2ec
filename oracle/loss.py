"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's scalar epilogues.

  * score_fit            /root/reference/config/config.py:176-226
  * curve_fit residual   /root/reference/paramest/normest.py:403-423 (model_func) and the
                         sigma weighting SciPy applies: r = (f(x) - y) / sigma, cost = sum r^2
  * Morris scalar Y      /root/reference/sensitivity/analysis.py:89-176 (_compute_Y)
Pinned against the unmodified reference by tests/golden/ (see oracle/gen_golden.py).
"""
import numpy as np

Y_METRICS = ("total_signal", "mean_activity", "variance", "dynamics", "l2_norm")


def score_fit(params, target, prediction, alpha=1.0, beta=1.0, gamma=1.0, delta=1.0, mu=1.0):
    params = np.asarray(params, float)
    r = np.abs(np.asarray(target, float) - np.asarray(prediction, float)) / np.size(target)
    mse = np.sum(r ** 2)
    rmse = np.sqrt(np.mean(r ** 2))
    mae = np.mean(r)
    var = np.var(r)
    l2 = np.linalg.norm(params, ord=2) / len(params)
    return delta * mse + alpha * rmse + beta * mae + gamma * var + mu * l2


def weighted_ssr(params, flat, target, sigma=None, lam=0.0):
    """Sum of squared curve_fit residuals of normest's regularised model_func:
    [flat | lam/P * params^2] against [target | 0], each divided by sigma."""
    params = np.asarray(params, float)
    P = len(params)
    model = np.concatenate([flat, lam / P * params ** 2])
    tgt = np.concatenate([target, np.zeros(P)])
    sig = np.ones_like(model) if sigma is None else np.asarray(sigma, float)
    if sig.shape[0] == flat.shape[0]:
        sig = np.concatenate([sig, np.ones(P)])
    return float(np.sum(((model - tgt) / sig) ** 2))


def compute_Y(sol, num_psites, metric="total_signal"):
    """Scalar Morris output over columns 0 (mRNA), 1 (protein) and 2..2+ns-1."""
    cols = np.asarray(sol, float)[:, :2 + num_psites]
    n_t = cols.shape[0]
    length = 2 * n_t + n_t * num_psites
    total = cols.sum()
    if metric == "total_signal":
        return total
    if metric == "mean_activity":
        return total / length
    if metric == "variance":
        mean = total / length
        return ((cols - mean) ** 2).sum() / length
    if metric == "dynamics":
        return (np.diff(cols, axis=0) ** 2).sum()
    if metric == "l2_norm":
        return np.sqrt((cols ** 2).sum())
    raise ValueError("Unknown Y_METRIC")
