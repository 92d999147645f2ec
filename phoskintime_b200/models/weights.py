"""Residual weights (sigma vectors) of the estimation path — host mirror of the reference's `models/weights.py`.

`curve_fit(..., sigma=sigma)` divides every residual by sigma; the reference builds up to 17 alternative sigma vectors
per protein (`get_weight_options`, models/weights.py:166-240), always in the layout of `full_weight` (:148-163): 9 ones
for the RNA block, the option's values for the protein + site block, and `reg_len` ones for the regularisation rows.
These are a few hundred numbers per protein, computed once on the host; the fits that consume them run on the device
(`paramest.find_best_lambda`, `paramest.fit_multistart`).  `get_protein_weights` (:80-146) reads the measurement
uncertainties from the reference's CSV files — file I/O, outside this path: the caller passes `ms_gauss_weights`.
"""
import numpy as np

USE_CUSTOM_WEIGHTS = False          # config/constants.py: only 'uncertainties_from_data' is used unless switched on


def early_emphasis(pr_data, p_data, time_points, num_psites):
    """models/weights.py:10-76: 1/(|value| + 1e-5), times 1/(dt + 1e-5) for the first eight time points (dt = spacing to
    the previous point; the first point keeps time weight 1).  Returns [n_times + num_psites*n_times]."""
    p_data = np.atleast_2d(np.asarray(p_data, dtype=np.float64))
    pr_data = np.atleast_2d(np.asarray(pr_data, dtype=np.float64))
    t = np.asarray(time_points, dtype=np.float64)
    n_times = t.shape[0]
    time_w = np.ones(n_times)
    time_w[1:] = 1.0 / (np.diff(t) + 1e-5)
    time_w[8:] = 1.0
    w_pr = time_w / (np.abs(pr_data[0, :n_times]) + 1e-5)
    w_p = time_w[None, :] / (np.abs(p_data[:num_psites, :n_times]) + 1e-5)
    return np.concatenate([w_pr, w_p.reshape(-1)])


def full_weight(p_data_weight, use_regularization, reg_len):
    """models/weights.py:148-163."""
    base = np.concatenate([np.ones(9), np.asarray(p_data_weight, dtype=np.float64)])
    if use_regularization:
        base = np.concatenate([base, np.ones(reg_len)])
    return base


def _uniform_filter3(x):
    """uniform_filter1d(x, 3) of SciPy's ndimage with its default 'reflect' boundary (the edge sample repeats)."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return x.copy()
    pad = np.concatenate([x[:1], x, x[-1:]])
    return (pad[:-2] + pad[1:-1] + pad[2:]) / 3.0


def get_weight_options(target, t_target, num_psites, use_regularization, reg_len, early_weights, ms_gauss_weights,
                       use_custom_weights=None):
    """models/weights.py:166-240: dict name -> sigma vector (insertion order of the reference)."""
    target = np.asarray(target, dtype=np.float64)
    time_indices = np.tile(np.arange(1, len(t_target) + 1), num_psites)
    log_scale = np.log1p(np.abs(target))
    sqrt_signal = np.sqrt(np.maximum(np.abs(target), 1e-5))
    if len(target) >= 2:
        flat_region_penalty = 1 / np.maximum(np.abs(np.gradient(target)), 1e-5)
    else:
        flat_region_penalty = 1 / np.maximum(np.abs(target), 1e-5)
    fw = lambda v: full_weight(v, use_regularization, reg_len)
    tail = target[9:]
    n_ti = len(time_indices)
    opts = {
        "inverse": fw(1 / np.maximum(np.abs(tail), 1e-5)),
        "exponential_decay": fw(np.exp(-0.5 * tail)),
        "inverse_log_scale": fw(1 / np.maximum(log_scale[9:], 1e-5)),
        "inverse_time_diff": fw(1 / np.maximum(np.abs(np.diff(tail, prepend=tail[0])), 1e-5)),
        "inverse_moving_avg": fw(1 / np.maximum(np.abs(tail - _uniform_filter3(tail)), 1e-5)),
        "sigmoid_decay": fw(1 / (1 + np.exp((time_indices - 5)))),
        "exponential_early_decay": fw(np.exp(-0.5 * time_indices)),
        "polynomial_time_decay": fw(1 / (1 + 0.5 * time_indices)),
        "signal_noise": fw(1 / sqrt_signal[9:]),
        "inverse_variance": fw(1 / (np.maximum(np.abs(tail), 1e-5) ** 0.7)),
        "flat_penalty": fw(flat_region_penalty[9:]) if flat_region_penalty.shape[0] == target.shape[0] else flat_region_penalty,
        "steady_decay": fw(np.exp(-0.1 * time_indices)),
        "inverse_square_root_data": fw(1 / sqrt_signal[9:]),
        "early_moderate_decay": fw(np.linspace(1.0, 0.3, n_ti)),
        "early_steep_decay": fw(np.concatenate([np.full(min(8, n_ti), 0.05), np.full(min(2, max(n_ti - 8, 0)), 0.2),
                                                np.ones(max(n_ti - 10, 0))])),
        "early_emphasis": fw(early_weights),
        "uncertainties_from_data": fw(ms_gauss_weights),
    }
    custom = USE_CUSTOM_WEIGHTS if use_custom_weights is None else use_custom_weights
    if not custom:
        opts = {"uncertainties_from_data": opts["uncertainties_from_data"]}
    return opts
