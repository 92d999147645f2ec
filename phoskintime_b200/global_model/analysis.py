"""Long-horizon steady-state check of the global network — the compute part of the reference's
`global_model/analysis.py:29-67` (`simulate_until_steady`); its plotting/report half (:70-…) is out of scope
(SURVEY.md §2).  The 1000-point log grid is just another `t_eval`: the kernel lands on every requested time, so
no interpolation error enters the convergence diagnostic."""
import numpy as np

from .simulate import simulate_batch, simulate_odeint


def steady_time_grid(t_max=1440.0, n_points=1000):
    """analysis.py:47-50: t = 0 followed by n_points-1 log-spaced times from 1e-3 to t_max."""
    return np.concatenate(([0.0], np.logspace(np.log10(1e-3), np.log10(t_max), n_points - 1)))


def simulate_until_steady(sys, t_max=1440.0, n_points=1000):
    """Reference signature (analysis.py:29): current parameters of `sys` -> (t_eval, Y[n_points, state_dim]) at
    rtol=1e-6, atol=1e-8, mxstep=50000 (analysis.py:56)."""
    t_eval = steady_time_grid(t_max, n_points)
    Y = simulate_odeint(sys, t_eval, rtol=1e-6, atol=1e-8, mxstep=50000)
    return t_eval, Y


def final_rate_of_change(t_eval, Y):
    """analysis.py:58-62: ||Y[-1] - Y[-2]|| / (t[-1] - t[-2]); Y may carry leading batch axes."""
    return np.linalg.norm(Y[..., -1, :] - Y[..., -2, :], axis=-1) / (t_eval[-1] - t_eval[-2])


def steady_check_batch(sys, params, t_max=1440.0, n_points=1000, engine=None, **kw):
    """The same check for B parameter vectors in one launch: (t_eval, Y[B,T,n], rate[B], status[B])."""
    t_eval = steady_time_grid(t_max, n_points)
    kw.setdefault("rtol", 1e-6)
    kw.setdefault("atol", 1e-8)
    r = simulate_batch(sys, params, t_eval, ("Y",), mxstep=50000, engine=engine, **kw)
    return t_eval, r["Y"], final_rate_of_change(t_eval, r["Y"]), r["status"]
