"""TEST INFRASTRUCTURE ONLY — restatement of the Morris screening the reference calls.

The reference calls SALib 1.5.1 (poetry.lock:1760-1761), which is NOT installed in this
image and not vendored in /root/reference, and no reference test pins its output:
**parity unpinned at this boundary**.  Call sites: sensitivity/analysis.py:223 (`sample`,
N=1000, num_levels=400, unseeded) and :264-265 (`analyze(..., num_levels=400,
conf_level=0.99, scaled=True)`).  What follows is the published algorithm (Morris 1991;
Campolongo et al. 2007; sigma-scaling of Sin & Gernaey 2009):

  sample : N trajectories of D+1 points on a p-level grid in [0,1]^D, step
           delta = p / (2(p-1)), one coordinate changed per move in random order and
           random direction, then scaled to the parameter bounds.
  analyze: EE_i = (Y(x_i up) - Y(x_i down)) / delta;  mu = mean, mu* = mean|EE|,
           sigma = std(ddof=1).  scaled=True multiplies dY/dx_i by std(x_i)/std(Y).

Ranking parity is evaluated by feeding the SAME X to both the oracle and the GPU path.
"""
import numpy as np


def compute_bound(value, perturbation=0.5):
    """sensitivity/analysis.py:20-35."""
    if abs(value) < 1e-6:
        return [0.0, 0.1]
    return [max(0.0, value * (1 - perturbation)), value * (1 + perturbation)]


def delta_of(num_levels):
    return num_levels / (2.0 * (num_levels - 1))


def sample(bounds, N, num_levels=4, seed=None):
    """-> X[N*(D+1), D] in parameter units."""
    rng = np.random.default_rng(seed)
    bounds = np.asarray(bounds, float)
    D = bounds.shape[0]
    delta = delta_of(num_levels)
    grid = np.linspace(0.0, 1.0 - delta, num_levels // 2)
    lower = np.tril(np.ones((D + 1, D)), -1)          # row k has k leading ones
    out = np.empty((N, D + 1, D))
    for r in range(N):
        base = rng.choice(grid, D)
        order = rng.permutation(D)
        direction = rng.choice([-1.0, 1.0], D)
        moved = lower[:, np.argsort(order)]           # coordinate order[k] moves at step k+1
        # direction +1: base -> base+delta ; -1: start at base+delta and come down
        out[r] = base + delta * np.where(direction > 0, moved, 1.0 - moved)
    X01 = out.reshape(-1, D)
    return bounds[:, 0] + X01 * (bounds[:, 1] - bounds[:, 0])


def elementary_effects(X, Y, D, num_levels, bounds=None, scaled=False):
    """-> EE[N, D].  X in parameter units; the unit-cube step is recovered from bounds
    (unscaled) or the raw dx is used with std(x_i)/std(Y) (scaled)."""
    X = np.asarray(X, float).reshape(-1, D + 1, D)
    Yt = np.asarray(Y, float).reshape(-1, D + 1)
    N = X.shape[0]
    dX = X[:, 1:, :] - X[:, :-1, :]                   # [N, D, D] one nonzero per row
    dY = Yt[:, 1:] - Yt[:, :-1]                       # [N, D]
    which = np.abs(dX).argmax(axis=2)                 # coordinate moved at each step
    step = np.take_along_axis(dX, which[:, :, None], axis=2)[:, :, 0]
    ee = np.empty((N, D))
    rows = np.arange(N)[:, None]
    if scaled:
        sx = X.reshape(-1, D).std(axis=0)
        sy = Yt.std()
        ee[rows, which] = dY / step * (sx[which] / sy)
    else:
        ee[rows, which] = np.sign(step) * dY / delta_of(num_levels)
    return ee


def analyze(X, Y, D, num_levels=4, scaled=False):
    ee = elementary_effects(X, Y, D, num_levels, scaled=scaled)
    return {"mu": ee.mean(axis=0), "mu_star": np.abs(ee).mean(axis=0),
            "sigma": ee.std(axis=0, ddof=1), "ee": ee}
