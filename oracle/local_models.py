"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's local-model path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (phoskintime_b200/) never does.

Restates, in plain numpy/scipy (numba-jitted RHS when numba is importable, as the
reference does), what the reference computes for one `solve_ode` call:

  * distributive RHS      /root/reference/models/distmod.py:6-65
  * successive RHS        /root/reference/models/succmod.py:8-90
  * random (2^n) RHS      /root/reference/models/randmod.py:8-85 (tables), :121-247 (RHS)
  * the wrapper           /root/reference/models/distmod.py:93-134, succmod.py:114-152,
                          randmod.py:249-305: odeint (LSODA, default tolerances) -> clip at 0
                          -> flat = [R(t[5:]) | P(t) | site columns transposed]

The integrator arithmetic is the third-party SciPy ODEPACK/LSODA (reference pin
scipy 1.15.2 in poetry.lock; this image has scipy 1.18.x) — the very same dependency
the reference calls, so it is *called*, not restated.  Parity of this restatement
with the unmodified reference is pinned by tests/golden/*.npz, produced by
oracle/gen_golden.py from the real reference modules (tests/test_oracle.py).

Three oracles:
  O1 "stock"  : solve_ode(...)                      default LSODA tolerances
  O2 "tight"  : solve_ode(..., rtol=1e-12, atol=1e-12)  same RHS, tight tolerances
  O3 "exact"  : exact_linear(...)                   matrix exponential (models are linear)
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import odeint
from scipy.linalg import expm

try:  # the reference jits its RHS with numba; do the same when it is there
    from numba import njit
    HAVE_NUMBA = True
except Exception:  # pragma: no cover
    HAVE_NUMBA = False

    def njit(*a, **k):
        def deco(f):
            return f
        return deco if not (a and callable(a[0])) else a[0]

MODELS = ("distmod", "succmod", "randmod")
TIME_POINTS = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0,
                        120.0, 240.0, 480.0, 960.0])  # config/constants.py:56-62
RNA_OFFSET = 5  # distmod.py:125 / randmod.py:286 — sol[5:, 0]


def n_states(model: str, ns: int) -> int:
    return 2 + ns if model != "randmod" else 2 + (1 << ns) - 1


def n_params(model: str, ns: int) -> int:
    return 4 + 2 * ns if model != "randmod" else 4 + ns + (1 << ns) - 1


# ----------------------------------------------------------------------------- RHS
@njit(cache=False)
def _rhs_dist(y, t, p, ns):
    # distmod.py:39-65 ; p = [A,B,C,D,S_1..S_ns,Dr_1..Dr_ns] (distmod.py:85-90)
    out = np.empty_like(y)
    out[0] = p[0] - p[1] * y[0]
    sS = 0.0
    sY = 0.0
    for i in range(ns):
        sS += p[4 + i]
        sY += y[2 + i]
    out[1] = p[2] * y[0] - (p[3] + sS) * y[1] + sY
    for i in range(ns):
        out[2 + i] = p[4 + i] * y[1] - (1.0 + p[4 + ns + i]) * y[2 + i]
    return out


@njit(cache=False)
def _rhs_succ(y, t, p, ns):
    # succmod.py:33-90 (same operation order, so LSODA sees bit-identical derivatives)
    out = np.empty_like(y)
    out[0] = p[0] - p[1] * y[0]
    dP = p[2] * y[0] - p[3] * y[1]
    if ns > 0:
        dP -= p[4] * y[1]
        dP += y[2]
    out[1] = dP
    for i in range(ns):
        S_i = p[4 + i]
        Dr_i = p[4 + ns + i]
        if i == ns - 1:  # last (or only) site: no onward phosphorylation
            out[2 + i] = S_i * y[1 + i] - (1 + Dr_i) * y[2 + i]
        else:
            out[2 + i] = S_i * y[1 + i] - (1 + p[4 + i + 1] + Dr_i) * y[2 + i] + y[3 + i]
    return out


@njit(cache=False)
def _rhs_rand(y, t, p, ns):
    # randmod.py:152-247, written in per-state form.  State index s (bitmask 1..m) lives in
    # y[1+s]; Ddeg is indexed by bitmask-1 (randmod.py:231).  Forward rate into target `tgt`
    # is S[lsb(tgt)] (randmod.py:201), whatever bit was actually added.
    m = (1 << ns) - 1
    out = np.zeros_like(y)
    R = y[0]
    P = y[1]
    out[0] = p[0] - p[1] * R
    dP = p[2] * R - p[3] * P
    for k in range(ns):  # randmod.py:174-185
        r = p[4 + k] * P
        out[1 + (1 << k)] += r
        dP -= r
    for s in range(1, m + 1):
        x = y[1 + s]
        for j in range(ns):  # forward phosphorylation of every unset bit (randmod.py:193-210)
            bit = 1 << j
            if not (s & bit):
                tgt = s | bit
                lsb = 0
                while not (tgt >> lsb) & 1:
                    lsb += 1
                r = p[4 + lsb] * x
                out[1 + tgt] += r
                out[1 + s] -= r
        for j in range(ns):  # dephosphorylation of every set bit, unit rate (randmod.py:213-228)
            bit = 1 << j
            if s & bit:
                low = s & ~bit
                if low == 0:
                    dP += x
                else:
                    out[1 + low] += x
                out[1 + s] -= x
        out[1 + s] -= p[4 + ns + s - 1] * x
    out[1] = dP
    return out


_RHS = {"distmod": _rhs_dist, "succmod": _rhs_succ, "randmod": _rhs_rand}


def rhs(model: str, y, params, ns: int):
    return _RHS[model](np.asarray(y, float), 0.0, np.asarray(params, float), int(ns))


# ------------------------------------------------------------------------- wrapper
def flat_from_sol(model: str, sol: np.ndarray, ns: int) -> np.ndarray:
    """flat = [sol[5:,0] | sol[:,1] | site columns, site-major] (distmod.py:124-134).
    randmod takes columns 2..2+ns-1 = bitmask states 1..ns (randmod.py:297-302)."""
    sites = sol[:, 2:2 + ns] if model == "randmod" else sol[:, 2:]
    return np.concatenate((sol[RNA_OFFSET:, 0], sol[:, 1], sites.T.ravel()))


def solve_ode(model: str, params, init_cond, num_psites: int, t, rtol=None, atol=None,
              mxstep=0, normalize=False):
    """One reference-equivalent solve: (sol[T,n], flat[L]).  rtol/atol None = SciPy defaults
    (~1.49e-8), as the reference calls odeint without tolerances (distmod.py:112)."""
    p = np.ascontiguousarray(params, dtype=float)
    y0 = np.ascontiguousarray(init_cond, dtype=float)
    t = np.atleast_1d(np.asarray(t, dtype=float))
    kw = {}
    if rtol is not None:
        kw["rtol"] = rtol
    if atol is not None:
        kw["atol"] = atol
    if mxstep:
        kw["mxstep"] = mxstep
    sol = np.clip(np.asarray(odeint(_RHS[model], y0, t, args=(p, int(num_psites)), **kw)), 0, None)
    if normalize:  # distmod.py:115-122 (NORMALIZE_MODEL_OUTPUT, default False)
        sol = sol * (1.0 / y0)[None, :]
    return sol, flat_from_sol(model, sol, num_psites)


def solve_ode_tight(model, params, init_cond, num_psites, t):
    return solve_ode(model, params, init_cond, num_psites, t, rtol=1e-12, atol=1e-12, mxstep=100000)


# --------------------------------------------------------------------- exact (O3)
def linear_form(model: str, params, ns: int):
    """(M, b) with f(y) = M y + b, probed from the restated RHS (models are linear)."""
    n = n_states(model, ns)
    b = rhs(model, np.zeros(n), params, ns)
    M = np.empty((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        M[:, j] = rhs(model, e, params, ns) - b
    return M, b


def exact_linear(model: str, params, init_cond, ns: int, t, clip=True):
    M, b = linear_form(model, params, ns)
    n = M.shape[0]
    Z = np.zeros((n + 1, n + 1))
    Z[:n, :n] = M
    Z[:n, n] = b
    y0 = np.asarray(init_cond, float)
    out = np.empty((len(t), n))
    for k, tk in enumerate(t):
        E = expm(Z * float(tk))
        out[k] = E[:n, :n] @ y0 + E[:n, n]
    return np.clip(out, 0, None) if clip else out


# ------------------------------------------------------------------ steady state
def initial_condition(model: str, ns: int) -> list:
    """All-ones-parameter steady state, as the reference's SLSQP formulation converges to
    (steady/initdist.py:9-50, initsucc.py:9-55 — which solves the *distributive* equations —
    and initrand.py:10-77 with its subset-size state ordering)."""
    if model in ("distmod", "succmod"):
        P = 1.0 / (1.0 + ns / 2.0)
        return [1.0, P] + [P / 2.0] * ns
    from itertools import combinations
    subsets = [c for k in range(1, ns + 1) for c in combinations(range(ns), k)]
    index = {s: i for i, s in enumerate(subsets)}
    n = 2 + len(subsets)
    Amat = np.zeros((n, n))
    rhs_v = np.zeros(n)
    Amat[0, 0] = -1.0
    rhs_v[0] = -1.0
    Amat[1, 0] = 1.0
    Amat[1, 1] = -(1.0 + ns)
    for i, sub in enumerate(subsets):
        row = 2 + i
        if len(sub) == 1:
            Amat[1, row] += 1.0
            Amat[row, 1] += 1.0
        else:
            for site in sub:
                red = tuple(x for x in sub if x != site)
                Amat[row, 2 + index[red]] += 1.0
        Amat[row, row] -= (ns - len(sub)) + len(sub) + 1.0
        for site in range(ns):
            if site not in sub:
                up = tuple(sorted(sub + (site,)))
                Amat[row, 2 + index[up]] += 1.0
    return np.linalg.solve(Amat, rhs_v).tolist()
