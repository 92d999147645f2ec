"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's global-network path
(models 0 distributive, 1 sequential, 2 combinatorial, 4 saturating).

Restates (only tests/, smoke() and bench.py's CPU legs may import this):
  * the odeint RHS wrappers   /root/reference/global_model/jacspeedup.py:175-237 (distributive),
                              :240-282 (sequential), :347-375 (saturating): kinase step input
                              (kin_eval_step :148-172), S = W.(K*c_k), live-drive P_vec, TF input with
                              the first squash (the saturating wrapper skips it, :371-373)
  * the block kernels         /root/reference/global_model/models.py:27-65 (synthesis rate, second
                              squash, 1e-6 in the denominator), :71-146, :149-212, :215-306
  * the combinatorial model   jacspeedup.py:287-343 wrapper (no live drive: every protein's total is the sum
                              of its pattern states; one squash), models.py:322-432 block kinetics
                              (dephosphorylation E per set bit, per-pattern decay sum(Dp_j + D), forward
                              transitions from the edge lists of models.py:435-485), rate table
                              jacspeedup.py:114-145, loss lossfn.py:248-382, observables simulate.py:135-158
  * the forward-difference dense Jacobian  jacspeedup.py:397-448 (h = 1e-8*max(1,|y_j|))
  * simulate_odeint default path           global_model/simulate.py:34-80 (LSODA + Dfun, col_deriv=False)
  * the 3-modality loss                    global_model/lossfn.py:28-246 (all 8 modes, EPS=1e-9 floors)
  * the scalar Morris metric               global_model/simulate.py:105-182 FC tables (1e-12 floors) and
                                           global_model/sensitivity.py:106-140
  * the prior penalty / objectives         global_model/optproblem.py:87-160
  * softplus parameter packing             global_model/params.py:106-132, utils.py:228-253

The reference compiles these kernels with numba fastmath=True (re-association allowed); the
restatement uses plain IEEE order, so agreement with the reference is ~1e-12, not bitwise.
Pinned by tests/golden/global_*.npz (oracle/gen_golden_global.py runs the unmodified reference).
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import odeint

try:
    from numba import njit
except Exception:  # pragma: no cover
    def njit(*a, **k):
        def deco(f):
            return f
        return deco if not (a and callable(a[0])) else a[0]

EPS = 1e-9
PARAM_KEYS = ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i")


@njit(cache=False)
def _bucket(t, grid):
    # kin_eval_step / time_bucket: grid[j] <= t < grid[j+1], clamped at both ends
    if t <= grid[0]:
        return 0
    if t >= grid[-1]:
        return grid.size - 1
    j = np.searchsorted(grid, t, side="right") - 1
    if j < 0:
        j = 0
    if j >= grid.size:
        j = grid.size - 1
    return j


@njit(cache=False)
def _synth(Ai, tf_scale, u_raw):
    u = u_raw / (1.0 + abs(u_raw))
    if u >= 0.0:
        return Ai * (1.0 + (tf_scale * u) / (1.0 + u + 1e-6))
    return Ai / (1.0 + tf_scale * abs(u))


@njit(cache=False)
def _rhs(model, y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, kin_Kmat,
         W_indptr, W_indices, W_data, n_W_rows, TF_indptr, TF_indices, TF_data, n_TF_rows,
         offset_y, offset_s, n_sites, tf_deg, driver_map):
    dy = np.zeros_like(y)
    jb = _bucket(t, kin_grid)
    Kt = kin_Kmat[:, jb] * c_k
    S_all = np.zeros(n_W_rows)
    for i in range(n_W_rows):
        s = 0.0
        for p in range(W_indptr[i], W_indptr[i + 1]):
            s += W_data[p] * Kt[W_indices[p]]
        S_all[i] = s
    P_vec = np.zeros(n_TF_rows)
    for i in range(n_TF_rows):
        d = driver_map[i]
        if d >= 0:
            P_vec[i] = Kt[d]
        else:
            st = offset_y[i]
            tot = y[st + 1]
            for j in range(n_sites[i]):
                tot += y[st + 2 + j]
            P_vec[i] = tot
    TF_in = np.zeros(n_TF_rows)
    for i in range(n_TF_rows):
        s = 0.0
        for p in range(TF_indptr[i], TF_indptr[i + 1]):
            s += TF_data[p] * P_vec[TF_indices[p]]
        v = s / tf_deg[i]
        TF_in[i] = v if model == 4 else v / (1.0 + abs(v))
    N = A_i.shape[0]
    for i in range(N):
        st = offset_y[i]
        ss = offset_s[i]
        ns = n_sites[i]
        R = y[st]
        P = y[st + 1]
        dy[st] = _synth(A_i[i], tf_scale, TF_in[i]) - B_i[i] * R
        if model == 0:
            if ns == 0:
                dy[st + 1] = C_i[i] * R - D_i[i] * P
            else:
                sum_S = 0.0
                back = 0.0
                for j in range(ns):
                    s_rate = S_all[ss + j]
                    ps = y[st + 2 + j]
                    sum_S += s_rate
                    back += E_i[i] * ps
                    dy[st + 2 + j] = s_rate * P - (E_i[i] + Dp_i[ss + j] + D_i[i]) * ps
                dy[st + 1] = C_i[i] * R - (D_i[i] + sum_S) * P + back
        elif model == 4:
            trans = (C_i[i] * R) / (1.0 + R)
            if ns == 0:
                dy[st + 1] = trans - D_i[i] * P
            else:
                flux = 0.0
                back = 0.0
                for j in range(ns):
                    fwd = (S_all[ss + j] * P) / (1.0 + P)
                    bwd = E_i[i] * y[st + 2 + j]
                    flux += fwd
                    back += bwd
                    dy[st + 2 + j] = fwd - (Dp_i[ss + j] + D_i[i]) * y[st + 2 + j] - bwd
                dy[st + 1] = trans - D_i[i] * P - flux + back
        else:  # sequential chain P0 -> P1 -> ... -> Pns with distributive back-flow E
            if ns == 0:
                dy[st + 1] = C_i[i] * R - D_i[i] * P
            else:
                Ei = E_i[i]
                Di = D_i[i]
                dy[st + 1] = C_i[i] * R - Di * P - S_all[ss] * P + Ei * y[st + 2]
                for j in range(ns):
                    idx = st + 2 + j
                    gain = S_all[ss + j] * y[idx - 1]
                    out = Ei + Dp_i[ss + j] + Di
                    if j < ns - 1:
                        gain += Ei * y[idx + 1]
                        out += S_all[ss + j + 1]
                    dy[idx] = gain - out * y[idx]
    return dy


@njit(cache=False)
def _fd_jac(model, y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, kin_Kmat,
            W_indptr, W_indices, W_data, n_W_rows, TF_indptr, TF_indices, TF_data, n_TF_rows,
            offset_y, offset_s, n_sites, tf_deg, driver_map):
    n = y.size
    J = np.empty((n, n))
    f0 = _rhs(model, y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, kin_Kmat,
              W_indptr, W_indices, W_data, n_W_rows, TF_indptr, TF_indices, TF_data, n_TF_rows,
              offset_y, offset_s, n_sites, tf_deg, driver_map)
    for j in range(n):
        yp = y.copy()
        aj = y[j]
        h = 1e-8 * (1.0 if abs(aj) < 1.0 else abs(aj))
        yp[j] = aj + h
        fj = _rhs(model, yp, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, kin_Kmat,
                  W_indptr, W_indices, W_data, n_W_rows, TF_indptr, TF_indices, TF_data, n_TF_rows,
                  offset_y, offset_s, n_sites, tf_deg, driver_map)
        J[:, j] = (fj - f0) * (1.0 / h)
    return J


@njit(cache=False)
def _rhs_comb(y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, S_cache,
              TF_indptr, TF_indices, TF_data, n_TF_rows, offset_y, offset_s, n_sites, n_states,
              trans_from, trans_to, trans_site, trans_off, trans_n, tf_deg, driver_map):
    dy = np.zeros_like(y)
    jb = _bucket(t, kin_grid)
    P_vec = np.zeros(n_TF_rows)
    for i in range(n_TF_rows):              # jacspeedup.py:318-325: driver_map is NOT consulted
        tot = 0.0
        for m in range(n_states[i]):
            tot += y[offset_y[i] + 1 + m]
        P_vec[i] = tot
    N = A_i.shape[0]
    for i in range(N):
        s = 0.0
        for p in range(TF_indptr[i], TF_indptr[i + 1]):
            s += TF_data[p] * P_vec[TF_indices[p]]
        v = s / tf_deg[i]
        u = v / (1.0 + abs(v))
        st = offset_y[i]
        ss = offset_s[i]
        ns = n_sites[i]
        R = y[st]
        base = st + 1
        dy[st] = _synth(A_i[i], tf_scale, u) - B_i[i] * R
        if ns == 0:
            dy[base] = C_i[i] * R - D_i[i] * y[base]
            continue
        dy[base] += C_i[i] * R
        dy[base] += -D_i[i] * y[base]
        for m in range(1, n_states[i]):
            Pm = y[base + m]
            if Pm == 0.0:
                continue
            dp_rate = 0.0
            for j in range(ns):
                if (m >> j) & 1:
                    flux = E_i[i] * Pm
                    dy[base + m] -= flux
                    dy[base + (m ^ (1 << j))] += flux
                    dp_rate += Dp_i[ss + j] + D_i[i]
            dy[base + m] -= dp_rate * Pm
        for k in range(trans_n[i]):
            frm = trans_from[trans_off[i] + k]
            to = trans_to[trans_off[i] + k]
            flux = S_cache[ss + trans_site[trans_off[i] + k], jb] * y[base + frm]
            dy[base + frm] -= flux
            dy[base + to] += flux
    return dy


@njit(cache=False)
def _fd_jac_comb(y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, S_cache,
                 TF_indptr, TF_indices, TF_data, n_TF_rows, offset_y, offset_s, n_sites, n_states,
                 trans_from, trans_to, trans_site, trans_off, trans_n, tf_deg, driver_map):
    n = y.size
    J = np.empty((n, n))
    f0 = _rhs_comb(y, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, S_cache,
                   TF_indptr, TF_indices, TF_data, n_TF_rows, offset_y, offset_s, n_sites, n_states,
                   trans_from, trans_to, trans_site, trans_off, trans_n, tf_deg, driver_map)
    for j in range(n):
        yp = y.copy()
        aj = y[j]
        h = 1e-8 * (1.0 if abs(aj) < 1.0 else abs(aj))
        yp[j] = aj + h
        fj = _rhs_comb(yp, t, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, kin_grid, S_cache,
                       TF_indptr, TF_indices, TF_data, n_TF_rows, offset_y, offset_s, n_sites, n_states,
                       trans_from, trans_to, trans_site, trans_off, trans_n, tf_deg, driver_map)
        J[:, j] = (fj - f0) * (1.0 / h)
    return J


def comb_tables(n_sites):
    """models.py:435-485 edge lists (m -> m | 1<<j for every unset bit, patterns then sites ascending)."""
    frm, to, site, off, cnt = [], [], [], [], []
    for ns in np.asarray(n_sites, int):
        off.append(len(frm))
        for m in range(1 << ns if ns > 0 else 0):
            for j in range(ns):
                if not (m >> j) & 1:
                    frm.append(m)
                    to.append(m | (1 << j))
                    site.append(j)
        cnt.append(len(frm) - off[-1])
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return i32(frm), i32(to), i32(site), i32(off), i32(cnt)


def s_cache(net, c_k):
    """jacspeedup.py:114-145."""
    Kc = np.asarray(net["kin_Kmat"]) * np.asarray(c_k)[:, None]
    out = np.zeros((int(net["n_W_rows"]), Kc.shape[1]))
    for i in range(out.shape[0]):
        for p in range(net["W_indptr"][i], net["W_indptr"][i + 1]):
            out[i] += net["W_data"][p] * Kc[net["W_indices"][p]]
    return out


def args_tuple_comb(net, params=None):
    """The 27-tuple of network.py:475-505 without the two work buffers."""
    p = net["defaults"] if params is None else params
    c = lambda k: np.ascontiguousarray(p[k], float)
    n_states = (1 << np.asarray(net["n_sites"], np.int32)).astype(np.int32)
    return (c("c_k"), c("A_i"), c("B_i"), c("C_i"), c("D_i"), c("Dp_i"), c("E_i"), float(p["tf_scale"]),
            net["kin_grid"], s_cache(net, p["c_k"]), net["TF_indptr"], net["TF_indices"], net["TF_data"], int(net["N"]),
            net["offset_y"], net["offset_s"], net["n_sites"], n_states, *comb_tables(net["n_sites"]),
            net["tf_deg"], net["driver_map"])


def _dispatch(model, net, params):
    """(rhs(y, t, *args), jac(y, t, *args), args) for a kinetic model."""
    if int(model) == 2:
        return _rhs_comb, _fd_jac_comb, args_tuple_comb(net, params)
    m = int(model)
    return (lambda y, t, *a: _rhs(m, y, t, *a)), (lambda y, t, *a: _fd_jac(m, y, t, *a)), args_tuple(net, params)


def args_tuple(net, params=None):
    """The 23-tuple of System.odeint_args (network.py:508-526); `params` overrides net['defaults']."""
    p = net["defaults"] if params is None else params
    return (np.ascontiguousarray(p["c_k"], float), np.ascontiguousarray(p["A_i"], float),
            np.ascontiguousarray(p["B_i"], float), np.ascontiguousarray(p["C_i"], float),
            np.ascontiguousarray(p["D_i"], float), np.ascontiguousarray(p["Dp_i"], float),
            np.ascontiguousarray(p["E_i"], float), float(p["tf_scale"]),
            net["kin_grid"], net["kin_Kmat"],
            net["W_indptr"], net["W_indices"], net["W_data"], int(net["n_W_rows"]),
            net["TF_indptr"], net["TF_indices"], net["TF_data"], int(net["N"]),
            net["offset_y"], net["offset_s"], net["n_sites"], net["tf_deg"], net["driver_map"])


def rhs(model, y, t, net, params=None):
    f, _, args = _dispatch(model, net, params)
    return f(np.asarray(y, float), float(t), *args)


def simulate_odeint(model, net, t_eval, rtol, atol, mxstep, params=None, y0=None, tcrit=None):
    """simulate.py:69-79: LSODA with the dense forward-difference Jacobian, integrating THROUGH the
    kinase-bucket discontinuities.  `tcrit` (not used by the reference) lets the tight-tolerance
    variant tell LSODA where the RHS jumps."""
    f, jac, args = _dispatch(model, net, params)
    y0 = np.array(net["y0"] if y0 is None else y0, dtype=np.float64)
    kw = {} if tcrit is None else {"tcrit": np.asarray(tcrit, float)}
    xs = odeint(lambda y, t, *a: f(y, t, *a), y0, np.asarray(t_eval, np.float64), args=args,
                Dfun=lambda y, t, *a: jac(np.asarray(y, np.float64), t, *a), col_deriv=False,
                rtol=rtol, atol=atol, mxstep=mxstep, **kw)
    return np.ascontiguousarray(xs, dtype=np.float64)


def simulate_exact_buckets(model, net, t_eval, params=None, y0=None, rtol=1e-12, atol=1e-13):
    """O2 for the global path: the same RHS integrated bucket by bucket (restart at every kinase-grid
    point, so no step ever straddles a discontinuity) at tight tolerance."""
    f, _, args = _dispatch(model, net, params)
    t_eval = np.asarray(t_eval, float)
    grid = np.asarray(net["kin_grid"], float)
    stops = np.unique(np.concatenate([t_eval, grid[(grid > t_eval[0]) & (grid < t_eval[-1])]]))
    y = np.array(net["y0"] if y0 is None else y0, dtype=np.float64)
    out = {float(stops[0]): y.copy()}
    for a, b in zip(stops[:-1], stops[1:]):
        mid = 0.5 * (a + b)      # evaluate the piecewise-constant input inside the bucket
        ys = odeint(lambda yy, tt, *aa: f(yy, mid, *aa), y, [a, b], args=args, rtol=rtol, atol=atol,
                    mxstep=500000)
        y = ys[-1]
        out[float(b)] = y.copy()
    return np.array([out[float(t)] for t in t_eval])


# ---------------------------------------------------------------------------------- loss
def _atom(mode, diff, obs, pred):
    if mode == 0:
        return diff * diff
    if mode == 1:
        a = abs(diff)
        return 0.5 * diff * diff if a <= 0.5 else 0.5 * (a - 0.25)
    if mode == 2:
        d = np.log(diff + EPS) - np.log(obs + EPS)
        x = d / 0.5
        return 0.25 * ((1.0 + x * x) ** 0.5 - 1.0)
    if mode == 3:
        s = abs(diff)
        return s - 0.69314718056 if s > 20.0 else np.log(np.cosh(diff))
    if mode == 4:
        return np.log(1.0 + diff * diff)
    if mode == 5:
        return (diff * diff) / (abs(pred) + 1e-6)
    if mode == 6:
        return diff * diff / (diff * diff + 1.0)
    return (diff * diff + 1e-6) ** 0.5 - 1e-3


def loss_noncomb(Y, ld, mode=0):
    """(loss_p, loss_r, loss_ph) of lossfn.py:113-246."""
    fl = lambda v: v if v > EPS else EPS
    pm = ld["prot_map"]
    out = [0.0, 0.0, 0.0]
    with np.errstate(invalid="ignore", divide="ignore"):
        for k in range(ld["p_prot"].size):
            st, ns = pm[ld["p_prot"][k]]
            tot_t = Y[ld["t_prot"][k], st + 1:st + 2 + ns].sum()
            tot_b = Y[ld["prot_base_idx"], st + 1:st + 2 + ns].sum()
            pred = fl(tot_t) / fl(tot_b)
            out[0] += ld["w_prot"][k] * _atom(mode, ld["obs_prot"][k] - pred, ld["obs_prot"][k], pred)
        for k in range(ld["p_rna"].size):
            st = pm[ld["p_rna"][k], 0]
            pred = fl(Y[ld["t_rna"][k], st]) / fl(Y[ld["rna_base_idx"], st])
            out[1] += ld["w_rna"][k] * _atom(mode, ld["obs_rna"][k] - pred, ld["obs_rna"][k], pred)
        for k in range(ld["p_pho"].size):
            col = pm[ld["p_pho"][k], 0] + 2 + ld["s_pho"][k]
            pred = fl(Y[ld["t_pho"][k], col]) / fl(Y[ld["pho_base_idx"], col])
            out[2] += ld["w_pho"][k] * _atom(mode, ld["obs_pho"][k] - pred, ld["obs_pho"][k], pred)
    return tuple(out)


def loss_comb(Y, ld, mode=0):
    """(loss_p, loss_r, loss_ph) of lossfn.py:248-382: total protein = sum of all pattern states, site j =
    sum of the patterns with bit j set; prot_map[:, 1] = n_states."""
    fl = lambda v: v if v > EPS else EPS
    pm = ld["prot_map"]
    out = [0.0, 0.0, 0.0]
    with np.errstate(invalid="ignore", divide="ignore"):
        for k in range(ld["p_prot"].size):
            st, nst = pm[ld["p_prot"][k]]
            pred = fl(Y[ld["t_prot"][k], st + 1:st + 1 + nst].sum()) / fl(Y[ld["prot_base_idx"], st + 1:st + 1 + nst].sum())
            out[0] += ld["w_prot"][k] * _atom(mode, ld["obs_prot"][k] - pred, ld["obs_prot"][k], pred)
        for k in range(ld["p_rna"].size):
            st = pm[ld["p_rna"][k], 0]
            pred = fl(Y[ld["t_rna"][k], st]) / fl(Y[ld["rna_base_idx"], st])
            out[1] += ld["w_rna"][k] * _atom(mode, ld["obs_rna"][k] - pred, ld["obs_rna"][k], pred)
        for k in range(ld["p_pho"].size):
            st, nst = pm[ld["p_pho"][k]]
            sel = st + 1 + np.flatnonzero((np.arange(nst) >> ld["s_pho"][k]) & 1)
            pred = fl(Y[ld["t_pho"][k], sel].sum()) / fl(Y[ld["pho_base_idx"], sel].sum())
            out[2] += ld["w_pho"][k] * _atom(mode, ld["obs_pho"][k] - pred, ld["obs_pho"][k], pred)
    return tuple(out)


def loss(model, Y, ld, mode=0):
    """LOSS_FN as the reference binds it at import time (lossfn.py:386)."""
    return loss_comb(Y, ld, mode) if int(model) == 2 else loss_noncomb(Y, ld, mode)


def objectives(losses, ld, params, defaults, lambdas=(1.0, 1.0, 1.0), lambda_prior=0.1):
    """F[3] of GlobalODE_MOO._evaluate (optproblem.py:105-160)."""
    acc, cnt = 0.0, 0
    for k in ("A_i", "B_i", "C_i", "D_i", "E_i"):
        d = (np.asarray(params[k]) - np.asarray(defaults[k])) / (np.asarray(defaults[k]) + 1e-6)
        acc += float(np.sum(d ** 2))
        cnt += d.size
    prior = lambda_prior * (acc / max(1, cnt))
    norms = [1.0 / max(1e-6, float(np.sum(ld[w]))) for w in ("w_prot", "w_rna", "w_pho")]
    return np.array([losses[i] * norms[i] * lambdas[i] + prior for i in range(3)])


def scalar_metric(Y, net, obs, metric="total_signal"):
    """Morris scalar of the global path: all fold-changes simulate_and_measure tabulates
    (simulate.py:105-182: every protein at the protein times, every protein's RNA at the RNA times,
    every site at the phospho times; floors 1e-12; bases t=0 / t=4 / t=0) reduced as in
    sensitivity.py:106-140.  `obs` = dict(t_prot, t_rna, t_pho index arrays, prot_b, rna_b, pho_b)."""
    vals = []
    fl = lambda a: np.maximum(a, 1e-12)
    if int(net.get("model", 0)) == 2:                  # simulate.py:135-158
        for i in range(net["N"]):
            st, ns = int(net["offset_y"][i]), int(net["n_sites"][i])
            tot = Y[:, st + 1:st + 1 + (1 << ns)].sum(axis=1)
            vals.append(fl(tot[obs["t_prot"]]) / fl(tot[obs["prot_b"]]))
        for i in range(net["N"]):
            R = Y[:, int(net["offset_y"][i])]
            vals.append(fl(R[obs["t_rna"]]) / fl(R[obs["rna_b"]]))
        for i in range(net["N"]):
            st, ns = int(net["offset_y"][i]), int(net["n_sites"][i])
            for j in range(ns):
                ph = Y[:, st + 1 + np.flatnonzero((np.arange(1 << ns) >> j) & 1)].sum(axis=1)
                vals.append(fl(ph[obs["t_pho"]]) / fl(ph[obs["pho_b"]]))
        return _reduce_metric(np.concatenate(vals), metric)
    for i in range(net["N"]):
        st, ns = int(net["offset_y"][i]), int(net["n_sites"][i])
        tot = Y[:, st + 1:st + 2 + ns].sum(axis=1)
        vals.append(fl(tot[obs["t_prot"]]) / fl(tot[obs["prot_b"]]))
    for i in range(net["N"]):
        R = Y[:, int(net["offset_y"][i])]
        vals.append(fl(R[obs["t_rna"]]) / fl(R[obs["rna_b"]]))
    for i in range(net["N"]):
        st, ns = int(net["offset_y"][i]), int(net["n_sites"][i])
        for j in range(ns):
            ph = Y[:, st + 2 + j]
            vals.append(fl(ph[obs["t_pho"]]) / fl(ph[obs["pho_b"]]))
    return _reduce_metric(np.concatenate(vals), metric)


def fc_tables(Y, net, obs):
    """The fold-change values of simulate_and_measure (simulate.py:105-182) as three arrays
    (fc_prot[N, len(t_prot)], fc_rna[N, len(t_rna)], fc_pho[total_sites, len(t_pho)])."""
    fl = lambda a: np.maximum(a, 1e-12)
    comb = int(net.get("model", 0)) == 2
    P, R, PH = [], [], []
    for i in range(net["N"]):
        st, ns = int(net["offset_y"][i]), int(net["n_sites"][i])
        tot = Y[:, st + 1:st + 1 + (1 << ns)].sum(axis=1) if comb else Y[:, st + 1:st + 2 + ns].sum(axis=1)
        P.append(fl(tot[obs["t_prot"]]) / fl(tot[obs["prot_b"]]))
        Rv = Y[:, st]
        R.append(fl(Rv[obs["t_rna"]]) / fl(Rv[obs["rna_b"]]))
        for j in range(ns):
            ph = (Y[:, st + 1 + np.flatnonzero((np.arange(1 << ns) >> j) & 1)].sum(axis=1) if comb else Y[:, st + 2 + j])
            PH.append(fl(ph[obs["t_pho"]]) / fl(ph[obs["pho_b"]]))
    return np.array(P), np.array(R), np.array(PH).reshape(-1, len(obs["t_pho"]))


def _reduce_metric(c, metric):
    if metric == "mean":
        return float(np.mean(c))
    if metric == "variance":
        return float(np.var(c))
    if metric == "l2_norm":
        return float(np.linalg.norm(c))
    return float(np.sum(c))


# --------------------------------------------------------------------------- parameters
def softplus(x):
    x = np.asarray(x, float)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def pack_params(p):
    """Flat physical vector [c_k | A_i | B_i | C_i | D_i | Dp_i | E_i | tf_scale] (params.py:60-101 order)."""
    return np.concatenate([np.asarray(p[k], float).ravel() for k in PARAM_KEYS] + [[float(p["tf_scale"])]])


def unpack_params(vec, net):
    sizes = [net["K"], net["N"], net["N"], net["N"], net["N"], net["total_sites"], net["N"]]
    out, o = {}, 0
    for k, s in zip(PARAM_KEYS, sizes):
        out[k] = np.array(vec[o:o + s], float)
        o += s
    out["tf_scale"] = float(vec[o])
    return out
