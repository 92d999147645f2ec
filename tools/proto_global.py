"""Dev prototype (CPU, numpy) of the global-network integrator the CUDA kernel implements:
staged RODAS4 with the ANALYTIC Jacobian, bucket-landing, and the Schur-complement solve
(block tree elimination + dense |Q|x|Q| system over the regulator set).  Used only to validate the
math and the step counts against tests/golden/global_*.npz before writing csrc/global.cuh.
Not part of the product path, not the oracle."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import global_models as og
from phoskintime_b200.global_model import synthetic_system

GAMMA = 0.25
A = {(2, 1): 0.1544000000000000e+01,
     (3, 1): 0.9466785280815826e+00, (3, 2): 0.2557011698983284e+00,
     (4, 1): 0.3314825187068521e+01, (4, 2): 0.2896124015972201e+01, (4, 3): 0.9986419139977817e+00,
     (5, 1): 0.1221224509226641e+01, (5, 2): 0.6019134481288629e+01, (5, 3): 0.1253708332932087e+02,
     (5, 4): -0.6878860361058950e+00}
Cc = {(2, 1): -0.5668800000000000e+01,
      (3, 1): -0.2430093356833875e+01, (3, 2): -0.2063599157091915e+00,
      (4, 1): -0.1073529058151375e+00, (4, 2): -0.9594562251023355e+01, (4, 3): -0.2047028614809616e+02,
      (5, 1): 0.7496443313967647e+01, (5, 2): -0.1024680431464352e+02, (5, 3): -0.3399990352819905e+02,
      (5, 4): 0.1170890893206160e+02,
      (6, 1): 0.8083246795921522e+01, (6, 2): -0.7981132988064893e+01, (6, 3): -0.3152159432874371e+02,
      (6, 4): 0.1631930543123136e+02, (6, 5): -0.6058818238834054e+01}
for j in range(1, 5):
    A[(6, j)] = A[(5, j)]
A[(6, 5)] = 1.0


class Net:
    def __init__(self, s, params):
        self.s = s
        self.model = s.model
        self.idx = s.idx
        self.p = params
        self.N = s.idx.N
        self.n = s.idx.state_dim

    def bucket_inputs(self, jb):
        s, p = self.s, self.p
        Kt = s.kin_Kmat[:, jb] * p["c_k"]
        S = np.zeros(s.idx.total_sites)
        for i in range(s.n_W_rows):
            sl = slice(s.W_indptr[i], s.W_indptr[i + 1])
            S[i] = np.dot(s.W_data[sl], Kt[s.W_indices[sl]])
        return Kt, S

    def rhs_jac(self, y, Kt, S, want_jac=True):
        """f(y) and the pieces of the analytic Jacobian: per-state block coefficients + G (N x N)."""
        s, p, idx, N = self.s, self.p, self.idx, self.N
        f = np.zeros(self.n)
        pvec = np.zeros(N)
        for i in range(N):
            d = s.driver_map[i]
            st, ns = idx.offset_y[i], idx.n_sites[i]
            pvec[i] = Kt[d] if d >= 0 else y[st + 1:st + 2 + ns].sum()
        G = np.zeros((N, N))
        blocks = []
        for i in range(N):
            st, ss, ns = idx.offset_y[i], idx.offset_s[i], idx.n_sites[i]
            sl = slice(s.TF_indptr[i], s.TF_indptr[i + 1])
            v = np.dot(s.TF_data[sl], pvec[s.TF_indices[sl]]) / s.tf_deg[i]
            if self.model == 4:
                u_raw, du_raw = v, 1.0
            else:
                u_raw, du_raw = v / (1 + abs(v)), 1.0 / (1 + abs(v)) ** 2
            u = u_raw / (1 + abs(u_raw))
            du = 1.0 / (1 + abs(u_raw)) ** 2
            Ai, tfs = p["A_i"][i], p["tf_scale"]
            if u >= 0:
                synth = Ai * (1 + tfs * u / (1 + u + 1e-6))
                ds = Ai * tfs * (1 + 1e-6) / (1 + u + 1e-6) ** 2
            else:
                synth = Ai / (1 + tfs * abs(u))
                ds = Ai * tfs / (1 + tfs * abs(u)) ** 2
            g = ds * du * du_raw / s.tf_deg[i]
            for q in range(s.TF_indptr[i], s.TF_indptr[i + 1]):
                j = s.TF_indices[q]
                if s.driver_map[j] < 0:
                    G[i, j] += g * s.TF_data[q]
            R, P = y[st], y[st + 1]
            Bi, Ci, Di, Ei = p["B_i"][i], p["C_i"][i], p["D_i"][i], p["E_i"][i]
            f[st] = synth - Bi * R
            ps = y[st + 2:st + 2 + ns]
            Sj = S[ss:ss + ns]
            Dp = p["Dp_i"][ss:ss + ns]
            if self.model == 0:
                f[st + 1] = Ci * R - (Di + Sj.sum()) * P + Ei * ps.sum()
                f[st + 2:st + 2 + ns] = Sj * P - (Ei + Dp + Di) * ps
                blk = dict(cPR=Ci, dPP=-(Di + Sj.sum()), lo=Sj.copy(), up=np.full(ns, Ei), dg=-(Ei + Dp + Di))
            elif self.model == 4:
                f[st + 1] = Ci * R / (1 + R) - Di * P - (Sj * P / (1 + P)).sum() + Ei * ps.sum()
                f[st + 2:st + 2 + ns] = Sj * P / (1 + P) - (Dp + Di) * ps - Ei * ps
                q2 = 1.0 / (1 + P) ** 2
                blk = dict(cPR=Ci / (1 + R) ** 2, dPP=-Di - Sj.sum() * q2, lo=Sj * q2, up=np.full(ns, Ei), dg=-(Dp + Di + Ei))
            else:
                lo, up, dg = np.zeros(ns), np.zeros(ns), np.zeros(ns)
                if ns == 0:
                    f[st + 1] = Ci * R - Di * P
                    dPP = -Di
                else:
                    f[st + 1] = Ci * R - Di * P - Sj[0] * P + Ei * ps[0]
                    dPP = -Di - Sj[0]
                    for j in range(ns):
                        prev = P if j == 0 else ps[j - 1]
                        gain = Sj[j] * prev
                        out = Ei + Dp[j] + Di
                        if j < ns - 1:
                            gain += Ei * ps[j + 1]
                            out += Sj[j + 1]
                        f[st + 2 + j] = gain - out * ps[j]
                        lo[j], up[j], dg[j] = Sj[j], Ei, -out
                blk = dict(cPR=Ci, dPP=dPP, lo=lo, up=up, dg=dg)
            blk["B"] = Bi
            blocks.append(blk)
        return f, blocks, G

    def dense_jac(self, blocks, G):
        idx, n, N = self.idx, self.n, self.N
        J = np.zeros((n, n))
        chain = self.model == 1
        for i in range(N):
            st, ns = idx.offset_y[i], idx.n_sites[i]
            b = blocks[i]
            J[st, st] = -b["B"]
            J[st + 1, st] = b["cPR"]
            J[st + 1, st + 1] = b["dPP"]
            for j in range(ns):
                par = st + 1 + j if chain else st + 1
                J[st + 2 + j, par] += b["lo"][j]
                J[par, st + 2 + j] += b["up"][j]
                J[st + 2 + j, st + 2 + j] += b["dg"][j]
            for j2 in range(N):
                if G[i, j2] != 0.0:
                    s2, n2 = idx.offset_y[j2], idx.n_sites[j2]
                    J[st, s2 + 1:s2 + 2 + n2] += G[i, j2]
        return J


class SchurSolver:
    """(I - c J) x = b through per-protein tree elimination + dense system on the regulator set Q."""

    def __init__(self, net, blocks, G, c):
        self.net, self.c = net, c
        idx, N = net.idx, net.N
        chain = net.model == 1
        self.fac = []
        self.w = np.zeros(net.n)          # A_blk^-1 e_R  (response of each block to a unit mRNA-row input)
        self.m = np.zeros(N)
        for i in range(N):
            ns = idx.n_sites[i]
            b = blocks[i]
            piv = np.concatenate([[1 - c * b["dPP"]], 1 - c * b["dg"]])      # index 0 = P0, 1+j = site j
            mult = np.zeros(ns)
            for j in range(ns - 1, -1, -1):
                par = j if chain else 0                                       # pivot index of the parent
                mult[j] = (-c * b["up"][j]) / piv[1 + j]
                piv[par] -= mult[j] * (-c * b["lo"][j])
            self.fac.append(dict(piv=piv, mult=mult, iR=1.0 / (1 + c * b["B"]), cPR=c * b["cPR"], clo=c * b["lo"], chain=chain))
        e = np.zeros(net.n)
        e[idx.offset_y] = 1.0
        self.w = self.block_solve(e)
        for i in range(N):
            st, ns = idx.offset_y[i], idx.n_sites[i]
            self.m[i] = self.w[st + 1:st + 2 + ns].sum()
        self.Q = np.array([j for j in range(N) if np.any(G[:, j] != 0.0)], dtype=int)
        self.G = G
        Sc = np.eye(len(self.Q)) - c * self.m[self.Q][:, None] * G[np.ix_(self.Q, self.Q)]
        self.Sc = Sc

    def block_solve(self, b):
        idx, N = self.net.idx, self.net.N
        x = np.array(b, float)
        for i in range(N):
            st, ns = idx.offset_y[i], idx.n_sites[i]
            F = self.fac[i]
            x[st] *= F["iR"]
            x[st + 1] += F["cPR"] * x[st]
            for j in range(ns - 1, -1, -1):
                par = st + 1 + j if F["chain"] else st + 1
                x[par] -= F["mult"][j] * x[st + 2 + j]
            x[st + 1] /= F["piv"][0]
            for j in range(ns):
                par = st + 1 + j if F["chain"] else st + 1
                x[st + 2 + j] = (x[st + 2 + j] + F["clo"][j] * x[par]) / F["piv"][1 + j]
        return x

    def solve(self, b):
        idx, N, c = self.net.idx, self.net.N, self.c
        x0 = self.block_solve(b)
        z0 = np.array([x0[idx.offset_y[i] + 1:idx.offset_y[i] + 2 + idx.n_sites[i]].sum() for i in range(N)])
        zQ = np.linalg.solve(self.Sc, z0[self.Q]) if len(self.Q) else np.zeros(0)
        Gz = self.G[:, self.Q] @ zQ
        x = x0.copy()
        for i in range(N):
            st, ns = idx.offset_y[i], idx.n_sites[i]
            x[st:st + 2 + ns] += c * Gz[i] * self.w[st:st + 2 + ns]
        return x


def integrate(net, y0, t_eval, rtol, atol, use_schur=False, verbose=False):
    s = net.s
    grid = s.kin_grid
    t_eval = np.asarray(t_eval, float)
    stops = np.unique(np.concatenate([t_eval, grid[(grid > t_eval[0]) & (grid < t_eval[-1])]]))
    y = np.array(y0, float)
    out = {float(stops[0]): y.copy()}
    nst = nrej = 0
    h = None
    t = stops[0]
    n = net.n
    hacc, erracc, nacc = 0.0, 1.0, 0
    for a, b in zip(stops[:-1], stops[1:]):
        jb = int(og._bucket(0.5 * (a + b), grid))
        Kt, S = net.bucket_inputs(jb)
        if h is None:
            f0, _, _ = net.rhs_jac(y, Kt, S)
            sc = atol + rtol * np.abs(y)
            d0, d1 = np.max(np.abs(y) / sc), np.max(np.abs(f0) / sc)
            h = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        rejected = False
        while t < b:
            rem = b - t
            hh, land = h, False
            if 1.03 * hh >= rem:
                hh, land = rem, True
            elif hh > 0.5 * rem:
                hh = 0.5 * rem
            f, blocks, G = net.rhs_jac(y, Kt, S)
            c = hh * GAMMA
            if use_schur:
                sol = SchurSolver(net, blocks, G, c)
                solve = sol.solve
            else:
                J = net.dense_jac(blocks, G)
                Wm = np.eye(n) - c * J
                solve = lambda r: np.linalg.solve(Wm, r)
            U = {}
            for i in range(1, 7):
                if i == 1:
                    fi = f
                else:
                    arg = y + sum(A[(i, j)] * U[j] for j in range(1, i) if (i, j) in A)
                    fi = net.rhs_jac(arg, Kt, S)[0]
                r = fi + sum(Cc[(i, j)] / hh * U[j] for j in range(1, i) if (i, j) in Cc)
                U[i] = solve(c * r)
            ynew = y + sum(A[(5, j)] * U[j] for j in range(1, 5)) + U[5] + U[6]
            err = np.max(np.abs(U[6]) / (atol + rtol * np.maximum(np.abs(y), np.abs(ynew))))
            fac = max(1 / 6, min(5.0, err ** 0.25 / 0.9))
            if err <= 1.0 and np.all(np.isfinite(ynew)):
                nst += 1
                if nacc > 0:
                    fg = hacc / hh * (err * err / erracc) ** 0.25 / 0.9
                    fac = max(fac, max(1 / 6, min(5.0, fg)))
                hacc, erracc, nacc = hh, max(1e-2, err), nacc + 1
                hnew = hh / fac
                if rejected:
                    hnew = min(hnew, hh)
                rejected = False
                h = max(hnew, min(h, 6 * hh)) if hh < h else hnew
                y = ynew
                t = b if land else t + hh
            else:
                nrej += 1
                rejected = True
                h = hh / fac
        out[float(b)] = y.copy()
    return np.array([out[float(tt)] for tt in t_eval]), nst, nrej


if __name__ == "__main__":
    for f in ("global_m0_N10", "global_m1_N14", "global_m4_N14", "global_m0_N36"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f + ".npz"))
        s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]), model=int(g["model"]))
        for b in (0, 1):
            p = s.unpack_params(g["params"][b])
            net = Net(s, p)
            # analytic Jacobian vs finite differences of the ORACLE rhs, Schur solve vs dense solve
            y = g["Y_tight"][b][6]
            Kt, S = net.bucket_inputs(3)
            tmid = 0.5 * (s.kin_grid[3] + s.kin_grid[4])
            f0, blocks, G = net.rhs_jac(y, Kt, S)
            fo = og.rhs(s.model, y, tmid, s.as_dict(), p)
            J = net.dense_jac(blocks, G)
            Jfd = np.zeros_like(J)
            for j in range(net.n):
                e = np.zeros(net.n); hfd = 1e-6 * max(1.0, abs(y[j])); e[j] = hfd
                Jfd[:, j] = (og.rhs(s.model, y + e, tmid, s.as_dict(), p) - og.rhs(s.model, y - e, tmid, s.as_dict(), p)) / (2 * hfd)
            rb = np.random.default_rng(0).standard_normal(net.n)
            c = 3.7
            xs = SchurSolver(net, blocks, G, c).solve(rb)
            xd = np.linalg.solve(np.eye(net.n) - c * J, rb)
            print(f"{f} b={b}: |f-oracle| {np.abs(f0 - fo).max():.2e} |J-Jfd| {np.abs(J - Jfd).max():.2e} |schur-dense| {np.abs(xs - xd).max():.2e}", flush=True)
            for rtol, atol in ((1e-6, 1e-9), (1e-7, 1e-10)):
                Y, nst, nrej = integrate(net, g["y0"], g["t"], rtol, atol, use_schur=(b == 1))
                et = np.abs(Y - g["Y_tight"][b]) / (1e-6 * np.abs(g["Y_tight"][b]) + 1e-9)
                es = np.abs(Y - g["Y"][b]) / (1e-6 * np.abs(g["Y"][b]) + 1e-7)
                est = np.abs(g["Y"][b] - g["Y_tight"][b]) / (1e-6 * np.abs(g["Y_tight"][b]) + 1e-9)
                print(f"   rtol {rtol:g}: steps {nst} rej {nrej} | vs tight {et.max():.3g} | vs stock {es.max():.3g} | stock vs tight {est.max():.3g}", flush=True)
