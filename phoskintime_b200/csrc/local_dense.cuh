// Warp-per-system RODAS4 kernel with a dense shared-memory LU: the random (2^ns-state hypercube)
// model (models/randmod.py) and distributive / successive systems too large for the
// register-resident kernel.  One warp owns one system: W = I/(h*gamma) - J is assembled in shared
// memory from the analytic Jacobian (the models are linear, J is the transition-rate matrix),
// factorised in place (right-looking LU, lanes own columns of the trailing block, no pivoting:
// for non-negative rates W is strictly column diagonally dominant once the decoupled mRNA row is
// removed), and the six stage systems are solved by substitution with the stage vector in shared
// memory.  Warps pull systems from the global queue as they finish.
#pragma once
#include "pk_common.cuh"

namespace pk {

struct DenseLayout {       // per-warp shared-memory carve-up, in doubles
    int n, ld, P, nobs;
    __host__ __device__ int W() const { return 0; }
    __host__ __device__ int vec(int k) const { return n * ld + k * n; }   // k = 0..3: y, v, y_new, err
    __host__ __device__ int par() const { return n * ld + 4 * n; }
    __host__ __device__ int prev() const { return par() + P; }
    __host__ __device__ int total() const { return prev() + nobs; }
};

// ------------------------------------------------------------------ model: rhs and W assembly
// p = physical parameters in shared memory.  All functions are warp-cooperative (lane strided).
template <int MODEL>
__device__ __forceinline__ void dense_rhs(int ns, int n, const double* p, const double* y, double* f, int lane) {
    if (MODEL == 2) {
        const double* S = p + 4;
        const double* Dd = p + 4 + ns;
        const double Pv = y[1];
        for (int r = lane; r < n; r += 32) {
            double v;
            if (r == 0) v = fma(-p[1], y[0], p[0]);
            else if (r == 1) {
                double sS = 0.0, back = 0.0;
                for (int k = 0; k < ns; ++k) { sS += S[k]; back += y[1 + (1 << k)]; }
                v = fma(p[2], y[0], -(p[3] + sS) * Pv) + back;
            } else {
                const int s = r - 1;
                const double rate_in = S[__ffs(s) - 1];
                double gain = 0.0, out = Dd[s - 1];
                for (int j = 0; j < ns; ++j) {
                    const int bit = 1 << j;
                    if (s & bit) {
                        const int src = s & ~bit;
                        gain = fma(rate_in, src ? y[1 + src] : Pv, gain);
                        out += 1.0;
                    } else {
                        const int up = s | bit;
                        gain += y[1 + up];
                        out += S[__ffs(up) - 1];
                    }
                }
                v = fma(-out, y[r], gain);
            }
            f[r] = v;
        }
    } else {
        const double* S = p + 4;
        const double* Dr = p + 4 + ns;
        for (int r = lane; r < n; r += 32) {
            double v;
            if (r == 0) v = fma(-p[1], y[0], p[0]);
            else if (MODEL == 0) {
                if (r == 1) {
                    double sS = 0.0, back = 0.0;
                    for (int k = 0; k < ns; ++k) { sS += S[k]; back += y[2 + k]; }
                    v = fma(p[2], y[0], -(p[3] + sS) * y[1]) + back;
                } else {
                    v = fma(S[r - 2], y[1], -(1.0 + Dr[r - 2]) * y[r]);
                }
            } else {
                if (r == 1) v = fma(p[2], y[0], fma(-(p[3] + S[0]), y[1], y[2]));
                else {
                    const int i = r - 2;
                    double d = 1.0 + Dr[i] + (i < ns - 1 ? S[i + 1] : 0.0);
                    v = fma(S[i], y[r - 1], -d * y[r]);
                    if (i < ns - 1) v += y[r + 1];
                }
            }
            f[r] = v;
        }
    }
}

// W = I - c*J (row-major, leading dimension ld), c = h*gamma.  Each lane assembles whole rows.
template <int MODEL>
__device__ __forceinline__ void dense_fillW(int ns, int n, int ld, const double* p, double c, double* W, int lane) {
    for (int idx = lane; idx < n * ld; idx += 32) W[idx] = 0.0;
    __syncwarp();
    const double* S = p + 4;
    for (int r = lane; r < n; r += 32) {
        double* row = W + r * ld;
        if (r == 0) { row[0] = fma(c, p[1], 1.0); continue; }
        if (MODEL == 2) {
            const double* Dd = p + 4 + ns;
            if (r == 1) {
                double sS = 0.0;
                for (int k = 0; k < ns; ++k) { sS += S[k]; row[1 + (1 << k)] = -c; }
                row[0] = -c * p[2];
                row[1] = fma(c, p[3] + sS, 1.0);
            } else {
                const int s = r - 1;
                const double rate_in = S[__ffs(s) - 1];
                double out = Dd[s - 1];
                for (int j = 0; j < ns; ++j) {
                    const int bit = 1 << j;
                    if (s & bit) { row[1 + (s & ~bit)] = -c * rate_in; out += 1.0; }
                    else { const int up = s | bit; row[1 + up] = -c; out += S[__ffs(up) - 1]; }
                }
                row[r] = fma(c, out, 1.0);
            }
        } else {
            const double* Dr = p + 4 + ns;
            if (MODEL == 0) {
                if (r == 1) {
                    double sS = 0.0;
                    for (int k = 0; k < ns; ++k) { sS += S[k]; row[2 + k] = -c; }
                    row[0] = -c * p[2];
                    row[1] = fma(c, p[3] + sS, 1.0);
                } else {
                    row[1] = -c * S[r - 2];
                    row[r] = fma(c, 1.0 + Dr[r - 2], 1.0);
                }
            } else {
                if (r == 1) { row[0] = -c * p[2]; row[1] = fma(c, p[3] + S[0], 1.0); row[2] = -c; }
                else {
                    const int i = r - 2;
                    row[r - 1] = -c * S[i];
                    row[r] = fma(c, 1.0 + Dr[i] + (i < ns - 1 ? S[i + 1] : 0.0), 1.0);
                    if (i < ns - 1) row[r + 1] = -c;
                }
            }
        }
    }
    __syncwarp();
}

// In-place LU (unit lower), no pivoting.  Lanes own columns of the trailing block; the pivot row
// element stays in a register while the lane walks down its column.
__device__ __forceinline__ void dense_lu(int n, int ld, double* W, int lane) {
    for (int k = 0; k < n - 1; ++k) {
        const double ipiv = 1.0 / W[k * ld + k];
        for (int i = k + 1 + lane; i < n; i += 32) W[i * ld + k] *= ipiv;
        __syncwarp();
        for (int j = k + 1 + lane; j < n; j += 32) {
            const double ukj = W[k * ld + j];
            if (ukj != 0.0) {
                for (int i = k + 1; i < n; ++i) {
                    double* wij = W + i * ld + j;
                    *wij = fma(-W[i * ld + k], ukj, *wij);
                }
            }
        }
        __syncwarp();
    }
}

// Solve W x = r in place (x in shared memory).
__device__ __forceinline__ void dense_solve(int n, int ld, const double* W, double* x, int lane) {
    for (int k = 0; k < n - 1; ++k) {              // L y = r
        const double xk = x[k];
        for (int i = k + 1 + lane; i < n; i += 32) x[i] = fma(-W[i * ld + k], xk, x[i]);
        __syncwarp();
    }
    for (int k = n - 1; k >= 0; --k) {             // U x = y
        if (lane == 0) x[k] /= W[k * ld + k];
        __syncwarp();
        const double xk = x[k];
        for (int i = lane; i < k; i += 32) x[i] = fma(-W[i * ld + k], xk, x[i]);
        __syncwarp();
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int MODEL>
__global__ void __launch_bounds__(32) local_dense_kernel(const LocalArgs a, const DenseLayout lay) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const int n = a.n, ns = a.ns, ld = lay.ld, T = a.T, P = a.P, nobs = lay.nobs;
    double* W = smem + lay.W();
    double* y = smem + lay.vec(0);
    double* v = smem + lay.vec(1);
    double* w = smem + lay.vec(2);
    double* E = smem + lay.vec(3);
    double* p = smem + lay.par();
    double* prev = smem + lay.prev();
    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;

    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if ((long long)idx >= a.B) break;
        const size_t sys = (size_t)idx;

        // ---------------------------------------------------------------------------- init
        double p2l = 0.0;
        for (int i = lane; i < P; i += 32) {
            double v = a.params[sys * P + i];
            if (a.log_params) v = exp(v);
            p[i] = v;
            p2l = fma(v, v, p2l);
        }
        const double* y0 = a.y0 + (a.y0_stride ? sys * (size_t)a.y0_stride : 0);
        for (int i = lane; i < n; i += 32) y[i] = y0[i];
        __syncwarp();
        const int grp = a.group ? a.group[sys] : 0;
        const double* tg = a.target ? a.target + (size_t)grp * a.L : nullptr;
        const double* sg = a.sigma ? a.sigma + (size_t)grp * a.sigma_len : nullptr;
        EpiAcc e{0, 0, 0, 0, 0, 0};
        double t = a.t[0];
        int nst = 0, nrej = 0, status = 0, kout = 1;

        // lane-parallel emit of output index k from vector src (nullptr -> NaN)
        auto emit = [&](int k, const double* src) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            for (int i = lane; i < n; i += 32) {
                double v = src ? fmax(src[i], 0.0) : qnan;
                if (a.normalize) v *= 1.0 / y0[i];
                if (a.out_sol) a.out_sol[(sys * T + k) * n + i] = v;
                if (i < nobs) {
                    int fi = (i == 0) ? (k >= RNA_OFFSET ? k - RNA_OFFSET : -1)
                                      : (i == 1 ? rna_len + k : rna_len + T + (i - 2) * T + k);
                    if (fi >= 0) {
                        if (a.out_flat) a.out_flat[sys * a.L + fi] = v;
                        if (want_loss) {
                            double dlt = v - __ldg(tg + fi);
                            double ww = sg ? dlt / __ldg(sg + fi) : dlt;
                            e.ssr = fma(ww, ww, e.ssr);
                            e.sr += fabs(dlt);
                            e.sr2 = fma(dlt, dlt, e.sr2);
                        }
                    }
                    if (want_y) {
                        e.s1 += v;
                        e.s2 = fma(v, v, e.s2);
                        if (a.y_metric == 3) {
                            if (k > 0) { double dd = v - prev[i]; e.dyn = fma(dd, dd, e.dyn); }
                            prev[i] = v;
                        }
                    }
                }
            }
        };

        // initial step from the error-weighted time scale |y|/|f|
        dense_rhs<MODEL>(ns, n, p, y, v, lane);
        __syncwarp();
        double d0 = 0.0, d1 = 0.0;
        for (int i = lane; i < n; i += 32) {
            double sc = 1.0 / fma(a.rtol, fabs(y[i]), a.atol);
            d0 = fmax(d0, fabs(y[i]) * sc);
            d1 = fmax(d1, fabs(v[i]) * sc);
        }
        d0 = warp_max(d0);
        d1 = warp_max(d1);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        StepCtl ctl{h0, (float)h0, 1.0f, 0, 0};
        emit(0, y);

        // ---------------------------------------------------------------------- time loop
        while (kout < T && status == 0) {
            const double tout = a.t[kout];
            const double rem = tout - t;
            if (!(rem > 0.0)) { emit(kout, y); ++kout; continue; }
            double hh = ctl.h;
            bool land = false;
            if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
            else if (hh > 0.5 * rem) hh = 0.5 * rem;
            dense_fillW<MODEL>(ns, n, ld, p, hh * a.m.gamma, W, lane);
            dense_lu(n, ld, W, lane);

            // v_0 = h f(y); v_k = A^-1 v_{k-1}; y_new = y + sum MU_k v_k; err = sum EPS_k v_k
            dense_rhs<MODEL>(ns, n, p, y, v, lane);
            for (int i = lane; i < n; i += 32) v[i] *= hh;
            __syncwarp();
            dense_solve(n, ld, W, v, lane);
            for (int i = lane; i < n; i += 32) w[i] = fma(a.m.mu[0], v[i], y[i]);
            dense_solve(n, ld, W, v, lane);
            for (int i = lane; i < n; i += 32) { w[i] = fma(a.m.mu[1], v[i], w[i]); E[i] = a.m.eps[1] * v[i]; }
            dense_solve(n, ld, W, v, lane);
            for (int i = lane; i < n; i += 32) { w[i] = fma(a.m.mu[2], v[i], w[i]); E[i] = fma(a.m.eps[2], v[i], E[i]); }
            dense_solve(n, ld, W, v, lane);
            for (int i = lane; i < n; i += 32) { w[i] = fma(a.m.mu[3], v[i], w[i]); E[i] = fma(a.m.eps[3], v[i], E[i]); }
            dense_solve(n, ld, W, v, lane);
            for (int i = lane; i < n; i += 32) { w[i] = fma(a.m.mu[4], v[i], w[i]); E[i] = fma(a.m.eps[4], v[i], E[i]); }
            dense_solve(n, ld, W, v, lane);
            float err = 0.0f;
            bool bad = false;
            for (int i = lane; i < n; i += 32) {
                const double yn = fma(a.m.mu[5], v[i], w[i]);
                const double ei = fma(a.m.eps[5], v[i], E[i]);
                w[i] = yn;
                const float q = err_ratio(ei, y[i], yn, a.rtol, a.atol);
                bad |= !(q < 3.0e38f) || !(fabs(yn) < 1.0e300);
                err = fmaxf(err, q);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
            bad = __any_sync(0xffffffffu, bad);
            __syncwarp();

            if (bad) { status = 3; break; }
            if (err <= 1.0f) {
                ++nst;
                const double hprop = ctl.h;
                const double hnew = ctl_accept(ctl, hh, err, a.m.expo);
                ctl.h = (hh < hprop) ? fmax(hnew, fmin(hprop, 6.0 * hh)) : hnew;
                for (int i = lane; i < n; i += 32) y[i] = w[i];
                __syncwarp();
                if (land) { t = tout; emit(kout, y); ++kout; }
                else t += hh;
            } else {
                ++nrej;
                ctl.h = ctl_reject(ctl, hh, err, a.m.expo);
                if (ctl.h < 1e-14 * fmax(1.0, fabs(t))) status = 2;
            }
            if (status == 0 && kout < T && nst + nrej >= a.max_steps) status = 1;
        }
        for (int k = kout; k < T; ++k) emit(k, nullptr);       // failed system: NaN tail

        // ------------------------------------------------------------------------ finish
        if (lane == 0) {
            if (a.out_status) a.out_status[sys] = status;
            if (a.out_nsteps) a.out_nsteps[sys] = nst;
            if (a.out_nrej) a.out_nrej[sys] = nrej;
        }
        if (want_loss) {
            double ssr = e.ssr;
            if (a.lam != 0.0) {
                for (int i = lane; i < P; i += 32) {
                    double th = a.params[sys * P + i];
                    double ww = a.lam / (double)P * th * th;
                    if (sg && a.sigma_len > a.L) ww /= __ldg(sg + a.L + i);
                    ssr = fma(ww, ww, ssr);
                }
            }
            ssr = warp_sum(ssr);
            const double sr = warp_sum(e.sr), sr2 = warp_sum(e.sr2), p2 = warp_sum(p2l);
            if (lane == 0) {
                if (a.out_ssr) a.out_ssr[sys] = ssr;
                if (a.out_score) {
                    const double invL = 1.0 / (double)a.L;
                    const double r1 = sr * invL, r2 = sr2 * invL * invL;      // sum r, sum r^2, r = |d|/L
                    const double mean_r2 = r2 * invL, mae = r1 * invL;
                    a.out_score[sys] = a.w_delta * r2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                       a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * sqrt(p2) / (double)P;
                }
            }
        }
        if (want_y) {
            const double s1 = warp_sum(e.s1), s2 = warp_sum(e.s2), dyn = warp_sum(e.dyn);
            if (lane == 0) {
                const double len = (double)(T * nobs), mean = s1 / len;
                double yv;
                switch (a.y_metric) {
                    case 0: yv = s1; break;
                    case 1: yv = mean; break;
                    case 2: yv = s2 / len - mean * mean; break;
                    case 3: yv = dyn; break;
                    default: yv = sqrt(s2); break;
                }
                a.out_Y[sys] = yv;
            }
        }
        __syncwarp();
    }
}

}  // namespace pk
