// Thread-per-system RODAS4 kernel for the small local models (distributive, successive).
//
// One lane integrates one system with its whole working set (y, the Krylov vector v, y_new, err
// and the factorised I - h*gamma*M) in registers; the Jacobian structure is exploited analytically:
//   distributive (models/distmod.py:57-63): arrow matrix  -> Schur pivot on the protein row
//   successive   (models/succmod.py:33-90): tridiagonal in (P, site_1..site_ns) -> Thomas, with the
//                pivots obtained from the continuant recurrence so that they are independent
// mRNA (row 0) is decoupled in both and eliminated first.  All reciprocals of one factorisation come
// from ONE FP64 division (batch inversion by prefix products); the error ratio and the step-size
// controller run in FP32.  Rarely touched per-system state (loss / Y accumulators) lives in shared
// memory.  Lanes pull systems from a global queue (warp-aggregated atomicAdd) as they finish, so a
// warp never idles on its slowest member.  The epilogue (clip, flat layout, weighted residual /
// score_fit, Morris Y) is fused: it runs at the step that lands on each requested output time.
#pragma once
#include "pk_common.cuh"

namespace pk {

// a[i] <- 1/a[i] for i < M with one division (Montgomery's trick).
template <int M>
__device__ __forceinline__ void batch_invert(double (&a)[M]) {
    double pre[M];
    pre[0] = a[0];
#pragma unroll
    for (int i = 1; i < M; ++i) pre[i] = pre[i - 1] * a[i];
    double inv = 1.0 / pre[M - 1];
#pragma unroll
    for (int i = M - 1; i > 0; --i) {
        double ai = a[i];
        a[i] = inv * pre[i - 1];
        inv *= ai;
    }
    a[0] = inv;
}

// ------------------------------------------------------------------------------------ models
// Each model provides rhs(y) = M y + b, factor(c) of A = I - c M (c = h*gamma) and an in-place
// solve A x = r.
template <int NS_>
struct DistModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = 2 * NS_ + 4;
    double A, Bm, C, kP, S[NS], k[NS];
    __device__ __forceinline__ void load(const double* p) {
        A = p[0]; Bm = p[1]; C = p[2];
        double sS = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) { S[i] = p[4 + i]; sS += S[i]; k[i] = 1.0 + p[4 + NS + i]; }
        kP = p[3] + sS;
    }
    __device__ __forceinline__ void rhs(const double (&y)[N], double (&f)[N]) const {
        f[0] = fma(-Bm, y[0], A);
        double acc = fma(C, y[0], -kP * y[1]);
#pragma unroll
        for (int i = 0; i < NS; ++i) { acc += y[2 + i]; f[2 + i] = fma(S[i], y[1], -k[i] * y[2 + i]); }
        f[1] = acc;
    }
    // rows: (1+cB) x0 = r0 ; -cC x0 + (1+c kP) x1 - c sum x_{2+i} = r1 ; -c S_i x1 + q_i x_{2+i} = r_{2+i}
    // F = [1/q0, 1/pivot, cC, c, 1/q_i (NS), c S_i / q_i (NS)]
    __device__ __forceinline__ void factor(double c, double (&F)[NF]) const {
        double q[NS], pre[NS], suf[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) q[i] = fma(c, k[i], 1.0);
        pre[0] = 1.0;
#pragma unroll
        for (int i = 1; i < NS; ++i) pre[i] = pre[i - 1] * q[i - 1];
        suf[NS - 1] = 1.0;
#pragma unroll
        for (int i = NS - 2; i >= 0; --i) suf[i] = suf[i + 1] * q[i + 1];
        const double Q = pre[NS - 1] * q[NS - 1];
        double sq = 0.0;                                   // sum_i S_i prod_{j != i} q_j
#pragma unroll
        for (int i = 0; i < NS; ++i) { pre[i] *= suf[i]; sq = fma(S[i], pre[i], sq); }
        double inv[3] = {fma(c, Bm, 1.0), Q, fma(fma(c, kP, 1.0), Q, -(c * c) * sq)};   // q0, Q, pivot*Q
        batch_invert<3>(inv);
        F[0] = inv[0];
        F[1] = Q * inv[2];
        F[2] = c * C;
        F[3] = c;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double iq = pre[i] * inv[1];
            F[4 + i] = iq;
            F[4 + NS + i] = c * S[i] * iq;
        }
    }
    __device__ __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) const {
        x[0] *= F[0];
        double sz = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) { x[2 + i] *= F[4 + i]; sz += x[2 + i]; }
        x[1] = fma(F[3], sz, fma(F[2], x[0], x[1])) * F[1];
#pragma unroll
        for (int i = 0; i < NS; ++i) x[2 + i] = fma(F[4 + NS + i], x[1], x[2 + i]);
    }
};

template <int NS_>
struct SuccModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = 3 * NS_ + 3;
    // d[0] = D + S_0 (protein), d[1+i] = 1 + Dr_i + S_{i+1} (site i; no S term for the last site)
    double A, Bm, C, S[NS], d[NS + 1];
    __device__ __forceinline__ void load(const double* p) {
        A = p[0]; Bm = p[1]; C = p[2];
#pragma unroll
        for (int i = 0; i < NS; ++i) S[i] = p[4 + i];
        d[0] = p[3] + S[0];
#pragma unroll
        for (int i = 0; i < NS; ++i) d[1 + i] = 1.0 + p[4 + NS + i] + (i < NS - 1 ? S[i < NS - 1 ? i + 1 : i] : 0.0);
    }
    __device__ __forceinline__ void rhs(const double (&y)[N], double (&f)[N]) const {
        f[0] = fma(-Bm, y[0], A);
        f[1] = fma(C, y[0], fma(-d[0], y[1], y[2]));
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = fma(S[i], y[1 + i], -d[1 + i] * y[2 + i]);
            if (i < NS - 1) v += y[i < NS - 1 ? 3 + i : 2 + i];
            f[2 + i] = v;
        }
    }
    // Tridiagonal block over x_1..x_{1+NS}: diag a_j = 1 + c d_j, sub -c S_{j-1}, super -c.
    // Continuants theta_j = a_j theta_{j-1} - (c S_{j-1}) c theta_{j-2} give the Thomas pivots
    // p_j = theta_j / theta_{j-1} without a sequential chain of divisions (no pivoting needed: the
    // block is strictly column diagonally dominant for non-negative rates).
    // F = [1/q0, cC, 1/p_j (NS+1), l_j = c S_{j-1}/p_{j-1} (NS), c/p_j (NS)]
    __device__ __forceinline__ void factor(double c, double (&F)[NF]) const {
        double th[NS + 2];                       // th[0] = q0, th[1+j] = theta_j
        th[0] = fma(c, Bm, 1.0);
        th[1] = fma(c, d[0], 1.0);
        double cs[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) cs[j] = c * S[j];
        th[2] = fma(fma(c, d[1], 1.0), th[1], -(cs[0] * c));
#pragma unroll
        for (int j = 2; j <= NS; ++j) th[1 + j] = fma(fma(c, d[j], 1.0), th[j], -(cs[j - 1] * c) * th[j - 1]);
        double inv[NS + 2];
#pragma unroll
        for (int j = 0; j < NS + 2; ++j) inv[j] = th[j];
        batch_invert<NS + 2>(inv);
        F[0] = inv[0];
        F[1] = c * C;
        F[2] = inv[1];                                          // 1/p_0 = 1/theta_0
#pragma unroll
        for (int j = 1; j <= NS; ++j) F[2 + j] = th[j] * inv[1 + j];   // theta_{j-1} / theta_j
#pragma unroll
        for (int j = 1; j <= NS; ++j) F[2 + NS + j] = cs[j - 1] * F[1 + j];       // l_j
#pragma unroll
        for (int j = 0; j < NS; ++j) F[3 + 2 * NS + j] = c * F[2 + j];            // c / p_j
    }
    __device__ __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) const {
        x[0] *= F[0];
        x[1] = fma(F[1], x[0], x[1]);
#pragma unroll
        for (int j = 1; j <= NS; ++j) x[1 + j] = fma(F[2 + NS + j], x[j], x[1 + j]);        // forward
        x[1 + NS] *= F[2 + NS];
#pragma unroll
        for (int j = NS - 1; j >= 0; --j) x[1 + j] = fma(F[3 + 2 * NS + j], x[2 + j], x[1 + j] * F[2 + j]);
    }
};

// ------------------------------------------------------------------------------------ kernel
constexpr int TPS_BLOCK = 128;
constexpr int TPS_COLD = 7;      // per-lane shared-memory doubles: ssr, sr, sr2, s1, s2, dyn, |p|^2

template <class M, int MIN_BLOCKS>
__global__ void __launch_bounds__(TPS_BLOCK, MIN_BLOCKS) local_tps_kernel(const LocalArgs a) {
    constexpr int N = M::N, NS = M::NS, NF = M::NF, P = M::P;
    using namespace rodas4;
    extern __shared__ double smem[];
    double* tgrid = smem;                                     // [T]
    double* cold = smem + a.T + threadIdx.x;                  // [TPS_COLD][TPS_BLOCK]
    double* prev = cold + TPS_COLD * TPS_BLOCK;               // [N][TPS_BLOCK] (dynamics metric only)
    for (int i = threadIdx.x; i < a.T; i += TPS_BLOCK) tgrid[i] = a.t[i];
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int T = a.T;
    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;

    bool active = false, exhausted = false;
    long long sys = -1;
    M mdl;
    double y[N];
    double t = 0.0;
    StepCtl ctl;
    int kout = 0, nst = 0, nrej = 0, status = 0;

    // outputs of time index k for state y (or NaN for a failed system)
    auto emit = [&](int k, bool failed) {
        double v[N];
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = failed ? qnan : fmax(y[i], 0.0);      // np.clip(sol, 0, None)
        if (a.normalize) {                                                       // NORMALIZE_MODEL_OUTPUT
            const double* y0 = a.y0 + (a.y0_stride ? (size_t)sys * a.y0_stride : 0);
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] *= 1.0 / y0[i];       // sol *= 1/init (distmod.py:118-122)
        }
        if (a.out_sol) {
            double* o = a.out_sol + ((size_t)sys * T + k) * N;
#pragma unroll
            for (int i = 0; i < N; ++i) o[i] = v[i];
        }
        if (a.out_flat) {
            double* o = a.out_flat + (size_t)sys * a.L;
            if (k >= RNA_OFFSET) o[k - RNA_OFFSET] = v[0];
            o[rna_len + k] = v[1];
#pragma unroll
            for (int i = 0; i < NS; ++i) o[rna_len + T + i * T + k] = v[2 + i];
        }
        if (want_loss) {
            const int g = a.group ? a.group[sys] : 0;
            const double* tg = a.target + (size_t)g * a.L;
            const double* sg = a.sigma ? a.sigma + (size_t)g * a.sigma_len : nullptr;
            double ssr = cold[0 * TPS_BLOCK], sr = cold[1 * TPS_BLOCK], sr2 = cold[2 * TPS_BLOCK];
            auto point = [&](int fi, double val) {
                double dlt = val - __ldg(tg + fi);
                double w = sg ? dlt / __ldg(sg + fi) : dlt;
                ssr = fma(w, w, ssr);
                sr += fabs(dlt);
                sr2 = fma(dlt, dlt, sr2);
            };
            if (k >= RNA_OFFSET) point(k - RNA_OFFSET, v[0]);
            point(rna_len + k, v[1]);
#pragma unroll
            for (int i = 0; i < NS; ++i) point(rna_len + T + i * T + k, v[2 + i]);
            cold[0 * TPS_BLOCK] = ssr; cold[1 * TPS_BLOCK] = sr; cold[2 * TPS_BLOCK] = sr2;
        }
        if (want_y) {
            double s1 = cold[3 * TPS_BLOCK], s2 = cold[4 * TPS_BLOCK], dyn = cold[5 * TPS_BLOCK];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s1 += v[i];
                s2 = fma(v[i], v[i], s2);
                if (a.y_metric == 3) {
                    if (k > 0) { double dd = v[i] - prev[i * TPS_BLOCK]; dyn = fma(dd, dd, dyn); }
                    prev[i * TPS_BLOCK] = v[i];
                }
            }
            cold[3 * TPS_BLOCK] = s1; cold[4 * TPS_BLOCK] = s2; cold[5 * TPS_BLOCK] = dyn;
        }
    };
    // final scalars of a system; called from exactly one site to keep the kernel inside the
    // instruction cache (every extra inlined copy of emit/finish costs ~10 KB of SASS)
    auto finish = [&]() {
        if (a.out_status) a.out_status[sys] = status;
        if (a.out_nsteps) a.out_nsteps[sys] = nst;
        if (a.out_nrej) a.out_nrej[sys] = nrej;
        if (want_loss) {
            double ssr = cold[0 * TPS_BLOCK];
            if (a.lam != 0.0) {
                // regularisation rows of normest's model_func: lam/P * theta^2, target 0
                const double* pr = a.params + (size_t)sys * P;
                const int g = a.group ? a.group[sys] : 0;
                const double* sg = (a.sigma && a.sigma_len > a.L) ? a.sigma + (size_t)g * a.sigma_len + a.L : nullptr;
#pragma unroll 1
                for (int i = 0; i < P; ++i) {
                    double th = pr[i];
                    double w = a.lam / (double)P * th * th;
                    if (sg) w /= __ldg(sg + i);
                    ssr = fma(w, w, ssr);
                }
            }
            if (a.out_ssr) a.out_ssr[sys] = ssr;
            if (a.out_score) {
                // score_fit (config/config.py:176-226) with r = |target - pred| / L
                const double Ld = (double)a.L, invL = 1.0 / Ld;
                const double sr = cold[1 * TPS_BLOCK] * invL, sr2 = cold[2 * TPS_BLOCK] * invL * invL;
                const double mean_r2 = sr2 * invL, mae = sr * invL;
                const double l2 = sqrt(cold[6 * TPS_BLOCK]) / (double)P;
                a.out_score[sys] = a.w_delta * sr2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                   a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * l2;
            }
        }
        if (want_y) {
            const double s1 = cold[3 * TPS_BLOCK], s2 = cold[4 * TPS_BLOCK];
            const double len = (double)(T * N), mean = s1 / len;
            double yv;
            switch (a.y_metric) {
                case 0: yv = s1; break;
                case 1: yv = mean; break;
                case 2: yv = s2 / len - mean * mean; break;
                case 3: yv = cold[5 * TPS_BLOCK]; break;
                default: yv = sqrt(s2); break;
            }
            a.out_Y[sys] = yv;
        }
        active = false;
    };

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        unsigned need = __ballot_sync(FULL, !active && !exhausted);
        if (need) {
            int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!active && !exhausted) {
                long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (idx < a.B) {
                    sys = idx;
                    active = true;
                    const double* pr = a.params + (size_t)sys * P;
                    double pv[P];
                    double p2 = 0.0;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        double v = pr[i];
                        if (a.log_params) v = exp(v);
                        pv[i] = v;
                        p2 = fma(v, v, p2);
                    }
                    mdl.load(pv);
                    const double* y0 = a.y0 + (a.y0_stride ? (size_t)sys * a.y0_stride : 0);
#pragma unroll
                    for (int i = 0; i < N; ++i) y[i] = y0[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) cold[i * TPS_BLOCK] = 0.0;
                    cold[6 * TPS_BLOCK] = p2;
                    t = tgrid[0];
                    nst = nrej = status = 0;
                    // initial step: 1% of the time scale |y|/|f| in the error-weighted norm
                    double f0[N];
                    mdl.rhs(y, f0);
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        double sc = 1.0 / fma(a.rtol, fabs(y[i]), a.atol);
                        d0 = fmax(d0, fabs(y[i]) * sc);
                        d1 = fmax(d1, fabs(f0[i]) * sc);
                    }
                    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                    ctl = StepCtl{h0, (float)h0, 1.0f, 0, 0};
                    kout = 0;                 // output index 0 (the initial state) is emitted below
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULL, !active)) break;
        if (!active) continue;

        // ------------------------------------------------------------- one step attempt, or an
        // emission without a step (initial state, repeated output time, NaN tail of a failed system)
        bool do_emit = false;
        const double tout = tgrid[kout];
        const double rem = tout - t;
        if (status != 0 || kout == 0 || !(rem > 0.0)) {
            do_emit = true;
        } else {
            double hh = ctl.h;
            bool land = false;
            if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
            else if (hh > 0.5 * rem) hh = 0.5 * rem;

            double F[NF];
            mdl.factor(hh * GAMMA, F);
            double v[N], yn[N], er[N];
            mdl.rhs(y, v);
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] *= hh;
            mdl.solve(F, v);
#pragma unroll
            for (int i = 0; i < N; ++i) yn[i] = fma(MU1, v[i], y[i]);
            mdl.solve(F, v);
#pragma unroll
            for (int i = 0; i < N; ++i) { yn[i] = fma(MU2, v[i], yn[i]); er[i] = EPS2 * v[i]; }
            mdl.solve(F, v);
#pragma unroll
            for (int i = 0; i < N; ++i) { yn[i] = fma(MU3, v[i], yn[i]); er[i] = fma(EPS3, v[i], er[i]); }
            mdl.solve(F, v);
#pragma unroll
            for (int i = 0; i < N; ++i) { yn[i] = fma(MU4, v[i], yn[i]); er[i] = fma(EPS4, v[i], er[i]); }
            mdl.solve(F, v);
#pragma unroll
            for (int i = 0; i < N; ++i) { yn[i] = fma(MU5, v[i], yn[i]); er[i] = fma(EPS5, v[i], er[i]); }
            mdl.solve(F, v);
            float err = 0.0f;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                yn[i] = fma(MU6, v[i], yn[i]);
                er[i] = fma(EPS6, v[i], er[i]);
                err = fmaxf(err, err_ratio(er[i], y[i], yn[i], a.rtol, a.atol));
            }
            double chk = 0.0;                      // NaN/inf anywhere in y_new poisons the sum
#pragma unroll
            for (int i = 0; i < N; ++i) chk += yn[i];

            if (!(fabs(chk) < 1.0e300) || !(err < 3.0e38f)) {
                status = 3;
            } else if (err <= 1.0f) {
                ++nst;
                const double hprop = ctl.h;
                const double hnew = ctl_accept(ctl, hh, err);
                ctl.h = (hh < hprop) ? fmax(hnew, fmin(hprop, 6.0 * hh)) : hnew;
#pragma unroll
                for (int i = 0; i < N; ++i) y[i] = yn[i];
                if (land) { t = tout; do_emit = true; }
                else t += hh;
            } else {
                ++nrej;
                ctl.h = ctl_reject(ctl, hh, err);
                if (ctl.h < 1e-14 * fmax(1.0, fabs(t))) status = 2;
            }
            if (status == 0 && !do_emit && nst + nrej >= a.max_steps) status = 1;
        }
        if (do_emit) {
            emit(kout, status != 0);
            if (++kout >= T) finish();
        }
    }
}

}  // namespace pk
