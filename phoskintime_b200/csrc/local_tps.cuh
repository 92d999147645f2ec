// Thread-per-system RODAS4 kernel for the small local models (distributive, successive).
//
// One lane integrates one system with its whole working set (y, the Krylov vector v, y_new, err
// and the factorised I - h*gamma*M) in registers; the Jacobian structure is exploited analytically:
//   distributive (models/distmod.py:57-63): arrow matrix  -> Schur pivot on the protein row
//   successive   (models/succmod.py:33-90): tridiagonal in (P, site_1..site_ns) -> Thomas, with the
//                pivots obtained from the continuant recurrence so that they are independent
// mRNA (row 0) is decoupled in both and eliminated first.  All reciprocals of one factorisation come
// from ONE FP64 division (batch inversion by prefix products); the error ratio and the step-size
// controller run in FP32.  Lanes pull systems from a global queue (warp-aggregated atomicAdd) as they finish, so a
// warp never idles on its slowest member.  The epilogue (clip, flat layout, weighted residual /
// score_fit, Morris Y) is fused and runs warp-cooperatively when a system completes.
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
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = 2 * NS_ + 4;
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
    // F = [1/q0, cC, 1/p_j (NS+1), l_j = c S_{j-1}/p_{j-1} (NS), c]
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
        F[3 + 2 * NS] = c;
    }
    __device__ __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) const {
        x[0] *= F[0];
        x[1] = fma(F[1], x[0], x[1]);
#pragma unroll
        for (int j = 1; j <= NS; ++j) x[1 + j] = fma(F[2 + NS + j], x[j], x[1 + j]);        // forward
        x[1 + NS] *= F[2 + NS];
#pragma unroll
        for (int j = NS - 1; j >= 0; --j) x[1 + j] = fma(F[3 + 2 * NS], x[2 + j], x[1 + j]) * F[2 + j];   // (x_j + c x_{j+1}) / p_j
    }
};

// ------------------------------------------------------------------------------------ kernel
constexpr int TPS_BLOCK = 128;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Step loop: every lane owns one system.  A lane that lands on an output time only stores its raw
// state (N doubles) into its private trajectory slot `traj` (one [T][N] slot per resident lane,
// ~45 MB in total, L2-resident and reused for every system the lane integrates).  When a lane has
// produced its last output the WHOLE WARP runs the epilogue for it: 32 lanes walk the T*N stored
// values coalesced, clip / normalise, write sol and flat rows coalesced, and reduce the weighted
// residual, score_fit and Morris-Y sums with shuffles.  The divergent part of the loop is thus a
// handful of stores; the expensive epilogue runs at full warp width once per system.
template <class M, int MIN_BLOCKS>
__global__ void __launch_bounds__(TPS_BLOCK, MIN_BLOCKS) local_tps_kernel(const LocalArgs a) {
    constexpr int N = M::N, NF = M::NF, P = M::P;
    extern __shared__ double smem[];
    double* tgrid = smem;                                     // [T]
    short* fmap = (short*)(smem + a.T);                       // [L]: trajectory index (k*N + i) of every flat entry
    for (int i = threadIdx.x; i < a.T; i += TPS_BLOCK) tgrid[i] = a.t[i];
    {
        const int rl = a.T > RNA_OFFSET ? a.T - RNA_OFFSET : 0;
        for (int fi = threadIdx.x; fi < a.L; fi += TPS_BLOCK) {
            int k, i;
            if (fi < rl) { k = fi + RNA_OFFSET; i = 0; }
            else if (fi < rl + a.T) { k = fi - rl; i = 1; }
            else { const int j = fi - rl - a.T; i = 2 + j / a.T; k = j - (i - 2) * a.T; }
            fmap[fi] = (short)(k * N + i);
        }
    }
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int T = a.T;
    const int TN = T * N;
    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;
    double* const warp_traj = a.traj + ((size_t)blockIdx.x * TPS_BLOCK + (threadIdx.x & ~31)) * TN;
    double* const my_traj = warp_traj + (size_t)lane * TN;

    bool active = false, exhausted = false;
    long long sys = -1;
    M mdl;
    double y[N];
    double t = 0.0;
    StepCtl ctl;
    int kout = 0, nst = 0, nrej = 0, status = 0;
    double p2own = 0.0;

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
#pragma unroll
                    for (int i = 0; i < P; ++i) pv[i] = pr[i];
                    if (a.log_params) {
#pragma unroll 1
                        for (int i = 0; i < P; ++i) pv[i] = exp(pv[i]);
                    }
                    mdl.load(pv);
                    p2own = 0.0;                      // |physical params|^2 for score_fit's l2 term
#pragma unroll
                    for (int i = 0; i < P; ++i) p2own = fma(pv[i], pv[i], p2own);
                    const double* y0 = a.y0 + (a.y0_stride ? (size_t)sys * a.y0_stride : 0);
#pragma unroll
                    for (int i = 0; i < N; ++i) { y[i] = y0[i]; my_traj[i] = y[i]; }
                    t = tgrid[0];
                    nst = nrej = status = 0;
                    kout = 1;
                    // initial step: 1% of the time scale |y|/|f| in the error-weighted norm
                    double f0[N];
                    mdl.rhs(y, f0);
                    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        float sc = (float)fma(a.rtol, fabs(y[i]), a.atol);
                        d0 = fmaxf(d0, __fdividef((float)fabs(y[i]), sc));
                        d1 = fmaxf(d1, __fdividef((float)fabs(f0[i]), sc));
                    }
                    double h0 = (d0 < 1e-5f || d1 < 1e-5f || !(d1 < 3.0e38f)) ? 1e-6 : 0.01 * (double)__fdividef(d0, d1);
                    ctl = StepCtl{h0, (float)h0, 1.0f, 0, 0};
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULL, !active)) break;

        // ------------------------------------------------------------------ one step attempt
        bool finished = false;
        if (active) {
            if (kout < T) {
                const double tout = tgrid[kout];
                const double rem = tout - t;
                bool store = false;
                if (!(rem > 0.0)) {
                    store = true;                                  // repeated output time
                } else {
                    double hh = ctl.h;
                    bool land = false;
                    if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
                    else if (hh > 0.5 * rem) hh = 0.5 * rem;

                    double F[NF];
                    mdl.factor(hh * a.m.gamma, F);
                    double v[N], yn[N], er[N];
                    mdl.rhs(y, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) v[i] *= hh;
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) yn[i] = fma(a.m.mu[0], v[i], y[i]);
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[1], v[i], yn[i]); er[i] = a.m.eps[1] * v[i]; }
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[2], v[i], yn[i]); er[i] = fma(a.m.eps[2], v[i], er[i]); }
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[3], v[i], yn[i]); er[i] = fma(a.m.eps[3], v[i], er[i]); }
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[4], v[i], yn[i]); er[i] = fma(a.m.eps[4], v[i], er[i]); }
                    mdl.solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[5], v[i], yn[i]); er[i] = fma(a.m.eps[5], v[i], er[i]); }
                    if (a.m.nsol > 6) {                    // ROS6L: seventh solve (uniform over the grid)
                        mdl.solve(F, v);
#pragma unroll
                        for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[6], v[i], yn[i]); er[i] = fma(a.m.eps[6], v[i], er[i]); }
                    }
                    float err = 0.0f;
                    double chk = 0.0;                      // NaN/inf anywhere in y_new poisons the sum
                    const float rtolf = (float)a.rtol, floorf_ = (float)a.rtol_floor, kapf = (float)a.kappa, atolf = (float)a.atol;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        err = fmaxf(err, err_ratio_inc(er[i], y[i], yn[i], rtolf, floorf_, kapf, atolf));
                        chk += yn[i];
                    }
                    if (!(fabs(chk) < 3.0e38) || !(err < 3.0e38f)) {     // NaN/inf, or beyond the FP32 range of the error scale
                        status = 3;
                    } else if (err <= 1.0f) {
                        ++nst;
                        const double hprop = ctl.h;
                        const double hnew = ctl_accept(ctl, hh, err, a.m.expo);
                        ctl.h = (hh < hprop) ? fmax(hnew, fmin(hprop, 6.0 * hh)) : hnew;
#pragma unroll
                        for (int i = 0; i < N; ++i) y[i] = yn[i];
                        if (land) { t = tout; store = true; }
                        else t += hh;
                    } else {
                        ++nrej;
                        ctl.h = ctl_reject(ctl, hh, err, a.m.expo);
                        if (ctl.h < 1e-14 * fmax(1.0, fabs(t))) status = 2;
                    }
                    if (status == 0 && !store && nst + nrej >= a.max_steps) status = 1;
                }
                if (store) {
                    double* o = my_traj + kout * N;
#pragma unroll
                    for (int i = 0; i < N; ++i) o[i] = y[i];
                    ++kout;
                }
            }
            finished = (kout >= T) || (status != 0);
        }

        // ------------------------------------------ warp-cooperative epilogue of finished systems
        unsigned fin = __ballot_sync(FULL, finished);
        while (fin) {
            const int f = __ffs(fin) - 1;
            fin &= fin - 1;
            const size_t fsys = (size_t)__shfl_sync(FULL, sys, f);
            const int fstatus = __shfl_sync(FULL, status, f);
            const int fvalid = __shfl_sync(FULL, kout, f);          // outputs 0..fvalid-1 were produced
            const int fnst = __shfl_sync(FULL, nst, f), fnrej = __shfl_sync(FULL, nrej, f);
            __syncwarp();                                           // the owner's trajectory stores are visible
            const double* tr = warp_traj + (size_t)f * TN;
            const double* y0 = a.y0 + (a.y0_stride ? fsys * (size_t)a.y0_stride : 0);
            const int g = (want_loss && a.group) ? a.group[fsys] : 0;
            const double* tg = want_loss ? a.target + (size_t)g * a.L : nullptr;
            const double* sg = (want_loss && a.sigma) ? a.sigma + (size_t)g * a.sigma_len : nullptr;
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            double ssr = 0.0, sr = 0.0, sr2 = 0.0, s1 = 0.0, s2 = 0.0, dyn = 0.0;
            // Fast path (loss and/or flat only, successful system): walk the L flat entries through the index table —
            // no (k, i) arithmetic, no per-element validity test, 3 instead of 4 sweeps for 5 sites.
            const bool fast = fstatus == 0 && !a.out_sol && !want_y && !a.normalize;
            if (fast) {
                for (int base = lane; base < a.L; base += 128) {
                    // four sweeps' loads are issued back to back (the trajectory slot lives in L2: one latency, not four)
                    double raw[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int fi = base + 32 * u;
                        raw[u] = tr[fi < a.L ? fmap[fi] : 0];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int fi = base + 32 * u;
                        if (fi < a.L) {
                            const double v = fmax(raw[u], 0.0);                   // np.clip(sol, 0, None)
                            if (a.out_flat) a.out_flat[fsys * a.L + fi] = v;
                            if (want_loss) {
                                const double dlt = v - __ldg(tg + fi);
                                const double w = sg ? dlt / __ldg(sg + fi) : dlt;
                                ssr = fma(w, w, ssr);
                                sr += fabs(dlt);
                                sr2 = fma(dlt, dlt, sr2);
                            }
                        }
                    }
                }
            } else
            for (int idx = lane; idx < TN; idx += 32) {
                const int k = idx / N, i = idx - k * N;
                double v = (k < fvalid) ? fmax(tr[idx], 0.0) : qnan;            // np.clip(sol, 0, None)
                const double nrm = a.normalize ? 1.0 / y0[i] : 1.0;             // NORMALIZE_MODEL_OUTPUT
                v *= nrm;
                if (a.out_sol) a.out_sol[fsys * TN + idx] = v;
                const int fi = (i == 0) ? (k >= RNA_OFFSET ? k - RNA_OFFSET : -1)
                                        : (i == 1 ? rna_len + k : rna_len + T + (i - 2) * T + k);
                if (fi >= 0) {
                    if (a.out_flat) a.out_flat[fsys * a.L + fi] = v;
                    if (want_loss) {
                        const double dlt = v - __ldg(tg + fi);
                        const double w = sg ? dlt / __ldg(sg + fi) : dlt;
                        ssr = fma(w, w, ssr);
                        sr += fabs(dlt);
                        sr2 = fma(dlt, dlt, sr2);
                    }
                }
                if (want_y) {
                    s1 += v;
                    s2 = fma(v, v, s2);
                    if (a.y_metric == 3 && k > 0) {
                        const double pv = (k - 1 < fvalid) ? fmax(tr[idx - N], 0.0) * nrm : qnan;
                        dyn = fma(v - pv, v - pv, dyn);
                    }
                }
            }
            const double p2 = __shfl_sync(FULL, p2own, f);           // computed by the owner when it loaded the parameters
            if (want_loss) {
                if (a.lam != 0.0) {
                    // the lam/P*theta^2 rows of normest's model_func (paramest/normest.py:403-423)
                    const double* sgr = (sg && a.sigma_len > a.L) ? sg + a.L : nullptr;
                    for (int i = lane; i < P; i += 32) {
                        const double th = a.params[fsys * P + i];
                        double w = a.lam / (double)P * th * th;
                        if (sgr) w /= __ldg(sgr + i);
                        ssr = fma(w, w, ssr);
                    }
                }
                ssr = warp_sum_d(ssr); sr = warp_sum_d(sr); sr2 = warp_sum_d(sr2);
            }
            if (want_y) { s1 = warp_sum_d(s1); s2 = warp_sum_d(s2); dyn = warp_sum_d(dyn); }
            if (lane == 0) {
                if (a.out_status) a.out_status[fsys] = fstatus;
                if (a.out_nsteps) a.out_nsteps[fsys] = fnst;
                if (a.out_nrej) a.out_nrej[fsys] = fnrej;
                if (a.out_ssr) a.out_ssr[fsys] = ssr;
                if (a.out_score) {
                    // score_fit (config/config.py:176-226) with r = |target - pred| / L
                    const double invL = 1.0 / (double)a.L;
                    const double r1 = sr * invL, r2 = sr2 * invL * invL;
                    const double mean_r2 = r2 * invL, mae = r1 * invL;
                    a.out_score[fsys] = a.w_delta * r2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                        a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * sqrt(p2) / (double)P;
                }
                if (want_y) {
                    const double len = (double)TN, mean = s1 / len;
                    double yv;
                    switch (a.y_metric) {
                        case 0: yv = s1; break;
                        case 1: yv = mean; break;
                        case 2: yv = s2 / len - mean * mean; break;
                        case 3: yv = dyn; break;
                        default: yv = sqrt(s2); break;
                    }
                    a.out_Y[fsys] = yv;
                }
            }
        }
        if (finished) active = false;
    }
}

}  // namespace pk
