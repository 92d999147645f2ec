// Thread-per-system RODAS4 kernel for the small local models (distributive, successive).
//
// One lane integrates one system with its whole state (y, five stage vectors, the factorised
// W = I/(h*gamma) - J) in registers; the Jacobian structure is exploited analytically:
//   distributive (models/distmod.py:57-63): arrow matrix  -> O(n) elimination via the P pivot
//   successive   (models/succmod.py:33-90): tridiagonal in (P, site_1..site_ns) -> Thomas
// mRNA (row 0) is decoupled in both and eliminated first.
// Lanes pull systems from a global queue (warp-aggregated atomicAdd) as they finish, so a warp
// never idles on its slowest member: per-system adaptivity costs no SIMT efficiency except in
// the tail.  The epilogue (clip, flat layout, weighted residual / score_fit, Morris Y) is fused:
// it runs at the step that lands on each requested output time.
#pragma once
#include "pk_common.cuh"

namespace pk {

// ------------------------------------------------------------------------------------ models
template <int NS_>
struct DistModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = NS_ + 2;
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
    // F[0] = 1/(g+B); F[1] = 1/schur pivot of P; F[2+i] = 1/(g+k_i)
    __device__ __forceinline__ void factor(double g, double (&F)[NF]) const {
        F[0] = 1.0 / (g + Bm);
        double piv = g + kP;
#pragma unroll
        for (int i = 0; i < NS; ++i) { F[2 + i] = 1.0 / (g + k[i]); piv = fma(-S[i], F[2 + i], piv); }
        F[1] = 1.0 / piv;
    }
    __device__ __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) const {
        x[0] *= F[0];
        double s = fma(C, x[0], x[1]);
#pragma unroll
        for (int i = 0; i < NS; ++i) s = fma(x[2 + i], F[2 + i], s);
        x[1] = s * F[1];
#pragma unroll
        for (int i = 0; i < NS; ++i) x[2 + i] = fma(S[i], x[1], x[2 + i]) * F[2 + i];
    }
};

template <int NS_>
struct SuccModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = 2 * NS_ + 2;
    // d[0] = D + S_0 (protein), d[1+i] = 1 + Dr_i + S_{i+1} (site i; no S term for the last site)
    double A, Bm, C, S[NS], d[NS + 1];
    __device__ __forceinline__ void load(const double* p) {
        A = p[0]; Bm = p[1]; C = p[2];
#pragma unroll
        for (int i = 0; i < NS; ++i) S[i] = p[4 + i];
        d[0] = p[3] + S[0];
#pragma unroll
        for (int i = 0; i < NS; ++i) d[1 + i] = 1.0 + p[4 + NS + i] + (i < NS - 1 ? S[i + 1] : 0.0);
    }
    __device__ __forceinline__ void rhs(const double (&y)[N], double (&f)[N]) const {
        f[0] = fma(-Bm, y[0], A);
        f[1] = fma(C, y[0], fma(-d[0], y[1], y[2]));
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = fma(S[i], y[1 + i], -d[1 + i] * y[2 + i]);
            if (i < NS - 1) v += y[3 + i];
            f[2 + i] = v;
        }
    }
    // Tridiagonal (x_1..x_{n-1}): diag g+d[j], sub -S[j-1], super -1.  Thomas without pivoting
    // (column diagonally dominant for non-negative rates).  F[0] = 1/(g+B); F[1+j] = 1/pivot_j;
    // F[NS+2+j] unused slot kept for alignment of indices (only first NS+2 are pivots).
    __device__ __forceinline__ void factor(double g, double (&F)[NF]) const {
        F[0] = 1.0 / (g + Bm);
        double piv = g + d[0];
        F[1] = 1.0 / piv;
#pragma unroll
        for (int j = 1; j <= NS; ++j) {
            // eliminate sub-diagonal -S[j-1] with row j-1: l = -S[j-1]/piv_{j-1}; piv_j = diag_j - l*(-1)...
            double l = S[j - 1] * F[j];          // multiplier magnitude
            F[NS + 1 + j] = l;                   // store for the forward sweep
            piv = (g + d[j]) - l;                // diag_j - (S[j-1]/piv_{j-1}) * 1
            F[1 + j] = 1.0 / piv;
        }
    }
    __device__ __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) const {
        x[0] *= F[0];
        x[1] = fma(C, x[0], x[1]);
#pragma unroll
        for (int j = 1; j <= NS; ++j) x[1 + j] = fma(F[NS + 1 + j], x[j], x[1 + j]);   // forward
        x[1 + NS] *= F[1 + NS];
#pragma unroll
        for (int j = NS - 1; j >= 0; --j) x[1 + j] = (x[1 + j] + x[2 + j]) * F[1 + j];  // backward
    }
};

// ---------------------------------------------------------------------------------- epilogue
__device__ __forceinline__ void epi_loss_point(const LocalArgs& a, EpiAcc& e, const double* tg,
                                               const double* sg, int fi, double v, double invL) {
    double dlt = v - __ldg(tg + fi);
    double w = sg ? dlt / __ldg(sg + fi) : dlt;
    e.ssr = fma(w, w, e.ssr);
    double r = fabs(dlt) * invL;
    e.sr += r;
    e.sr2 = fma(r, r, e.sr2);
}

// ------------------------------------------------------------------------------------ kernel
template <class M>
__global__ void __launch_bounds__(128) local_tps_kernel(const LocalArgs a) {
    constexpr int N = M::N, NS = M::NS, NF = M::NF, P = M::P;
    using namespace rodas4;
    extern __shared__ double smem[];
    double* tgrid = smem;                                     // [T]
    double* prev = smem + a.T;                                // [N][blockDim] (dynamics metric only)
    for (int i = threadIdx.x; i < a.T; i += blockDim.x) tgrid[i] = a.t[i];
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int T = a.T;
    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    const double invL = 1.0 / (double)a.L;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;

    bool active = false, exhausted = false;
    long long sys = -1;
    M mdl;
    double y[N], inv0[N];
    double t = 0.0, p2 = 0.0;
    StepCtl ctl;
    int kout = 0, nst = 0, nrej = 0, status = 0;
    EpiAcc e;
    const double* tg = nullptr;
    const double* sg = nullptr;

    // write outputs of time index k for state vector v (already clipped / normalised)
    auto emit = [&](int k, const double (&v)[N]) {
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
            if (k >= RNA_OFFSET) epi_loss_point(a, e, tg, sg, k - RNA_OFFSET, v[0], invL);
            epi_loss_point(a, e, tg, sg, rna_len + k, v[1], invL);
#pragma unroll
            for (int i = 0; i < NS; ++i) epi_loss_point(a, e, tg, sg, rna_len + T + i * T + k, v[2 + i], invL);
        }
        if (want_y) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                e.s1 += v[i];
                e.s2 = fma(v[i], v[i], e.s2);
                if (a.y_metric == 3) {
                    double* pv = prev + i * blockDim.x + threadIdx.x;
                    if (k > 0) { double dd = v[i] - *pv; e.dyn = fma(dd, dd, e.dyn); }
                    *pv = v[i];
                }
            }
        }
    };
    auto emit_state = [&](int k) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            v[i] = fmax(y[i], 0.0);                       // np.clip(sol, 0, None)
            if (a.normalize) v[i] *= inv0[i];             // NORMALIZE_MODEL_OUTPUT
        }
        emit(k, v);
    };
    auto finish = [&]() {
        if (status != 0) {                                // failed system: NaN for what is missing
            double v[N];
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = qnan;
            for (int k = kout; k < T; ++k) emit(k, v);
        }
        if (a.out_status) a.out_status[sys] = status;
        if (a.out_nsteps) a.out_nsteps[sys] = nst;
        if (a.out_nrej) a.out_nrej[sys] = nrej;
        if (want_loss) {
            // regularisation rows of normest's model_func: lam/P * theta^2, target 0, sigma from tail
            const double* pr = a.params + (size_t)sys * P;
            double ssr = e.ssr;
            if (a.lam != 0.0) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double th = pr[i];
                    double w = a.lam / (double)P * th * th;
                    if (sg && a.sigma_len > a.L) w /= __ldg(sg + a.L + i);
                    ssr = fma(w, w, ssr);
                }
            }
            if (a.out_ssr) a.out_ssr[sys] = ssr;
            if (a.out_score) {
                double Ld = (double)a.L;
                double mse = e.sr2, mean_r2 = e.sr2 / Ld, mae = e.sr / Ld;
                double var = mean_r2 - mae * mae;
                double l2 = sqrt(p2) / (double)P;
                a.out_score[sys] = a.w_delta * mse + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                   a.w_gamma * var + a.w_mu * l2;
            }
        }
        if (want_y) {
            double len = (double)(T * N), yv;
            double mean = e.s1 / len;
            switch (a.y_metric) {
                case 0: yv = e.s1; break;
                case 1: yv = mean; break;
                case 2: yv = e.s2 / len - mean * mean; break;
                case 3: yv = e.dyn; break;
                default: yv = sqrt(e.s2); break;
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
                    p2 = 0.0;
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
                    for (int i = 0; i < N; ++i) { y[i] = y0[i]; inv0[i] = a.normalize ? 1.0 / y[i] : 1.0; }
                    int g = a.group ? a.group[sys] : 0;
                    tg = a.target ? a.target + (size_t)g * a.L : nullptr;
                    sg = a.sigma ? a.sigma + (size_t)g * a.sigma_len : nullptr;
                    e = EpiAcc{0, 0, 0, 0, 0, 0};
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
                    ctl = StepCtl{h0, h0, 1.0, 0, 0};
                    emit_state(0);
                    kout = 1;
                    if (T <= 1) finish();
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULL, !active)) break;
        if (!active) continue;

        // ------------------------------------------------------------------ one step attempt
        const double tout = tgrid[kout];
        const double rem = tout - t;
        if (!(rem > 0.0)) {            // repeated output time
            emit_state(kout);
            if (++kout >= T) finish();
            continue;
        }
        double hh = ctl.h;
        bool land = false;
        if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
        else if (hh > 0.5 * rem) hh = 0.5 * rem;
        const double ih = 1.0 / hh;
        const double g = ih * (1.0 / GAMMA);

        double F[NF];
        mdl.factor(g, F);
        double U1[N], U2[N], U3[N], U4[N], U5[N], w[N], E[N];
        mdl.rhs(y, U1);
        mdl.solve(F, U1);
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = fma(A21, U1[i], y[i]);
        mdl.rhs(w, U2);
#pragma unroll
        for (int i = 0; i < N; ++i) U2[i] = fma(C21 * ih, U1[i], U2[i]);
        mdl.solve(F, U2);
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = fma(A32, U2[i], fma(A31, U1[i], y[i]));
        mdl.rhs(w, U3);
#pragma unroll
        for (int i = 0; i < N; ++i) U3[i] = fma(C32 * ih, U2[i], fma(C31 * ih, U1[i], U3[i]));
        mdl.solve(F, U3);
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = fma(A43, U3[i], fma(A42, U2[i], fma(A41, U1[i], y[i])));
        mdl.rhs(w, U4);
#pragma unroll
        for (int i = 0; i < N; ++i)
            U4[i] = fma(C43 * ih, U3[i], fma(C42 * ih, U2[i], fma(C41 * ih, U1[i], U4[i])));
        mdl.solve(F, U4);
#pragma unroll
        for (int i = 0; i < N; ++i)
            w[i] = fma(A54, U4[i], fma(A53, U3[i], fma(A52, U2[i], fma(A51, U1[i], y[i]))));
        mdl.rhs(w, U5);
#pragma unroll
        for (int i = 0; i < N; ++i)
            U5[i] = fma(C54 * ih, U4[i], fma(C53 * ih, U3[i], fma(C52 * ih, U2[i], fma(C51 * ih, U1[i], U5[i]))));
        mdl.solve(F, U5);
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] += U5[i];
        mdl.rhs(w, E);
#pragma unroll
        for (int i = 0; i < N; ++i)
            E[i] = fma(C65 * ih, U5[i],
                       fma(C64 * ih, U4[i], fma(C63 * ih, U3[i], fma(C62 * ih, U2[i], fma(C61 * ih, U1[i], E[i])))));
        mdl.solve(F, E);
        double err = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            w[i] += E[i];
            double sc = fma(a.rtol, fmax(fabs(y[i]), fabs(w[i])), a.atol);
            err = fmax(err, fabs(E[i]) / sc);
        }

        if (!(err < 1.0e300)) {                     // NaN or inf
            status = 3;
            finish();
        } else if (err <= 1.0) {
            ++nst;
            double hprop = ctl.h;
            double hnew = ctl_accept(ctl, hh, err);
            ctl.h = (hh < hprop) ? fmax(hnew, fmin(hprop, 6.0 * hh)) : hnew;
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] = w[i];
            if (land) {
                t = tout;
                emit_state(kout);
                if (++kout >= T) finish();
            } else {
                t += hh;
            }
            if (active && nst + nrej >= a.max_steps) { status = 1; finish(); }
        } else {
            ++nrej;
            ctl.h = ctl_reject(ctl, hh, err);
            if (ctl.h < 1e-14 * fmax(1.0, fabs(t))) { status = 2; finish(); }
            else if (nst + nrej >= a.max_steps) { status = 1; finish(); }
        }
    }
}

}  // namespace pk
