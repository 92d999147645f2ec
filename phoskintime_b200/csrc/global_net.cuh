// CTA-per-system kernel for the global coupled kinase-TF-protein network
// (reference: global_model/jacspeedup.py:148-375 RHS wrappers, global_model/models.py:27-306 block
// kinetics, simulate.py:34-80 integration, lossfn.py:113-246 loss, simulate.py:105-182 +
// sensitivity.py:106-140 Morris scalar; models 0 distributive, 1 sequential, 4 saturating).
//
// Structure of the problem (what the design exploits):
//   * protein i owns the block [mRNA, P0, site_1..site_ns]; phosphorylation rates S = W.(Kmat[:,bucket]*c_k)
//     are piecewise constant in time (13 buckets) and do not depend on the state;
//   * the ONLY coupling between blocks is mRNA synthesis: dR_i/dt = synth_i(TF_in_i) - B_i R_i with
//     TF_in_i a sparse combination of the total protein p_j = P0_j + sum(sites_j) of its regulators.
//   So  J = J_blk + E_R G E_p  with J_blk block diagonal (arrow for models 0/4, tridiagonal chain for
//   model 1, mRNA row decoupled inside the block), G = d synth / d p  (N x N, TF's sparsity) and
//   E_p the "total protein" selector.  (I - cJ) x = b is solved EXACTLY by
//       x = A^-1 b + c (G z) . w,      w = A^-1 e_R,  A = I - c J_blk  (tree elimination per protein)
//       (I - c diag(m) G) z = z0,      z0_i = 1^T [A^-1 b]_protein i,   m_i = 1^T w_i
//   i.e. one dense LU of size |Q| x |Q| (Q = non-driven proteins that regulate someone, ~N) instead
//   of state_dim x state_dim (~4-5 N): ~70x fewer flops than the reference's dense Jacobian route.
//
// Integrator: staged RODAS4 (Hairer & Wanner, 6 stages, order 4(3), stiffly accurate, L-stable) with
// the analytic Jacobian above, one factorisation per step; steps land exactly on every output time
// and on every kinase-bucket boundary (the RHS is discontinuous there).
#pragma once
#include "pk_common.cuh"

namespace pk {

struct GlobalTopoDev {
    int model, N, K, nb, n, S, nQ, pad;
    const int *offset_y, *offset_s, *n_sites;
    const int *W_indptr, *W_indices;
    const double* W_data;
    const int *TF_indptr, *TF_indices;
    const double* TF_data;
    const double *kin_grid, *kin_Kmat, *tf_deg;
    const int* driver_map;
    const int *qlist, *qpos;          // regulator set Q and its inverse map (-1 outside Q)
    // loss tables (lossfn.py:113-121)
    int n_prot, n_rna, n_pho, prot_base, rna_base, pho_base;
    const int *p_prot, *t_prot, *p_rna, *t_rna, *p_pho, *s_pho, *t_pho;
    const double *obs_prot, *w_prot, *obs_rna, *w_rna, *obs_pho, *w_pho;
    const double* defaults;           // packed physical prior centre [P] or nullptr
    double norm[3];                   // 1/max(1e-6, sum w) per modality (optproblem.py:83-85)
};

struct GlobalSmem {                   // offsets in doubles
    int par, Kt, Sall, y, arg, U, w, facA, mult, clo, pvec, g, m, z, Sc, idiag, red, perm, total;
    int ld;
};

struct GlobalArgs {
    GlobalTopoDev tp;
    GlobalSmem sm;
    long long B;
    int T, P, theta_mode, n_stops;
    const double* params;
    const double* y0;
    long long y0_stride;
    const double* stop_t;             // [n_stops] union of t_eval and interior kinase-grid points
    const int* stop_out;              // [n_stops] output index or -1
    const int* stop_bucket;           // [n_stops] kinase bucket of the interval that STARTS here
    double rtol, atol;
    int max_steps, loss_mode, metric;
    int n_mt_prot, n_mt_rna, n_mt_pho, mb_prot, mb_rna, mb_pho;
    const int *mt_prot, *mt_rna, *mt_pho;
    double lam[3], lam_prior;
    double *out_Y, *out_loss, *out_F, *out_metric;
    int *out_status, *out_nsteps, *out_nrej;
    double* traj;                     // [grid][T][n] scratch when out_Y is not requested
    unsigned long long* counter;
};

constexpr int GLOBAL_BLOCK = 256;
constexpr int GLOBAL_WARPS = GLOBAL_BLOCK / 32;

// RODAS4 (Hairer & Wanner II, RODAS METH=1), gamma = 1/4
__constant__ double G_A[5][4] = {
    {0.1544000000000000e+01, 0, 0, 0},
    {0.9466785280815826e+00, 0.2557011698983284e+00, 0, 0},
    {0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00}};   // stage 6 adds U5
__constant__ double G_C[5][5] = {
    {-0.5668800000000000e+01, 0, 0, 0, 0},
    {-0.2430093356833875e+01, -0.2063599157091915e+00, 0, 0, 0},
    {-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02, 0, 0},
    {0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0},
    {0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
     -0.6058818238834054e+01}};
constexpr double G_GAMMA = 0.25;

__device__ __forceinline__ double softplus_d(double x) {       // global_model/utils.py:228-253
    return x > 20.0 ? x : log1p(exp(x));
}

__device__ __forceinline__ double block_max_f(float v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int w = 1; w < GLOBAL_WARPS; ++w) r = fmax(r, red[w]);
    return r;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int w = 1; w < GLOBAL_WARPS; ++w) r += red[w];
    return r;
}

// mRNA synthesis rate and its derivative w.r.t. the raw TF input v = (TF.p)_i / tf_deg_i.
// models 0/1: the wrapper squashes once (jacspeedup.py:225-228), the kernel again (models.py:52);
// model 4: only the kernel's squash (jacspeedup.py:371-373).  1e-6 in the denominator (models.py:58).
__device__ __forceinline__ void synth_rate(int model, double v, double Ai, double tfs, double& synth, double& dsdv) {
    double u_raw = v, du_raw = 1.0;
    if (model != 4) {
        const double q = 1.0 / (1.0 + fabs(v));
        u_raw = v * q;
        du_raw = q * q;
    }
    const double q2 = 1.0 / (1.0 + fabs(u_raw));
    const double u = u_raw * q2;
    double ds;
    if (u >= 0.0) {
        const double d = 1.0 / (1.0 + u + 1e-6);
        synth = Ai * (1.0 + tfs * u * d);
        ds = Ai * tfs * (1.0 + 1e-6) * d * d;
    } else {
        const double d = 1.0 / (1.0 + tfs * fabs(u));
        synth = Ai * d;
        ds = Ai * tfs * d * d;
    }
    dsdv = ds * q2 * q2 * du_raw;
}

// lossfn.py:28-110 residual atoms as dispatched at lossfn.py:150-246
__device__ __forceinline__ double loss_atom(int mode, double diff, double obs, double pred) {
    switch (mode) {
        case 0: return diff * diff;
        case 1: { const double a = fabs(diff); return a <= 0.5 ? 0.5 * diff * diff : 0.5 * (a - 0.25); }
        case 2: { const double x = (log(diff + 1e-9) - log(obs + 1e-9)) / 0.5; return 0.25 * (sqrt(1.0 + x * x) - 1.0); }
        case 3: { const double s = fabs(diff); return s > 20.0 ? s - 0.69314718056 : log(cosh(diff)); }
        case 4: return log(1.0 + diff * diff);
        case 5: return diff * diff / (fabs(pred) + 1e-6);
        case 6: return diff * diff / (diff * diff + 1.0);
        default: return sqrt(diff * diff + 1e-6) - 1e-3;
    }
}

// (loss_p, loss_r, loss_ph) of lossfn.py:113-246 from one trajectory Y[T][n] (whole CTA cooperates):
// predicted fold change = max(x_t, 1e-9) / max(x_base, 1e-9), x = total protein / mRNA / site.
__device__ __forceinline__ void loss_sums(const GlobalTopoDev& tp, const double* traj, int n, int mode, double* red,
                                          double& lp, double& lr, double& lph) {
    lp = 0.0; lr = 0.0; lph = 0.0;
    for (int k = threadIdx.x; k < tp.n_prot; k += GLOBAL_BLOCK) {
        const int i = tp.p_prot[k], st = tp.offset_y[i], ns = tp.n_sites[i];
        const double* rt = traj + (size_t)tp.t_prot[k] * n + st + 1;
        const double* rb = traj + (size_t)tp.prot_base * n + st + 1;
        double a1 = 0.0, b1 = 0.0;
        for (int j = 0; j <= ns; ++j) { a1 += rt[j]; b1 += rb[j]; }
        const double pred = fmax(a1, 1e-9) / fmax(b1, 1e-9), obs = tp.obs_prot[k];
        lp = fma(tp.w_prot[k], loss_atom(mode, obs - pred, obs, pred), lp);
    }
    for (int k = threadIdx.x; k < tp.n_rna; k += GLOBAL_BLOCK) {
        const int st = tp.offset_y[tp.p_rna[k]];
        const double pred = fmax(traj[(size_t)tp.t_rna[k] * n + st], 1e-9) / fmax(traj[(size_t)tp.rna_base * n + st], 1e-9);
        const double obs = tp.obs_rna[k];
        lr = fma(tp.w_rna[k], loss_atom(mode, obs - pred, obs, pred), lr);
    }
    for (int k = threadIdx.x; k < tp.n_pho; k += GLOBAL_BLOCK) {
        const int col = tp.offset_y[tp.p_pho[k]] + 2 + tp.s_pho[k];
        const double pred = fmax(traj[(size_t)tp.t_pho[k] * n + col], 1e-9) / fmax(traj[(size_t)tp.pho_base * n + col], 1e-9);
        const double obs = tp.obs_pho[k];
        lph = fma(tp.w_pho[k], loss_atom(mode, obs - pred, obs, pred), lph);
    }
    lp = block_sum_d(lp, red);
    lr = block_sum_d(lr, red);
    lph = block_sum_d(lph, red);
}

// Morris scalar of the global path: fold changes (floor 1e-12) of every protein / mRNA / site at the
// requested time indices (simulate.py:105-182) reduced as sensitivity.py:106-140.
__device__ __forceinline__ double metric_value(const GlobalTopoDev& tp, const double* traj, int n, int metric,
                                               int n_mt_prot, int n_mt_rna, int n_mt_pho, const int* mt_prot,
                                               const int* mt_rna, const int* mt_pho, int mb_prot, int mb_rna, int mb_pho,
                                               double* red) {
    const int N = tp.N;
    double s1 = 0.0, s2 = 0.0;
    const int np_ = N * n_mt_prot, nr_ = N * n_mt_rna;
    for (int k = threadIdx.x; k < np_; k += GLOBAL_BLOCK) {
        const int i = k / n_mt_prot, ti = mt_prot[k - i * n_mt_prot];
        const int st = tp.offset_y[i], ns = tp.n_sites[i];
        const double* rt = traj + (size_t)ti * n + st + 1;
        const double* rb = traj + (size_t)mb_prot * n + st + 1;
        double a1 = 0.0, b1 = 0.0;
        for (int j = 0; j <= ns; ++j) { a1 += rt[j]; b1 += rb[j]; }
        const double fc = fmax(a1, 1e-12) / fmax(b1, 1e-12);
        s1 += fc;
        s2 = fma(fc, fc, s2);
    }
    for (int k = threadIdx.x; k < nr_; k += GLOBAL_BLOCK) {
        const int i = k / n_mt_rna, ti = mt_rna[k - i * n_mt_rna];
        const int st = tp.offset_y[i];
        const double fc = fmax(traj[(size_t)ti * n + st], 1e-12) / fmax(traj[(size_t)mb_rna * n + st], 1e-12);
        s1 += fc;
        s2 = fma(fc, fc, s2);
    }
    if (n_mt_pho > 0) {
        for (int i = 0; i < N; ++i) {
            const int st = tp.offset_y[i], ns = tp.n_sites[i];
            for (int k = threadIdx.x; k < ns * n_mt_pho; k += GLOBAL_BLOCK) {
                const int j = k / n_mt_pho, ti = mt_pho[k - j * n_mt_pho];
                const int col = st + 2 + j;
                const double fc = fmax(traj[(size_t)ti * n + col], 1e-12) / fmax(traj[(size_t)mb_pho * n + col], 1e-12);
                s1 += fc;
                s2 = fma(fc, fc, s2);
            }
        }
    }
    s1 = block_sum_d(s1, red);
    s2 = block_sum_d(s2, red);
    const double cnt = (double)np_ + (double)nr_ + (double)tp.S * n_mt_pho;
    if (cnt == 0.0) return 0.0;
    if (metric == 1) return s1 / cnt;
    if (metric == 2) return s2 / cnt - (s1 / cnt) * (s1 / cnt);
    if (metric == 3) return sqrt(s2);
    return s1;
}

// LOSS_FN(Y, ...) on trajectories that already exist (lossfn.py:113-121): one CTA per trajectory.
__global__ void __launch_bounds__(GLOBAL_BLOCK) global_loss_kernel(const GlobalTopoDev tp, const double* Y, long long B, int T,
                                                                   int mode, double* out_loss) {
    __shared__ double red[2 * GLOBAL_WARPS];
    for (long long sys = blockIdx.x; sys < B; sys += gridDim.x) {
        double lp, lr, lph;
        loss_sums(tp, Y + (size_t)sys * T * tp.n, tp.n, mode, red, lp, lr, lph);
        if (threadIdx.x == 0) { out_loss[sys * 3] = lp; out_loss[sys * 3 + 1] = lr; out_loss[sys * 3 + 2] = lph; }
    }
}

struct GlobalCtx {
    const GlobalTopoDev& tp;
    double *par, *Kt, *Sall, *y, *arg, *U, *w, *facA, *mult, *clo, *pvec, *g, *m, *z, *Sc, *idiag, *red;
    int* perm;
    int ld, n, N;
    const double *cA, *cB, *cC, *cD, *cDp, *cE;      // views into par
    double tfs;
};

// f(src) -> dst.  With FACTOR: also the transcription gains g_i, the per-protein tree factorisation of
// A = I - c J_blk, the unit responses w = A^-1 e_R and m_i.  Two phases, thread per protein.
template <bool FACTOR>
__device__ __forceinline__ void eval_rhs(const GlobalCtx& cx, const double* src, double* dst, double c) {
    const GlobalTopoDev& tp = cx.tp;
    const int N = cx.N, model = tp.model;
    for (int i = threadIdx.x; i < N; i += GLOBAL_BLOCK) {
        const int d = tp.driver_map[i];
        double pv;
        if (d >= 0) pv = cx.Kt[d];                                   // live drive (jacspeedup.py:210-221)
        else {
            const int st = tp.offset_y[i], ns = tp.n_sites[i];
            pv = src[st + 1];
            for (int j = 0; j < ns; ++j) pv += src[st + 2 + j];
        }
        cx.pvec[i] = pv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += GLOBAL_BLOCK) {
        const int st = tp.offset_y[i], ss = tp.offset_s[i], ns = tp.n_sites[i];
        double v = 0.0;
        for (int q = tp.TF_indptr[i]; q < tp.TF_indptr[i + 1]; ++q) v = fma(tp.TF_data[q], cx.pvec[tp.TF_indices[q]], v);
        v /= tp.tf_deg[i];
        double synth, dsdv;
        synth_rate(model, v, cx.cA[i], cx.tfs, synth, dsdv);
        const double R = src[st], P = src[st + 1];
        const double Bi = cx.cB[i], Ci = cx.cC[i], Di = cx.cD[i], Ei = cx.cE[i];
        dst[st] = fma(-Bi, R, synth);
        const bool chain = model == 1;
        double cPR = Ci, dPP;
        if (model == 0) {
            double sumS = 0.0, back = 0.0;
            for (int j = 0; j < ns; ++j) {
                const double s = cx.Sall[ss + j], ps = src[st + 2 + j];
                sumS += s;
                back += ps;
                dst[st + 2 + j] = fma(s, P, -(Ei + cx.cDp[ss + j] + Di) * ps);
            }
            dst[st + 1] = fma(Ci, R, fma(-(Di + sumS), P, Ei * back));
            dPP = -(Di + sumS);
        } else if (model == 4) {
            const double iP = 1.0 / (1.0 + P), iR = 1.0 / (1.0 + R);
            double sumS = 0.0, back = 0.0;
            for (int j = 0; j < ns; ++j) {
                const double s = cx.Sall[ss + j], ps = src[st + 2 + j];
                sumS += s;
                back += ps;
                dst[st + 2 + j] = fma(s * P, iP, -(cx.cDp[ss + j] + Di + Ei) * ps);
            }
            dst[st + 1] = fma(Ci * R, iR, fma(-Di, P, fma(-sumS * P, iP, Ei * back)));
            cPR = Ci * iR * iR;
            dPP = -Di - sumS * iP * iP;
        } else {
            if (ns == 0) {
                dst[st + 1] = fma(Ci, R, -Di * P);
                dPP = -Di;
            } else {
                const double s0 = cx.Sall[ss];
                dst[st + 1] = fma(Ci, R, fma(-(Di + s0), P, Ei * src[st + 2]));
                dPP = -(Di + s0);
                for (int j = 0; j < ns; ++j) {
                    double gain = cx.Sall[ss + j] * src[st + 1 + j];
                    double out = Ei + cx.cDp[ss + j] + Di;
                    if (j < ns - 1) { gain = fma(Ei, src[st + 3 + j], gain); out += cx.Sall[ss + j + 1]; }
                    dst[st + 2 + j] = fma(-out, src[st + 2 + j], gain);
                }
            }
        }
        if (FACTOR) {
            cx.g[i] = dsdv / tp.tf_deg[i];
            // pivots: facA[st+1] (P0), facA[st+2+j] (site j); children are eliminated before parents
            const double lo_scale = (model == 4) ? 1.0 / ((1.0 + P) * (1.0 + P)) : 1.0;
            cx.facA[st + 1] = fma(-c, dPP, 1.0);
            for (int j = 0; j < ns; ++j) {
                double dg;
                if (model == 1) dg = -(Ei + cx.cDp[ss + j] + Di + (j < ns - 1 ? cx.Sall[ss + j + 1] : 0.0));
                else dg = -(Ei + cx.cDp[ss + j] + Di);
                cx.facA[st + 2 + j] = fma(-c, dg, 1.0);
            }
            for (int j = ns - 1; j >= 0; --j) {
                const int par = chain ? st + 1 + j : st + 1;
                const double ip = 1.0 / cx.facA[st + 2 + j];
                const double clo = c * cx.Sall[ss + j] * lo_scale;
                const double mu = -c * Ei * ip;                         // A(par, j) / pivot_j
                cx.facA[par] = fma(mu, clo, cx.facA[par]);              // pivot_par -= mu * A(j, par), A(j,par) = -clo
                cx.facA[st + 2 + j] = ip;
                cx.mult[st + 2 + j] = mu;
                cx.clo[st + 2 + j] = clo;
            }
            const double ipP = 1.0 / cx.facA[st + 1];
            cx.facA[st + 1] = ipP;
            cx.mult[st + 1] = c * cPR;
            const double iRr = 1.0 / fma(c, Bi, 1.0);
            cx.facA[st] = iRr;
            // w = A^-1 e_R for this block, m_i = total-protein response
            cx.w[st] = iRr;
            double xp = c * cPR * iRr * ipP;
            cx.w[st + 1] = xp;
            double msum = xp;
            for (int j = 0; j < ns; ++j) {
                const int par = chain ? st + 1 + j : st + 1;
                const double xj = cx.clo[st + 2 + j] * cx.w[par] * cx.facA[st + 2 + j];
                cx.w[st + 2 + j] = xj;
                msum += xj;
            }
            cx.m[i] = msum;
        }
    }
    __syncthreads();
}

// Dense LU with partial pivoting of the nQ x nQ Schur matrix in shared memory (row-major, ld).
__device__ __forceinline__ void schur_factor(const GlobalCtx& cx, double c) {
    const GlobalTopoDev& tp = cx.tp;
    const int nQ = tp.nQ, ld = cx.ld;
    double* Sc = cx.Sc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int idx = threadIdx.x; idx < nQ * ld; idx += GLOBAL_BLOCK) Sc[idx] = 0.0;
    __syncthreads();
    for (int qi = threadIdx.x; qi < nQ; qi += GLOBAL_BLOCK) {
        const int i = tp.qlist[qi];
        double* row = Sc + qi * ld;
        const double f = -c * cx.m[i] * cx.g[i];
        for (int q = tp.TF_indptr[i]; q < tp.TF_indptr[i + 1]; ++q) {
            const int qj = tp.qpos[tp.TF_indices[q]];
            if (qj >= 0) row[qj] = fma(f, tp.TF_data[q], row[qj]);
        }
        row[qi] += 1.0;
        cx.perm[qi] = qi;
    }
    __syncthreads();
    for (int k = 0; k < nQ; ++k) {
        if (warp == 0) {                                    // pivot search down column k
            double best = -1.0;
            int bi = k;
            for (int i = k + lane; i < nQ; i += 32) {
                const double a = fabs(Sc[i * ld + k]);
                if (a > best) { best = a; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                cx.red[GLOBAL_WARPS] = (double)bi;
                const int t = cx.perm[k]; cx.perm[k] = cx.perm[bi]; cx.perm[bi] = t;
            }
        }
        __syncthreads();
        const int p = (int)cx.red[GLOBAL_WARPS];
        if (p != k) {
            for (int j = threadIdx.x; j < nQ; j += GLOBAL_BLOCK) {
                const double a = Sc[k * ld + j];
                Sc[k * ld + j] = Sc[p * ld + j];
                Sc[p * ld + j] = a;
            }
            __syncthreads();
        }
        const double ipv = 1.0 / Sc[k * ld + k];
        if (threadIdx.x == 0) cx.idiag[k] = ipv;
        for (int i = k + 1 + warp; i < nQ; i += GLOBAL_WARPS) {
            double* row = Sc + i * ld;
            const double l = row[k] * ipv;
            __syncwarp();
            if (lane == 0) row[k] = l;
            for (int j = k + 1 + lane; j < nQ; j += 32) row[j] = fma(-l, Sc[k * ld + j], row[j]);
        }
        __syncthreads();
    }
}

// x (vector in shared memory, holds the right-hand side b) <- (I - cJ)^-1 b
__device__ __forceinline__ void schur_solve(const GlobalCtx& cx, double* x, double c) {
    const GlobalTopoDev& tp = cx.tp;
    const int N = cx.N, nQ = tp.nQ, ld = cx.ld;
    const bool chain = tp.model == 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // block solves x0 = A^-1 b and z0 (into pvec)
    for (int i = threadIdx.x; i < N; i += GLOBAL_BLOCK) {
        const int st = tp.offset_y[i], ns = tp.n_sites[i];
        const double xr = x[st] * cx.facA[st];
        x[st] = xr;
        double xp = fma(cx.mult[st + 1], xr, x[st + 1]);
        if (chain) {
            for (int j = ns - 1; j >= 1; --j) x[st + 1 + j] = fma(-cx.mult[st + 2 + j], x[st + 2 + j], x[st + 1 + j]);
            if (ns > 0) xp = fma(-cx.mult[st + 2], x[st + 2], xp);
        } else {
            for (int j = 0; j < ns; ++j) xp = fma(-cx.mult[st + 2 + j], x[st + 2 + j], xp);
        }
        xp *= cx.facA[st + 1];
        x[st + 1] = xp;
        double zs = xp, prev = xp;
        for (int j = 0; j < ns; ++j) {
            const double xj = fma(cx.clo[st + 2 + j], chain ? prev : xp, x[st + 2 + j]) * cx.facA[st + 2 + j];
            x[st + 2 + j] = xj;
            prev = xj;
            zs += xj;
        }
        cx.pvec[i] = zs;
        cx.z[i] = 0.0;
    }
    __syncthreads();
    if (warp == 0 && nQ > 0) {
        double* xq = cx.red + 2 * GLOBAL_WARPS;                      // [nQ] scratch behind the reduction slots
        for (int k = lane; k < nQ; k += 32) xq[k] = cx.pvec[tp.qlist[cx.perm[k]]];
        __syncwarp();
        const double* Sc = cx.Sc;
        for (int k = 0; k < nQ - 1; ++k) {                           // L y = P z0
            const double xk = xq[k];
            for (int i = k + 1 + lane; i < nQ; i += 32) xq[i] = fma(-Sc[i * ld + k], xk, xq[i]);
            __syncwarp();
        }
        for (int k = nQ - 1; k >= 0; --k) {                          // U z = y
            if (lane == 0) xq[k] *= cx.idiag[k];
            __syncwarp();
            const double xk = xq[k];
            for (int i = lane; i < k; i += 32) xq[i] = fma(-Sc[i * ld + k], xk, xq[i]);
            __syncwarp();
        }
        for (int k = lane; k < nQ; k += 32) cx.z[tp.qlist[k]] = xq[k];
    }
    __syncthreads();
    // x += c (G z)_i w_i
    for (int i = threadIdx.x; i < N; i += GLOBAL_BLOCK) {
        double gz = 0.0;
        for (int q = tp.TF_indptr[i]; q < tp.TF_indptr[i + 1]; ++q) gz = fma(tp.TF_data[q], cx.z[tp.TF_indices[q]], gz);
        gz *= c * cx.g[i];
        const int st = tp.offset_y[i], ns = tp.n_sites[i];
        for (int s = st; s < st + 2 + ns; ++s) x[s] = fma(gz, cx.w[s], x[s]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(GLOBAL_BLOCK, 1) global_net_kernel(const GlobalArgs a) {
    extern __shared__ double smem[];
    const GlobalTopoDev& tp = a.tp;
    const GlobalSmem& L = a.sm;
    const int n = tp.n, N = tp.N, K = tp.K, S = tp.S, T = a.T, P = a.P;
    __shared__ long long s_sys;
    GlobalCtx cx{tp,
                 smem + L.par, smem + L.Kt, smem + L.Sall, smem + L.y, smem + L.arg, smem + L.U, smem + L.w,
                 smem + L.facA, smem + L.mult, smem + L.clo, smem + L.pvec, smem + L.g, smem + L.m, smem + L.z,
                 smem + L.Sc, smem + L.idiag, smem + L.red, (int*)(smem + L.perm), L.ld, n, N,
                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0};
    cx.cA = cx.par + K;
    cx.cB = cx.cA + N;
    cx.cC = cx.cB + N;
    cx.cD = cx.cC + N;
    cx.cDp = cx.cD + N;
    cx.cE = cx.cDp + S;
    double* const y = cx.y;
    double* const arg = cx.arg;
    double* const U = cx.U;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_sys = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long sys = s_sys;
        if (sys >= a.B) break;

        // ------------------------------------------------------------------------ load
        const double* pr = a.params + (size_t)sys * P;
        for (int i = threadIdx.x; i < P; i += GLOBAL_BLOCK) {
            const double v = pr[i];
            cx.par[i] = a.theta_mode ? softplus_d(v) : v;             // params.py:106-132
        }
        const double* y0 = a.y0 + (a.y0_stride ? (size_t)sys * a.y0_stride : 0);
        double* traj = a.out_Y ? a.out_Y + (size_t)sys * T * n : a.traj + (size_t)blockIdx.x * T * n;
        for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) y[i] = y0[i];
        __syncthreads();
        cx.tfs = cx.par[P - 1];

        double t = a.stop_t[0];
        double h = 0.0;
        float hacc = 0.f, erracc = 1.f;
        int naccpt = 0, rejected_last = 0;
        int nst = 0, nrej = 0, status = 0;
        if (a.stop_out[0] >= 0)
            for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) traj[(size_t)a.stop_out[0] * n + i] = y[i];

        for (int si = 0; si + 1 < a.n_stops && status == 0; ++si) {
            const double tend = a.stop_t[si + 1];
            const int jb = a.stop_bucket[si];
            // kinase input of this bucket: Kt = Kmat[:, jb] * c_k;  S = W . Kt   (jacspeedup.py:148-172, 70-113)
            if (si == 0 || jb != a.stop_bucket[si - 1]) {
                __syncthreads();
                for (int k = threadIdx.x; k < K; k += GLOBAL_BLOCK) cx.Kt[k] = tp.kin_Kmat[(size_t)k * tp.nb + jb] * cx.par[k];
                __syncthreads();
                for (int s = threadIdx.x; s < S; s += GLOBAL_BLOCK) {
                    double acc = 0.0;
                    for (int q = tp.W_indptr[s]; q < tp.W_indptr[s + 1]; ++q) acc = fma(tp.W_data[q], cx.Kt[tp.W_indices[q]], acc);
                    cx.Sall[s] = acc;
                }
                __syncthreads();
            }
            if (si == 0) {
                // initial step: 1% of the error-weighted time scale |y|/|f|
                eval_rhs<false>(cx, y, arg, 0.0);
                float d0 = 0.f, d1 = 0.f;
                for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) {
                    const double sc = 1.0 / fma(a.rtol, fabs(y[i]), a.atol);
                    d0 = fmaxf(d0, (float)(fabs(y[i]) * sc));
                    d1 = fmaxf(d1, (float)(fabs(arg[i]) * sc));
                }
                const double D0 = block_max_f(d0, cx.red), D1 = block_max_f(d1, cx.red);
                h = (D0 < 1e-5 || D1 < 1e-5 || !(D1 < 3.0e38)) ? 1e-6 : 0.01 * D0 / D1;
                hacc = (float)h;
            }
            while (status == 0) {
                const double rem = tend - t;
                if (!(rem > 0.0)) break;
                double hh = h;
                bool land = false;
                if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
                else if (hh > 0.5 * rem) hh = 0.5 * rem;
                const double c = hh * G_GAMMA;
                const double ih = 1.0 / hh;

                // stage 1: f(y), Jacobian pieces, factorisation
                eval_rhs<true>(cx, y, U, c);
                schur_factor(cx, c);
                for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) U[i] *= c;
                __syncthreads();
                schur_solve(cx, U, c);
                // stages 2..6:  (I - cJ) U_s = c ( f(y + sum a_sj U_j) + sum c_sj/h U_j )
                for (int s = 1; s < 6; ++s) {
                    for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) {
                        double v = y[i];
                        for (int j = 0; j < s && j < 4; ++j) v = fma(G_A[s - 1][j], U[j * n + i], v);
                        if (s == 5) v += U[4 * n + i];
                        arg[i] = v;
                    }
                    __syncthreads();
                    double* Us = U + s * n;
                    eval_rhs<false>(cx, arg, Us, c);
                    for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) {
                        double v = 0.0;
                        for (int j = 0; j < s; ++j) v = fma(G_C[s - 1][j], U[j * n + i], v);
                        Us[i] = c * fma(v, ih, Us[i]);
                    }
                    __syncthreads();
                    schur_solve(cx, Us, c);
                }
                // y_new = arg_6 + U_6, err = U_6
                float err = 0.f;
                bool bad = false;
                for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) {
                    const double e = U[5 * n + i];
                    const double yn = arg[i] + e;
                    arg[i] = yn;
                    const float q = err_ratio(e, y[i], yn, a.rtol, a.atol);
                    bad |= !(q < 3.0e38f) || !(fabs(yn) < 1.0e300);
                    err = fmaxf(err, q);
                }
                if (bad) err = __int_as_float(0x7f800000);
                err = (float)block_max_f(err, cx.red);
                if (!(err < 3.0e38f)) {
                    // non-finite stage values: treat as a rejected step with maximal shrink unless h is already tiny
                    ++nrej;
                    rejected_last = 1;
                    h = hh * 0.2;
                    if (h < 1e-14 * fmax(1.0, fabs(t))) status = 3;
                } else if (err <= 1.0f) {
                    ++nst;
                    float fac = ctl_factor(err, 0.25f);
                    const float hf = (float)hh;
                    if (naccpt > 0) {
                        const float r = __fdividef(err * err, erracc);
                        float fg = __fdividef(hacc, hf) * __powf(r, 0.25f) * CTL_INV_SAFE;
                        fg = fmaxf(CTL_FAC_GROW, fminf(CTL_FAC_SHRINK, fg));
                        fac = fmaxf(fac, fg);
                    }
                    hacc = hf;
                    erracc = fmaxf(1.0e-2f, err);
                    ++naccpt;
                    double hnew = hh / (double)fac;
                    if (rejected_last) hnew = fmin(hnew, hh);
                    rejected_last = 0;
                    h = (hh < h) ? fmax(hnew, fmin(h, 6.0 * hh)) : hnew;
                    for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) y[i] = arg[i];
                    t = land ? tend : t + hh;
                    __syncthreads();
                } else {
                    ++nrej;
                    rejected_last = 1;
                    h = hh / (double)ctl_factor(err, 0.25f);
                    if (h < 1e-14 * fmax(1.0, fabs(t))) status = 2;
                }
                if (status == 0 && nst + nrej >= a.max_steps && t < tend) status = 1;
            }
            if (status == 0) {
                t = tend;
                const int ko = a.stop_out[si + 1];
                if (ko >= 0)
                    for (int i = threadIdx.x; i < n; i += GLOBAL_BLOCK) traj[(size_t)ko * n + i] = y[i];
            }
        }
        if (status != 0) {                                           // failed system: NaN trajectory
            for (int i = threadIdx.x; i < T * n; i += GLOBAL_BLOCK) traj[i] = qnan;
        }
        __syncthreads();

        // ---------------------------------------------------------------------- epilogue
        if (threadIdx.x == 0) {
            if (a.out_status) a.out_status[sys] = status;
            if (a.out_nsteps) a.out_nsteps[sys] = nst;
            if (a.out_nrej) a.out_nrej[sys] = nrej;
        }
        if (a.out_loss || a.out_F) {
            double lp, lr, lph;
            loss_sums(tp, traj, n, a.loss_mode, cx.red, lp, lr, lph);
            double prior = 0.0;
            if (a.out_F && tp.defaults) {
                // optproblem.py:105-114: mean over A,B,C,D,E of ((p - p0)/(p0 + 1e-6))^2
                double acc = 0.0;
                for (int i = threadIdx.x; i < 5 * N; i += GLOBAL_BLOCK) {
                    const int grp = i / N, j = i - grp * N;
                    const int pi = (grp < 4) ? K + grp * N + j : K + 4 * N + S + j;
                    const double p0 = tp.defaults[pi];
                    const double d = (cx.par[pi] - p0) / (p0 + 1e-6);
                    acc = fma(d, d, acc);
                }
                acc = block_sum_d(acc, cx.red);
                prior = a.lam_prior * acc / (double)(5 * N > 0 ? 5 * N : 1);
            }
            if (threadIdx.x == 0) {
                if (a.out_loss) { a.out_loss[sys * 3] = lp; a.out_loss[sys * 3 + 1] = lr; a.out_loss[sys * 3 + 2] = lph; }
                if (a.out_F) {
                    a.out_F[sys * 3] = lp * tp.norm[0] * a.lam[0] + prior;
                    a.out_F[sys * 3 + 1] = lr * tp.norm[1] * a.lam[1] + prior;
                    a.out_F[sys * 3 + 2] = lph * tp.norm[2] * a.lam[2] + prior;
                }
            }
        }
        if (a.out_metric) {
            const double mv = metric_value(tp, traj, n, a.metric, a.n_mt_prot, a.n_mt_rna, a.n_mt_pho, a.mt_prot, a.mt_rna,
                                           a.mt_pho, a.mb_prot, a.mb_rna, a.mb_pho, cx.red);
            if (threadIdx.x == 0) a.out_metric[sys] = mv;
        }
    }
}

}  // namespace pk
