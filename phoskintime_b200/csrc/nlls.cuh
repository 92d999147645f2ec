// Device side of the batched bounded non-linear least-squares driver (pk_local_nlls_batch).
//
// Reference shape (paramest/normest.py:79-89, 278-290, 494-509): one SciPy `curve_fit(model_func, ..., p0, bounds,
// sigma, x_scale='jac')` per start, i.e. bounded trust-region least squares on the residual
//     r(theta) = ([flat(solve_ode(theta)) | lam/P * theta^2] - [target | 0]) / sigma        (normest.py:403-423)
// with a 2-point finite-difference Jacobian: P+1 solves per Jacobian, one after the other, on one core.
// Here every start of every protein is one row of a batch and an iteration is
//     perturb (this file)  ->  ONE launch of the ODE kernel over B*(P+1) systems (flat + ssr)
//     lm_step  (this file)  ->  ONE launch of the ODE kernel over the B trial points (ssr)  ->  lm_accept (this file)
// Only problems that are still running take part: lm_accept appends the survivors to a compact index list, the next
// iteration's batches hold nA*(P+1) and nA systems (nA = survivors), so converged starts cost nothing further.
// The optimiser POLICY is a projected Levenberg-Marquardt (Nielsen damping, Jacobian column scaling as MINPACK's
// x_scale='jac'), not a transcription of SciPy's TRF: SURVEY.md section 8(c) places the optimiser outside the parity
// contract; what is checked is that the minima found agree with SciPy's on the same residual (tests/test_gpu_nlls.py).
#pragma once
#include <cuda_runtime.h>

namespace pk {

struct NllsState {           // per problem, device resident between iterations
    double mu, nu;           // damping and its growth factor (Nielsen)
    double cost;             // 0.5 * ssr at theta
    double pred;             // predicted reduction of the trial step
    double snorm, xnorm;     // scaled norms of the trial step and of theta
    int status;              // 0 running, 1 gtol, 2 ftol, 3 xtol, 4 max_iter, -1 numerical failure
    int iters, nfev, pad;
};

struct NllsArgs {
    long long B;
    int P, L, sigma_len, log_params;
    double lam, fd_rel, ftol, xtol, gtol;
    const double* lam_group;       // NULL or [n_groups]
    const double *lb, *ub;   // [P]
    const double* target;    // [G,L]
    const double* sigma;     // [G,sigma_len] or nullptr
    const int* group;        // [B] or nullptr
    double* theta;           // [B,P]
    double* pert;            // [B*(P+1),P]
    int* pert_group;         // [B*(P+1)] or nullptr
    double* hstep;           // [B,P] signed forward-difference steps
    const double* flat;      // [B*(P+1),L]
    const double* ssr;       // [B*(P+1)]
    const int* solve_status; // [B*(P+1)]
    double* dscale;          // [B,P] running column scale
    double* trial;           // [B,P]
    const double* trial_ssr; // [B]
    const int* trial_status; // [B]
    NllsState* st;           // [B]
    int* n_running;          // device counter of problems still running (written by lm_accept)
    // compaction: row a of the perturbed / trial batches belongs to problem idx[a], a < nA
    long long nA;
    const int* idx;          // [nA] problems still running, this iteration
    int* idx_next;           // [<=nA] survivors, written by lm_accept
    int* trial_group;        // [nA] or nullptr
    const double* y0;        // per-problem initial conditions [B, y0_stride] (y0_stride > 0) ...
    long long y0_stride;
    int n;
    double* y0_pert;         // ... gathered to [nA*(P+1), n]
    double* y0_trial;        // ... and to [nA, n]
};

// theta -> the P+1 parameter rows of the forward-difference Jacobian.  Step: fd_rel * max(1, |theta_j|), flipped when it
// would leave the box (SciPy's 2-point scheme does the same, scipy/optimize/_numdiff.py `_adjust_scheme_to_bounds`).
__global__ void nlls_perturb_kernel(const NllsArgs a) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int P = a.P;
    if (idx >= a.nA * (P + 1)) return;
    const long long ai = idx / (P + 1);
    const long long b = a.idx[ai];
    const int k = (int)(idx - ai * (P + 1));
    const double* th = a.theta + b * P;
    double* row = a.pert + idx * P;
    for (int j = 0; j < P; ++j) row[j] = th[j];
    if (a.pert_group) a.pert_group[idx] = a.group[b];
    if (a.y0_pert)
        for (int i = 0; i < a.n; ++i) a.y0_pert[idx * a.n + i] = a.y0[b * a.y0_stride + i];
    if (k > 0) {
        const int j = k - 1;
        double h = a.fd_rel * fmax(1.0, fabs(th[j]));
        if (th[j] + h > a.ub[j] && th[j] - h >= a.lb[j]) h = -h;
        row[j] = th[j] + h;
        a.hstep[b * P + j] = row[j] - th[j];     // the step actually representable in FP64
    }
}

__device__ __forceinline__ double nlls_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per problem.  Shared memory per warp: J [P][Lp] (column major, Lp odd), r [L], A [P][P], g, d, dl, act [P].
__global__ void nlls_step_kernel(const NllsArgs a, int warps_per_cta, int Lp) {
    extern __shared__ double nsm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long ai = blockIdx.x * (long long)warps_per_cta + wid;
    if (ai >= a.nA) return;
    const long long b = a.idx[ai];
    NllsState& S = a.st[b];
    const int P = a.P, L = a.L;
    double* th = a.theta + b * P;
    double* tr = a.trial + ai * P;
    if (lane == 0 && a.trial_group) a.trial_group[ai] = a.group[b];
    if (a.y0_trial)
        for (int i = lane; i < a.n; i += 32) a.y0_trial[ai * a.n + i] = a.y0[b * a.y0_stride + i];
    const size_t per_warp = (size_t)P * Lp + L + (size_t)P * P + 4 * (size_t)P;
    double* J = nsm + wid * per_warp;
    double* r = J + (size_t)P * Lp;
    double* A = r + L;
    double* g = A + (size_t)P * P;
    double* dsc = g + P;
    double* dl = dsc + P;
    double* act = dl + P;
    const long long row0 = ai * (P + 1);
    const int grp = a.group ? a.group[b] : 0;
    const double* tg = a.target + (size_t)grp * L;
    const double* sg = a.sigma ? a.sigma + (size_t)grp * a.sigma_len : nullptr;
    const double* f0 = a.flat + row0 * L;
    bool bad = a.solve_status[row0] != 0;
    for (int j = 0; j < P; ++j) bad |= a.solve_status[row0 + 1 + j] != 0;
    if (bad) {                                // the integrator failed on the base point or a perturbed one
        if (lane == 0) { S.status = -1; }
        for (int j = lane; j < P; j += 32) tr[j] = th[j];
        return;
    }
    // reciprocal forward-difference steps once per problem (dl is free until the normal equations are damped)
    for (int j = lane; j < P; j += 32) dl[j] = 1.0 / a.hstep[b * P + j];
    __syncwarp();
    {
        const double* __restrict__ fl = a.flat + (row0 + 1) * L;     // the P perturbed rows, [P][L]
        for (int l = lane; l < L; l += 32) {
            const double is = sg ? 1.0 / sg[l] : 1.0;
            const double base = f0[l];
            r[l] = (base - tg[l]) * is;
            int j = 0;
            for (; j + 4 <= P; j += 4) {                              // four independent loads in flight per lane
                const double v0 = fl[(size_t)j * L + l], v1 = fl[(size_t)(j + 1) * L + l];
                const double v2 = fl[(size_t)(j + 2) * L + l], v3 = fl[(size_t)(j + 3) * L + l];
                J[(size_t)j * Lp + l] = (v0 - base) * is * dl[j];
                J[(size_t)(j + 1) * Lp + l] = (v1 - base) * is * dl[j + 1];
                J[(size_t)(j + 2) * Lp + l] = (v2 - base) * is * dl[j + 2];
                J[(size_t)(j + 3) * Lp + l] = (v3 - base) * is * dl[j + 3];
            }
            for (; j < P; ++j) J[(size_t)j * Lp + l] = (fl[(size_t)j * L + l] - base) * is * dl[j];
        }
    }
    __syncwarp();
    // A = J^T J, g = J^T r  (+ the diagonal regularisation rows lam/P*theta_j^2 / sigma_{L+j})
    const int npairs = P * (P + 1) / 2;
    for (int q = lane; q < npairs; q += 32) {
        int i = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        while (i * (i + 1) / 2 > q) --i;
        const int j = q - i * (i + 1) / 2;    // j <= i
        const double* ci = J + (size_t)i * Lp;
        const double* cj = J + (size_t)j * Lp;
        double acc = 0.0;
        for (int l = 0; l < L; ++l) acc = fma(ci[l], cj[l], acc);
        A[i * P + j] = acc;
        A[j * P + i] = acc;
    }
    for (int j = lane; j < P; j += 32) {
        const double* cj = J + (size_t)j * Lp;
        double acc = 0.0;
        for (int l = 0; l < L; ++l) acc = fma(cj[l], r[l], acc);
        g[j] = acc;
    }
    __syncwarp();
    for (int j = lane; j < P; j += 32) {
        const double lam = a.lam_group ? a.lam_group[grp] : a.lam;
        if (lam != 0.0) {
            const double is = (sg && a.sigma_len > L) ? 1.0 / sg[L + j] : 1.0;
            const double rr = lam / (double)P * th[j] * th[j] * is;
            const double dj = 2.0 * lam / (double)P * th[j] * is;
            A[j * P + j] = fma(dj, dj, A[j * P + j]);
            g[j] = fma(dj, rr, g[j]);
        }
        // MINPACK-style running column scale (x_scale='jac')
        double d = fmax(a.dscale[b * P + j], sqrt(A[j * P + j]));
        if (!(d > 0.0)) d = 1.0;
        a.dscale[b * P + j] = d;
        dsc[j] = d;
        // active bounds: at a bound with the gradient pushing outwards
        const double tol = 1e-12 * fmax(1.0, fabs(th[j]));
        const bool fixed = (th[j] <= a.lb[j] + tol && g[j] > 0.0) || (th[j] >= a.ub[j] - tol && g[j] < 0.0);
        act[j] = fixed ? 0.0 : 1.0;
    }
    __syncwarp();
    // first-order optimality as in MINPACK lmder: max_j |J_j . r| / (|J_j| |r|) over the free variables
    const double cost0 = 0.5 * a.ssr[row0];
    double gm = 0.0, xn = 0.0;
    for (int j = lane; j < P; j += 32) {
        gm = fmax(gm, act[j] * fabs(g[j]) / dsc[j]);
        xn = fma(dsc[j] * th[j], dsc[j] * th[j], xn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    gm = (cost0 > 0.0) ? gm / sqrt(2.0 * cost0) : 0.0;
    xn = sqrt(nlls_warp_sum(xn));
    if (lane == 0) {
        S.cost = cost0;
        S.xnorm = xn;
        ++S.iters;
        S.nfev += P + 1;
    }
    if (!(gm > a.gtol)) {                      // also catches NaN
        if (lane == 0) S.status = (gm <= a.gtol) ? 1 : -1;
        for (int j = lane; j < P; j += 32) tr[j] = th[j];
        return;
    }
    // damped normal equations over the free variables: (A + mu D^2) dl = -g, Cholesky in place in the LOWER triangle
    // (the strict upper triangle keeps J^T J for the predicted reduction)
    const double mu = S.mu;
    for (int j = lane; j < P; j += 32) {
        dl[j] = A[j * P + j];                  // keep the undamped diagonal
        if (act[j] == 0.0) {
            for (int i = 0; i < P; ++i)
                if (i > j) A[i * P + j] = 0.0; else if (i < j) A[j * P + i] = 0.0;
        }
    }
    __syncwarp();
    for (int j = lane; j < P; j += 32) A[j * P + j] = (act[j] == 0.0) ? 1.0 : fma(mu * dsc[j], dsc[j], dl[j]);
    __syncwarp();
    bool chol_ok = true;
    for (int k = 0; k < P; ++k) {
        const double akk = A[k * P + k];
        if (!(akk > 0.0)) { chol_ok = false; break; }
        const double d = sqrt(akk);
        __syncwarp();
        for (int i = k + 1 + lane; i < P; i += 32) A[i * P + k] /= d;
        if (lane == 0) A[k * P + k] = d;
        __syncwarp();
        for (int i = k + 1 + lane; i < P; i += 32) {
            const double lik = A[i * P + k];
            for (int j = k + 1; j <= i; ++j) A[i * P + j] = fma(-lik, A[j * P + k], A[i * P + j]);
        }
        __syncwarp();
    }
    if (!chol_ok) {
        if (lane == 0) S.status = -1;
        for (int j = lane; j < P; j += 32) tr[j] = th[j];
        return;
    }
    // forward / backward substitution by lane 0 (P <= ~140: a few thousand flops)
    if (lane == 0) {
        double* x = act;                        // reuse: act is re-derived below from g
        for (int i = 0; i < P; ++i) {
            double s = (x[i] == 0.0) ? 0.0 : -g[i];
            for (int j = 0; j < i; ++j) s = fma(-A[i * P + j], x[j], s);
            x[i] = s / A[i * P + i];
        }
        for (int i = P - 1; i >= 0; --i) {
            double s = x[i];
            for (int j = i + 1; j < P; ++j) s = fma(-A[j * P + i], x[j], s);
            x[i] = s / A[i * P + i];
        }
    }
    __syncwarp();
    // clip the step to the box, predicted reduction with the UNDAMPED model  -(g.d + 0.5 d^T (J^T J) d)
    for (int j = lane; j < P; j += 32) {
        const double t = fmin(fmax(th[j] + act[j], a.lb[j]), a.ub[j]);
        tr[j] = t;
        act[j] = t - th[j];
    }
    __syncwarp();
    double lin = 0.0, quad = 0.0, sn = 0.0;
    for (int i = lane; i < P; i += 32) {
        const double di = act[i];
        lin = fma(g[i], di, lin);
        double rowacc = 0.5 * dl[i] * di;      // diagonal of J^T J (saved before damping)
        for (int j = i + 1; j < P; ++j) rowacc = fma(A[i * P + j], act[j], rowacc);   // strict upper triangle: J^T J
        quad = fma(di, rowacc, quad);
        sn = fma(dsc[i] * di, dsc[i] * di, sn);
    }
    lin = nlls_warp_sum(lin);
    quad = nlls_warp_sum(quad);
    sn = sqrt(nlls_warp_sum(sn));
    if (lane == 0) {
        S.pred = -(lin + quad);
        S.snorm = sn;
    }
}

// Gain ratio, acceptance, damping update (Nielsen 1999) and the ftol / xtol tests; one thread per problem.
__global__ void nlls_accept_kernel(const NllsArgs a, int last_iter) {
    const long long ai = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ai >= a.nA) return;
    const long long b = a.idx[ai];
    NllsState& S = a.st[b];
    if (S.status != 0) return;                // finished inside lm_step (gtol / numerical failure)
    const int P = a.P;
    const double ct = 0.5 * a.trial_ssr[ai];
    const bool ok = a.trial_status[ai] == 0 && ct == ct && S.pred > 0.0;
    ++S.nfev;
    const double rho = ok ? (S.cost - ct) / S.pred : -1.0;
    if (ok && rho > 1e-4 && ct < S.cost) {
        const double dc = S.cost - ct;
        for (int j = 0; j < P; ++j) a.theta[b * P + j] = a.trial[ai * P + j];
        const double f = 2.0 * rho - 1.0;
        S.mu *= fmax(1.0 / 3.0, 1.0 - f * f * f);
        S.nu = 2.0;
        if (dc <= a.ftol * S.cost) S.status = 2;
        else if (S.snorm <= a.xtol * (a.xtol + S.xnorm)) S.status = 3;
        S.cost = ct;
    } else {
        S.mu *= S.nu;
        S.nu *= 2.0;
        if (S.snorm <= a.xtol * (a.xtol + S.xnorm)) S.status = 3;       // the step has shrunk to nothing
        if (!(S.mu < 1e30)) S.status = 3;
    }
    if (S.status == 0 && last_iter) S.status = 4;
    if (S.status == 0) a.idx_next[atomicAdd(a.n_running, 1)] = (int)b;
}

__global__ void nlls_init_kernel(const NllsArgs a, double mu0, int* idx0) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    idx0[b] = (int)b;
    NllsState& S = a.st[b];
    S.mu = mu0; S.nu = 2.0; S.cost = 0.0; S.pred = 0.0; S.snorm = 0.0; S.xnorm = 0.0;
    S.status = 0; S.iters = 0; S.nfev = 0; S.pad = 0;
    for (int j = 0; j < a.P; ++j) {
        a.dscale[b * a.P + j] = 0.0;
        a.theta[b * a.P + j] = fmin(fmax(a.theta[b * a.P + j], a.lb[j]), a.ub[j]);   // curve_fit requires p0 inside the box
    }
}

}  // namespace pk
