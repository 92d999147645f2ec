// Host driver of pk_local_nlls_batch (include/phoskin_b200.h): batched bounded least squares over the local models.
// Replaces a loop of scipy.optimize.curve_fit calls (paramest/normest.py:79-89, 278-290, 494-509), each of which
// evaluates its residual model (normest.py:403-423) and a 2-point Jacobian one solve_ode at a time.
// The iteration loop lives here (C++), every solve is pk_local_solve_batch on device buffers, the linear algebra
// of a step is nlls_step_kernel (csrc/nlls.cuh).  Per iteration 2 ODE launches + 3 small kernels, 4 bytes to the host
// (the number of problems still running, which sizes the next iteration's batches).
#include <algorithm>
#include <cmath>

#include "pk_internal.hpp"

#include "nlls.cuh"

using pkh::fail;

namespace {

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T*)p;
        return e;
    }
};

}  // namespace

extern "C" {

void pk_nlls_job_init(pk_nlls_job* j) {
    memset(j, 0, sizeof(*j));
    j->max_iter = 100;
    j->ftol = 1e-8;            /* curve_fit / least_squares defaults */
    j->xtol = 1e-8;
    j->gtol = 1e-8;
    j->fd_rel = 1e-4;
    j->rtol = 2e-7;            /* one decade below the library default: the cost must be smooth enough to difference */
    j->atol = 2e-10;
    j->n_groups = 1;
    for (int i = 0; i < 5; ++i) j->score_w[i] = 1.0;
}

int pk_sizeof_nlls_job(void) { return (int)sizeof(pk_nlls_job); }

int pk_local_nlls_batch(pk_handle_t h, const pk_nlls_job* j) {
    if (!h || !j) return fail("null handle or job");
    int n, P, L;
    if (pk_local_dims(j->model, j->n_sites, j->T, &n, &P, &L)) return -1;
    if (j->B < 0) return fail("B < 0");
    if (!j->theta || !j->y0 || !j->t || !j->target || !j->lb || !j->ub) return fail("theta, y0, t, target, lb and ub are required");
    if (j->n_groups < 1) return fail("n_groups < 1");
    if (j->sigma && j->sigma_len != L && j->sigma_len != L + P) return fail("sigma_len must be L or L+P");
    if (j->T <= 5) return fail("the residual model needs T > 5 (flat layout, models/distmod.py:124-134)");
    if (j->max_iter < 1) return fail("max_iter < 1");
    for (int k = 0; k < P; ++k)
        if (!(j->lb[k] < j->ub[k]) || !std::isfinite(j->lb[k]) || !std::isfinite(j->ub[k]))
            return fail("bounds must be finite with lb < ub (normest.py:216-218)");
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (j->B == 0) return 0;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t B = (size_t)j->B, R = B * (size_t)(P + 1);
    const bool host = j->memspace == PK_HOST;
    if (B > 0x7fffffffull / (size_t)(P + 1)) return fail("pk_local_nlls_batch: B*(P+1) exceeds 2^31");
    const size_t G = (size_t)j->n_groups;
    const size_t y0_elems = j->y0_stride ? (B - 1) * (size_t)j->y0_stride + n : (size_t)n;

    // shared memory of the step kernel: one warp per problem
    const int Lp = L | 1;
    const size_t per_warp = ((size_t)P * Lp + L + (size_t)P * P + 4 * (size_t)P) * sizeof(double);
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    if (per_warp > (size_t)max_optin) return fail("pk_local_nlls_batch: problem too large for the step kernel (P*L)");
    const int wpc = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)max_optin / 2 / per_warp));
    CK(cudaFuncSetAttribute(pk::nlls_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wpc * per_warp)));

    Scratch ws;
    pk::NllsArgs a;
    memset(&a, 0, sizeof(a));
    a.B = j->B; a.P = P; a.L = L; a.sigma_len = j->sigma ? j->sigma_len : 0; a.log_params = j->log_params;
    a.lam = j->lam; a.fd_rel = j->fd_rel > 0 ? j->fd_rel : 1e-4;
    a.ftol = j->ftol; a.xtol = j->xtol; a.gtol = j->gtol;
    double *d_lb, *d_ub, *d_t, *d_y0, *d_target, *d_sigma = nullptr, *d_theta, *d_pert, *d_h, *d_flat, *d_ssr, *d_dscale,
           *d_trial, *d_tssr, *d_score = nullptr, *d_lamg = nullptr;
    int *d_group = nullptr, *d_pgroup = nullptr, *d_tgroup = nullptr, *d_sstat, *d_tstat, *d_run, *d_idx[2];
    pk::NllsState* d_state;
    CK(ws.get(&d_lb, P)); CK(ws.get(&d_ub, P)); CK(ws.get(&d_pert, R * P)); CK(ws.get(&d_h, B * P));
    CK(ws.get(&d_flat, R * L)); CK(ws.get(&d_ssr, R)); CK(ws.get(&d_dscale, B * P)); CK(ws.get(&d_trial, B * P));
    CK(ws.get(&d_tssr, B)); CK(ws.get(&d_sstat, R)); CK(ws.get(&d_tstat, B)); CK(ws.get(&d_run, 1));
    CK(ws.get(&d_state, B));
    if (j->group) { CK(ws.get(&d_pgroup, R)); CK(ws.get(&d_tgroup, B)); }
    CK(ws.get(&d_idx[0], B)); CK(ws.get(&d_idx[1], B));
    CK(cudaMemcpyAsync(d_lb, j->lb, P * sizeof(double), cudaMemcpyHostToDevice, st));      // lb/ub/t: always host
    CK(cudaMemcpyAsync(d_ub, j->ub, P * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(ws.get(&d_t, j->T));
    CK(cudaMemcpyAsync(d_t, j->t, (size_t)j->T * sizeof(double), cudaMemcpyHostToDevice, st));
    if (host) {
        CK(ws.get(&d_theta, B * P)); CK(ws.get(&d_y0, y0_elems)); CK(ws.get(&d_target, G * L));
        CK(cudaMemcpyAsync(d_theta, j->theta, B * P * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_y0, j->y0, y0_elems * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_target, j->target, G * L * sizeof(double), cudaMemcpyHostToDevice, st));
        if (j->sigma) {
            CK(ws.get(&d_sigma, G * (size_t)j->sigma_len));
            CK(cudaMemcpyAsync(d_sigma, j->sigma, G * (size_t)j->sigma_len * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (j->group) {
            CK(ws.get(&d_group, B));
            CK(cudaMemcpyAsync(d_group, j->group, B * sizeof(int), cudaMemcpyHostToDevice, st));
        }
        if (j->lam_group) {
            CK(ws.get(&d_lamg, G));
            CK(cudaMemcpyAsync(d_lamg, j->lam_group, G * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (j->out_score) CK(ws.get(&d_score, B));
    } else {
        d_theta = j->theta; d_y0 = (double*)j->y0; d_target = (double*)j->target; d_sigma = (double*)j->sigma;
        d_group = (int*)j->group; d_score = j->out_score; d_lamg = (double*)j->lam_group;
    }
    // per-system y0 rows follow their problem into the (compacted) perturbed and trial batches
    double *d_y0p = nullptr, *d_y0t = nullptr;
    if (j->y0_stride) {
        CK(ws.get(&d_y0p, R * (size_t)n));
        CK(ws.get(&d_y0t, B * (size_t)n));
    }
    a.lb = d_lb; a.ub = d_ub; a.target = d_target; a.sigma = d_sigma; a.group = d_group; a.lam_group = d_lamg;
    a.theta = d_theta; a.pert = d_pert; a.pert_group = d_pgroup; a.hstep = d_h; a.flat = d_flat; a.ssr = d_ssr;
    a.solve_status = d_sstat; a.dscale = d_dscale; a.trial = d_trial; a.trial_ssr = d_tssr; a.trial_status = d_tstat;
    a.st = d_state; a.n_running = d_run;
    a.trial_group = d_tgroup; a.y0 = d_y0; a.y0_stride = j->y0_stride; a.n = n; a.y0_pert = d_y0p; a.y0_trial = d_y0t;

    pk_local_job base;
    pk_local_job_init(&base);
    base.model = j->model; base.n_sites = j->n_sites; base.T = j->T; base.memspace = PK_DEVICE;
    base.t = d_t; base.rtol = j->rtol; base.atol = j->atol; base.max_steps = j->max_steps; base.log_params = j->log_params;
    base.method = j->method; base.target = d_target; base.sigma = d_sigma; base.n_groups = j->n_groups;
    base.sigma_len = j->sigma ? j->sigma_len : 0; base.lam = j->lam; base.lam_group = d_lamg;
    for (int i = 0; i < 5; ++i) base.score_w[i] = j->score_w[i];

    const int TB = 256;
    const unsigned gB = (unsigned)((B + TB - 1) / TB), gR = (unsigned)((R + TB - 1) / TB);
    // timing events of this call: destroyed on every exit path
    struct EventPair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } evp;
    CK(cudaEventCreate(&evp.a)); CK(cudaEventCreate(&evp.b));
    cudaEvent_t& e0 = evp.a; cudaEvent_t& e1 = evp.b;
    CK(cudaEventRecord(e0, st));
    pk::nlls_init_kernel<<<gB, TB, 0, st>>>(a, j->mu0 > 0 ? j->mu0 : 1e-3, d_idx[0]);
    int launches = 1, iters_done = 0;
    int rc = 0;
    size_t nA = B;                                    // problems still running
    (void)gR;
    for (int it = 0; it < j->max_iter; ++it) {
        const size_t RA = nA * (size_t)(P + 1);
        a.nA = (long long)nA; a.idx = d_idx[it & 1]; a.idx_next = d_idx[(it + 1) & 1];
        pk::nlls_perturb_kernel<<<(unsigned)((RA + TB - 1) / TB), TB, 0, st>>>(a);
        pk_local_job jj = base;
        jj.B = (int64_t)RA; jj.params = d_pert; jj.group = d_pgroup;
        jj.y0 = j->y0_stride ? d_y0p : d_y0; jj.y0_stride = j->y0_stride ? n : 0;
        jj.out_flat = d_flat; jj.out_ssr = d_ssr; jj.out_status = d_sstat;
        if ((rc = pk_local_solve_batch(h, &jj)) != 0) break;
        launches += 1 + h->last_launches;
        pk::nlls_step_kernel<<<(unsigned)((nA + wpc - 1) / wpc), wpc * 32, wpc * per_warp, st>>>(a, wpc, Lp);
        pk_local_job jt = base;
        jt.B = (int64_t)nA; jt.params = d_trial; jt.group = d_tgroup;
        jt.y0 = j->y0_stride ? d_y0t : d_y0; jt.y0_stride = j->y0_stride ? n : 0;
        jt.out_ssr = d_tssr; jt.out_status = d_tstat;
        if ((rc = pk_local_solve_batch(h, &jt)) != 0) break;
        launches += 1 + h->last_launches;
        CK(cudaMemsetAsync(d_run, 0, sizeof(int), st));
        pk::nlls_accept_kernel<<<(unsigned)((nA + TB - 1) / TB), TB, 0, st>>>(a, it + 1 == j->max_iter);
        ++launches;
        int running = 0;
        CK(cudaMemcpyAsync(&running, d_run, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        iters_done = it + 1;
        nA = (size_t)running;
        if (nA == 0) break;
    }
    if (rc != 0) return rc;
    // final evaluation at the optimum: cost and score_fit (normest.py:293-306 ranks the starts by score_fit)
    {
        pk_local_job jf = base;
        jf.B = j->B; jf.params = d_theta; jf.group = d_group; jf.y0 = d_y0; jf.y0_stride = j->y0_stride;
        jf.out_ssr = d_tssr; jf.out_status = d_tstat; jf.out_score = d_score;
        if ((rc = pk_local_solve_batch(h, &jf)) != 0) return rc;
        launches += h->last_launches;
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(std::string("pk_local_nlls_batch: ") + cudaGetErrorString(le));

    // results: theta (in place), cost = 0.5*ssr at theta, score, per-problem status / iterations / evaluations
    std::vector<pk::NllsState> hs(B);
    CK(cudaMemcpy(hs.data(), d_state, B * sizeof(pk::NllsState), cudaMemcpyDeviceToHost));
    std::vector<double> hssr(B);
    CK(cudaMemcpy(hssr.data(), d_tssr, B * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<double> cost(B);
    std::vector<int> stv(B), itv(B), nfv(B);
    for (size_t b = 0; b < B; ++b) { cost[b] = 0.5 * hssr[b]; stv[b] = hs[b].status; itv[b] = hs[b].iters; nfv[b] = hs[b].nfev; }
    const cudaMemcpyKind back = host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice;
    if (host) {
        CK(cudaMemcpy(j->theta, d_theta, B * P * sizeof(double), cudaMemcpyDeviceToHost));
        if (j->out_score) CK(cudaMemcpy(j->out_score, d_score, B * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (j->out_cost) CK(cudaMemcpy(j->out_cost, cost.data(), B * sizeof(double), back));
    if (j->out_status) CK(cudaMemcpy(j->out_status, stv.data(), B * sizeof(int), back));
    if (j->out_iters) CK(cudaMemcpy(j->out_iters, itv.data(), B * sizeof(int), back));
    if (j->out_nfev) CK(cudaMemcpy(j->out_nfev, nfv.data(), B * sizeof(int), back));
    h->last_launches = launches;
    h->last_ms = ms;
    (void)iters_done;
    return 0;
}

}  // extern "C"
