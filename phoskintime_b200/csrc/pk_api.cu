// C ABI of libphoskin_b200.so (see include/phoskin_b200.h).  Host side: handle, workspaces,
// staging copies, kernel dispatch by (model, n_sites), Morris elementary-effects reduction,
// FP64 peak probe and the NCCL all-gather (NCCL resolved with dlopen so that the library loads on
// hosts without it and reuses the copy PyTorch already mapped).
#include "pk_internal.hpp"

#include <cstdlib>

#include "pk_common.cuh"

namespace pkh {
thread_local std::string g_err;
}
using pkh::fail;
using pkh::DevBuf;
using pkh::ncclComm_t;

namespace {

// NCCL through dlopen
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllGather && CommDestroy;
    }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble

}  // namespace

// ---------------------------------------------------------------------------------- kernels
namespace pk {

// FP64 FMA peak probe: 8 independent register chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

// Residual tables of the thread-per-system kernels, one launch per job that asks for the fused loss:
//   isig[g][q] = 1 / sigma[g][q]                                    (q < sigma_len; only when sigma is given)
//   tw[g][k*n + i] = {target, weight} of trajectory entry (output k, state i) in TRAJECTORY order, weight = 1/sigma or 1;
//                    entries that are not part of flat (the RNA column before sol[5:, 0] starts) get {0, 0}
// so that a lane that lands on output k reads the n pairs of its row with ONE base address (16-byte loads).
__global__ void tps_prep_kernel(const double* target, const double* sigma, int G, int T, int n, int L, int sigma_len,
                                double* isig, double2* tw) {
    const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int TN = T * n;
    if (sigma && id < (long long)G * sigma_len) isig[id] = 1.0 / sigma[id];
    if (id < (long long)G * TN) {
        const int g = (int)(id / TN), r = (int)(id - (long long)g * TN);
        const int k = r / n, i = r - k * n;
        const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;
        const int fi = (i == 0) ? (k >= RNA_OFFSET ? k - RNA_OFFSET : -1) : rna_len + (i - 1) * T + k;
        double2 v = make_double2(0.0, 0.0);
        if (fi >= 0) {
            v.x = target[(size_t)g * L + fi];
            v.y = sigma ? 1.0 / sigma[(size_t)g * sigma_len + fi] : 1.0;
        }
        tw[id] = v;
    }
}

// ---- Morris elementary effects (SALib.analyze.morris as called at sensitivity/analysis.py:264)
// column statistics: block c < D reduces X[:,c]; block D reduces Y.  out[c] = population std.
__global__ void __launch_bounds__(256) morris_std_kernel(const double* X, const double* Y, long long rows, int D,
                                                         double* out_std) {
    __shared__ double sh[256];
    __shared__ double mean_s;
    const int c = blockIdx.x;
    const double* base = c < D ? X + c : Y;
    const long long stride = c < D ? D : 1;
    double s = 0.0;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) s += base[r * stride];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) mean_s = sh[0] / (double)rows;
    __syncthreads();
    const double m = mean_s;
    s = 0.0;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) { double d = base[r * stride] - m; s = fma(d, d, s); }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out_std[c] = sqrt(sh[0] / (double)rows);
}

// one thread per (trajectory r, move k): ee[r, moved coordinate]
__global__ void morris_ee_kernel(const double* X, const double* Y, long long N, int D, double inv_delta,
                                 int scaled, const double* stds, double* ee) {
    long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (id >= N * D) return;
    long long r = id / D;
    int k = (int)(id % D);
    const double* x0 = X + (r * (D + 1) + k) * D;
    const double* x1 = x0 + D;
    int which = 0;
    double best = -1.0, step = 0.0;
    for (int c = 0; c < D; ++c) {
        double d = x1[c] - x0[c];
        if (fabs(d) > best) { best = fabs(d); which = c; step = d; }
    }
    double dY = Y[r * (D + 1) + k + 1] - Y[r * (D + 1) + k];
    double v;
    if (scaled) v = dY / step * (stds[which] / stds[D]);
    else v = (step > 0.0 ? dY : -dY) * inv_delta;
    ee[r * D + which] = v;
}

// block c: statistics of ee[:, c] over N trajectories
__global__ void __launch_bounds__(256) morris_stats_kernel(const double* ee, long long N, int D, double* mu,
                                                           double* mu_star, double* sigma) {
    __shared__ double sh[256], sh2[256];
    __shared__ double mean_s;
    const int c = blockIdx.x;
    double s = 0.0, sa = 0.0;
    for (long long r = threadIdx.x; r < N; r += blockDim.x) { double v = ee[r * D + c]; s += v; sa += fabs(v); }
    sh[threadIdx.x] = s;
    sh2[threadIdx.x] = sa;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sh[threadIdx.x] += sh[threadIdx.x + o]; sh2[threadIdx.x] += sh2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        mean_s = sh[0] / (double)N;
        if (mu) mu[c] = mean_s;
        if (mu_star) mu_star[c] = sh2[0] / (double)N;
    }
    __syncthreads();
    const double m = mean_s;
    s = 0.0;
    for (long long r = threadIdx.x; r < N; r += blockDim.x) { double d = ee[r * D + c] - m; s = fma(d, d, s); }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0 && sigma) sigma[c] = N > 1 ? sqrt(sh[0] / (double)(N - 1)) : 0.0;
}

}  // namespace pk

// --------------------------------------------------------------------------------- dispatch
namespace {

// Register-resident kernel limits: the largest site counts that compile with ZERO local-memory
// spill (csrc/ptxas.log); larger systems take the shared-memory dense path.  The kernels are instantiated in their own
// translation units (pk_tps_dist.cu, pk_tps_succ.cu, pk_dense.cu) so that the library builds in parallel.
constexpr int TPS_MAX_NS_DIST = pkh::TPS_MAX_NS;
constexpr int TPS_MAX_NS_SUCC = pkh::TPS_MAX_NS;
constexpr int PIPE_MAX_CHUNKS = pk_handle_s::MAX_CHUNKS;
constexpr size_t PIPE_MIN_CHUNK = 100000;   // host-path batches >= 2x/4x this are pipelined in 2/4 chunks

int dims(int model, int ns, int T, int* n, int* P, int* L) {
    if (ns < 1) return fail("n_sites must be >= 1");
    if (model == PK_DISTMOD || model == PK_SUCCMOD) {
        if (ns > 60) return fail("n_sites too large");
        *n = 2 + ns;
        *P = 4 + 2 * ns;
    } else if (model == PK_RANDMOD) {
        if (ns > 7) return fail("randmod supports n_sites <= 7 (2^ns states in shared memory)");
        *n = 2 + (1 << ns) - 1;
        *P = 4 + ns + (1 << ns) - 1;
    } else {
        return fail("unknown model id");
    }
    *L = (T > pk::RNA_OFFSET ? T - pk::RNA_OFFSET : 0) + T + ns * T;
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

int pk_abi_version(void) { return PK_ABI_VERSION; }

// host evaluation of the run-time ROS5L(gamma) coefficient construction the dense kernel uses (for tests)
int pk_ros6l_coeffs(double gamma, double* mu7, double* eps7) {
    if (!(gamma > 0.0) || !mu7 || !eps7) return fail("pk_ros6l_coeffs: bad arguments");
    double scratch[64];
    pk::rosl_coeffs<7>(gamma, scratch);
    for (int k = 0; k < 7; ++k) { mu7[k] = scratch[48 + k]; eps7[k] = scratch[55 + k]; }
    return 0;
}
int pk_ros5l_coeffs(double gamma, double* mu6, double* eps6) {
    if (!(gamma > 0.0) || !mu6 || !eps6) return fail("pk_ros5l_coeffs: bad arguments");
    double scratch[48];
    pk::ros5l_coeffs(gamma, scratch);
    for (int k = 0; k < 6; ++k) { mu6[k] = scratch[36 + k]; eps6[k] = scratch[42 + k]; }
    return 0;
}
const char* pk_last_error(void) { return pkh::g_err.c_str(); }

int pk_device_count(int* out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *out = n;
    return 0;
}

int pk_create(int device, pk_handle_t* out) {
    if (!out) return fail("pk_create: out is NULL");
    *out = nullptr;
    CK(cudaSetDevice(device));
    pk_handle_s* h = new pk_handle_s();
    h->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    h->clock_khz = khz;
    snprintf(h->name, sizeof(h->name), "%.127s", prop.name);
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    {   // copy-out / collective stream at the highest priority: a collective that becomes runnable together with the
        // next piece's persistent compute grid must get its SM slots first (pk_local_solve_allgather)
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->s_out, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < pk_handle_s::MAX_CHUNKS; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreate(&h->evr0));
    CK(cudaEventCreate(&h->evr1));
    CK(cudaMalloc(&h->counter, sizeof(unsigned long long)));
    *out = h;
    return 0;
}

int pk_destroy(pk_handle_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    DevBuf* bufs[] = {&h->params, &h->y0, &h->t, &h->sol, &h->flat, &h->Y, &h->ssr, &h->score, &h->status,
                      &h->nsteps, &h->nrej, &h->target, &h->sigma, &h->group, &h->scratch, &h->traj, &h->ag_stage, &h->isig, &h->tw, &h->bar, &h->lamg};
    for (DevBuf* b : bufs) b->release();
    DevBuf* gbufs[] = {&h->g_params, &h->g_y0, &h->g_t, &h->g_stops, &h->g_Y, &h->g_loss, &h->g_F, &h->g_metric,
                       &h->g_status, &h->g_nsteps, &h->g_nrej, &h->g_traj, &h->g_binv, &h->g_fc, &h->g_ovf};
    for (DevBuf* b : gbufs) b->release();
    pkh::release_global_topologies(h);
    pk_sym_free(h);
    if (h->counter) cudaFree(h->counter);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evr0) cudaEventDestroy(h->evr0);
    if (h->evr1) cudaEventDestroy(h->evr1);
    for (int i = 0; i < pk_handle_s::MAX_CHUNKS; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int pk_device_info(pk_handle_t h, int* sm_count, int* clock_khz, char* name, int name_len) {
    if (!h) return fail("null handle");
    if (sm_count) *sm_count = h->sm_count;
    if (clock_khz) *clock_khz = h->clock_khz;
    if (name && name_len > 0) snprintf(name, name_len, "%s", h->name);
    return 0;
}

int pk_local_dims(int model, int n_sites, int T, int* n_states, int* n_params, int* flat_len) {
    int n, P, L;
    if (dims(model, n_sites, T, &n, &P, &L)) return -1;
    if (n_states) *n_states = n;
    if (n_params) *n_params = P;
    if (flat_len) *flat_len = L;
    return 0;
}

void pk_local_job_init(pk_local_job* job) {
    memset(job, 0, sizeof(*job));
    job->y_metric = PK_Y_NONE;
    job->n_groups = 1;
    for (int i = 0; i < 5; ++i) job->score_w[i] = 1.0;
}

int pk_sizeof_local_job(void) { return (int)sizeof(pk_local_job); }

int pk_region_begin(pk_handle_t h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->evr0, h->stream));
    return 0;
}

int pk_region_end(pk_handle_t h, float* ms) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->evr1, h->stream));
    CK(cudaEventSynchronize(h->evr1));
    float v = 0.f;
    CK(cudaEventElapsedTime(&v, h->evr0, h->evr1));
    if (ms) *ms = v;
    return 0;
}

int pk_last_launch_info(pk_handle_t h, int* n_launches, float* kernel_ms) {
    if (!h) return fail("null handle");
    if (n_launches) *n_launches = h->last_launches;
    if (kernel_ms) *kernel_ms = h->last_ms;
    return 0;
}

int pk_local_solve_batch(pk_handle_t h, const pk_local_job* j) {
    if (!h || !j) return fail("null handle or job");
    int n, P, L;
    if (dims(j->model, j->n_sites, j->T, &n, &P, &L)) return -1;
    if (j->B < 0) return fail("B < 0");
    if (j->T < 1) return fail("T < 1");
    if (!j->params || !j->y0 || !j->t) return fail("params, y0 and t are required");
    const bool want_loss = j->out_ssr || j->out_score;
    if (want_loss && !j->target) return fail("out_ssr/out_score need target");
    if (want_loss && j->n_groups < 1) return fail("n_groups < 1");
    if (want_loss && j->sigma && j->sigma_len != L && j->sigma_len != L + P)
        return fail("sigma_len must be L or L+P");
    if (j->out_Y && (j->y_metric < 0 || j->y_metric > 4)) return fail("out_Y needs a valid y_metric");
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (j->B == 0) return 0;
    CK(cudaSetDevice(h->device));
    const size_t B = (size_t)j->B;
    const bool host = j->memspace == PK_HOST;
    cudaStream_t st = h->stream;

    pk::LocalArgs a;
    memset(&a, 0, sizeof(a));
    a.B = j->B; a.T = j->T; a.ns = j->n_sites; a.n = n; a.P = P; a.L = L;
    if (j->method < 0 || j->method > 3) return fail("unknown method");
    const bool tps_path = (j->model == PK_DISTMOD && j->n_sites <= TPS_MAX_NS_DIST) ||
                          (j->model == PK_SUCCMOD && j->n_sites <= TPS_MAX_NS_SUCC);
    // default method: ROS6L (7 solves, order 6(5)) on the thread-per-system kernels.  The dense kernel supports it too
    // (run-time ROS6L(gamma') family, rosl_coeffs<7>) but keeps ROS5L as its default: measured on rand-6 the 31 % fewer
    // steps buy +5 % (U draws) to +33 % (harsh draws) throughput — inversions, not mat-vecs, dominate there — at 0.16-0.22
    // instead of 0.04-0.10 of the parity bound (tools/dense_method_scan.py).
    const bool ros6 = j->method == PK_METHOD_ROS6L || (j->method == PK_METHOD_DEFAULT && tps_path);
    a.m = (j->method == PK_METHOD_RODAS4) ? pk::METHOD_RODAS4 : (ros6 ? pk::METHOD_ROS6L : pk::METHOD_ROS5L);
    // default tolerances: chosen per method from the measured error against the reference's tight
    // solution — ROS5L/RODAS4 2e-6/2e-9, ROS6L 2e-5/2e-9 (error <= 0.15 of the parity bound at ~40 % fewer steps; the
    // absolute part stays at 2e-9: the parity bound's own absolute term is 1e-9, components that decay towards zero need it)
    a.rtol = j->rtol > 0 ? j->rtol : (ros6 ? 2e-5 : 2e-6);
    // ROS6L: atol 2e-11 — measured on the 1 M succ-5 bench systems (tools/dev/atol_scan.py): states down to 2e-8 occur, and the
    // max PURE relative error (north star: <= 1e-6) is 2.5e-5 at atol 2e-9, 2.1e-6 at 1e-10, 3.4e-7 at 2e-11, for 42.1 ->
    // 42.7 steps per solve (+2 % kernel time)
    a.atol = j->atol > 0 ? j->atol : (ros6 ? 2e-11 : 2e-9);
    // ROS6L: tolerance tightened to rtol/20 where the solution is (nearly) stationary — err_ratio_inc, pk_common.cuh.
    // (kappa, floor) scanned on B200 over 2^18 restarts near steady state (tools/dbg_semigroup.py): worst deviation
    // 1.25 / 0.82 / 0.33 of the parity bound at (1e-2, 1/10) / (4e-3, 1/10) / (4e-3, 1/20); ROS5L at 2e-6: 0.55.
    a.rtol_floor = ros6 ? 0.05 * a.rtol : a.rtol;
    a.kappa = ros6 ? 4e-3 : 0.0;
    // dense kernel: power-of-two step grid.  Measured (tools/dense_grid_scan.py, rand-6): ratio 2 -> 1.68e5 solves/s at 85 steps,
    // ratio 4 -> 1.76e5 at 119 steps (inversions saved ~ mat-vecs added), ratio 2.83 -> 1.67e5; ratios above the
    // controller's growth limit (6) stall.
    a.hgrid_log2 = 1.0;
    if (const char* e = getenv("PK_DENSE_HGRID")) a.hgrid_log2 = atof(e);      // development override
    a.max_steps = j->max_steps > 0 ? j->max_steps : 100000;
    a.normalize = j->normalize; a.log_params = j->log_params;
    a.y_metric = j->out_Y ? j->y_metric : -1;
    a.sigma_len = j->sigma ? j->sigma_len : 0;
    a.lam = j->lam;
    a.w_alpha = j->score_w[0]; a.w_beta = j->score_w[1]; a.w_gamma = j->score_w[2];
    a.w_delta = j->score_w[3]; a.w_mu = j->score_w[4];
    a.y0_stride = j->y0_stride;
    a.counter = h->counter;
    // the flat-index table of the trajectory-slot epilogue holds k*n + i as a short
    if (tps_path && (long long)j->T * n > 32767) return fail("T * n_states must be <= 32767 on the thread-per-system path");

    const size_t y0_elems = j->y0_stride ? (B - 1) * (size_t)j->y0_stride + n : (size_t)n;
    const size_t G = want_loss ? (size_t)j->n_groups : 0;
    const size_t TN = (size_t)j->T * n;

    // thread-per-system kernels: reciprocal weights and the {target, weight} table in trajectory order (tps_prep_kernel)
    auto tps_prep = [&](const double* d_target, const double* d_sigma) -> cudaError_t {
        const long long n_is = d_sigma ? (long long)G * j->sigma_len : 0, n_tw = (long long)G * TN;
        cudaError_t e = h->isig.ensure((size_t)(n_is + 1) * sizeof(double));
        if (e != cudaSuccess) return e;
        e = h->tw.ensure((size_t)n_tw * 2 * sizeof(double));
        if (e != cudaSuccess) return e;
        const long long nthr = std::max(n_is, n_tw);
        pk::tps_prep_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(d_target, d_sigma, (int)G, j->T, n, L, j->sigma_len,
                                                                          (double*)h->isig.p, (double2*)h->tw.p);
        a.isigma = d_sigma ? (const double*)h->isig.p : nullptr;
        a.tw = (const double2*)h->tw.p;
        a.n_groups = (int)G;
        return cudaGetLastError();
    };

    auto launch = [&](const pk::LocalArgs& ac) -> cudaError_t {
        const bool tps = tps_path;
        if (tps) return (j->model == PK_DISTMOD) ? pkh::launch_tps_dist(h, ac) : pkh::launch_tps_succ(h, ac);
        return pkh::launch_dense_model(h, ac, j->model == PK_DISTMOD ? 0 : (j->model == PK_SUCCMOD ? 1 : 2));
    };

    if (!host) {
        a.params = j->params; a.y0 = j->y0; a.t = j->t;
        a.target = j->target; a.sigma = j->sigma; a.group = j->group;
        a.lam_group = want_loss ? j->lam_group : nullptr;
        if (tps_path && want_loss) CK(tps_prep(a.target, a.sigma));
        if (h->p2p_which >= 0) {
            if (!tps_path) return fail("pk_local_solve_gather_p2p: only the thread-per-system kernels store to peer memory");
            a.n_peer = h->world;
            a.peer_which = h->p2p_which;
            a.peer_base = (long long)h->rank * h->sym_slots;
            for (int r = 0; r < h->world; ++r) a.peer[r] = h->sym_peer[r];
        }
        a.out_sol = j->out_sol; a.out_flat = j->out_flat; a.out_Y = j->out_Y; a.out_ssr = j->out_ssr;
        a.out_score = j->out_score; a.out_status = j->out_status; a.out_nsteps = j->out_nsteps;
        a.out_nrej = j->out_nrej;
        // pk_local_solve_allgather: integrate in pieces; the all-gather of piece c (copy stream) overlaps piece c+1
        double* const ag = h->ag_recv;
        int nch = ag ? std::min(h->ag_chunks, PIPE_MAX_CHUNKS) : 1;
        if ((size_t)nch > B) nch = (int)B;
        if (nch < 1) nch = 1;
        const size_t W = (size_t)h->world;
        if (ag && W > 1) CK(h->ag_stage.ensure(W * B * sizeof(double)));
        CK(cudaEventRecord(h->ev0, st));
        // piece boundaries: equal pieces for a plain solve; for the fused gather every piece is half the previous one
        // (2/3 + 1/3 for two pieces), so that the exposed gather of the LAST piece is short while the earlier, larger
        // gathers hide behind the following pieces' integration
        size_t bound[PIPE_MAX_CHUNKS + 1];
        bound[0] = 0;
        {
            const double total = ag ? (double)((1u << nch) - 1u) : (double)nch;
            double acc = 0.0;
            for (int c = 0; c < nch; ++c) {
                acc += ag ? (double)(1u << (nch - 1 - c)) : 1.0;
                bound[c + 1] = (c + 1 == nch) ? B : std::min(B, (size_t)((double)B * acc / total));
                if (bound[c + 1] <= bound[c] && bound[c] < B) bound[c + 1] = bound[c] + 1;   // no empty piece
            }
        }
        for (int c = 0; c < nch; ++c) {
            const size_t o = bound[c], cnt = bound[c + 1] - bound[c];
            if (cnt == 0) continue;
            pk::LocalArgs ac = a;
            ac.B = (long long)cnt;
            ac.params += o * P;
            ac.peer_base += (long long)o;
            if (j->y0_stride) ac.y0 += o * j->y0_stride;
            if (ac.group) ac.group += o;
            if (ac.out_sol) ac.out_sol += o * TN;
            if (ac.out_flat) ac.out_flat += o * L;
            if (ac.out_Y) ac.out_Y += o;
            if (ac.out_ssr) ac.out_ssr += o;
            if (ac.out_score) ac.out_score += o;
            if (ac.out_status) ac.out_status += o;
            if (ac.out_nsteps) ac.out_nsteps += o;
            if (ac.out_nrej) ac.out_nrej += o;
            CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned long long), st));
            cudaError_t e = launch(ac);
            if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
            h->last_launches++;
            if (ag) {
                const double* src = h->ag_which == 0 ? ac.out_score : (h->ag_which == 1 ? ac.out_ssr : ac.out_Y);
                CK(cudaEventRecord(h->ev_k[c], st));
                CK(cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
                if (W == 1) {
                    CK(cudaMemcpyAsync(ag + o, src, cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->s_out));
                } else {
                    double* stage = (double*)h->ag_stage.p + W * o;            // [world][cnt]
                    int r = g_nccl.AllGather(src, stage, cnt, NCCL_FLOAT64, h->comm, h->s_out);
                    if (r != 0) return fail(std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
                    // rank-major placement: row r of the landing area goes to recv[r*B + o .. + cnt)
                    CK(cudaMemcpy2DAsync(ag + o, B * sizeof(double), stage, cnt * sizeof(double), cnt * sizeof(double), W,
                                         cudaMemcpyDeviceToDevice, h->s_out));
                }
            }
        }
        CK(cudaEventRecord(h->ev1, st));
        if (h->p2p_which >= 0 && h->world > 1) {
            // Closing rendezvous of the peer-memory gather: a kernel's peer stores are performed when the kernel
            // completes, so once EVERY rank has passed this point every rank's buffer is complete.  One 8-byte NCCL
            // all-gather on the compute stream is that rendezvous (the payload itself never goes through NCCL).
            CK(h->bar.ensure(2 * (size_t)h->world * sizeof(double)));
            double* bb = (double*)h->bar.p;
            int r = g_nccl.AllGather(bb + h->world + h->rank, bb, 1, NCCL_FLOAT64, h->comm, st);
            if (r != 0) return fail(std::string("ncclAllGather (rendezvous): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
        }
        if (ag) {                                   // join the collective stream back into the compute stream
            CK(cudaEventRecord(h->ev_in[0], h->s_out));
            CK(cudaStreamWaitEvent(st, h->ev_in[0], 0));
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
        return 0;
    }

    // ---- host memory: stage through device workspaces.  Large batches are pipelined in chunks:
    // H2D of chunk c+1 (copy-in stream) overlaps the kernel of chunk c (compute stream), whose
    // results leave on the copy-out stream while chunk c+1 computes.
#define WS(buf, need, bytes, dstfield)                                                        \
    do {                                                                                      \
        if (need) { CK(h->buf.ensure(bytes)); dstfield = (decltype(dstfield))h->buf.p; }      \
    } while (0)
    WS(params, true, B * P * sizeof(double), a.params);
    WS(y0, true, y0_elems * sizeof(double), a.y0);
    WS(t, true, (size_t)j->T * sizeof(double), a.t);
    WS(target, want_loss, G * L * sizeof(double), a.target);
    WS(sigma, want_loss && j->sigma, G * (size_t)j->sigma_len * sizeof(double), a.sigma);
    WS(group, want_loss && j->group, B * sizeof(int32_t), a.group);
    WS(lamg, want_loss && j->lam_group, G * sizeof(double), a.lam_group);
    WS(sol, j->out_sol, B * TN * sizeof(double), a.out_sol);
    WS(flat, j->out_flat, B * L * sizeof(double), a.out_flat);
    WS(Y, j->out_Y, B * sizeof(double), a.out_Y);
    WS(ssr, j->out_ssr, B * sizeof(double), a.out_ssr);
    WS(score, j->out_score, B * sizeof(double), a.out_score);
    WS(status, j->out_status, B * sizeof(int32_t), a.out_status);
    WS(nsteps, j->out_nsteps, B * sizeof(int32_t), a.out_nsteps);
    WS(nrej, j->out_nrej, B * sizeof(int32_t), a.out_nrej);
#undef WS
    // Large batches are uploaded in chunks.  Round 1 (kernel 2.65 ms per 1 M succ-5 systems > upload ~2.1 ms: compute bound, the
    // FIRST upload is what nothing hides): 5 chunks, each 1.2x the previous one.  Since round 2 the kernel (1.66 ms) is faster
    // than the upload, so the pipeline is PCIe bound and the exposed part is the LAST chunk's kernel + read-back: 8 equal
    // chunks.  Measured at 1 M systems (tools/e2e_pipe_scan.py, ms per call): 5 x1.2 2.82, 6 equal 2.84, 7 equal 2.92,
    // 7 x1.1 2.78, 8 equal 2.72, 8 x0.95 2.90, 8 x1.1 2.89, 9 equal 2.88, 10 equal 3.06, 12 equal 3.61 (every chunk is a launch
    // with its own tail and ~10 API calls; shrinking chunks lose more to that than they gain at the end).
    const int nchunks = B >= 8 * PIPE_MIN_CHUNK ? 8 : (B >= 4 * PIPE_MIN_CHUNK ? 5 : (B >= 2 * PIPE_MIN_CHUNK ? 2 : 1));
    const double growth = nchunks == 5 ? 1.2 : 1.0;
    cudaStream_t sin = h->s_in, sout = h->s_out;
    // small shared inputs first, then the per-system inputs chunk by chunk on the copy-in stream
    CK(cudaMemcpyAsync((void*)a.t, j->t, (size_t)j->T * sizeof(double), cudaMemcpyHostToDevice, sin));
    if (!j->y0_stride) CK(cudaMemcpyAsync((void*)a.y0, j->y0, n * sizeof(double), cudaMemcpyHostToDevice, sin));
    if (want_loss) {
        CK(cudaMemcpyAsync((void*)a.target, j->target, G * L * sizeof(double), cudaMemcpyHostToDevice, sin));
        if (j->sigma)
            CK(cudaMemcpyAsync((void*)a.sigma, j->sigma, G * (size_t)j->sigma_len * sizeof(double),
                               cudaMemcpyHostToDevice, sin));
        if (j->lam_group)
            CK(cudaMemcpyAsync((void*)a.lam_group, j->lam_group, G * sizeof(double), cudaMemcpyHostToDevice, sin));
        CK(cudaEventRecord(h->ev_in[PIPE_MAX_CHUNKS - 1], sin));
    }
    size_t lo[PIPE_MAX_CHUNKS + 1];
    {
        double total = 0.0, w = 1.0, acc = 0.0;
        for (int c = 0; c < nchunks; ++c) { total += w; w *= growth; }
        w = 1.0;
        lo[0] = 0;
        for (int c = 0; c < nchunks; ++c) {
            acc += w;
            w *= growth;
            lo[c + 1] = (c + 1 == nchunks) ? B : std::min(B, (size_t)((double)B * acc / total));
        }
    }
    for (int c = 0; c < nchunks; ++c) {
        const size_t o = lo[c], cnt = lo[c + 1] - lo[c];
        CK(cudaMemcpyAsync((double*)a.params + o * P, j->params + o * P, cnt * P * sizeof(double),
                           cudaMemcpyHostToDevice, sin));
        if (j->y0_stride)
            CK(cudaMemcpyAsync((double*)a.y0 + o * j->y0_stride, j->y0 + o * j->y0_stride,
                               ((cnt - 1) * (size_t)j->y0_stride + n) * sizeof(double), cudaMemcpyHostToDevice, sin));
        if (a.group)
            CK(cudaMemcpyAsync((int*)a.group + o, j->group + o, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, sin));
        CK(cudaEventRecord(h->ev_in[c], sin));
    }
    if (tps_path && want_loss) {
        CK(cudaStreamWaitEvent(st, h->ev_in[PIPE_MAX_CHUNKS - 1], 0));       // recorded right after the target / sigma upload
        CK(tps_prep(a.target, a.sigma));
    }
    CK(cudaEventRecord(h->ev0, st));
    for (int c = 0; c < nchunks; ++c) {
        const size_t o = lo[c], cnt = lo[c + 1] - lo[c];
        pk::LocalArgs ac = a;
        ac.B = (long long)cnt;
        ac.params += o * P;
        if (j->y0_stride) ac.y0 += o * j->y0_stride;
        if (ac.group) ac.group += o;
        if (ac.out_sol) ac.out_sol += o * TN;
        if (ac.out_flat) ac.out_flat += o * L;
        if (ac.out_Y) ac.out_Y += o;
        if (ac.out_ssr) ac.out_ssr += o;
        if (ac.out_score) ac.out_score += o;
        if (ac.out_status) ac.out_status += o;
        if (ac.out_nsteps) ac.out_nsteps += o;
        if (ac.out_nrej) ac.out_nrej += o;
        CK(cudaStreamWaitEvent(st, h->ev_in[c], 0));
        CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned long long), st));
        cudaError_t e = launch(ac);
        if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
        CK(cudaEventRecord(h->ev_k[c], st));
        h->last_launches++;
        CK(cudaStreamWaitEvent(sout, h->ev_k[c], 0));
#define COPY_OUT(field, user, per)                                                            \
        do {                                                                                  \
            if (user) CK(cudaMemcpyAsync(user + o * (per), ac.field, cnt * (per) * sizeof(*user), \
                                         cudaMemcpyDeviceToHost, sout));                      \
        } while (0)
        COPY_OUT(out_sol, j->out_sol, TN);
        COPY_OUT(out_flat, j->out_flat, (size_t)L);
        COPY_OUT(out_Y, j->out_Y, (size_t)1);
        COPY_OUT(out_ssr, j->out_ssr, (size_t)1);
        COPY_OUT(out_score, j->out_score, (size_t)1);
        COPY_OUT(out_status, j->out_status, (size_t)1);
        COPY_OUT(out_nsteps, j->out_nsteps, (size_t)1);
        COPY_OUT(out_nrej, j->out_nrej, (size_t)1);
#undef COPY_OUT
    }
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(sout));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

int pk_morris_ee(pk_handle_t h, int memspace, const double* X, const double* Y, int64_t N, int32_t D,
                 int32_t num_levels, int32_t scaled, double* out_mu, double* out_mu_star, double* out_sigma,
                 double* out_ee) {
    if (!h || !X || !Y) return fail("null argument");
    if (N < 1 || D < 1 || num_levels < 2) return fail("bad N, D or num_levels");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t rows = (size_t)N * (D + 1);
    const bool host = memspace == PK_HOST;
    // scratch layout: [X | Y] (host mode) | ee[N*D] | stds[D+1] | mu[D] | mu*[D] | sigma[D]
    size_t off_x = 0, off_y = host ? rows * D : 0, off_ee = host ? rows * (D + 1) : 0;
    size_t off_std = off_ee + (size_t)N * D, off_mu = off_std + D + 1;
    size_t total = off_mu + 3 * (size_t)D;
    CK(h->scratch.ensure(total * sizeof(double)));
    double* base = (double*)h->scratch.p;
    const double* dX = X;
    const double* dY = Y;
    if (host) {
        CK(cudaMemcpyAsync(base + off_x, X, rows * D * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(base + off_y, Y, rows * sizeof(double), cudaMemcpyHostToDevice, st));
        dX = base + off_x;
        dY = base + off_y;
    }
    double* ee = base + off_ee;
    double* stds = base + off_std;
    double* mu = base + off_mu;
    double* mus = mu + D;
    double* sig = mus + D;
    CK(cudaMemsetAsync(ee, 0, (size_t)N * D * sizeof(double), st));
    CK(cudaEventRecord(h->ev0, st));
    h->last_launches = 0;
    if (scaled) {
        pk::morris_std_kernel<<<D + 1, 256, 0, st>>>(dX, dY, (long long)rows, D, stds);
        h->last_launches++;
    }
    const double inv_delta = 2.0 * (num_levels - 1) / (double)num_levels;
    long long nthreads = (long long)N * D;
    pk::morris_ee_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(dX, dY, N, D, inv_delta, scaled, stds, ee);
    pk::morris_stats_kernel<<<D, 256, 0, st>>>(ee, N, D, mu, mus, sig);
    h->last_launches += 2;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, st));
    cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (out_mu) CK(cudaMemcpyAsync(out_mu, mu, D * sizeof(double), kind, st));
    if (out_mu_star) CK(cudaMemcpyAsync(out_mu_star, mus, D * sizeof(double), kind, st));
    if (out_sigma) CK(cudaMemcpyAsync(out_sigma, sig, D * sizeof(double), kind, st));
    if (out_ee) CK(cudaMemcpyAsync(out_ee, ee, (size_t)N * D * sizeof(double), kind, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

int pk_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail("bad argument");
    CK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
int pk_host_free(void* ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return 0;
}

int pk_measure_fp64_peak(pk_handle_t h, double* tflops, float* ms_out) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    CK(h->scratch.ensure(64));
    const int iters = 4096, block = 256;
    const int grid = h->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(h->ev0, h->stream));
        pk::dfma_peak_kernel<<<grid, block, 0, h->stream>>>((double*)h->scratch.p, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    double flops = 2.0 * 64.0 * (double)iters * (double)grid * block;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

int pk_nccl_unique_id(char* out128) {
    if (!g_nccl.load()) return fail("libnccl.so.2 not found");
    ncclUniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    memcpy(out128, id.internal, 128);
    return 0;
}

int pk_nccl_init(pk_handle_t h, const char* id128, int world, int rank) {
    if (!h) return fail("null handle");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad world/rank");
    h->world = world;
    h->rank = rank;
    if (world == 1) return 0;
    if (!g_nccl.load()) return fail("libnccl.so.2 not found");
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    int r = g_nccl.CommInitRank(&h->comm, world, id, rank);
    if (r != 0) return fail(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    return 0;
}

int pk_local_solve_allgather(pk_handle_t h, const pk_local_job* job, int32_t which, int32_t chunks, double* recv_dev) {
    if (!h || !job || !recv_dev) return fail("pk_local_solve_allgather: null argument");
    if (job->memspace != PK_DEVICE) return fail("pk_local_solve_allgather needs device buffers");
    if (which < 0 || which > 2) return fail("pk_local_solve_allgather: which must be 0 (score), 1 (ssr) or 2 (Y)");
    if ((which == 0 && !job->out_score) || (which == 1 && !job->out_ssr) || (which == 2 && !job->out_Y))
        return fail("pk_local_solve_allgather: the gathered output must be requested in the job");
    if (h->world > 1 && !h->comm) return fail("pk_nccl_init was not called");
    h->ag_recv = recv_dev;
    h->ag_which = which;
    h->ag_chunks = chunks > 0 ? chunks : 4;
    const int rc = pk_local_solve_batch(h, job);
    h->ag_recv = nullptr;
    return rc;
}

// ---- peer-memory gather: symmetric buffers mapped through CUDA IPC, filled by the solve kernel itself
int pk_sym_free(pk_handle_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (int r = 0; r < pk_handle_s::MAX_PEERS; ++r) {
        if (h->sym_peer[r] && h->sym_peer[r] != h->sym_local) cudaIpcCloseMemHandle(h->sym_peer[r]);
        h->sym_peer[r] = nullptr;
    }
    if (h->sym_local) cudaFree(h->sym_local);
    h->sym_local = nullptr;
    h->sym_slots = 0;
    h->sym_mapped = false;
    return 0;
}

int pk_sym_alloc(pk_handle_t h, int64_t slots_per_rank, char* handle_out64) {
    if (!h || slots_per_rank < 1 || !handle_out64) return fail("pk_sym_alloc: bad argument");
    if (h->world > pk_handle_s::MAX_PEERS) return fail("pk_sym_alloc: at most 8 ranks (one NVSwitch domain)");
    pk_sym_free(h);
    CK(cudaSetDevice(h->device));
    CK(cudaMalloc(&h->sym_local, (size_t)h->world * (size_t)slots_per_rank * sizeof(double)));
    h->sym_slots = slots_per_rank;
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == 64, "CUDA IPC handles are 64 bytes");
    CK(cudaIpcGetMemHandle(&mh, h->sym_local));
    memcpy(handle_out64, &mh, 64);
    if (h->world == 1) { h->sym_peer[0] = h->sym_local; h->sym_mapped = true; }
    return 0;
}

int pk_sym_open(pk_handle_t h, const char* handles) {
    if (!h || !handles) return fail("pk_sym_open: bad argument");
    if (!h->sym_local) return fail("pk_sym_open: pk_sym_alloc first");
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { h->sym_peer[r] = h->sym_local; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->sym_peer[r] = (double*)p;
    }
    h->sym_mapped = true;
    return 0;
}

int pk_sym_buffer(pk_handle_t h, double** local_dev, int64_t* slots_per_rank) {
    if (!h) return fail("null handle");
    if (local_dev) *local_dev = h->sym_local;
    if (slots_per_rank) *slots_per_rank = h->sym_slots;
    return 0;
}

int pk_local_solve_gather_p2p(pk_handle_t h, const pk_local_job* job, int32_t which) {
    if (!h || !job) return fail("pk_local_solve_gather_p2p: null argument");
    if (job->memspace != PK_DEVICE) return fail("pk_local_solve_gather_p2p needs device buffers");
    if (which < 0 || which > 2) return fail("pk_local_solve_gather_p2p: which must be 0 (score), 1 (ssr) or 2 (Y)");
    if ((which == 0 && !job->out_score) || (which == 1 && !job->out_ssr) || (which == 2 && !job->out_Y))
        return fail("pk_local_solve_gather_p2p: the gathered output must be requested in the job");
    if (!h->sym_mapped) return fail("pk_local_solve_gather_p2p: pk_sym_alloc / pk_sym_open first");
    if (job->B > h->sym_slots) return fail("pk_local_solve_gather_p2p: job->B exceeds the slots per rank of the symmetric buffer");
    if (h->world > 1 && !h->comm) return fail("pk_nccl_init was not called");
    h->p2p_which = which;
    const int rc = pk_local_solve_batch(h, job);
    h->p2p_which = -1;
    return rc;
}

int pk_allgather_f64(pk_handle_t h, const double* send_dev, int64_t count, double* recv_dev) {
    if (!h || !send_dev || !recv_dev || count < 0) return fail("bad argument");
    CK(cudaSetDevice(h->device));
    if (h->world == 1) {
        if (send_dev != recv_dev)
            CK(cudaMemcpyAsync(recv_dev, send_dev, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (!h->comm) return fail("pk_nccl_init was not called");
    int r = g_nccl.AllGather(send_dev, recv_dev, (size_t)count, NCCL_FLOAT64, h->comm, h->stream);
    if (r != 0) return fail(std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
