// Host-side internals shared by the translation units of libphoskin_b200.so: error string,
// device workspace buffers, the NCCL entry points (resolved with dlopen) and the handle.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phoskin_b200.h"

namespace pkh {

extern thread_local std::string g_err;

inline int fail(const std::string& m) {
    g_err = m;
    return -1;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return pkh::fail(std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

typedef struct ncclComm* ncclComm_t;

struct GlobalTopoHost;   // pk_global.cu

}  // namespace pkh

struct pk_handle_s {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    static constexpr int MAX_CHUNKS = 16;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evr0 = nullptr, evr1 = nullptr;
    pkh::DevBuf params, y0, t, sol, flat, Y, ssr, score, status, nsteps, nrej, target, sigma, group, scratch, traj, isig, tw, lamg;
    pkh::DevBuf g_params, g_y0, g_t, g_stops, g_Y, g_loss, g_F, g_metric, g_status, g_nsteps, g_nrej, g_traj, g_binv, g_fc, g_ovf;
    unsigned long long* counter = nullptr;
    int last_launches = 0;
    float last_ms = 0.f;
    pkh::DevBuf ag_stage;                      // pk_local_solve_allgather: [world][chunk] landing area of one piece
    double* ag_recv = nullptr;                 // set for the duration of one pk_local_solve_allgather call
    int ag_which = 0, ag_chunks = 1;
    pkh::ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    // symmetric result buffer of the peer-memory gather (pk_sym_*, pk_local_solve_gather_p2p): every rank owns
    // [world][sym_slots] doubles and maps every peer's copy through CUDA IPC
    static constexpr int MAX_PEERS = 8;
    double* sym_local = nullptr;
    double* sym_peer[MAX_PEERS] = {};
    long long sym_slots = 0;
    bool sym_mapped = false;
    int p2p_which = -1;                        // >= 0 for the duration of one pk_local_solve_gather_p2p call
    pkh::DevBuf bar;                           // 2 * world doubles: the closing rendezvous of the peer-memory gather
    std::vector<pkh::GlobalTopoHost*> topos;   // uploaded global networks (index = topology id)
};

namespace pk {
struct LocalArgs;
}
namespace pkh {
void release_global_topologies(pk_handle_s* h);   // pk_global.cu
constexpr int TPS_MAX_NS = 8;                     // thread-per-system kernels: 1..8 sites (pk_tps_launch.cuh)
cudaError_t launch_tps_dist(pk_handle_s* h, const pk::LocalArgs& a);               // pk_tps_dist.cu
cudaError_t launch_tps_succ(pk_handle_s* h, const pk::LocalArgs& a);               // pk_tps_succ.cu
cudaError_t launch_dense_model(pk_handle_s* h, const pk::LocalArgs& a, int model); // pk_dense.cu (0 dist, 1 succ, 2 rand)
}
