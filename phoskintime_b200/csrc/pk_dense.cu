// Dense (shared-memory matrix) kernels: random model and the distributive / successive models beyond 8 sites.
#include "pk_internal.hpp"

#include <cstdlib>

#include "local_dense.cuh"

namespace pkh {
namespace {

template <int MODEL, int NT, int TR = 0, int TC = 0>
cudaError_t launch_dense_nt(pk_handle_s* h, const pk::LocalArgs& a) {
    pk::DenseLayout lay;
    lay.n = a.n;
    lay.ld = (a.n & 1) ? a.n : a.n + 1;   // odd leading dimension: conflict-free column walks
    lay.P = a.P;
    lay.nobs = 2 + a.ns;
    constexpr int NW = NT / 32;
    lay.xtra = TR ? (2 + NW) * 32 * TR + 6 * NT : 0;   // exchange buffers + one EpiAcc (6 doubles) per thread
    lay.nv = TR ? NW * TC : a.n;
    if (TR && (a.n > NW * TC || a.n > 32 * TR)) return cudaErrorInvalidValue;
    size_t smem = (size_t)lay.total() * sizeof(double);
    auto kern = pk::local_dense_kernel<MODEL, NT, TR, TC>;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, NT, smem, h->stream>>>(a, lay);
    return cudaGetLastError();
}

// one system per 1, 2 or 4 warps: the shared-memory matrix limits the systems resident on an SM (5 at 65 states),
// so larger systems get more threads each to keep the SM's issue slots busy
template <int MODEL>
cudaError_t launch_dense(pk_handle_s* h, const pk::LocalArgs& a) {
    // 41..68 states (rand-6: 65): the inverse lives in registers (3 x 17 tile per thread), see local_dense.cuh
    if constexpr (MODEL == 2) {
        // (8 warps per system with a 3 x 9 tile — the functions are generic in the warp count — measured SLOWER: 3.09e5
        //  against 3.57e5 solves/s on rand-6; the per-column overhead is paid by twice as many warps for half the FMAs each)
        if (a.n > 40 && a.n <= 68 && !getenv("PK_DENSE_SMEM")) return launch_dense_nt<MODEL, 128, 3, 17>(h, a);
    }
    if (a.n >= 40) return launch_dense_nt<MODEL, 128>(h, a);
    if (a.n >= 24) return launch_dense_nt<MODEL, 64>(h, a);
    return launch_dense_nt<MODEL, 32>(h, a);
}

}  // namespace

cudaError_t launch_dense_model(pk_handle_s* h, const pk::LocalArgs& a, int model) {
    if (model == 0) return launch_dense<0>(h, a);
    if (model == 1) return launch_dense<1>(h, a);
    return launch_dense<2>(h, a);
}

}  // namespace pkh
