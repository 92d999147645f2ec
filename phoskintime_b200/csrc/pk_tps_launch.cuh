// Host-side launch of the thread-per-system kernel (shared by pk_tps_dist.cu / pk_tps_succ.cu: one translation unit per
// model family so that the library builds in parallel).
#pragma once
#include "pk_internal.hpp"

#include "local_tps.cuh"

namespace pkh {

// Launch configuration per model size: resident CTAs per SM the register allocator must allow with ZERO spill and no
// stack frame (checked in ptxas.log by the Makefile), and whether the model's rate coefficients live in shared memory
// (SmemCoef: 2P registers less per lane).  n <= 5 states fit 128 registers with everything in registers (4 CTAs x 128
// lanes); 6-7 states reach 4 CTAs with the coefficients in shared memory (16 instead of 12 warps per SM for the
// latency-bound step loop); larger systems keep 2 CTAs.
struct TpsCfg { int min_blocks; bool csmem; };
template <class M> constexpr TpsCfg tps_cfg() {
#ifdef PK_TPS_CFG_OVERRIDE
    return TpsCfg{PK_TPS_CFG_OVERRIDE};
#else
    return M::N <= 4 ? TpsCfg{4, false} : (M::N <= 7 ? TpsCfg{4, true} : TpsCfg{2, false});
#endif
}

template <class M, bool SCALAR>
cudaError_t launch_tps_mode(pk_handle_s* h, pk::LocalArgs a) {
    using pk::TPS_BLOCK;
    constexpr TpsCfg cfg = tps_cfg<M>();
    const size_t smem = pk::tps_smem_bytes(M::P, a.T, a.L, SCALAR, cfg.csmem);
    auto kern = pk::local_tps_kernel<M, cfg.min_blocks, SCALAR, cfg.csmem>;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPS_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long need = (a.B + TPS_BLOCK - 1) / TPS_BLOCK;
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (!SCALAR) {
        // one [T][n] trajectory slot per resident lane (L2-resident, reused for every system of the lane)
        e = h->traj.ensure((size_t)grid * TPS_BLOCK * a.T * M::N * sizeof(double));
        if (e != cudaSuccess) return e;
        a.traj = (double*)h->traj.p;
    }
    kern<<<(unsigned)grid, TPS_BLOCK, smem, h->stream>>>(a);
    return cudaGetLastError();
}

// SCALAR instantiation: only per-system scalars leave the kernel (residual sums XOR Morris-Y sums accumulated in
// registers); everything else takes the trajectory-slot instantiation with the warp-cooperative epilogue.
template <class M>
cudaError_t launch_tps(pk_handle_s* h, pk::LocalArgs a) {
    const bool want_loss = a.out_ssr || a.out_score;
    const bool want_y = a.out_Y != nullptr;
    const bool scalar = !a.out_sol && !a.out_flat && !a.normalize && (want_loss != want_y) && !(want_y && a.y_metric == 3);
    if (scalar) {
        a.ymode = want_y ? 1 : 0;
        return launch_tps_mode<M, true>(h, a);
    }
    a.ymode = 0;
    return launch_tps_mode<M, false>(h, a);
}

template <template <int> class M>
cudaError_t dispatch_tps(pk_handle_s* h, const pk::LocalArgs& a) {
    switch (a.ns) {
        case 1: return launch_tps<M<1>>(h, a);
        case 2: return launch_tps<M<2>>(h, a);
        case 3: return launch_tps<M<3>>(h, a);
        case 4: return launch_tps<M<4>>(h, a);
        case 5: return launch_tps<M<5>>(h, a);
        case 6: return launch_tps<M<6>>(h, a);
        case 7: return launch_tps<M<7>>(h, a);
        case 8: return launch_tps<M<8>>(h, a);
        default: return cudaErrorInvalidValue;
    }
}
static_assert(TPS_MAX_NS == 8, "dispatch_tps covers 1..8 sites");

}  // namespace pkh
