// Thread-per-system Rosenbrock kernel for the small local models (distributive, successive).
//
// One lane integrates one system with its whole working set (y, the Krylov vector v, y_new, err
// and the factorised I - h*gamma*M) in registers; the Jacobian structure is exploited analytically:
//   distributive (models/distmod.py:57-63): arrow matrix  -> Schur pivot on the protein row
//   successive   (models/succmod.py:33-90): tridiagonal in (P, site_1..site_ns) -> TWISTED factorisation: eliminated
//                from both ends towards a middle row and substituted back outwards, so that every solve carries two
//                independent dependency chains of half the length (the kernel is bound by DFMA latency, not by the
//                FP64 pipe); the pivots come from the two continuant recurrences, so they are independent too
// mRNA (row 0) is decoupled in both and eliminated first.  All reciprocals of one factorisation come
// from ONE reciprocal (batch inversion by prefix products; MUFU seed + two Newton steps); the error ratio and the
// step-size controller run in FP32.
// Work distribution: every warp claims batches of up to 32 consecutive systems with ONE atomicAdd and prefetches their
// parameter rows into a double-buffered shared-memory stash with cp.async while its lanes are still integrating; a lane
// that finishes takes the next system from the stash (no global atomic, no exposed HBM latency).  Near the end of the
// queue the batches shrink (guided self-scheduling) so that no warp sits on a long private tail.
// Outputs: two instantiations per model.
//   SCALAR   only per-system scalars are requested (ssr / score_fit / Morris Y): the lane accumulates the weighted
//            residual sums in registers at the moment it lands on an output time; finished systems are parked in a
//            per-warp shared-memory ring and up to 32 of them at a time are finished (t = 0 terms, square roots,
//            divisions, stores) by 32 lanes in lock-step.  No trajectory ever leaves the registers.
//   !SCALAR  sol / flat rows requested: the lane stores its raw state into a private L2-resident trajectory slot and
//            the whole warp runs the epilogue of a finished system (coalesced rows, shuffle reductions).
#pragma once
#include <type_traits>

#include "pk_common.cuh"

namespace pk {

// a[i] <- 1/a[i] for i < M with one reciprocal (Montgomery's trick).
template <int M>
__device__ __forceinline__ void batch_invert(double (&a)[M]) {
    double pre[M];
    pre[0] = a[0];
#pragma unroll
    for (int i = 1; i < M; ++i) pre[i] = pre[i - 1] * a[i];
    double inv = fast_rcp(pre[M - 1]);
#pragma unroll
    for (int i = M - 1; i > 0; --i) {
        double ai = a[i];
        a[i] = inv * pre[i - 1];
        inv *= ai;
    }
    a[0] = inv;
}

// ------------------------------------------------------------------------------------ models
// Each model provides rhs(y) = M y + b, factor(c) of A = I - c M (c = h*gamma) and an in-place solve A x = r.
// Its P derived rate coefficients live either in registers (RegCoef) or in a per-lane shared-memory column (SmemCoef:
// 2P registers less per lane, which buys the resident warps the latency-bound step loop needs; the coefficients are
// read once per step, conflict-free at [coefficient][lane]).
constexpr int TPS_BLOCK = 128;
template <int NC>
struct RegCoef {
    double v[NC];
    __device__ __forceinline__ double get(int k) const { return v[k]; }
    __device__ __forceinline__ void set(int k, double x) { v[k] = x; }
};
struct SmemCoef {
    double* p;                       // &coef[0][lane]; coefficient k at p[k * TPS_BLOCK]
    __device__ __forceinline__ double get(int k) const { return p[k * TPS_BLOCK]; }
    __device__ __forceinline__ void set(int k, double x) { p[k * TPS_BLOCK] = x; }
};

template <int NS_>
struct DistModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NF = 2 * NS_ + 4, NC = P;
    // coefficients: [A, B, C, kP = D + sum S, S_i (NS), k_i = 1 + D_i (NS)]
    static constexpr int cA = 0, cB = 1, cC = 2, cKP = 3, cS = 4, cK = 4 + NS_;
    template <class Co>
    __device__ static __forceinline__ void load(Co& c, const double* p) {
        c.set(cA, p[0]); c.set(cB, p[1]); c.set(cC, p[2]);
        double sS = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) { c.set(cS + i, p[4 + i]); sS += p[4 + i]; c.set(cK + i, 1.0 + p[4 + NS + i]); }
        c.set(cKP, p[3] + sS);
    }
    template <class Co>
    __device__ static __forceinline__ void rhs(const Co& c, const double (&y)[N], double (&f)[N]) {
        f[0] = fma(-c.get(cB), y[0], c.get(cA));
        double acc = fma(c.get(cC), y[0], -c.get(cKP) * y[1]);
#pragma unroll
        for (int i = 0; i < NS; ++i) { acc += y[2 + i]; f[2 + i] = fma(c.get(cS + i), y[1], -c.get(cK + i) * y[2 + i]); }
        f[1] = acc;
    }
    // rows: (1+cB) x0 = r0 ; -cC x0 + (1+c kP) x1 - c sum x_{2+i} = r1 ; -c S_i x1 + q_i x_{2+i} = r_{2+i}
    // F = [1/q0, 1/pivot, cC, c, 1/q_i (NS), c S_i / q_i (NS)]
    template <class Co>
    __device__ static __forceinline__ void factor(const Co& co, double c, double (&F)[NF]) {
        double q[NS], pre[NS], suf[NS], S[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) { q[i] = fma(c, co.get(cK + i), 1.0); S[i] = co.get(cS + i); }
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
        double inv[3] = {fma(c, co.get(cB), 1.0), Q, fma(fma(c, co.get(cKP), 1.0), Q, -(c * c) * sq)};   // q0, Q, pivot*Q
        batch_invert<3>(inv);
        F[0] = inv[0];
        F[1] = Q * inv[2];
        F[2] = c * co.get(cC);
        F[3] = c;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double iq = pre[i] * inv[1];
            F[4 + i] = iq;
            F[4 + NS + i] = c * S[i] * iq;
        }
    }
    __device__ static __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) {
        x[0] *= F[0];
        double sz0 = 0.0, sz1 = 0.0;                        // two partial sums: half the dependent chain
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            x[2 + i] *= F[4 + i];
            if (i & 1) sz1 += x[2 + i]; else sz0 += x[2 + i];
        }
        x[1] = fma(F[3], sz0 + sz1, fma(F[2], x[0], x[1])) * F[1];
#pragma unroll
        for (int i = 0; i < NS; ++i) x[2 + i] = fma(F[4 + NS + i], x[1], x[2 + i]);
    }
};

template <int NS_>
struct SuccModel {
    static constexpr int NS = NS_, N = NS_ + 2, P = 4 + 2 * NS_, NC = P;
    // Twisted factorisation of the tridiagonal block over the unknowns j = 0..NS (x[1+j]): rows 0..MID-1 are eliminated
    // downwards, rows NS..MID+1 upwards, the middle row MID last; back substitution runs outwards from MID.
    static constexpr int MID = (NS + 1) / 2, NT = MID, NB = NS - MID;
    // F = [1/q0, cC, c, 1/d_MID, 1/p_j (NT), lt_j = c S_j / p_j (NT: row j into row j+1),
    //      1/q_j (NB: j = MID+1..NS), lb_j = c / q_j (NB: row j into row j-1), c S_{j-1} (NB: back substitution downwards)]
    static constexpr int NF = 4 + 2 * NT + 3 * NB;
    static constexpr int O_IP = 4, O_LT = 4 + NT, O_IQ = 4 + 2 * NT, O_LB = 4 + 2 * NT + NB, O_CS = 4 + 2 * NT + 2 * NB;
    // coefficients: [A, B, C, S_i (NS), d_j (NS+1)] with d_0 = D + S_0 (protein), d_{1+i} = 1 + Dr_i + S_{i+1} (site i; no S
    // term for the last site)
    static constexpr int cA = 0, cB = 1, cC = 2, cS = 3, cD = 3 + NS_;
    template <class Co>
    __device__ static __forceinline__ void load(Co& c, const double* p) {
        c.set(cA, p[0]); c.set(cB, p[1]); c.set(cC, p[2]);
#pragma unroll
        for (int i = 0; i < NS; ++i) c.set(cS + i, p[4 + i]);
        c.set(cD, p[3] + p[4]);
#pragma unroll
        for (int i = 0; i < NS; ++i) c.set(cD + 1 + i, 1.0 + p[4 + NS + i] + (i < NS - 1 ? p[i < NS - 1 ? 5 + i : 4] : 0.0));
    }
    template <class Co>
    __device__ static __forceinline__ void rhs(const Co& c, const double (&y)[N], double (&f)[N]) {
        f[0] = fma(-c.get(cB), y[0], c.get(cA));
        f[1] = fma(c.get(cC), y[0], fma(-c.get(cD), y[1], y[2]));
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = fma(c.get(cS + i), y[1 + i], -c.get(cD + 1 + i) * y[2 + i]);
            if (i < NS - 1) v += y[i < NS - 1 ? 3 + i : 2 + i];
            f[2 + i] = v;
        }
    }
    // Tridiagonal block: diag a_j = 1 + c d_j, sub (row j, col j-1) -c S_{j-1}, super -c.  No pivoting is needed: for
    // non-negative rates the block is strictly column diagonally dominant.
    //   top continuants    th_j = a_j th_{j-1} - (c S_{j-1}) c th_{j-2}   (th_{-1} = 1)      p_j = th_j / th_{j-1}
    //   bottom continuants ph_j = a_j ph_{j+1} - (c S_j) c ph_{j+2}       (ph_{NS+1} = 1)    q_j = ph_j / ph_{j+1}
    //   middle pivot       d_M  = Delta / (th_{M-1} ph_{M+1}),
    //                      Delta = a_M th_{M-1} ph_{M+1} - c^2 S_{M-1} th_{M-2} ph_{M+1} - c^2 S_M th_{M-1} ph_{M+2}
    // One reciprocal for all of {q0, th_0..th_{M-1}, ph_{M+1}..ph_NS, Delta} (NS + 2 numbers).
    template <class Co>
    __device__ static __forceinline__ void factor(const Co& co, double c, double (&F)[NF]) {
        double cs[NS], cc[NS], d[NS + 1];
#pragma unroll
        for (int j = 0; j < NS; ++j) { cs[j] = c * co.get(cS + j); cc[j] = cs[j] * c; }
#pragma unroll
        for (int j = 0; j <= NS; ++j) d[j] = co.get(cD + j);
        double th[NT];
        th[0] = fma(c, d[0], 1.0);
        if constexpr (NT > 1) th[1] = fma(fma(c, d[1], 1.0), th[0], -cc[0]);
#pragma unroll
        for (int j = 2; j < NT; ++j) th[j] = fma(fma(c, d[j], 1.0), th[j - 1], -(cc[j - 1] * th[j - 2]));
        double ph[NB + 1];                                   // ph[k] = ph_{MID+1+k}; ph[NB] = 1
        ph[NB] = 1.0;
        if constexpr (NB > 0) ph[NB - 1] = fma(c, d[NS], 1.0);
        if constexpr (NB > 1) ph[NB - 2] = fma(fma(c, d[NS - 1], 1.0), ph[NB - 1], -cc[NS - 1]);
#pragma unroll
        for (int k = NB - 3; k >= 0; --k) ph[k] = fma(fma(c, d[MID + 1 + k], 1.0), ph[k + 1], -(cc[MID + 1 + k] * ph[k + 2]));
        const double thm2 = (NT > 1) ? th[NT > 1 ? NT - 2 : 0] : 1.0;      // th_{M-2}
        const double tp = th[NT - 1] * ph[0];                              // th_{M-1} ph_{M+1}
        double delta = fma(fma(c, d[MID], 1.0), tp, -(cc[MID - 1] * thm2) * ph[0]);
        if constexpr (NB > 0) delta = fma(-(cc[MID < NS ? MID : 0] * th[NT - 1]), ph[1], delta);
        double inv[NS + 2];
        inv[0] = fma(c, co.get(cB), 1.0);
#pragma unroll
        for (int j = 0; j < NT; ++j) inv[1 + j] = th[j];
#pragma unroll
        for (int k = 0; k < NB; ++k) inv[1 + NT + k] = ph[k];
        inv[NS + 1] = delta;
        batch_invert<NS + 2>(inv);
        F[0] = inv[0];
        F[1] = c * co.get(cC);
        F[2] = c;
        F[3] = tp * inv[NS + 1];
        F[O_IP] = inv[1];
#pragma unroll
        for (int j = 1; j < NT; ++j) F[O_IP + j] = th[j - 1] * inv[1 + j];
#pragma unroll
        for (int j = 0; j < NT; ++j) F[O_LT + j] = cs[j] * F[O_IP + j];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            F[O_IQ + k] = ph[k + 1] * inv[1 + NT + k];
            F[O_LB + k] = c * F[O_IQ + k];
            F[O_CS + k] = cs[MID + k];
        }
    }
    __device__ static __forceinline__ void solve(const double (&F)[NF], double (&x)[N]) {
        x[0] *= F[0];
        x[1] = fma(F[1], x[0], x[1]);
#pragma unroll
        for (int j = 1; j < NT; ++j) x[1 + j] = fma(F[O_LT + j - 1], x[j], x[1 + j]);                       // downwards
#pragma unroll
        for (int j = NS - 1; j > MID; --j) x[1 + j] = fma(F[O_LB + j - MID], x[2 + j], x[1 + j]);            // upwards
        double xm = fma(F[O_LT + NT - 1], x[MID], x[1 + MID]);
        if constexpr (NB > 0) xm = fma(F[O_LB], x[2 + MID], xm);
        x[1 + MID] = xm * F[3];
#pragma unroll
        for (int j = MID - 1; j >= 0; --j) x[1 + j] = fma(F[2], x[2 + j], x[1 + j]) * F[O_IP + j];           // back, upwards
#pragma unroll
        for (int j = MID + 1; j <= NS; ++j) x[1 + j] = fma(F[O_CS + j - MID - 1], x[j], x[1 + j]) * F[O_IQ + j - MID - 1];
    }
};

// ------------------------------------------------------------------------------------ kernel
constexpr int TPS_WARPS = TPS_BLOCK / 32;
constexpr int TPS_STASH = 24;          // systems per claimed batch (upper bound; two buffers of this size per warp)
constexpr int TPS_RING = 32;           // finished systems parked per warp before they are written out

// dynamic shared memory of one CTA (bytes); the host side mirrors this (launch_tps)
__host__ __device__ constexpr size_t tps_smem_bytes(int P, int T, int L, bool scalar, bool csmem) {
    size_t b = (size_t)TPS_WARPS * 2 * TPS_STASH * P * 8;        // parameter stash
    if (csmem) b += (size_t)TPS_BLOCK * P * 8;                   // per-lane model coefficients (SmemCoef)
    b += (size_t)TPS_BLOCK * 16;                                 // per-lane system index and |params|^2
    b += (size_t)TPS_BLOCK * 16;                                 // per-lane controller record (TpsCtlRec)
    b += (size_t)TPS_WARPS * 48;                                 // per-warp queue state (TpsQueue)
    if (scalar) b += (size_t)TPS_WARPS * TPS_RING * (4 * 8 + 8 + 16);
    if (scalar) b += (size_t)TPS_BLOCK * 4 * 8;                  // per-lane residual sums (3) and group index
    b += (size_t)T * 8;
    if (scalar) b += (size_t)T * (P / 2) * 16;                   // {target, weight} table of a single-group job (n = P/2 states)
    if (!scalar) b += ((size_t)L * 2 + 7) & ~(size_t)7;
    return b;
}

// Work-queue state of one warp that is only touched when a batch is claimed or swapped in: kept in shared memory (every
// lane writes the same value) so that it does not occupy registers across the step body.
struct TpsQueue {
    long long cur_base, nxt_base, head_seen;
    int nxt_cnt, pend_want;
    int pad[4];
};
static_assert(sizeof(TpsQueue) == 48, "tps_smem_bytes reserves 48 bytes per warp");

// Step-controller memory of one lane (last accepted step and its error, first-step / rejected-last flags, rejected-step
// count): read and written once per step attempt, so it lives in shared memory, not in registers.
struct __align__(16) TpsCtlRec {
    float hacc, erracc;
    int flags;                    // bit 0: at least one step accepted, bit 1: the last attempt was rejected
    int nrej;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <class M, int MIN_BLOCKS, bool SCALAR, bool CSMEM>
__global__ void __launch_bounds__(TPS_BLOCK, MIN_BLOCKS) local_tps_kernel(const LocalArgs a) {
    constexpr int N = M::N, NF = M::NF, P = M::P;
    static_assert(2 * N == P, "tps_smem_bytes sizes the residual table with n = P/2");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int T = a.T;
    const int TN = T * N;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;

    // ---- shared-memory carve-up (tps_smem_bytes)
    unsigned char* sp = smem_raw;
    double* const stash = (double*)sp + (size_t)wid * 2 * TPS_STASH * P;          // [2][TPS_STASH][P] of this warp
    sp += (size_t)TPS_WARPS * 2 * TPS_STASH * P * 8;
    double* const lane_coef = (double*)sp;                                        // CSMEM: [P][TPS_BLOCK]
    if constexpr (CSMEM) sp += (size_t)TPS_BLOCK * P * 8;
    long long* const lane_sys = (long long*)sp;  sp += TPS_BLOCK * 8;             // system index of every lane
    double* const lane_p2 = (double*)sp;         sp += TPS_BLOCK * 8;             // |physical params|^2 (score_fit's l2 term)
    TpsCtlRec* const lrec = (TpsCtlRec*)sp + threadIdx.x;  sp += TPS_BLOCK * sizeof(TpsCtlRec);
    volatile TpsQueue* const q = (volatile TpsQueue*)sp + wid;  sp += TPS_WARPS * sizeof(TpsQueue);
    double* ring_v = nullptr;                    // [TPS_RING][4]: ssr, sum |r|, sum r^2, |params|^2
    long long* ring_sys = nullptr;
    int* ring_i = nullptr;                       // [TPS_RING][4]: status, accepted, rejected, -
    if constexpr (SCALAR) {
        ring_v = (double*)sp + (size_t)wid * TPS_RING * 4;       sp += (size_t)TPS_WARPS * TPS_RING * 32;
        ring_sys = (long long*)sp + (size_t)wid * TPS_RING;      sp += (size_t)TPS_WARPS * TPS_RING * 8;
        ring_i = (int*)sp + (size_t)wid * TPS_RING * 4;          sp += (size_t)TPS_WARPS * TPS_RING * 16;
    }
    // SCALAR: the residual sums of the system a lane is integrating live in shared memory ([4][TPS_BLOCK] columns: weighted
    // SSR, sum |r|, sum r^2 (Morris Y: -, sum v, sum v^2), group index) — touched only when the lane lands on an output time
    double* const lacc = (double*)sp + threadIdx.x;
    if constexpr (SCALAR) sp += (size_t)TPS_BLOCK * 4 * 8;
    double2* const tw_s = (double2*)sp;          // SCALAR: the {target, weight} rows of a single-group job
    if constexpr (SCALAR) sp += (size_t)TN * 16;
    double* const tgrid = (double*)sp;           sp += (size_t)T * 8;
    short* const fmap = (short*)sp;              // !SCALAR: trajectory index (k*N + i) of every flat entry
    for (int i = threadIdx.x; i < T; i += TPS_BLOCK) tgrid[i] = a.t[i];
    const bool tw_in_smem = SCALAR && a.tw && a.n_groups == 1;
    if (tw_in_smem)
        for (int i = threadIdx.x; i < TN; i += TPS_BLOCK) tw_s[i] = a.tw[i];
    if constexpr (!SCALAR) {
        for (int fi = threadIdx.x; fi < a.L; fi += TPS_BLOCK) {
            int k, i;
            if (fi < rna_len) { k = fi + RNA_OFFSET; i = 0; }
            else if (fi < rna_len + T) { k = fi - rna_len; i = 1; }
            else { const int j = fi - rna_len - T; i = 2 + j / T; k = j - (i - 2) * T; }
            fmap[fi] = (short)(k * N + i);
        }
    }
    __syncthreads();

    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    // !SCALAR: trajectory slot of lane l of this CTA = a.traj + (blockIdx.x * TPS_BLOCK + l) * TN (recomputed where it is
    // used: two pointers less to keep across the step body)
    auto traj_slot = [&](int thread) { return a.traj + ((size_t)blockIdx.x * TPS_BLOCK + thread) * TN; };

    // ---- work queue state of the warp (uniform across its lanes; the rarely touched part lives in *q)
    int cur_cnt = 0, cur_taken = 0, pb = 0;
    int pf = 0;                               // 0 idle, 1 claim in flight, 2 copies in flight, 3 queue drained
    const bool al16 = ((reinterpret_cast<unsigned long long>(a.params) & 15ull) == 0);
    int nfin = 0;                             // SCALAR: parked finished systems
    unsigned long long pend = 0;              // lane 0: result of the claim in flight (consumed an iteration later)
    if (lane == 0) { q->cur_base = 0; q->nxt_base = 0; q->head_seen = 0; q->nxt_cnt = 0; q->pend_want = 0; }
    __syncwarp();

    // claim in flight -> copies in flight (pf 1 -> 2), or queue drained (pf 1 -> 3)
    auto consume_claim = [&]() {
        const long long b = (long long)__shfl_sync(FULL, pend, 0);
        const long long c = min((long long)q->pend_want, a.B - b);
        __syncwarp();
        if (c <= 0) pf = 3;
        else {
            if (lane == 0) { q->nxt_base = b; q->nxt_cnt = (int)c; q->head_seen = b + c; }
            double* dst = stash + (size_t)(pb ^ 1) * TPS_STASH * P;
            const double* src = a.params + (size_t)b * P;
            if (al16) { for (int e = lane; e < (int)c * (P / 2); e += 32) cp_async16(dst + 2 * e, src + 2 * e); }
            else { for (int e = lane; e < (int)c * P; e += 32) cp_async8(dst + e, src + e); }
            pf = 2;
        }
    };
    auto issue_claim = [&](int want) {
        if (lane == 0) { q->pend_want = want; pend = atomicAdd(a.counter, (unsigned long long)want); }
        __syncwarp();
        pf = 1;
    };

    // ---- per-lane state
    bool active = false, exhausted = false;
    typename std::conditional<CSMEM, SmemCoef, RegCoef<M::NC>>::type co;
    if constexpr (CSMEM) co.p = lane_coef + threadIdx.x;
    double y[N];
    double t = 0.0;
    double hprop = 0.0;                       // step-size proposal of the controller
    int kout = 0, natt = 0, status = 0;       // next output row, step attempts (accepted + rejected), status

    // SCALAR: write out the parked systems, one per lane
    auto flush_ring = [&]() {
        if constexpr (SCALAR) {
            __syncwarp();
            if (lane < nfin) {
                const long long fs = ring_sys[lane];
                double ssr = ring_v[lane * 4 + 0], s1 = ring_v[lane * 4 + 1], s2 = ring_v[lane * 4 + 2];
                const double p2 = ring_v[lane * 4 + 3];
                const int fstatus = ring_i[lane * 4 + 0];
                // the t = t[0] row (the initial state) is added here, at full warp width
                const double* y0 = a.y0 + (a.y0_stride ? (size_t)fs * a.y0_stride : 0);
                if (a.ymode) {
#pragma unroll
                    for (int i = 0; i < N; ++i) { const double v = fmax(y0[i], 0.0); s1 += v; s2 = fma(v, v, s2); }
                } else {
                    const int g = a.group ? a.group[fs] : 0;
                    const double* tg = a.target + (size_t)g * a.L;
                    const double* sg = a.isigma ? a.isigma + (size_t)g * a.sigma_len : nullptr;
#pragma unroll
                    for (int i = 1; i < N; ++i) {              // row 0 of the RNA column is not part of flat (sol[5:, 0])
                        const int fi = rna_len + (i - 1) * T;
                        const double dlt = fmax(y0[i], 0.0) - __ldg(tg + fi);
                        s1 += fabs(dlt);
                        s2 = fma(dlt, dlt, s2);
                        if (sg) { const double w = dlt * __ldg(sg + fi); ssr = fma(w, w, ssr); }
                    }
                    if (RNA_OFFSET == 0 && T > 0) {
                        const double dlt = fmax(y0[0], 0.0) - __ldg(tg);
                        s1 += fabs(dlt);
                        s2 = fma(dlt, dlt, s2);
                        if (sg) { const double w = dlt * __ldg(sg); ssr = fma(w, w, ssr); }
                    }
                    if (!sg) ssr += s2;                        // unit weights: the weighted and the plain sums coincide
                }
                const double qnan = __longlong_as_double(0x7ff8000000000000LL);
                if (fstatus != 0) { ssr = qnan; s1 = qnan; s2 = qnan; }
                if (a.out_status) a.out_status[fs] = fstatus;
                if (a.out_nsteps) a.out_nsteps[fs] = ring_i[lane * 4 + 1];
                if (a.out_nrej) a.out_nrej[fs] = ring_i[lane * 4 + 2];
                if (a.out_ssr) a.out_ssr[fs] = ssr;
                double shared_v = ssr;                             // the scalar that also goes to the peers
                if (a.out_score) {
                    // score_fit (config/config.py:176-226) with r = |target - pred| / L
                    const double invL = 1.0 / (double)a.L;
                    const double r1 = s1 * invL, r2 = s2 * invL * invL;
                    const double mean_r2 = r2 * invL, mae = r1 * invL;
                    const double sc = a.w_delta * r2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                      a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * sqrt(p2) / (double)P;
                    a.out_score[fs] = sc;
                    if (a.peer_which == 0) shared_v = sc;
                }
                if (a.out_Y) {
                    const double len = (double)TN, mean = s1 / len;
                    double yv;
                    switch (a.y_metric) {
                        case 0: yv = s1; break;
                        case 1: yv = mean; break;
                        case 2: yv = s2 / len - mean * mean; break;
                        default: yv = sqrt(s2); break;
                    }
                    a.out_Y[fs] = yv;
                    if (a.peer_which == 2) shared_v = yv;
                }
                // fused gather: 8-byte stores straight into every rank's result buffer (NVLink peer memory)
                for (int r = 0; r < a.n_peer; ++r) a.peer[r][a.peer_base + fs] = shared_v;
            }
            __syncwarp();
            nfin = 0;
        }
    };

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes from the stash
        unsigned need = __ballot_sync(FULL, !active && !exhausted);
        while (need) {
            if (cur_taken == cur_cnt) {
                // the current batch is used up: make the prefetched one current
                if (pf == 0) issue_claim(TPS_STASH);                   // (only before the first batch)
                if (pf == 1) consume_claim();
                if (pf == 3) {
                    if (!active) exhausted = true;
                    break;
                }
                cp_async_wait_all();
                __syncwarp();
                cur_cnt = q->nxt_cnt; cur_taken = 0; pb ^= 1; pf = 0;
                __syncwarp();
                if (lane == 0) q->cur_base = q->nxt_base;
                if (a.log_params) {
                    double* cur = stash + (size_t)pb * TPS_STASH * P;
                    for (int e = lane; e < cur_cnt * P; e += 32) cur[e] = exp(cur[e]);
                }
                __syncwarp();
            }
            const int avail = cur_cnt - cur_taken;
            const int rank = __popc(need & lt_mask);
            if (((need >> lane) & 1u) && rank < avail) {
                const int slot = cur_taken + rank;
                const long long sys = q->cur_base + slot;
                const double* pr = stash + ((size_t)pb * TPS_STASH + slot) * P;
                double p2 = 0.0;                      // |physical params|^2 for score_fit's l2 term
                {
                    double pv[P];
#pragma unroll
                    for (int i = 0; i < P; ++i) { pv[i] = pr[i]; p2 = fma(pv[i], pv[i], p2); }
                    M::load(co, pv);
                }
                lane_sys[threadIdx.x] = sys;
                lane_p2[threadIdx.x] = p2;
                const double* y0 = a.y0 + (a.y0_stride ? (size_t)sys * a.y0_stride : 0);
#pragma unroll
                for (int i = 0; i < N; ++i) y[i] = y0[i];
                if constexpr (!SCALAR) {
#pragma unroll
                    for (int i = 0; i < N; ++i) traj_slot(threadIdx.x)[i] = y[i];
                } else {
                    double acc_w = 0.0;
                    const int grp = (want_loss && a.group) ? a.group[sys] : 0;
                    const double lam = (want_loss && a.lam_group) ? a.lam_group[grp] : a.lam;
                    if (want_loss && lam != 0.0) {
                        // the lam/P*theta^2 rows of normest's model_func (paramest/normest.py:403-423); theta = the caller's
                        // (possibly logarithmic) parameters
                        const double* th = a.params + (size_t)sys * P;
                        const double* sgr = (a.isigma && a.sigma_len > a.L) ? a.isigma + (size_t)grp * a.sigma_len + a.L : nullptr;
                        for (int i = 0; i < P; ++i) {
                            double w = lam / (double)P * th[i] * th[i];
                            if (sgr) w *= __ldg(sgr + i);
                            acc_w = fma(w, w, acc_w);
                        }
                    }
                    lacc[0] = acc_w; lacc[TPS_BLOCK] = 0.0; lacc[2 * TPS_BLOCK] = 0.0;
                    ((int*)(lacc + 3 * TPS_BLOCK))[0] = grp;
                }
                t = tgrid[0];
                natt = status = 0;
                kout = 1;
                // initial step: 1% of the time scale max|y| / max|f|.  The maxima are taken on the high words of the doubles
                // (monotone for non-negative values, 20 mantissa bits: plenty for a first guess the controller corrects).
                double f0[N];
                M::rhs(co, y, f0);
                int my = 0, mf = 0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    my = max(my, __double2hiint(y[i]) & 0x7fffffff);
                    mf = max(mf, __double2hiint(f0[i]) & 0x7fffffff);
                }
                const float d0 = (float)__hiloint2double(my, 0), d1 = (float)__hiloint2double(mf, 0);
                double h0 = (!(d0 > 1e-30f) || !(d1 > 1e-30f) || !(d0 < 3.0e38f) || !(d1 < 3.0e38f)) ? 1e-6 : 0.01 * (double)(d0 * rcp_ftz(d1));
                hprop = h0;
                *lrec = TpsCtlRec{(float)h0, 1.0f, 0, 0};
                active = true;
            }
            cur_taken += min(__popc(need), avail);
            need = __ballot_sync(FULL, !active && !exhausted);
        }
        if (__all_sync(FULL, !active)) break;

        // ------------------------------------------------------------------ background prefetch of the next batch
        if (pf == 0) {
            // guided self-scheduling: full batches until ~64 systems per warp are left, then half of the per-warp share
            const long long nwarps = (long long)gridDim.x * TPS_WARPS;
            const long long rem = a.B - q->head_seen;
            issue_claim(rem > 64 * nwarps ? TPS_STASH : (int)max(1LL, min((long long)TPS_STASH, rem / (2 * nwarps))));
        } else if (pf == 1) {
            consume_claim();
        }

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
                    double hh = hprop;
                    bool land = false;
                    if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
                    else if (hh > 0.5 * rem) hh = 0.5 * rem;

                    double F[NF];
                    M::factor(co, hh * a.m.gamma, F);
                    double v[N], yn[N], er[N];
                    M::rhs(co, y, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) v[i] *= hh;
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) yn[i] = fma(a.m.mu[0], v[i], y[i]);
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[1], v[i], yn[i]); er[i] = a.m.eps[1] * v[i]; }
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[2], v[i], yn[i]); er[i] = fma(a.m.eps[2], v[i], er[i]); }
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[3], v[i], yn[i]); er[i] = fma(a.m.eps[3], v[i], er[i]); }
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[4], v[i], yn[i]); er[i] = fma(a.m.eps[4], v[i], er[i]); }
                    M::solve(F, v);
#pragma unroll
                    for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[5], v[i], yn[i]); er[i] = fma(a.m.eps[5], v[i], er[i]); }
                    if (a.m.nsol > 6) {                    // ROS6L: seventh solve (uniform over the grid)
                        M::solve(F, v);
#pragma unroll
                        for (int i = 0; i < N; ++i) { yn[i] = fma(a.m.mu[6], v[i], yn[i]); er[i] = fma(a.m.eps[6], v[i], er[i]); }
                    }
                    float err = 0.0f;
                    int ymax = 0;                          // largest |y_new| by its high word (integer pipe): NaN/inf and the FP32 range
                    const float rtolf = (float)a.rtol, floorf_ = (float)a.rtol_floor, kapf = (float)a.kappa, atolf = (float)a.atol;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        err = fmaxf(err, err_ratio_inc(er[i], y[i], yn[i], rtolf, floorf_, kapf, atolf));
                        ymax = max(ymax, __double2hiint(yn[i]) & 0x7fffffff);
                    }
                    // Accept / reject without branches (a branch here splits the warp into landing and non-landing lanes
                    // that then walk the same code twice): everything is a select on `acc`.
                    // 0x47ec0000 = high word of 2.98e38: beyond it (or NaN/inf) the FP32 error scale is meaningless
                    const bool finite = (ymax < 0x47ec0000) && (err < 3.0e38f);
                    const bool acc = finite && (err <= 1.0f);
                    // step controller (pk_common.cuh: elementary + Gustafsson's predictive factor), state in *lrec
                    TpsCtlRec rec = *lrec;
                    float fac = ctl_factor(err, a.m.expo);
                    const float hf = (float)hh;
                    {
                        const float r = err * err * rcp_ftz(rec.erracc);
                        float fg = rec.hacc * rcp_ftz(hf) * pow_ftz(r, a.m.expo) * CTL_INV_SAFE;
                        fg = fmaxf(CTL_FAC_GROW, fminf(CTL_FAC_SHRINK, fg));
                        fac = (acc && (rec.flags & 1)) ? fmaxf(fac, fg) : fac;
                    }
                    double hnew = hh * (double)rcp_ftz(fac);
                    hnew = (acc && (rec.flags & 2)) ? fmin(hnew, hh) : hnew;
                    if (acc && hh < hprop) hnew = fmax(hnew, fmin(hprop, 6.0 * hh));     // the output grid shortened this step
                    hprop = hnew;
                    rec.hacc = acc ? hf : rec.hacc;
                    rec.erracc = acc ? fmaxf(1.0e-2f, err) : rec.erracc;
                    rec.flags = acc ? 1 : (rec.flags | 2);
                    rec.nrej += acc ? 0 : 1;
                    *lrec = rec;
#pragma unroll
                    for (int i = 0; i < N; ++i) y[i] = acc ? yn[i] : y[i];
                    t = acc ? (land ? tout : t + hh) : t;
                    store = acc && land;
                    ++natt;
                    if (!finite) status = 3;
                    else if (!acc && hnew < 1e-14 * fmax(1.0, fabs(t))) status = 2;
                }
                if (store) {
                    if constexpr (SCALAR) {
                        // the residual sums of this output row, straight from the registers.  np.clip(sol, 0, None) on the
                        // sign bit (integer pipe; NaN stays NaN as in numpy).
                        double acc_1 = lacc[TPS_BLOCK], acc_2 = lacc[2 * TPS_BLOCK];
                        if (a.ymode) {
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double v = __double2hiint(y[i]) < 0 ? 0.0 : y[i];
                                acc_1 += v; acc_2 = fma(v, v, acc_2);
                            }
                        } else {
                            // row kout of the {target, weight} table (trajectory order: one base address, 16-byte loads);
                            // flat (models/distmod.py:124-134) = [sol[5:,0] | sol[:,1] | sol[:,2:].T]: the RNA column only from row 5
                            // (generic pointer: the shared-memory copy of a single-group job, the L1-cached global table otherwise)
                            const int grp = ((const int*)(lacc + 3 * TPS_BLOCK))[0];
                            const double2* row = tw_in_smem ? tw_s + kout * N : a.tw + ((size_t)grp * TN + kout * N);
                            if (a.isigma != nullptr) {
                                double acc_w = lacc[0];
#pragma unroll
                                for (int i = 0; i < N; ++i) {
                                    if (i == 0 && kout < RNA_OFFSET) continue;
                                    const double2 e = row[i];
                                    const double dlt = (__double2hiint(y[i]) < 0 ? 0.0 : y[i]) - e.x;
                                    const double w = dlt * e.y;
                                    acc_1 += fabs(dlt);
                                    acc_2 = fma(dlt, dlt, acc_2);
                                    acc_w = fma(w, w, acc_w);
                                }
                                lacc[0] = acc_w;
                            } else {                       // unit weights: the weighted sum is acc_2 (added when the system is written out)
#pragma unroll
                                for (int i = 0; i < N; ++i) {
                                    if (i == 0 && kout < RNA_OFFSET) continue;
                                    const double dlt = (__double2hiint(y[i]) < 0 ? 0.0 : y[i]) - row[i].x;
                                    acc_1 += fabs(dlt);
                                    acc_2 = fma(dlt, dlt, acc_2);
                                }
                            }
                        }
                        lacc[TPS_BLOCK] = acc_1; lacc[2 * TPS_BLOCK] = acc_2;
                    } else {
                        double* o = traj_slot(threadIdx.x) + kout * N;
#pragma unroll
                        for (int i = 0; i < N; ++i) o[i] = y[i];
                    }
                    ++kout;
                }
                if (status == 0 && kout < T && natt >= a.max_steps) status = 1;      // outputs still missing
            }
            finished = (kout >= T) || (status != 0);
        }

        unsigned fin = __ballot_sync(FULL, finished);
        if constexpr (SCALAR) {
            // ------------------------------------------ park finished systems; 32 of them are written out together
            if (fin) {
                const int nnew = __popc(fin);
                if (nfin + nnew > TPS_RING) flush_ring();
                if (finished) {
                    const int slot = nfin + __popc(fin & lt_mask);
                    ring_sys[slot] = lane_sys[threadIdx.x];
                    ring_v[slot * 4 + 0] = lacc[0]; ring_v[slot * 4 + 1] = lacc[TPS_BLOCK]; ring_v[slot * 4 + 2] = lacc[2 * TPS_BLOCK];
                    ring_v[slot * 4 + 3] = lane_p2[threadIdx.x];
                    ring_i[slot * 4 + 0] = status; ring_i[slot * 4 + 1] = natt - lrec->nrej; ring_i[slot * 4 + 2] = lrec->nrej;
                    active = false;
                }
                nfin += nnew;
            }
        } else {
        // ------------------------------------------ warp-cooperative epilogue of finished systems
        while (fin) {
            const int f = __ffs(fin) - 1;
            fin &= fin - 1;
            const int fstatus = __shfl_sync(FULL, status, f);
            const int fvalid = __shfl_sync(FULL, kout, f);          // outputs 0..fvalid-1 were produced
            const int fnatt = __shfl_sync(FULL, natt, f);
            __syncwarp();                                           // the owner's trajectory stores are visible
            const int fnrej = lrec[f - lane].nrej, fnst = fnatt - fnrej;
            const size_t fsys = (size_t)lane_sys[(threadIdx.x & ~31) + f];
            const double* tr = traj_slot((threadIdx.x & ~31) + f);
            const double* y0 = a.y0 + (a.y0_stride ? fsys * (size_t)a.y0_stride : 0);
            const int g = (want_loss && a.group) ? a.group[fsys] : 0;
            const double* tg = want_loss ? a.target + (size_t)g * a.L : nullptr;
            const double* sg = (want_loss && a.isigma) ? a.isigma + (size_t)g * a.sigma_len : nullptr;
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
                                const double w = sg ? dlt * __ldg(sg + fi) : dlt;
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
                        const double w = sg ? dlt * __ldg(sg + fi) : dlt;
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
            const double p2 = lane_p2[(threadIdx.x & ~31) + f];      // computed by the owner when it loaded the parameters
            if (want_loss) {
                const double lam = a.lam_group ? a.lam_group[g] : a.lam;
                if (lam != 0.0) {
                    // the lam/P*theta^2 rows of normest's model_func (paramest/normest.py:403-423)
                    const double* sgr = (sg && a.sigma_len > a.L) ? sg + a.L : nullptr;
                    for (int i = lane; i < P; i += 32) {
                        const double th = a.params[fsys * P + i];
                        double w = lam / (double)P * th * th;
                        if (sgr) w *= __ldg(sgr + i);
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
                double shared_v = ssr;
                if (a.out_score) {
                    // score_fit (config/config.py:176-226) with r = |target - pred| / L
                    const double invL = 1.0 / (double)a.L;
                    const double r1 = sr * invL, r2 = sr2 * invL * invL;
                    const double mean_r2 = r2 * invL, mae = r1 * invL;
                    const double sc = a.w_delta * r2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                      a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * sqrt(p2) / (double)P;
                    a.out_score[fsys] = sc;
                    if (a.peer_which == 0) shared_v = sc;
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
                    if (a.peer_which == 2) shared_v = yv;
                }
                for (int r = 0; r < a.n_peer; ++r) a.peer[r][a.peer_base + fsys] = shared_v;
            }
        }
        if (finished) active = false;
        }
    }
    flush_ring();
}

}  // namespace pk
