// Dev micro-benchmark (GPU box): the production gj_invert<TILE> (csrc/global_net.cuh) on a synthetic
// Schur-like matrix, one system per CTA, with optional cycle stamps of the phases of every column step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_gj tools/bench_gj.cu
//   /tmp/bench_gj [nQ] [ctas] [reps]
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ long long g_trace[8192];
__device__ int g_trace_on;
#define GJ_TRACE_DECL int _ti = 0; const bool _tr_on = g_trace_on && blockIdx.x == 0 && threadIdx.x == 32;
#define GJ_TRACE(slot) if (_tr_on && _ti < 8192) g_trace[_ti++] = clock64();
#include "../phoskintime_b200/csrc/global_net.cuh"

template <int TILE, int V>
__device__ __forceinline__ void gj_exp(const pk::GlobalCtx& cx, double (&A)[TILE][TILE]) {
    constexpr int GP = 16 * TILE;            // padded order
    constexpr int NW = (TILE + 1) / 2;       // pivot candidates per lane
    const int tr = threadIdx.x & 15, tc = threadIdx.x >> 4, lane = threadIdx.x & 31;
    const int nQ = cx.nQ;
    unsigned mymask = 0;                     // bit w: physical row lane + 32 w has already served as a pivot row
    
    // Thread (tr, tc) = (tid & 15, tid >> 4): the 16 owners of one matrix ROW segment sit in one half-warp, so
    // the pivot row reaches every thread by a warp shuffle; only the pivot COLUMN crosses warps, through shared
    // memory, written by its 16 owner lanes at the end of the previous iteration (other parity buffer).
    // => ONE block barrier per eliminated column.
    // The loop body is kept SMALL (one copy, ~300 instructions): the column slots of the tile are rotated by one
    // after every 16 columns so that the active column always sits in slot 0 (TILE rotations = identity), and the
    // row slot of the pivot is resolved by one uniform branch tree.  (An unrolled body of 12 copies measured
    // ~2x slower: the column sweep became instruction-fetch bound.)
    if (nQ > 0 && tc == 0) {
#pragma unroll
        for (int a = 0; a < TILE; ++a) cx.colbuf[tr + 16 * a] = A[a][0];
    }
#pragma unroll 1
    for (int kb = 0; kb < TILE; ++kb) {
#pragma unroll 1
        for (int kk = 0; kk < 16; ++kk) {
            const int k = kb * 16 + kk;
            if (k >= nQ) break;
            const double* const colb = cx.colbuf + (kk & 1) * GP;
            __syncthreads();
            
            double cv[TILE];                 // this thread's rows of column k
#pragma unroll
            for (int a = 0; a < TILE; ++a) cv[a] = colb[tr + 16 * a];
            // Pivot search, redundantly in every warp.  Key = FP32 magnitude with the low 7 mantissa bits replaced
            // by (127 - row): ONE integer max (redux.sync) returns the largest entry and, on ties, the lowest row.
            // Used rows carry key 0; padded rows hold exact zeros and lose to any valid row.
            unsigned key = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int r = lane + 32 * w;
                if (r < GP) {
                    const unsigned kq = (__float_as_uint(fabsf((float)colb[r])) & ~127u) | (unsigned)(127 - r);
                    key = max(key, ((mymask >> w) & 1u) ? 0u : kq);
                }
            }
            if (V != 1 && V != 2 && V != 5) key = __reduce_max_sync(0xffffffffu, key);
            const int p = (V == 1 || V == 2 || V == 5) ? k : 127 - (int)(key & 127u);
            
            if ((p & 31) == lane) mymask |= 1u << (p >> 5);
            if (threadIdx.x == 0) { cx.piv[k] = p; cx.pinv[p] = k; }
            const double nip = (V == 4 || V == 5) ? -colb[p] : -pk::fast_rcp(colb[p]);        // -1/pivot
            const int src = (lane & 16) | (p & 15);       // lane of this half-warp that owns row p
            const bool prow = tr == (p & 15);
            const bool pcol = tc == kk;
            double rv[TILE];
            
            // Row p sits in register slot p >> 4 (uniform over the CTA).  Inside the dispatched block: fetch it,
            // and arrange the operands so that the generic rank-1 update below ALSO produces the special entries:
            //   column k  (threads tc == kk): A := 0, rv := 1        ->  fma(g, 1, 0) = g = -col/pivot
            //   pivot row (lanes tr == p&15): A := -rv*nip, col := 0 ->  fma(0, rv, A) = row/pivot, 1/pivot at (p,k)
            pk::dispatch_uniform<0, TILE>((V == 2 || V == 5) ? 0 : (p >> 4), [&](auto slot) {
                constexpr int a = decltype(slot)::value;
#pragma unroll
                for (int b = 0; b < TILE; ++b) rv[b] = __shfl_sync(0xffffffffu, A[a][b], src);
                if (pcol) {
                    rv[0] = 1.0;
#pragma unroll
                    for (int a2 = 0; a2 < TILE; ++a2) A[a2][0] = 0.0;
                }
                if (prow) {
#pragma unroll
                    for (int b = 0; b < TILE; ++b) A[a][b] = -rv[b] * nip;
                    cv[a] = 0.0;
                }
            });
            
#pragma unroll
            for (int a = 0; a < TILE; ++a) {
                const double g = cv[a] * nip;
#pragma unroll
                if (V != 3) for (int b = 0; b < TILE; ++b) A[a][b] = fma(g, rv[b], A[a][b]);
                else A[a][0] += g * rv[a];
            }
            
            // column k+1 for the next iteration (other parity: nobody reads that buffer any more); after column 15
            // of a block it is the first column of the NEXT slot (the rotation below has not happened yet)
            if (k + 1 < nQ) {
                double* const coln = cx.colbuf + ((kk + 1) & 1) * GP;
                if (kk < 15) {
                    if (tc == kk + 1) {
#pragma unroll
                        for (int a = 0; a < TILE; ++a) coln[tr + 16 * a] = A[a][0];
                    }
                } else if (tc == 0) {
#pragma unroll
                    for (int a = 0; a < TILE; ++a) coln[tr + 16 * a] = A[a][TILE > 1 ? 1 : 0];
                }
            }
            
        }
        // rotate the column slots: slot b <- slot b+1 (executed TILE times in total = identity)
#pragma unroll
        for (int a = 0; a < TILE; ++a) {
            const double t0 = A[a][0];
#pragma unroll
            for (int b = 0; b + 1 < TILE; ++b) A[a][b] = A[a][b + 1];
            A[a][TILE - 1] = t0;
        }
    }
    __syncthreads();
}


template <int TILE, int V>
__global__ void __launch_bounds__(256, TILE <= 6 ? 2 : 1) benchv(int nQ, int reps, long long* cyc, double* chk) {
    extern __shared__ double sm[];
    constexpr int GP = 16 * TILE;
    pk::GlobalTopoDev tp{};
    pk::GlobalCtx cx{tp};
    cx.nQ = nQ; cx.colbuf = sm; cx.rowbuf = sm + 2 * GP; cx.piv = (int*)(sm + 4 * GP); cx.pinv = cx.piv + 128;
    const int tr = threadIdx.x & 15, tc = threadIdx.x >> 4;
    double A[TILE][TILE];
    double acc = 0.0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int a = 0; a < TILE; ++a)
            for (int b = 0; b < TILE; ++b) {
                const int r = tr + 16 * a, c = tc + 16 * b;
                unsigned h = (r * 131u + c * 977u + blockIdx.x * 31u + rep * 7u) * 2654435761u;
                double v = 0.0;
                if (r < nQ && c < nQ) { if ((h >> 8) % 24 == 0) v = ((h >> 16) % 2001) * 1e-4 - 0.1; if (r == c) v += 1.0; }
                A[a][b] = v;
            }
        gj_exp<TILE, V>(cx, A);
        for (int a = 0; a < TILE; ++a) for (int b = 0; b < TILE; ++b) acc += A[a][b];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    chk[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int TILE>
__global__ void __launch_bounds__(256, TILE <= 6 ? 2 : 1) bench(int nQ, int reps, long long* cyc, double* chk) {
    extern __shared__ double sm[];
    constexpr int GP = 16 * TILE;
    pk::GlobalTopoDev tp{};
    pk::GlobalCtx cx{tp};
    cx.nQ = nQ;
    cx.colbuf = sm;
    cx.rowbuf = sm + 2 * GP;
    cx.piv = (int*)(sm + 4 * GP);
    cx.pinv = cx.piv + 128;
    const int tr = threadIdx.x & 15, tc = threadIdx.x >> 4;
    double A[TILE][TILE];
    double acc = 0.0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int a = 0; a < TILE; ++a)
            for (int b = 0; b < TILE; ++b) {
                const int r = tr + 16 * a, c = tc + 16 * b;
                unsigned h = (r * 131u + c * 977u + blockIdx.x * 31u + rep * 7u) * 2654435761u;
                double v = 0.0;
                if (r < nQ && c < nQ) {
                    if ((h >> 8) % 24 == 0) v = ((h >> 16) % 2001) * 1e-3 - 1.0;
                    if (r == c) v += 1.0;
                }
                A[a][b] = v;
            }
        pk::gj_invert<TILE>(cx, A);
        for (int a = 0; a < TILE; ++a)
            for (int b = 0; b < TILE; ++b) acc += A[a][b];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    chk[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int nQ = argc > 1 ? atoi(argv[1]) : 96, ctas = argc > 2 ? atoi(argv[2]) : 148, reps = argc > 3 ? atoi(argv[3]) : 50;
    long long* cyc; double* chk;
    cudaMalloc(&cyc, ctas * 8); cudaMalloc(&chk, ctas * 256 * 8);
    const int smem = (4 * 128 + 256) * 8;
    auto run = [&](int trace) {
        cudaMemcpyToSymbol(g_trace_on, &trace, 4);
        if (nQ <= 32) bench<2><<<ctas, 256, smem>>>(nQ, reps, cyc, chk);
        else if (nQ <= 64) bench<4><<<ctas, 256, smem>>>(nQ, reps, cyc, chk);
        else if (nQ <= 96) bench<6><<<ctas, 256, smem>>>(nQ, reps, cyc, chk);
        else bench<8><<<ctas, 256, smem>>>(nQ, reps, cyc, chk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    };
    run(0);
    run(0);
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (auto v : h) s += v;
    printf("nQ=%d ctas=%d: %.0f cycles per inversion (%.1f per column)\n", nQ, ctas, s / ctas / reps, s / ctas / reps / nQ);
    run(1);
    std::vector<long long> tr(8192);
    cudaMemcpyFromSymbol(tr.data(), g_trace, 8192 * 8);
    // 6 stamps per column: average the phase durations over the columns of the first inversion
    double ph[6] = {0, 0, 0, 0, 0, 0};
    int ncol = nQ;
    for (int k = 0; k < ncol; ++k)
        for (int j = 0; j < 6; ++j) {
            const long long a = tr[k * 6 + j], b = (j < 5) ? tr[k * 6 + j + 1] : (k + 1 < ncol ? tr[(k + 1) * 6] : tr[k * 6 + 5]);
            ph[j] += double(b - a);
        }
    printf("phases (cycles/column, warp 1): search %.0f | rcp %.0f | row shuffle %.0f | fma %.0f | fix+colpub %.0f | barrier %.0f\n",
           ph[0] / ncol, ph[1] / ncol, ph[2] / ncol, ph[3] / ncol, ph[4] / ncol, ph[5] / ncol);
    if (nQ > 64 && nQ <= 96) {
        auto rv = [&](auto kern, const char* name) {
            kern<<<ctas, 256, smem>>>(nQ, reps, cyc, chk);
            cudaDeviceSynchronize();
            cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
            double s2 = 0; for (auto v : h) s2 += v;
            printf("  variant %-28s %.1f cycles per column\n", name, s2 / ctas / reps / nQ);
        };
        rv(benchv<6, 0>, "baseline");
        rv(benchv<6, 1>, "no search (p=k), dispatch kept");
        rv(benchv<6, 2>, "no search, static slot");
        rv(benchv<6, 3>, "1 FMA/row instead of 6");
        rv(benchv<6, 4>, "no reciprocal");
        rv(benchv<6, 5>, "no search/dispatch/reciprocal");
    }
    return 0;
}
