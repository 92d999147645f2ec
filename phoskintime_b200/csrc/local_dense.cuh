// Dense-inverse kernels for the random (2^ns-state hypercube) model (models/randmod.py) and for distributive /
// successive systems too large for the thread-per-system kernel.  1, 2 or 4 warps own one system: W = I - h*gamma*M is
// assembled in shared memory from the analytic Jacobian (the models are linear, M is the transition-rate matrix) and
// INVERTED (Gauss-Jordan, no pivoting: for non-negative rates W is strictly column diagonally dominant once the decoupled
// mRNA row is removed).  The six (ROS5L) or seven (ROS6L) Krylov vectors of a step, v_k = W^-1 v_{k-1}, are then
// mat-vecs instead of sequential triangular solves, and - because the models are linear with constant coefficients -
// W^-1 depends on the step size only: proposed steps are quantised (mantissa bits cleared), forced shorter steps reuse
// the inverse through the run-time gamma' family of the method (rosl_coeffs), so a rand-6 solve needs ~19 inversions for
// ~154 steps.  Two storage variants:
//   * shared-memory inverse (dense_invert / dense_apply): random model up to 7 sites (129 states), distributive / successive
//     beyond 8 sites;
//   * register-resident inverse for 41..68 states (local_dense_kernel<MODEL, 128, 3, 17>: reg_load / reg_invert /
//     reg_apply below), the default for rand-6.
// Thread groups pull systems from the global queue as they finish.
#pragma once
#include "pk_common.cuh"

namespace pk {

struct DenseLayout {       // per-system shared-memory carve-up, in doubles
    int n, ld, P, nobs;
    int xtra;              // register-tile variant: exchange buffers of reg_invert / reg_apply (6 * 32 * TR doubles)
    int nv;                // stride of the work vectors (n; register-tile variant: 4 TC >= n, the tail stays zero so that the
                           // mat-vec can read whole column groups without a bounds test)
    __host__ __device__ int W() const { return 0; }
    __host__ __device__ int vec(int k) const { return n * ld + k * nv; }   // k = 0..4: y, v, y_new, err, v'
    __host__ __device__ int par() const { return n * ld + 5 * nv; }
    __host__ __device__ int prev() const { return par() + P; }
    __host__ __device__ int xbuf() const { return prev() + nobs; }
    __host__ __device__ int total() const { return xbuf() + xtra; }
};

// NT threads (1, 2 or 4 warps) cooperate on one system.
template <int NT>
__device__ __forceinline__ void dsync() {
    if (NT > 32) __syncthreads();
    else __syncwarp();
}

// ------------------------------------------------------------------ model: rhs and W assembly
// p = physical parameters in shared memory.  All functions are warp-cooperative (lane strided).
template <int MODEL, int NT>
__device__ __forceinline__ void dense_rhs(int ns, int n, const double* p, const double* y, double* f, int lane) {
    if (MODEL == 2) {
        const double* S = p + 4;
        const double* Dd = p + 4 + ns;
        const double Pv = y[1];
        for (int r = lane; r < n; r += NT) {
            double v;
            if (r == 0) v = fma(-p[1], y[0], p[0]);
            else if (r == 1) {
                double sS = 0.0, back = 0.0;
                for (int k = 0; k < ns; ++k) { sS += S[k]; back += y[1 + (1 << k)]; }
                v = fma(p[2], y[0], -(p[3] + sS) * Pv) + back;
            } else {
                const int s = r - 1;
                const double rate_in = S[__ffs(s) - 1];
                double gain = 0.0, out = Dd[s - 1];
                for (int j = 0; j < ns; ++j) {
                    const int bit = 1 << j;
                    if (s & bit) {
                        const int src = s & ~bit;
                        gain = fma(rate_in, src ? y[1 + src] : Pv, gain);
                        out += 1.0;
                    } else {
                        const int up = s | bit;
                        gain += y[1 + up];
                        out += S[__ffs(up) - 1];
                    }
                }
                v = fma(-out, y[r], gain);
            }
            f[r] = v;
        }
    } else {
        const double* S = p + 4;
        const double* Dr = p + 4 + ns;
        for (int r = lane; r < n; r += NT) {
            double v;
            if (r == 0) v = fma(-p[1], y[0], p[0]);
            else if (MODEL == 0) {
                if (r == 1) {
                    double sS = 0.0, back = 0.0;
                    for (int k = 0; k < ns; ++k) { sS += S[k]; back += y[2 + k]; }
                    v = fma(p[2], y[0], -(p[3] + sS) * y[1]) + back;
                } else {
                    v = fma(S[r - 2], y[1], -(1.0 + Dr[r - 2]) * y[r]);
                }
            } else {
                if (r == 1) v = fma(p[2], y[0], fma(-(p[3] + S[0]), y[1], y[2]));
                else {
                    const int i = r - 2;
                    double d = 1.0 + Dr[i] + (i < ns - 1 ? S[i + 1] : 0.0);
                    v = fma(S[i], y[r - 1], -d * y[r]);
                    if (i < ns - 1) v += y[r + 1];
                }
            }
            f[r] = v;
        }
    }
}

// W = I - c*J (row-major, leading dimension ld), c = h*gamma.  Each lane assembles whole rows.
template <int MODEL, int NT>
__device__ __forceinline__ void dense_fillW(int ns, int n, int ld, const double* p, double c, double* W, int lane) {
    for (int idx = lane; idx < n * ld; idx += NT) W[idx] = 0.0;
    dsync<NT>();
    const double* S = p + 4;
    for (int r = lane; r < n; r += NT) {
        double* row = W + r * ld;
        if (r == 0) { row[0] = fma(c, p[1], 1.0); continue; }
        if (MODEL == 2) {
            const double* Dd = p + 4 + ns;
            if (r == 1) {
                double sS = 0.0;
                for (int k = 0; k < ns; ++k) { sS += S[k]; row[1 + (1 << k)] = -c; }
                row[0] = -c * p[2];
                row[1] = fma(c, p[3] + sS, 1.0);
            } else {
                const int s = r - 1;
                const double rate_in = S[__ffs(s) - 1];
                double out = Dd[s - 1];
                for (int j = 0; j < ns; ++j) {
                    const int bit = 1 << j;
                    if (s & bit) { row[1 + (s & ~bit)] = -c * rate_in; out += 1.0; }
                    else { const int up = s | bit; row[1 + up] = -c; out += S[__ffs(up) - 1]; }
                }
                row[r] = fma(c, out, 1.0);
            }
        } else {
            const double* Dr = p + 4 + ns;
            if (MODEL == 0) {
                if (r == 1) {
                    double sS = 0.0;
                    for (int k = 0; k < ns; ++k) { sS += S[k]; row[2 + k] = -c; }
                    row[0] = -c * p[2];
                    row[1] = fma(c, p[3] + sS, 1.0);
                } else {
                    row[1] = -c * S[r - 2];
                    row[r] = fma(c, 1.0 + Dr[r - 2], 1.0);
                }
            } else {
                if (r == 1) { row[0] = -c * p[2]; row[1] = fma(c, p[3] + S[0], 1.0); row[2] = -c; }
                else {
                    const int i = r - 2;
                    row[r - 1] = -c * S[i];
                    row[r] = fma(c, 1.0 + Dr[i] + (i < ns - 1 ? S[i + 1] : 0.0), 1.0);
                    if (i < ns - 1) row[r + 1] = -c;
                }
            }
        }
    }
    dsync<NT>();
}

// largest value of the step-size grid 2^(j*lg), j integer, that is <= h
__device__ __forceinline__ double dense_quantize_h(double h, double lg) {
    if (!(h > 0.0) || !(h < 1.0e300)) return h;
    // power-of-two grid (the default): clear the mantissa — FP64 log2 / exp2 cost ~400 instructions per step otherwise
    if (lg == 1.0 && h > 1.0e-300) return __longlong_as_double(__double_as_longlong(h) & 0x7ff0000000000000LL);
    if ((lg == 2.0 || lg == 3.0 || lg == 4.0) && h > 1.0e-300) {       // grids of ratio 4, 8, 16: floor the exponent to a multiple
        const int L = (int)lg;
        const int E = (int)((__double_as_longlong(h) >> 52) & 0x7ff) - 1023;
        int q = E / L;
        if (E < 0 && q * L != E) --q;                                  // floor division
        return __longlong_as_double((long long)(q * L + 1023) << 52);
    }
    return exp2(floor(log2(h) / lg) * lg);
}

// In-place inverse by Gauss-Jordan elimination without pivoting.  Thread (tx, ty) = (tid & 31, tid >> 5) updates
// columns tx, tx+32, ... of rows ty, ty+NW, ...; column k (multipliers) and the scaled pivot row are staged in
// `colk` / `rowk` so that the whole matrix can be rewritten in one sweep.  Zero multipliers / zero pivot-row
// entries are skipped (the early columns of I - c M are sparse).
template <int NT>
__device__ __forceinline__ void dense_invert(int n, int ld, double* W, double* colk, double* rowk, int tid) {
    constexpr int NW = NT / 32;
    const int tx = tid & 31, ty = tid >> 5;
    for (int k = 0; k < n; ++k) {
        const double ipiv = fast_rcp(W[k * ld + k]);
        dsync<NT>();                                                   // everyone has the pivot before it is overwritten
        for (int i = tid; i < n; i += NT) {
            colk[i] = (i == k) ? 0.0 : W[i * ld + k];
            rowk[i] = (i == k) ? 0.0 : W[k * ld + i] * ipiv;
        }
        dsync<NT>();
        // Columns 1.. are dealt to the lanes as j = 1 + tx + 32 c (the random model has 2^ns - 1 + 2 states: 64
        // columns besides the mRNA column 0, i.e. exactly two per lane); this thread's pivot-row entries stay in
        // registers and the rows are walked branch-free with the multiplier loaded once per row.
        const double u0 = 1 + tx < n ? rowk[1 + tx] : 0.0, u1 = 33 + tx < n ? rowk[33 + tx] : 0.0,
                     u2 = 65 + tx < n ? rowk[65 + tx] : 0.0, u3 = 97 + tx < n ? rowk[97 + tx] : 0.0;
        const bool h1 = 33 + tx < n, h2 = 65 + tx < n, h3 = 97 + tx < n, h0 = 1 + tx < n;
        const bool any2 = n > 65;                                       // uniform: third/fourth column slots in use
#pragma unroll 2
        for (int i = ty; i < n; i += NW) {
            const double ci = -colk[i];                                 // uniform over the warp
            double* wi = W + i * ld + 1 + tx;
            if (h0) wi[0] = fma(ci, u0, wi[0]);
            if (h1) wi[32] = fma(ci, u1, wi[32]);
            if (any2) {
                if (h2) wi[64] = fma(ci, u2, wi[64]);
                if (h3) wi[96] = fma(ci, u3, wi[96]);
            }
        }
        {                                                               // column 0 (mRNA): one element per thread
            const double uc = rowk[0];
            if (uc != 0.0)
                for (int i = tid; i < n; i += NT) W[i * ld] = fma(-colk[i], uc, W[i * ld]);
        }
        for (int j = 129 + tx; j < n; j += 32) {                        // n > 129 (not reached by the shipped models)
            const double ukj = rowk[j];
            if (ukj != 0.0)
                for (int i = ty; i < n; i += NW) W[i * ld + j] = fma(-colk[i], ukj, W[i * ld + j]);
        }
        // The sweep above rewrites row k and column k with unchanged values (their multiplier / pivot-row entry is
        // zero); the write-back below gives them their new values.  With more than one warp per system these must
        // not interleave — a late "unchanged" store would overwrite a new value — hence the barrier.
        dsync<NT>();
        // column k <- -col/pivot, row k <- row/pivot, pivot <- 1/pivot
        for (int i = tid; i < n; i += NT) {
            if (i != k) {
                W[i * ld + k] = -colk[i] * ipiv;
                W[k * ld + i] = rowk[i];
            } else {
                W[k * ld + k] = ipiv;
            }
        }
        dsync<NT>();
    }
}

// dst = Winv * src.  Threads own rows (odd leading dimension -> conflict-free); when there are at least two threads
// per row the row is split between the two lanes of a pair and combined with one shuffle.
template <int NT>
__device__ __forceinline__ void dense_apply(int n, int ld, const double* Winv, const double* src, double* dst, int tid) {
    if (NT >= 64 && 2 * (n - 1) <= NT) {                // uniform condition: one lane pair per row 1..n-1
        // row 0 of W is (1 + c B, 0, ..., 0) in every model (mRNA is decoupled), so is row 0 of the inverse
        // lanes l and l^16 share a row, so the 16 lanes served together by one shared-memory pass all read
        // different rows at the same column offset: conflict-free with the odd leading dimension
        if (tid == NT - 1) dst[0] = Winv[0] * src[0];
        const int i = 1 + 16 * (tid >> 5) + (tid & 15), half = (tid >> 4) & 1;
        double acc0 = 0.0, acc1 = 0.0;
        if (i < n) {
            const int mid = (n + 1) >> 1;
            const int j0 = half ? mid : 0, j1 = half ? n : mid;
            const double* row = Winv + i * ld;
            int j = j0;
            for (; j + 1 < j1; j += 2) {
                acc0 = fma(row[j], src[j], acc0);
                acc1 = fma(row[j + 1], src[j + 1], acc1);
            }
            if (j < j1) acc0 = fma(row[j], src[j], acc0);
        }
        double acc = acc0 + acc1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (i < n && half == 0) dst[i] = acc;
    } else {
        for (int i = tid; i < n; i += NT) {
            const double* row = Winv + i * ld;
            double acc0 = 0.0, acc1 = 0.0;
            int j = 0;
            for (; j + 1 < n; j += 2) {
                acc0 = fma(row[j], src[j], acc0);
                acc1 = fma(row[j + 1], src[j + 1], acc1);
            }
            if (j < n) acc0 = fma(row[j], src[j], acc0);
            dst[i] = acc0 + acc1;
        }
    }
    dsync<NT>();
}

// ------------------------------------------------------------------ register-resident inverse (TR x TC tile per thread)
// 128 threads own the matrix as a 32 x 4 grid: lane l holds rows l + 32a (a < TR), warp w holds columns w + 4b (b < TC).
// The shared-memory kernel above is bound by shared-memory bandwidth (every FMA of the inversion and of the six
// mat-vecs per step reads and writes shared memory); here W is assembled in shared memory once per inversion, loaded
// into registers, inverted there (Gauss-Jordan without pivoting: the pivot ROW reaches the threads by warp shuffle —
// row k lives in lane k & 31 of every warp — the pivot COLUMN through a double-buffered 32*TR-entry shared array
// written by its owner warp at the end of the previous column step: one block barrier per column), and the six Krylov
// mat-vecs of a step run from registers (x by broadcast loads, the four column groups reduced through shared memory).
// The outer loop over the column slot is unrolled, so the register slots of the pivot row (b >> 3) and of the pivot
// column (b) are compile-time constants.
__device__ __forceinline__ int fresh_tid() {           // re-read where it is used: the compiler otherwise hoists every index
    int t;                                             // derived from the thread id out of the time loop and keeps dozens of
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));    // loop-invariant addresses alive in registers across the inversion
    return t;
}

template <int TR, int TC, int NW>
__device__ __forceinline__ void reg_load(int n, int ld, const double* W, double (&Wt)[TR][TC], int) {
    const int tid = fresh_tid();
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
        for (int b = 0; b < TC; ++b) {
            const int i = lane + 32 * a, j = warp + NW * b;
            Wt[a][b] = (i < n && j < n) ? W[i * ld + j] : 0.0;
        }
}

template <int TR, int TC>
__device__ __forceinline__ void reg_rotate_rows(double (&Wt)[TR][TC]) {
#pragma unroll
    for (int b = 0; b < TC; ++b) {
        const double t0 = Wt[0][b];
#pragma unroll
        for (int a = 0; a + 1 < TR; ++a) Wt[a][b] = Wt[a + 1][b];
        Wt[TR - 1][b] = t0;
    }
}

// cyclic shift of the column slots: slot b <- slot (b + SH) mod TC, in place along the cycles of the permutation (one
// temporary per cycle: a scratch copy of a whole row would cost 2 TC registers)
__host__ __device__ constexpr int reg_gcd(int a, int b) { return b == 0 ? a : reg_gcd(b, a % b); }
template <int TR, int TC, int SH>
__device__ __forceinline__ void reg_rotate_cols(double (&Wt)[TR][TC]) {
    constexpr int G = reg_gcd(TC, SH % TC == 0 ? TC : SH % TC), LEN = TC / G;
    if constexpr (SH % TC != 0) {
#pragma unroll
        for (int a = 0; a < TR; ++a) {
#pragma unroll
            for (int c = 0; c < G; ++c) {
                const double t0 = Wt[a][c];
                int j = c;
#pragma unroll
                for (int s = 0; s + 1 < LEN; ++s) {
                    const int nj = (j + SH) % TC;
                    Wt[a][j] = Wt[a][nj];
                    j = nj;
                }
                Wt[a][j] = t0;
            }
        }
    }
}

template <int TR, int TC, int NW>
__device__ __forceinline__ void reg_invert(int n, double (&Wt)[TR][TC], double* colbuf, int) {
    constexpr int RS = 32 * TR;
    const int tid = fresh_tid();
    const int lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < TR; ++a) colbuf[lane + 32 * a] = Wt[a][0];
    }
    __syncthreads();
    // A fully unrolled loop over the column slots is ~45 KB of code and starves on instruction fetch; a single body with
    // the pivot column always in slot 0 pays a register rotation (2 TR TC moves) per slot.  Middle ground: the body is
    // unrolled over U = 4 consecutive slots (static register indices 0..3), then the column slots rotate by U; the row
    // slots rotate when k passes a multiple of 32, so the pivot row is always physical row slot 0.  A final static
    // rotation restores the layout (total column shift = 0 mod TC, TR row rotations).
    constexpr int U = TC >= 4 ? 4 : 1;
    int rot_r = 0;                                     // physical row slot a holds logical slot (a + rot_r) mod TR
#pragma unroll 1
    for (int bq = 0; bq < TC; bq += U) {
        constexpr int SLOTS32 = 32 / NW;               // column slots per 32 columns (NW warps deal the columns round-robin)
        static_assert(SLOTS32 % U == 0, "row rotations happen at column slots 32/NW, 2*32/NW, ... (k = 32, 64, ...): U must divide 32/NW");
        // k = NW bq passes a multiple of 32: the next row slot becomes the pivot-row slot (once per 32/NW column slots — kept
        // out of the column loop, where the compiler turns it into ~100 predicated moves per column)
        if (bq > 0 && (bq & (SLOTS32 - 1)) == 0 && NW * bq < n) { reg_rotate_rows<TR, TC>(Wt); rot_r = rot_r + 1 == TR ? 0 : rot_r + 1; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int bk = bq + u;                     // logical column slot, held in physical slot u
#pragma unroll 1
            for (int wk = 0; wk < NW; ++wk) {
                const int k = wk + NW * bk;
                if (bk >= TC || k >= n) break;         // uniform over the block
                const double* cb = colbuf + (k & 1) * RS;
                const int lk = k & 31;
                const double ip = fast_rcp(cb[k]);
                int row[TR];
                double m[TR];
#pragma unroll
                for (int a = 0; a < TR; ++a) {
                    int la = a + rot_r;
                    if (la >= TR) la -= TR;
                    row[a] = lane + 32 * la;
                    // row k itself: W[k][j] <- W[k][j] / pivot = W[k][j] + (1/pivot - 1) W[k][j] — the same FMA as every other row
                    m[a] = (row[a] == k) ? ip - 1.0 : -cb[row[a]] * ip;
                }
                const bool own = warp == wk;
#pragma unroll
                for (int b = 0; b < TC; ++b) {
                    const double r = __shfl_sync(0xffffffffu, Wt[0][b], lk);     // W[k][.] before this step (pivot row: slot 0)
                    const double rr = (b == u && own) ? 0.0 : r;                 // the pivot column itself is rewritten below
#pragma unroll
                    for (int a = 0; a < TR; ++a) Wt[a][b] = fma(m[a], rr, Wt[a][b]);
                }
                if (own) {                             // column k <- -column / pivot, pivot <- 1 / pivot
#pragma unroll
                    for (int a = 0; a < TR; ++a) Wt[a][u] = (row[a] == k) ? ip : m[a];
                }
                const int k1 = k + 1;                  // publish column k + 1 (its values are final now) for the next step
                if (k1 < n && warp == (k1 & (NW - 1))) {
                    double* nb = colbuf + (k1 & 1) * RS;
#pragma unroll
                    for (int a = 0; a < TR; ++a) nb[row[a]] = (wk < NW - 1) ? Wt[a][u] : Wt[a][u + 1 < TC ? u + 1 : u];
                }
                __syncthreads();
            }
        }
        reg_rotate_cols<TR, TC, U>(Wt);                // the next U column slots become slots 0..U-1
    }
    constexpr int DONE = ((TC + U - 1) / U) * U % TC;  // net column shift so far
    if constexpr (DONE != 0) reg_rotate_cols<TR, TC, TC - DONE>(Wt);
    while (rot_r != 0) { reg_rotate_rows<TR, TC>(Wt); rot_r = rot_r + 1 == TR ? 0 : rot_r + 1; }
}

// dst = Winv * src (both in shared memory); red = NW * 32 * TR doubles
template <int TR, int TC, int NW>
__device__ __forceinline__ void reg_apply(int n, const double (&Wt)[TR][TC], const double* src, double* dst, double* red, int) {
    constexpr int RS = 32 * TR;
    const int tid = fresh_tid();
    const int lane = tid & 31, warp = tid >> 5;
    double acc[TR];
#pragma unroll
    for (int a = 0; a < TR; ++a) acc[a] = 0.0;
#pragma unroll
    for (int b = 0; b < TC; ++b) {
        const double xj = src[warp + NW * b];          // same address across the warp: a broadcast load; the padded tail
                                                       // of the vector is zero and so are the padded columns of Wt
#pragma unroll
        for (int a = 0; a < TR; ++a) acc[a] = fma(Wt[a][b], xj, acc[a]);
    }
#pragma unroll
    for (int a = 0; a < TR; ++a) red[warp * RS + lane + 32 * a] = acc[a];
    __syncthreads();
    for (int i = tid; i < n; i += 32 * NW) {
        double sum = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2 += 2) sum += red[w2 * RS + i] + red[(w2 + 1) * RS + i];
        dst[i] = sum;
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reductions through shared memory (NT = 32: plain warp shuffles)
template <int NT>
__device__ __forceinline__ double dense_sum(double v, double* red) {
    v = warp_sum(v);
    if (NT > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        v = red[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) v += red[w];
    }
    return v;
}
template <int NT>
__device__ __forceinline__ double dense_max(double v, double* red) {
    v = warp_max(v);
    if (NT > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        v = red[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) v = fmax(v, red[w]);
    }
    return v;
}

#ifndef PK_DENSE_REG_BLOCKS
#define PK_DENSE_REG_BLOCKS 2          // resident CTAs per SM of the register-tile variant (zero spill at <= 255 registers)
#endif
template <int MODEL, int NT, int TR = 0, int TC = 0>
__global__ void __launch_bounds__(NT, TR ? PK_DENSE_REG_BLOCKS : 640 / NT) local_dense_kernel(const LocalArgs a, const DenseLayout lay) {
    static_assert(TR == 0 || NT == 128 || NT == 256, "the register tile is laid out over 4 or 8 warps");
    constexpr int NW = NT / 32;
    constexpr bool REG = TR > 0;
    double Wt[REG ? TR : 1][REG ? TC : 1];        // REG: (I - h gamma M)^-1, register resident
    extern __shared__ double smem[];
    __shared__ double red[8];
    __shared__ double coef[64];           // scratch + results of rosl_coeffs<6|7>
    __shared__ double smu[8], seps[8];    // REG: coefficients of the step in flight
    __shared__ unsigned long long s_idx;
    const int lane = threadIdx.x;                 // thread index inside the group that owns one system
    const int n = a.n, ns = a.ns, ld = lay.ld, T = a.T, P = a.P, nobs = lay.nobs;
    double* W = smem + lay.W();
    double* y = smem + lay.vec(0);
    double* v = smem + lay.vec(1);
    double* w = smem + lay.vec(2);
    double* E = smem + lay.vec(3);
    double* v2 = smem + lay.vec(4);
    double* p = smem + lay.par();
    double* prev = smem + lay.prev();
    double* xbuf = smem + lay.xbuf();             // REG: [2][32 TR] pivot-column exchange | [4][32 TR] mat-vec partial sums
    const bool want_loss = (a.out_ssr != nullptr) || (a.out_score != nullptr);
    const bool want_y = a.out_Y != nullptr;
    const int rna_len = T > RNA_OFFSET ? T - RNA_OFFSET : 0;
    if constexpr (REG) {
        for (int i = lane; i < 5 * lay.nv; i += NT) smem[lay.vec(0) + i] = 0.0;     // incl. the padded tails (never written again)
    }

    for (;;) {
        dsync<NT>();
        if (lane == 0) s_idx = atomicAdd(a.counter, 1ull);
        dsync<NT>();
        const unsigned long long idx = s_idx;
        if ((long long)idx >= a.B) break;
        const size_t sys = (size_t)idx;

        // ---------------------------------------------------------------------------- init
        double p2l = 0.0;
        for (int i = lane; i < P; i += NT) {
            double v = a.params[sys * P + i];
            if (a.log_params) v = exp(v);
            p[i] = v;
            p2l = fma(v, v, p2l);
        }
        const double* y0 = a.y0 + (a.y0_stride ? sys * (size_t)a.y0_stride : 0);
        for (int i = lane; i < n; i += NT) y[i] = y0[i];
        dsync<NT>();
        const int grp = a.group ? a.group[sys] : 0;
        const double* tg = a.target ? a.target + (size_t)grp * a.L : nullptr;
        const double* sg = a.sigma ? a.sigma + (size_t)grp * a.sigma_len : nullptr;
        // fused-output accumulators of this thread: registers, or (register-tile variant, whose registers hold the
        // inverse) a shared-memory record behind the exchange buffers
        EpiAcc e_loc{0, 0, 0, 0, 0, 0};
        EpiAcc& e = [&]() -> EpiAcc& {
            if constexpr (REG) return ((EpiAcc*)(xbuf + (2 + NW) * 32 * TR))[lane];
            else return e_loc;
        }();
        if constexpr (REG) e = EpiAcc{0, 0, 0, 0, 0, 0};
        double t = a.t[0];
        int nst = 0, nrej = 0, status = 0, kout = 1;
        double h_inv = -1.0;            // step size whose (I - h gamma M)^-1 currently sits in W

        // lane-parallel emit of output index k from vector src (nullptr -> NaN)
        auto emit = [&](int k, const double* src) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            for (int i = lane; i < n; i += NT) {
                double v = src ? fmax(src[i], 0.0) : qnan;
                if (a.normalize) v *= 1.0 / y0[i];
                if (a.out_sol) a.out_sol[(sys * T + k) * n + i] = v;
                if (i < nobs) {
                    int fi = (i == 0) ? (k >= RNA_OFFSET ? k - RNA_OFFSET : -1)
                                      : (i == 1 ? rna_len + k : rna_len + T + (i - 2) * T + k);
                    if (fi >= 0) {
                        if (a.out_flat) a.out_flat[sys * a.L + fi] = v;
                        if (want_loss) {
                            double dlt = v - __ldg(tg + fi);
                            double ww = sg ? dlt / __ldg(sg + fi) : dlt;
                            e.ssr = fma(ww, ww, e.ssr);
                            e.sr += fabs(dlt);
                            e.sr2 = fma(dlt, dlt, e.sr2);
                        }
                    }
                    if (want_y) {
                        e.s1 += v;
                        e.s2 = fma(v, v, e.s2);
                        if (a.y_metric == 3) {
                            if (k > 0) { double dd = v - prev[i]; e.dyn = fma(dd, dd, e.dyn); }
                            prev[i] = v;
                        }
                    }
                }
            }
        };

        // initial step from the error-weighted time scale |y|/|f|
        dense_rhs<MODEL, NT>(ns, n, p, y, v, lane);
        dsync<NT>();
        double d0 = 0.0, d1 = 0.0;
        for (int i = lane; i < n; i += NT) {
            double sc = 1.0 / fma(a.rtol, fabs(y[i]), a.atol);
            d0 = fmax(d0, fabs(y[i]) * sc);
            d1 = fmax(d1, fabs(v[i]) * sc);
        }
        d0 = dense_max<NT>(d0, red);
        d1 = dense_max<NT>(d1, red);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        StepCtl ctl{dense_quantize_h(h0, a.hgrid_log2), (float)h0, 1.0f, 0, 0};
        emit(0, y);

        // ---------------------------------------------------------------------- time loop
        while (kout < T && status == 0) {
            const double tout = a.t[kout];
            const double rem = tout - t;
            if (!(rem > 0.0)) { emit(kout, y); ++kout; continue; }
            double hh = ctl.h;
            bool land = false;
            if (LAND_STRETCH * hh >= rem) { hh = rem; land = true; }
            else if (hh > 0.5 * rem) hh = 0.5 * rem;
            // Coefficients of this step.  W holds (I - h_inv*gamma*M)^-1.  A step of exactly h_inv uses it as is; a
            // step SHORTENED by the output grid (landing / halving) uses it too, as the ROS5L member with
            // gamma' = gamma*h_inv/hh (pk_common.cuh: ros5l_coeffs) — so inversions only happen when the controller
            // moves to another level of the step-size grid.
            // the step's coefficients: 28 registers in the shared-memory variant; the register-tile variant keeps them in
            // shared memory (smu / seps) — its registers hold the inverse
            double mu_r[REG ? 1 : 7], eps_r[REG ? 1 : 7];
            if constexpr (REG) {
                __syncthreads();
                if (lane < 7) { smu[lane] = a.m.mu[lane]; seps[lane] = a.m.eps[lane]; }
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) { mu_r[k] = a.m.mu[k]; eps_r[k] = a.m.eps[k]; }
            }
            auto MU = [&](int k) -> double { if constexpr (REG) return smu[k]; else return mu_r[k]; };
            auto EPS = [&](int k) -> double { if constexpr (REG) return seps[k]; else return eps_r[k]; };
            const bool seven = a.m.nsol > 6;          // ROS6L: seven solves per step
            const bool forced = hh != ctl.h;          // hh was set by the output grid, not by the controller
            bool reuse = hh == h_inv;
            if (!reuse && a.m.family == 1 && forced && h_inv > 0.0 && hh <= 1.03 * h_inv && hh * 16.0 >= h_inv) reuse = true;
            if (!reuse) {
                // invert at the controller's grid value when it covers this (forced) step, else at the step itself
                const double hnew = (a.m.family == 1 && forced && hh <= 1.03 * ctl.h && hh * 16.0 >= ctl.h) ? ctl.h : hh;
                dense_fillW<MODEL, NT>(ns, n, ld, p, hnew * a.m.gamma, W, lane);
                if constexpr (REG) {
                    reg_load<TR, TC, NW>(n, ld, W, Wt, lane);
                    reg_invert<TR, TC, NW>(n, Wt, xbuf, lane);
                } else {
                    dense_invert<NT>(n, ld, W, v2, v, lane);
                }
                h_inv = hnew;
#ifdef PK_DENSE_COUNT_INV
                ++nrej;
#endif
            }
            if (hh != h_inv) {                      // uniform over the system's threads
                if (lane == 0) {
                    if (seven) rosl_coeffs<7>(a.m.gamma * h_inv / hh, coef);
                    else rosl_coeffs<6>(a.m.gamma * h_inv / hh, coef);
                }
                dsync<NT>();
                if constexpr (REG) {
                    if (lane < 7) {
                        if (seven) { smu[lane] = coef[48 + lane]; seps[lane] = coef[55 + lane]; }
                        else if (lane < 6) { smu[lane] = coef[36 + lane]; seps[lane] = coef[42 + lane]; }
                    }
                } else {
                    if (seven) {
#pragma unroll
                        for (int k = 0; k < 7; ++k) { mu_r[k] = coef[48 + k]; eps_r[k] = coef[55 + k]; }
                    } else {
#pragma unroll
                        for (int k = 0; k < 6; ++k) { mu_r[k] = coef[36 + k]; eps_r[k] = coef[42 + k]; }
                    }
                }
                dsync<NT>();
            }
            if constexpr (REG) __syncthreads();

            auto apply = [&](const double* src, double* dst) {
                if constexpr (REG) reg_apply<TR, TC, NW>(n, Wt, src, dst, xbuf + 2 * 32 * TR, lane);
                else dense_apply<NT>(n, ld, W, src, dst, lane);
            };
            // v_0 = h f(y); v_k = W^-1 v_{k-1}; y_new = y + sum MU_k v_k; err = sum EPS_k v_k
            dense_rhs<MODEL, NT>(ns, n, p, y, v2, lane);
            for (int i = lane; i < n; i += NT) v2[i] *= hh;
            dsync<NT>();
            const double* vl;                         // last Krylov vector and its coefficients
            double mul, epl;
            if constexpr (REG) {
                // one rolled loop over the solves (a single inlined copy of the register mat-vec; the coefficients are
                // read from shared memory by index)
                const int nsol = seven ? 7 : 6;
                double* src = v2;
                double* dst = v;
#pragma unroll 1
                for (int ks = 0; ks < nsol; ++ks) {
                    reg_apply<TR, TC, NW>(n, Wt, src, dst, xbuf + 2 * 32 * TR, lane);
                    if (ks + 1 < nsol) {
                        const double muk = smu[ks], epk = seps[ks];           // eps_0 = 0: E starts at zero
                        for (int i = lane; i < n; i += NT) {
                            w[i] = fma(muk, dst[i], ks == 0 ? y[i] : w[i]);
                            E[i] = ks == 0 ? 0.0 : fma(epk, dst[i], E[i]);
                        }
                    }
                    double* tmp = src; src = dst; dst = tmp;
                }
                vl = src; mul = smu[nsol - 1]; epl = seps[nsol - 1];
            } else {
            apply(v2, v);
            for (int i = lane; i < n; i += NT) w[i] = fma(MU(0), v[i], y[i]);
            apply(v, v2);
            for (int i = lane; i < n; i += NT) { w[i] = fma(MU(1), v2[i], w[i]); E[i] = EPS(1) * v2[i]; }
            apply(v2, v);
            for (int i = lane; i < n; i += NT) { w[i] = fma(MU(2), v[i], w[i]); E[i] = fma(EPS(2), v[i], E[i]); }
            apply(v, v2);
            for (int i = lane; i < n; i += NT) { w[i] = fma(MU(3), v2[i], w[i]); E[i] = fma(EPS(3), v2[i], E[i]); }
            apply(v2, v);
            for (int i = lane; i < n; i += NT) { w[i] = fma(MU(4), v[i], w[i]); E[i] = fma(EPS(4), v[i], E[i]); }
            apply(v, v2);
            vl = v2;
            mul = MU(5); epl = EPS(5);
            if (seven) {
                for (int i = lane; i < n; i += NT) { w[i] = fma(MU(5), v2[i], w[i]); E[i] = fma(EPS(5), v2[i], E[i]); }
                apply(v2, v);
                vl = v; mul = MU(6); epl = EPS(6);
            }
            }
            float err = 0.0f;
            bool bad = false;
            const float rtolf = (float)a.rtol, floorf_ = (float)a.rtol_floor, kapf = (float)a.kappa, atolf = (float)a.atol;
            for (int i = lane; i < n; i += NT) {
                const double yn = fma(mul, vl[i], w[i]);
                const double ei = fma(epl, vl[i], E[i]);
                w[i] = yn;
                const float q = err_ratio_inc(ei, y[i], yn, rtolf, floorf_, kapf, atolf);
                bad |= !(q < 3.0e38f) || !(fabs(yn) < 3.0e38);          // NaN/inf, or beyond the FP32 range of the error scale
                err = fmaxf(err, q);
            }
            if (bad) err = __int_as_float(0x7f800000);
            err = (float)dense_max<NT>((double)err, red);          // +inf marks a non-finite stage value
            bad = !(err < 3.0e38f);
            dsync<NT>();

            if (bad) { status = 3; break; }
            if (err <= 1.0f) {
                ++nst;
                const double hprop = ctl.h;
                const double hnew = ctl_accept(ctl, hh, err, a.m.expo);
                const double hp = (hh < hprop) ? fmax(hnew, fmin(hprop, 6.0 * hh)) : hnew;
                double hq = dense_quantize_h(hp, a.hgrid_log2);
                // hysteresis: a proposal only slightly below the current grid step keeps it (the proposal already
                // carries the 0.9 safety factor), so the controller does not flip between two neighbouring levels
                if (hq < hprop && hp >= 0.85 * hprop) hq = hprop;
                ctl.h = hq;
                for (int i = lane; i < n; i += NT) y[i] = w[i];
                dsync<NT>();
                if (land) { t = tout; emit(kout, y); ++kout; }
                else t += hh;
            } else {
                ++nrej;
                ctl.h = dense_quantize_h(ctl_reject(ctl, hh, err, a.m.expo), a.hgrid_log2);
                if (ctl.h < 1e-14 * fmax(1.0, fabs(t))) status = 2;
            }
            if (status == 0 && kout < T && nst + nrej >= a.max_steps) status = 1;
        }
        for (int k = kout; k < T; ++k) emit(k, nullptr);       // failed system: NaN tail

        // ------------------------------------------------------------------------ finish
        if (lane == 0) {
            if (a.out_status) a.out_status[sys] = status;
            if (a.out_nsteps) a.out_nsteps[sys] = nst;
            if (a.out_nrej) a.out_nrej[sys] = nrej;
        }
        if (want_loss) {
            double ssr = e.ssr;
            const double lam = a.lam_group ? a.lam_group[grp] : a.lam;
            if (lam != 0.0) {
                for (int i = lane; i < P; i += NT) {
                    double th = a.params[sys * P + i];
                    double ww = lam / (double)P * th * th;
                    if (sg && a.sigma_len > a.L) ww /= __ldg(sg + a.L + i);
                    ssr = fma(ww, ww, ssr);
                }
            }
            ssr = dense_sum<NT>(ssr, red);
            const double sr = dense_sum<NT>(e.sr, red), sr2 = dense_sum<NT>(e.sr2, red), p2 = dense_sum<NT>(p2l, red);
            if (lane == 0) {
                if (a.out_ssr) a.out_ssr[sys] = ssr;
                if (a.out_score) {
                    const double invL = 1.0 / (double)a.L;
                    const double r1 = sr * invL, r2 = sr2 * invL * invL;      // sum r, sum r^2, r = |d|/L
                    const double mean_r2 = r2 * invL, mae = r1 * invL;
                    a.out_score[sys] = a.w_delta * r2 + a.w_alpha * sqrt(mean_r2) + a.w_beta * mae +
                                       a.w_gamma * (mean_r2 - mae * mae) + a.w_mu * sqrt(p2) / (double)P;
                }
            }
        }
        if (want_y) {
            const double s1 = dense_sum<NT>(e.s1, red), s2 = dense_sum<NT>(e.s2, red), dyn = dense_sum<NT>(e.dyn, red);
            if (lane == 0) {
                const double len = (double)(T * nobs), mean = s1 / len;
                double yv;
                switch (a.y_metric) {
                    case 0: yv = s1; break;
                    case 1: yv = mean; break;
                    case 2: yv = s2 / len - mean * mean; break;
                    case 3: yv = dyn; break;
                    default: yv = sqrt(s2); break;
                }
                a.out_Y[sys] = yv;
            }
        }
        dsync<NT>();
    }
}

}  // namespace pk
