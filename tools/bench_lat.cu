// Dev micro-probes (GPU box): latencies/throughputs that bound the per-column chain of gj_invert.
#include <cstdio>
#include <cstdlib>
__global__ void __launch_bounds__(256) probes(long long* out, double* sink, double seed) {
    __shared__ double sh[1024];
    __shared__ unsigned shu[256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = seed + i;
    shu[threadIdx.x] = threadIdx.x * 7u;
    __syncthreads();
    long long t0, t1;
    // (a) 36 independent DFMA accumulators, 64 rounds
    double a[36];
    for (int i = 0; i < 36; ++i) a[i] = seed * i;
    double g = seed + 1.0, r = seed + 2.0;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < 36; ++i) a[i] = fma(g, r, a[i]);
        g += 1e-9;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 64;
    double s = 0; for (int i = 0; i < 36; ++i) s += a[i];
    // (b) dependent DFMA chain
    double x = seed;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x = fma(x, g, r);
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / 64;      // per 16 dependent DFMA
    s += x;
    // (c) barrier round trip
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        sh[threadIdx.x] = x + it;
        __syncthreads();
        x += sh[(threadIdx.x + 37) & 255];
        __syncthreads();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / 64;      // STS + BAR + LDS + DADD + BAR
    s += x;
    // (d) dependent LDS chain (pointer chase)
    int idx = lane;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) idx = shu[idx & 255] & 255;
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / 64;
    s += idx;
    // (e) redux chain
    unsigned k = lane * 3u + (unsigned)seed;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) k = __reduce_max_sync(0xffffffffu, k + lane) ^ (unsigned)it;
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0) / 64;
    s += k;
    // (f) F2F f64->f32 dependent chain
    double y = seed + lane;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) { float f = (float)y; y = (double)f + 1.0; }
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0) / 64;      // F2F.F32.F64 + F2F.F64.F32 + DADD
    s += y;
    // (g) MUFU.RCP64H + 4 DFMA chain
    double z = seed + 3.0;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double q;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(z));
        double e = fma(-z, q, 1.0); q = fma(q, e, q); e = fma(-z, q, 1.0); q = fma(q, e, q);
        z = q + 2.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0) / 64;
    s += z;
    // (h) 36 DFMA + 6 DMUL fed by 12 LDS (the update phase without barriers)
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double rv[6], gg[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) rv[b] = sh[(threadIdx.x & 15) + 16 * b + (it & 1) * 96];
#pragma unroll
        for (int b = 0; b < 6; ++b) gg[b] = sh[200 + (threadIdx.x >> 4) + 16 * b + (it & 1) * 96] * g;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) a[i * 6 + j] = fma(gg[i], rv[j], a[i * 6 + j]);
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[7] = (t1 - t0) / 64;
    // (i) column-step skeleton: barrier, 6 LDS, 6 DMUL, 36 DFMA, owner publish (6 STS by 16 lanes)
    {
        const int tr = threadIdx.x & 15, tc = threadIdx.x >> 4;
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 64; ++it) {
            const double* colb = sh + (it & 1) * 96;
            __syncthreads();
            double cv[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) cv[i] = colb[tr + 16 * i];
            const double nip = colb[it % 96] + 2.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double gg = cv[i] * nip;
#pragma unroll
                for (int j = 0; j < 6; ++j) a[i * 6 + j] = fma(gg, r, a[i * 6 + j]);
            }
            if (tc == ((it + 1) & 15)) {
                double* coln = sh + ((it + 1) & 1) * 96;
#pragma unroll
                for (int i = 0; i < 6; ++i) coln[tr + 16 * i] = a[i * 6];
            }
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[8] = (t1 - t0) / 64;
    }
    // (j) same plus 12 shuffles (pivot-row gather) before the FMAs
    {
        const int tr = threadIdx.x & 15, tc = threadIdx.x >> 4;
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 64; ++it) {
            const double* colb = sh + (it & 1) * 96;
            __syncthreads();
            double cv[6], rv[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) cv[i] = colb[tr + 16 * i];
            const double nip = colb[it % 96] + 2.0;
            const int src = (lane & 16) | (it & 15);
#pragma unroll
            for (int j = 0; j < 6; ++j) rv[j] = __shfl_sync(0xffffffffu, a[j], src);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double gg = cv[i] * nip;
#pragma unroll
                for (int j = 0; j < 6; ++j) a[i * 6 + j] = fma(gg, rv[j], a[i * 6 + j]);
            }
            if (tc == ((it + 1) & 15)) {
                double* coln = sh + ((it + 1) & 1) * 96;
#pragma unroll
                for (int i = 0; i < 6; ++i) coln[tr + 16 * i] = a[i * 6];
            }
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[9] = (t1 - t0) / 64;
    }
    for (int i = 0; i < 36; ++i) s += a[i];
    sink[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 296 * 256 * 8);
    for (int ctas : {1, 148, 296}) {
        probes<<<ctas, 256>>>(out, sink, 1.5);
        cudaDeviceSynchronize();
        long long h[10];
        cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost);
        printf("ctas=%d (256 thr): 36 indep DFMA %lld | 16 dep DFMA %lld | STS+BAR+LDS+BAR %lld | dep LDS %lld | redux %lld | F2F x2+DADD %lld | rcp %lld | update(12 LDS+6 DMUL+36 DFMA) %lld | skeleton %lld | skeleton+shfl %lld cycles\n",
               ctas, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
    }
    return 0;
}
