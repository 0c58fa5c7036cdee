// Dependent-chain latencies (cycles per op, one warp, sm_100a) of the primitives the morphology kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run: ./lat
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
#define REP 256
template <typename F> __device__ long long timed(F f) {
    long long t0 = clock64();
    f();
    long long t1 = clock64();
    return t1 - t0;
}
__global__ void k(double* out, long long* cyc, double seed, unsigned useed) {
    __shared__ unsigned hist[512];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    double x = seed + lane * 1e-3;
    unsigned u = useed + lane;
    long long c;
    int s = 0;
    c = timed([&] { for (int i = 0; i < REP; i++) x = x + seed; }); if (threadIdx.x == 0) cyc[s] = c; s++;                 // 0 DADD
    c = timed([&] { for (int i = 0; i < REP; i++) x = fma(x, seed, seed); }); if (threadIdx.x == 0) cyc[s] = c; s++;       // 1 DFMA
    x = fabs(x) * 1e-300 + 2.0;
    c = timed([&] { for (int i = 0; i < REP; i++) x = sqrt(x + 2.0); }); if (threadIdx.x == 0) cyc[s] = c; s++;            // 2 DSQRT(+DADD)
    c = timed([&] { for (int i = 0; i < REP; i++) x = 1.0 / (x + 2.0); }); if (threadIdx.x == 0) cyc[s] = c; s++;          // 3 DDIV(+DADD)
    c = timed([&] { for (int i = 0; i < REP; i++) x = asin(x * 0.25); }); if (threadIdx.x == 0) cyc[s] = c; s++;           // 4 asin
    float fx = (float)x;
    c = timed([&] { for (int i = 0; i < REP; i++) x = (double)((float)x) + 1e-3; }); if (threadIdx.x == 0) cyc[s] = c; s++; // 5 F2F.32.64 + F2F.64.32 + DADD
    // MATCH.ANY: all lanes distinct / all equal / 4 groups
    unsigned m = 0;
    c = timed([&] { for (int i = 0; i < REP; i++) { m = __match_any_sync(0xffffffffu, u + (m >> 31)); } }); if (threadIdx.x == 0) cyc[s] = c; s++;          // 6 distinct
    c = timed([&] { for (int i = 0; i < REP; i++) { m = __match_any_sync(0xffffffffu, (m >> 31) + useed - 1u); } }); if (threadIdx.x == 0) cyc[s] = c; s++; // 7 equal
    c = timed([&] { for (int i = 0; i < REP; i++) { m = __match_any_sync(0xffffffffu, (lane & 3) + ((m >> 31) << 4)); } }); if (threadIdx.x == 0) cyc[s] = c; s++; // 8 four groups
    // same-digit mask from 9 ballots (multisplit), dependent through the digit
    c = timed([&] { for (int i = 0; i < REP; i++) {
        const unsigned d = (u * 2654435761u + m) >> 23;
        unsigned mm = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 9; b++) { const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u); mm &= ((d >> b) & 1u) ? v : ~v; }
        m = mm; } }); if (threadIdx.x == 0) cyc[15] = c;
    // ballot chain
    c = timed([&] { for (int i = 0; i < REP; i++) { m = __ballot_sync(0xffffffffu, (m >> lane) & 1u) + 1u; } }); if (threadIdx.x == 0) cyc[s] = c; s++; // 9 ballot
    // shared atomics with return, spread addresses / same address
    unsigned r = u;
    c = timed([&] { for (int i = 0; i < REP; i++) { r = atomicAdd(&hist[(lane * 7 + r) & 511], 1u); } }); if (threadIdx.x == 0) cyc[s] = c; s++;   // 10 ATOMS spread (dependent)
    c = timed([&] { for (int i = 0; i < REP; i++) { r = atomicAdd(&hist[(r >> 30) & 1], 1u); } }); if (threadIdx.x == 0) cyc[s] = c; s++;                   // 11 ATOMS same address
    c = timed([&] { for (int i = 0; i < REP; i++) { r = hist[(r + lane) & 511] + 1u; } }); if (threadIdx.x == 0) cyc[s] = c; s++;                   // 12 LDS dependent
    c = timed([&] { for (int i = 0; i < REP; i++) { r = __shfl_xor_sync(0xffffffffu, r, 1) + 1u; } }); if (threadIdx.x == 0) cyc[s] = c; s++;      // 13 SHFL
    c = timed([&] { for (int i = 0; i < REP; i++) { __syncthreads(); } }); if (threadIdx.x == 0) cyc[s] = c; s++;                                  // 14 BAR
    out[threadIdx.x] = x + m + r + fx;
}
int main() {
    const char* names[] = {"DADD", "DFMA", "DSQRT+DADD", "DDIV+DADD", "asin+DMUL", "F2F x2 + DADD", "MATCH.ANY distinct", "MATCH.ANY equal",
                           "MATCH.ANY 4 groups", "VOTE.BALLOT(+shift,add)", "ATOMS spread (dep)", "ATOMS same addr (dep)", "LDS (dep)", "SHFL (dep)", "BAR.SYNC", "9-ballot same-digit mask"};
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
    for (int threads : {32, 256}) {
        for (int it = 0; it < 2; it++) k<<<1, threads>>>(out, cyc, 1.000001, 12345u);
        long long h[64];
        cudaMemcpy(h, cyc, 64 * 8, cudaMemcpyDeviceToHost);
        printf("threads per CTA = %d (one CTA on the GPU)\n", threads);
        for (int i = 0; i < 16; i++) printf("  %-26s %8.1f cycles/op\n", names[i], (double)h[i] / REP);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
