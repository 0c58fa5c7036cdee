// Latency of the exact circle/pixel overlap (pm_block.cuh) as the kernel calls it: one warp alone, then 8 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../pawlikmorph-lsst_b200/csrc -I../../include -o overlap overlap.cu
#include <cstdio>
#include "pm_block.cuh"
#define REP 64
__global__ void k(double* out, long long* cyc, double r0) {
    const int lane = threadIdx.x & 31;
    // pixels on a circle of radius ~r0 in the first octant (one per lane), then on the axis (b = 0)
    double acc = 0.0, r = r0;
    int dx = (int)(r0 * cos(0.02 * lane + 0.4)), dy = (int)(r0 * sin(0.02 * lane + 0.4));
    long long t0 = clock64();
    for (int i = 0; i < REP; i++) { const double w = pmk::circle_weight(dx, dy, r); acc += w; r += w * 1e-9; }
    long long t1 = clock64();
    for (int i = 0; i < REP; i++) { const double w = pmk::circle_weight((int)r0, 0, r); acc += w; r += w * 1e-9; }
    long long t2 = clock64();
    for (int i = 0; i < REP; i++) { const double w = pmk::overlap_core(dx - 0.5, dy - 0.5, dx + 0.5, dy + 0.5, r); acc += w; r += w * 1e-9; }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    out[threadIdx.x] = acc;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
    for (int threads : {32, 256, 512}) {
        for (int it = 0; it < 2; it++) k<<<1, threads>>>(out, cyc, 30.3);
        long long h[4];
        cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
        printf("threads %4d: circle_weight(off-axis) %7.1f  circle_weight(on-axis) %7.1f  overlap_core %7.1f cycles per call (dependent)\n",
               threads, (double)h[0] / REP, (double)h[1] / REP, (double)h[2] / REP);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
