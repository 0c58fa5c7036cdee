// Throughput of FP64 tensor-core MMA (DMMA.8x8x4) against DFMA on one SM and on the whole GPU:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, b0 = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a0), "d"(b0));
    }
    double s = 0.0;
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, b0 = 1.0 + threadIdx.x * 1e-4;
    double c[16];
    for (int j = 0; j < 16; j++) c[j] = j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) c[j] = fma(a0, b0, c[j]);
    }
    double s = 0.0;
    for (int j = 0; j < 16; j++) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int which = 0; which < 2; which++) {
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) k_dmma<<<sms, warps * 32>>>(out, iters); else k_dfma<<<sms, warps * 32>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double fma_per_thread = which == 0 ? (double)iters * 8 * 256 / 32 : (double)iters * 16;
            const double total = fma_per_thread * warps * 32 * sms;
            printf("%s warps/SM %2d: %.3f ms, %.2f TFLOP/s (2 flops per FMA), %.1f FMA/clk/SM at 1.965 GHz\n", which == 0 ? "DMMA.8x8x4" : "DFMA      ", warps, ms,
                   2 * total / ms / 1e9, total / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    return 0;
}
