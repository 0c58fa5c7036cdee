// The minapix inner loop (pm_kernels.cuh, minapix_group) at different occupancies: does the instruction-bound phase
// gain from more resident warps if it is compiled with fewer registers?  One synthetic galaxy per CTA (disc-shaped map
// of radius 33 px in a 256 x 256 frame, window + raster list + aperture plane in shared memory), every warp sweeps the
// whole list for groups of four candidates.  Prints pixel-candidate evaluations per cycle per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../pawlikmorph-lsst_b200/csrc -I../../include -o minapix_occ minapix_occ.cu
#include <cstdio>
#include <vector>
#include "pm_kernels.cuh"
using namespace pmk;

template <int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) mb(const unsigned* aper_g, const float* mw_g, const unsigned* plist_g, const unsigned* cand_g,
                                                        int n_g, int n_cand, int r0, int c0, int bh, int bw, int reps, double* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char smem[];
    Ctx<float> c;
    c.g = make_geo(256, 3);
    unsigned* aper = (unsigned*)smem;
    float* mw = (float*)(smem + 8192);
    unsigned* plist = (unsigned*)(smem + 8192 + ((bh * bw * 4 + 15) & ~15));
    unsigned* cand = plist + ((n_g + 3) & ~3);
    for (int i = threadIdx.x; i < 2048; i += kThreads) aper[i] = aper_g[i];
    for (int i = threadIdx.x; i < bh * bw; i += kThreads) mw[i] = mw_g[i];
    for (int i = threadIdx.x; i < n_g; i += kThreads) plist[i] = plist_g[i];
    for (int i = threadIdx.x; i < n_cand; i += kThreads) cand[i] = cand_g[i];
    __syncthreads();
    c.aper = aper; c.mw = mw; c.plist = plist; c.posR = plist; c.n_g = n_g;
    c.r0 = r0; c.c0 = c0; c.r1 = r0 + bh - 1; c.c1 = c0 + bw - 1; c.bw = bw; c.sky = 100.0;
    c.ar.init(smem, 200000u, nullptr, 0);          // "everything is in shared memory"
    __shared__ double parts[kWarps * 8];
    double acc = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        for (int grp = 0; grp * kCandBlock < n_cand; grp++) {
            int cidx[kCandBlock];
#pragma unroll
            for (int j = 0; j < kCandBlock; j++) cidx[j] = min(grp * kCandBlock + j, n_cand - 1);
            minapix_group(c, cand, cidx, pm::warp() * 32 + pm::lane(), n_g, 32 * kWarps, parts + 8 * pm::warp());
            if (pm::lane() == 0) acc += parts[8 * pm::warp()];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main() {
    const int n = 256, R = 33, cy = 128, cx = 127;
    std::vector<unsigned> aper(2048, 0), plist, cand;
    const int r0 = cy - R, c0 = cx - R, bh = 2 * R + 1, bw = 2 * R + 1;
    std::vector<float> mw(bh * bw);
    for (int r = 0; r < n; r++)
        for (int col = 0; col < n; col++) {
            const int dr = r - 128, dc = col - 128;
            if (dr * dr + dc * dc <= 46 * 46) aper[r * 8 + (col >> 5)] |= 1u << (col & 31);
            const int er = r - cy, ec = col - cx;
            if (r >= r0 && r < r0 + bh && col >= c0 && col < c0 + bw) {
                const bool in = er * er + ec * ec <= R * R;
                mw[(r - r0) * bw + (col - c0)] = in ? 100.0f + 500.0f * expf(-(er * er + ec * ec) / 200.0f) + (float)((r * 131 + col * 71) % 17) : nanf("");
                if (in) plist.push_back(((unsigned)r << 16) | (unsigned)col);
            }
        }
    for (int k = 0; k < 72; k++) cand.push_back(((unsigned)(cy - 4 + k / 9) << 16) | (unsigned)(cx - 4 + k % 9));
    const int n_g = (int)plist.size(), n_cand = (int)cand.size();
    unsigned *d_aper, *d_plist, *d_cand; float* d_mw; double* d_out; long long* d_cyc;
    cudaMalloc(&d_aper, 8192); cudaMalloc(&d_plist, n_g * 4); cudaMalloc(&d_cand, n_cand * 4); cudaMalloc(&d_mw, bh * bw * 4);
    cudaMalloc(&d_out, 1024 * kThreads * 8); cudaMalloc(&d_cyc, 1024 * 8);
    cudaMemcpy(d_aper, aper.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(d_plist, plist.data(), n_g * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_cand, cand.data(), n_cand * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_mw, mw.data(), bh * bw * 4, cudaMemcpyHostToDevice);
    const int smem = 8192 + ((bh * bw * 4 + 15) & ~15) + ((n_g + 3) & ~3) * 4 + n_cand * 4 + 64;
    const int reps = 4;
    printf("map pixels %d, candidates %d, shared memory per CTA %d bytes, %d threads per CTA\n", n_g, n_cand, smem, kThreads);
    auto run = [&](auto kern, int ctas, const char* name) {
        // the dynamic shared memory request is padded so that exactly `ctas` CTAs fit an SM
        const int smem_req = ctas >= 4 ? smem : (232448 + 1024) / ctas - 1024 - 512;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_req);
        cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem_req);
        const int grid = 148 * occ;
        for (int it = 0; it < 2; it++) kern<<<grid, kThreads, smem_req>>>(d_aper, d_mw, d_plist, d_cand, n_g, n_cand, r0, c0, bh, bw, reps, d_out, d_cyc);
        cudaDeviceSynchronize();
        std::vector<long long> h(grid); cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : h) mean += (double)v; mean /= grid;
        const double evals = (double)reps * n_cand * n_g;                 // per CTA
        printf("%-22s regs %3d  CTAs/SM %d  cycles/CTA %.0f  evaluations/cycle/SM %.3f  (%s)\n", name, fa.numRegs, occ, mean, evals * occ / mean,
               cudaGetErrorString(cudaGetLastError()));
    };
    run(mb<1>, 1, "launch_bounds(256,1)");
    run(mb<2>, 2, "launch_bounds(256,2)");
    run(mb<3>, 3, "launch_bounds(256,3)");
    run(mb<4>, 4, "launch_bounds(256,4)");
    return 0;
}
