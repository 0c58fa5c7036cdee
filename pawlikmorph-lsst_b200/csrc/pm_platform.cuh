// pm_platform.cuh — the handful of CUDA execution-model primitives the kernels use,
// behind one small namespace so that the SAME kernel source can also be compiled by
// g++ against tests/emul/cuda_emul.h (a fibre-based CTA emulator used only by the
// CPU test-suite to exercise kernel logic on boxes without a GPU).  The product
// library is always built by nvcc for sm_100a and contains none of the emulator.
#pragma once
#include <stdint.h>

#if defined(PM_EMULATE)
#include "cuda_emul.h"          // tests/emul — never part of libpawlik_b200.so
#else
#include <cuda_runtime.h>
#include <math.h>
#define PM_DEV __device__ __forceinline__
#define PM_DEV_NOINLINE __device__ __noinline__
#define PM_HOSTDEV __host__ __device__ inline
#define PM_KERNEL __global__
// Execution groups.  The phase code is written against pm::tid() / pm::warp() / pm::sync() of a GROUP of kThreads
// threads that analyses one cutout.  Two groups exist: the whole CTA (the default: __syncthreads barriers), and — for
// translation units compiled with PM_GROUP_WARP — a single warp (32 threads, __syncwarp as the barrier, several
// independent groups per CTA, each with its own slice of the dynamic shared memory).  The latency-bound tail of the
// path (r20/r80, calcA, smoothness) runs warp-per-cutout so that many cutouts are in flight per SM.
namespace pm_common {
PM_DEV int lane() { return threadIdx.x & 31; }
PM_DEV void syncwarp() { __syncwarp(); }
PM_DEV unsigned ballot(int p) { return __ballot_sync(0xffffffffu, p); }
PM_DEV unsigned match_any(unsigned v) { return __match_any_sync(0xffffffffu, v); }
template <typename T> PM_DEV T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename T> PM_DEV T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <typename T> PM_DEV T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <typename T> PM_DEV T shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
// warp-wide reductions / scans (the emulator implements these in one rendez-vous)
PM_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
PM_DEV double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
PM_DEV double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
PM_DEV unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }
PM_DEV int warp_sum(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
PM_DEV unsigned warp_max(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
PM_DEV unsigned warp_min(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }
PM_DEV int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
PM_DEV int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
PM_DEV unsigned warp_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
// inclusive prefix sums across the warp
PM_DEV unsigned warp_incl_scan(unsigned v) {
    int l = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= o) v += t;
    }
    return v;
}
PM_DEV double warp_incl_scan(double v) {
    int l = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= o) v += t;
    }
    return v;
}
PM_DEV int popc(unsigned v) { return __popc(v); }
PM_DEV int clz(unsigned v) { return __clz(v); }
PM_DEV int ffs(unsigned v) { return __ffs(v); }
PM_DEV unsigned brev(unsigned v) { return __brev(v); }
// position (0-based) of the k-th (0-based) set bit of v; 32 if there is none
PM_DEV int nth_set_bit(unsigned v, int k) {
    unsigned r = __fns(v, 0, k + 1);
    return r == 0xffffffffu ? 32 : (int)r;
}
template <typename T> PM_DEV T ldg(const T* p) { return __ldg(p); }
PM_DEV unsigned atomic_add(unsigned* p, unsigned v) { return atomicAdd(p, v); }
PM_DEV unsigned atomic_or(unsigned* p, unsigned v) { return atomicOr(p, v); }
PM_DEV unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
PM_DEV int atomic_add(int* p, int v) { return atomicAdd(p, v); }
PM_DEV int atomic_min(int* p, int v) { return atomicMin(p, v); }
PM_DEV int atomic_max(int* p, int v) { return atomicMax(p, v); }
PM_DEV int atomic_cas(int* p, int expected, int desired) { return atomicCAS(p, expected, desired); }
// IEEE-exact, never contracted into FMA (bit-parity with numpy/scipy where it matters)
PM_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
PM_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
PM_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
PM_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PM_DEV double dsqrt(double a) { return __dsqrt_rn(a); }
PM_DEV long long d2bits(double a) { return __double_as_longlong(a); }
PM_DEV double bits2d(long long a) { return __longlong_as_double(a); }
PM_DEV unsigned f2bits(float a) { return __float_as_uint(a); }
PM_DEV float bits2f(unsigned a) { return __uint_as_float(a); }
// hint: pull [p, p+bytes) towards L2 ahead of use (bytes multiple of 16, p 16-byte aligned)
PM_DEV void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (one elected thread issues; all wait)
PM_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
PM_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PM_DEV void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
PM_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
}
// acc += x when pred (a predicated DADD, no select)
PM_DEV void dadd_if(bool pred, double& acc, double x) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(x), "r"((unsigned)pred));
}
PM_DEV long long clock() { return clock64(); }
// explicit shared-memory loads through 32-bit addresses (hot loops: cheaper than generic 64-bit addressing)
PM_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PM_DEV unsigned lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
PM_DEV float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
PM_DEV double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
// FP64 tensor-core product of the warp, D (8 x 8) += A (8 x 4) B (4 x 8) (SASS DMMA.8x8x4).  Lane 4 g + t holds A[g][t] in
// `a`, B[t][g] in `b`, and D[g][2 t], D[g][2 t + 1] in c0, c1.
PM_DEV void dmma_8x8x4(double a, double b, double& c0, double& c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// thread-block clusters: the CTAs of one cluster read each other's shared memory (distributed shared memory)
PM_DEV unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
PM_DEV unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
PM_DEV unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
PM_DEV unsigned cluster_count() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
// every thread of every CTA of the cluster; shared-memory writes made before it are visible cluster-wide after it
PM_DEV void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location (32-bit shared address `a` of this CTA) in the CTA of rank `rank`
PM_DEV unsigned dsmem_addr(unsigned a, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
// (volatile keeps them behind the cluster barrier, itself a volatile asm; no memory clobber: several may be in flight)
PM_DEV unsigned dsmem_ld_u32(unsigned a) { unsigned v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
PM_DEV unsigned long long dsmem_ld_u64(unsigned a) { unsigned long long v; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
PM_DEV double dsmem_ld_f64(unsigned a) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
}  // namespace pm_common
#if defined(PM_GROUP_WARP)
namespace pm_grp_warp {
using namespace pm_common;
PM_DEV int tid() { return threadIdx.x & 31; }
PM_DEV int nthreads() { return 32; }
PM_DEV int warp() { return 0; }
PM_DEV int nwarps() { return 1; }
PM_DEV int block() { return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); }
PM_DEV int nblocks() { return gridDim.x * (blockDim.x >> 5); }
PM_DEV void sync() { __syncwarp(); }
PM_DEV int sync_or(int p) { return __any_sync(0xffffffffu, p); }
PM_DEV unsigned char* dyn_smem(unsigned group_bytes) {
    extern __shared__ __align__(16) unsigned char pm_dyn_smem[];
    return pm_dyn_smem + (size_t)(threadIdx.x >> 5) * group_bytes;
}
}  // namespace pm_grp_warp
namespace pm = pm_grp_warp;
#else
namespace pm_grp_cta {
using namespace pm_common;
PM_DEV int tid() { return threadIdx.x; }
PM_DEV int nthreads() { return blockDim.x; }
PM_DEV int warp() { return threadIdx.x >> 5; }
PM_DEV int nwarps() { return blockDim.x >> 5; }
PM_DEV int block() { return blockIdx.x; }
PM_DEV int nblocks() { return gridDim.x; }
PM_DEV void sync() { __syncthreads(); }
PM_DEV int sync_or(int p) { return __syncthreads_or(p); }
PM_DEV unsigned char* dyn_smem(unsigned) {
    extern __shared__ __align__(16) unsigned char pm_dyn_smem[];
    return pm_dyn_smem;
}
}  // namespace pm_grp_cta
namespace pm = pm_grp_cta;
#endif
#endif
