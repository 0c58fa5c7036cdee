// cuda_emul.h — TEST INFRASTRUCTURE ONLY.
//
// A tiny CTA emulator: every CUDA thread of a block is a fibre run round-robin on
// one OS thread; __syncthreads and the warp collectives are rendez-vous points.
// It lets the CPU test-suite (pytest -m "not gpu") compile the kernel sources of
// pawlikmorph-lsst_b200/csrc with g++ and compare their logic with the oracle on
// boxes that have no GPU.  It is NOT a CPU fallback: the product package never
// loads it, and libpawlik_b200.so is nvcc-only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#define PM_DEV inline
#define PM_DEV_NOINLINE __attribute__((noinline))
#define PM_HOSTDEV inline
#define PM_KERNEL

namespace pm_emul {
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
};
struct WarpState {
    uint64_t slot[2][32];
    int arrived[2] = {0, 0};
    unsigned long long gen = 0;  // completed collectives
};
struct Block {
    int nthreads = 0, block_id = 0, nblocks = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::vector<unsigned long long> wgen;  // per-thread count of warp collectives entered
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
    void* sched_sp = nullptr;
    int cur = 0;
    // block barrier
    int bar_arrived = 0, bar_or = 0, bar_or_result = 0;
    unsigned long long bar_gen = 0;
    std::vector<unsigned long long> tgen;
    unsigned long long progress = 0;
    const std::function<void()>* body = nullptr;
};
extern thread_local Block* g_blk;
extern "C" void pm_emul_swap(void** save_sp, void* load_sp);

inline void yield_() {
    Block* b = g_blk;
    pm_emul_swap(&b->fibers[b->cur].sp, b->sched_sp);
}

void run_block(int block_id, int nblocks, int nthreads, size_t smem_bytes, const std::function<void()>& body);
}  // namespace pm_emul

namespace pm_common {
using pm_emul::g_blk;
PM_DEV int lane() { return g_blk->cur & 31; }

// block barrier (all fibres of the emulated CTA)
PM_DEV int block_sync_or_(int p) {
    pm_emul::Block* b = g_blk;
    int me = b->cur;
    unsigned long long my_gen = b->tgen[me]++;
    b->bar_or |= (p != 0);
    b->bar_arrived++;
    b->progress++;
    if (b->bar_arrived == b->nthreads) {
        b->bar_or_result = b->bar_or;
        b->bar_or = 0;
        b->bar_arrived = 0;
        b->bar_gen = my_gen + 1;
    }
    while (b->bar_gen <= my_gen) pm_emul::yield_();
    return b->bar_or_result;
}

// generic warp rendez-vous: publish a 64-bit value, return all 32 once everyone arrived
PM_DEV void warp_exchange_(uint64_t mine, uint64_t out[32]) {
    pm_emul::Block* b = g_blk;
    int me = b->cur, w = me >> 5, l = me & 31;
    pm_emul::WarpState& ws = b->warps[w];
    unsigned long long my_gen = b->wgen[me]++;
    int par = (int)(my_gen & 1);
    ws.slot[par][l] = mine;
    ws.arrived[par]++;
    b->progress++;
    int lanes = b->nthreads - w * 32 < 32 ? b->nthreads - w * 32 : 32;
    if (ws.arrived[par] == lanes) {
        ws.arrived[par] = 0;
        ws.gen = my_gen + 1;
    }
    while (ws.gen <= my_gen) pm_emul::yield_();
    memcpy(out, ws.slot[par], sizeof(uint64_t) * 32);
}
PM_DEV void syncwarp() { uint64_t o[32]; warp_exchange_(0, o); }
PM_DEV unsigned ballot(int p) {
    uint64_t o[32]; warp_exchange_(p != 0, o);
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= (unsigned)(o[i] & 1) << i; return r;
}
PM_DEV unsigned match_any(unsigned v) {
    uint64_t o[32]; warp_exchange_(v, o);
    unsigned r = 0; for (int i = 0; i < 32; i++) if ((unsigned)o[i] == v) r |= 1u << i; return r;
}
template <typename T> PM_DEV uint64_t to_bits_(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> PM_DEV T from_bits_(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <typename T> PM_DEV T shfl(T v, int src) { uint64_t o[32]; warp_exchange_(to_bits_(v), o); return from_bits_<T>(o[src & 31]); }
template <typename T> PM_DEV T shfl_xor(T v, int m) { uint64_t o[32]; warp_exchange_(to_bits_(v), o); return from_bits_<T>(o[(lane() ^ m) & 31]); }
template <typename T> PM_DEV T shfl_up(T v, int d) { uint64_t o[32]; warp_exchange_(to_bits_(v), o); int s = lane() - d; return s < 0 ? v : from_bits_<T>(o[s]); }
template <typename T> PM_DEV T shfl_down(T v, int d) { uint64_t o[32]; warp_exchange_(to_bits_(v), o); int s = lane() + d; return s > 31 ? v : from_bits_<T>(o[s]); }

// D (8 x 8) += A (8 x 4) B (4 x 8) of the warp, fragments as mma.sync.m8n8k4.f64 lays them out (pm_platform.cuh)
PM_DEV void dmma_8x8x4(double a, double b, double& c0, double& c1) {
    uint64_t oa[32], ob[32];
    warp_exchange_(to_bits_(a), oa);
    warp_exchange_(to_bits_(b), ob);
    const int g = lane() >> 2, t = lane() & 3;
    for (int k = 0; k < 4; k++) {
        const double A = from_bits_<double>(oa[4 * g + k]);
        c0 += A * from_bits_<double>(ob[(2 * t) * 4 + k]);
        c1 += A * from_bits_<double>(ob[(2 * t + 1) * 4 + k]);
    }
}

// warp reductions: one rendez-vous, then the same xor-butterfly order the GPU code uses
template <typename T, typename F> PM_DEV T warp_tree_(T v, F f) {
    uint64_t o[32]; warp_exchange_(to_bits_(v), o);
    T a[32];
    for (int i = 0; i < 32; i++) a[i] = from_bits_<T>(o[i]);
    for (int off = 16; off > 0; off >>= 1) {
        T b[32];
        for (int i = 0; i < 32; i++) b[i] = f(a[i], a[i ^ off]);
        for (int i = 0; i < 32; i++) a[i] = b[i];
    }
    return a[lane()];
}
PM_DEV double warp_sum(double v) { return warp_tree_(v, [](double x, double y) { return x + y; }); }
PM_DEV double warp_max(double v) { return warp_tree_(v, [](double x, double y) { return fmax(x, y); }); }
PM_DEV double warp_min(double v) { return warp_tree_(v, [](double x, double y) { return fmin(x, y); }); }
PM_DEV unsigned warp_sum(unsigned v) { return warp_tree_(v, [](unsigned x, unsigned y) { return x + y; }); }
PM_DEV int warp_sum(int v) { return warp_tree_(v, [](int x, int y) { return x + y; }); }
PM_DEV unsigned warp_max(unsigned v) { return warp_tree_(v, [](unsigned x, unsigned y) { return x > y ? x : y; }); }
PM_DEV unsigned warp_min(unsigned v) { return warp_tree_(v, [](unsigned x, unsigned y) { return x < y ? x : y; }); }
PM_DEV int warp_max(int v) { return warp_tree_(v, [](int x, int y) { return x > y ? x : y; }); }
PM_DEV int warp_min(int v) { return warp_tree_(v, [](int x, int y) { return x < y ? x : y; }); }
PM_DEV unsigned warp_or(unsigned v) { return warp_tree_(v, [](unsigned x, unsigned y) { return x | y; }); }
template <typename T> PM_DEV T warp_incl_scan_(T v) {
    uint64_t o[32]; warp_exchange_(to_bits_(v), o);
    T a[32];
    for (int i = 0; i < 32; i++) a[i] = from_bits_<T>(o[i]);
    for (int off = 1; off < 32; off <<= 1) {          // Hillis-Steele, same order as the GPU code
        T b[32];
        for (int i = 0; i < 32; i++) b[i] = i >= off ? a[i] + a[i - off] : a[i];
        for (int i = 0; i < 32; i++) a[i] = b[i];
    }
    return a[lane()];
}
PM_DEV unsigned warp_incl_scan(unsigned v) { return warp_incl_scan_(v); }
PM_DEV double warp_incl_scan(double v) { return warp_incl_scan_(v); }

PM_DEV int popc(unsigned v) { return __builtin_popcount(v); }
PM_DEV int clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
PM_DEV int ffs(unsigned v) { return __builtin_ffs((int)v); }
PM_DEV unsigned brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
PM_DEV int nth_set_bit(unsigned v, int k) {
    for (int i = 0; i < 32; i++)
        if ((v >> i) & 1u) { if (k == 0) return i; k--; }
    return 32;
}
template <typename T> PM_DEV T ldg(const T* p) { return *p; }
PM_DEV unsigned atomic_add(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
PM_DEV unsigned atomic_or(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
PM_DEV unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
PM_DEV int atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
PM_DEV int atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
PM_DEV int atomic_max(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
PM_DEV int atomic_cas(int* p, int expected, int desired) { int o = *p; if (o == expected) *p = desired; return o; }
// built with -ffp-contract=off: plain IEEE operations
PM_DEV double dadd(double a, double b) { return a + b; }
PM_DEV double dsub(double a, double b) { return a - b; }
PM_DEV double dmul(double a, double b) { return a * b; }
PM_DEV double ddiv(double a, double b) { return a / b; }
PM_DEV double dsqrt(double a) { return sqrt(a); }
PM_DEV long long d2bits(double a) { long long u; memcpy(&u, &a, 8); return u; }
PM_DEV double bits2d(long long a) { double u; memcpy(&u, &a, 8); return u; }
PM_DEV unsigned f2bits(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
PM_DEV float bits2f(unsigned a) { float u; memcpy(&u, &a, 4); return u; }
PM_DEV void prefetch_l2(const void*, unsigned) {}
PM_DEV void mbar_init(unsigned long long* bar, unsigned) { *bar = 0; }
PM_DEV void mbar_fence_init() {}
PM_DEV void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    // the hardware requires a SHARED destination and 16-byte alignment: catch violations on the CPU
    const unsigned char* d = (const unsigned char*)smem_dst;
    if (d < g_blk->smem || d + bytes > g_blk->smem + g_blk->smem_bytes || ((uintptr_t)gsrc & 15) || (bytes & 15) || ((uintptr_t)d & 15)) {
        fprintf(stderr, "pm_emul: illegal TMA bulk copy (dst not in shared memory or misaligned)\n");
        abort();
    }
    memcpy(smem_dst, gsrc, bytes);   // the emulated copy completes at once
    (*bar)++;
}
PM_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {   // phase with this parity completed <=> current parity differs
    while (((*bar) & 1ull) == (unsigned long long)parity) { g_blk->progress++; pm_emul::yield_(); }
}
PM_DEV void dadd_if(bool pred, double& acc, double x) { if (pred) acc += x; }
PM_DEV long long clock() { return 0; }
PM_DEV unsigned smem_addr(const void* p) {
    const unsigned char* q = (const unsigned char*)p;
    if (q < g_blk->smem || q >= g_blk->smem + g_blk->smem_bytes) { fprintf(stderr, "pm_emul: smem_addr of a non-shared pointer\n"); abort(); }
    return (unsigned)(q - g_blk->smem);
}
PM_DEV unsigned lds_u32(unsigned a) { unsigned v; memcpy(&v, g_blk->smem + a, 4); return v; }
PM_DEV float lds_f32(unsigned a) { float v; memcpy(&v, g_blk->smem + a, 4); return v; }
PM_DEV double lds_f64(unsigned a) { double v; memcpy(&v, g_blk->smem + a, 8); return v; }
}  // namespace pm_common
// execution groups, see pm_platform.cuh: the whole emulated CTA, or (PM_GROUP_WARP) one warp of it
#if defined(PM_GROUP_WARP)
namespace pm_grp_warp {
using namespace pm_common;
PM_DEV int tid() { return g_blk->cur & 31; }
PM_DEV int nthreads() { return 32; }
PM_DEV int warp() { return 0; }
PM_DEV int nwarps() { return 1; }
PM_DEV int block() { return g_blk->block_id * (g_blk->nthreads >> 5) + (g_blk->cur >> 5); }
PM_DEV int nblocks() { return g_blk->nblocks * (g_blk->nthreads >> 5); }
PM_DEV void sync() { syncwarp(); }
PM_DEV int sync_or(int p) { return ballot(p) != 0u; }
PM_DEV unsigned char* dyn_smem(unsigned group_bytes) { return g_blk->smem + (size_t)(g_blk->cur >> 5) * group_bytes; }
}  // namespace pm_grp_warp
namespace pm = pm_grp_warp;
#else
namespace pm_grp_cta {
using namespace pm_common;
PM_DEV int tid() { return g_blk->cur; }
PM_DEV int nthreads() { return g_blk->nthreads; }
PM_DEV int warp() { return g_blk->cur >> 5; }
PM_DEV int nwarps() { return g_blk->nthreads >> 5; }
PM_DEV int block() { return g_blk->block_id; }
PM_DEV int nblocks() { return g_blk->nblocks; }
PM_DEV int sync_or(int p) { return block_sync_or_(p); }
PM_DEV void sync() { (void)block_sync_or_(0); }
PM_DEV unsigned char* dyn_smem(unsigned) { return g_blk->smem; }
// thread-block clusters: the emulator runs clusters of one CTA
PM_DEV unsigned cluster_rank() { return 0u; }
PM_DEV unsigned cluster_size() { return 1u; }
PM_DEV unsigned cluster_id() { return (unsigned)g_blk->block_id; }
PM_DEV unsigned cluster_count() { return (unsigned)g_blk->nblocks; }
PM_DEV void cluster_sync() { (void)block_sync_or_(0); }
PM_DEV unsigned dsmem_addr(unsigned a, unsigned) { return a; }
PM_DEV unsigned dsmem_ld_u32(unsigned a) { return lds_u32(a); }
PM_DEV unsigned long long dsmem_ld_u64(unsigned a) { unsigned long long v; memcpy(&v, g_blk->smem + a, 8); return v; }
PM_DEV double dsmem_ld_f64(unsigned a) { return lds_f64(a); }
}  // namespace pm_grp_cta
namespace pm = pm_grp_cta;
#endif

// what kernels use from the CUDA headers
#ifndef __CUDACC__
#define __restrict__ __restrict
#define __launch_bounds__(...)
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }
#endif
