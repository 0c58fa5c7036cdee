// pm_k_stats.cu — C ABI and launches of the order-statistics kernels (pm_stats.cuh): pm_sky_stats_batch (the per-pixel half
// of skyBackground._calcSkybgr) and pm_clipped_stats_batch (astropy sigma_clipped_stats of whole cutouts).  512-thread
// CTAs, two per SM; optionally a cutout is spread over a thread-block cluster and held in its shared memory.
#ifndef PM_STAT_THREADS
#define PM_STAT_THREADS 512
#endif
#ifndef PM_STAT_MIN_CTAS
#define PM_STAT_MIN_CTAS 2
#endif
#define PM_THREADS PM_STAT_THREADS
#define PM_NS pmk_s
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pawlik_b200.h"
#include "pm_stats.cuh"

using namespace pmk_s;

extern "C" void pm_set_error(const char* msg);
extern "C" void pm_count_launch(void);

#ifndef PM_EMULATE
template <typename T> __global__ void __launch_bounds__(kThreads, PM_STAT_MIN_CTAS)
pm_sky_stats_kernel(const T* images, long long B, int n, const double* ellipse, double* out, int resident) {
    sky_body<T>(images, B, n, ellipse, out, resident);
}
template <typename T> __global__ void __launch_bounds__(kThreads, PM_STAT_MIN_CTAS)
pm_clipstats_kernel(const T* images, long long B, int n, double sigma, int maxiters, double* out, int resident, int snap_at, double* out_snap) {
    clipstats_body<T>(images, B, n, sigma, maxiters, out, resident, snap_at, out_snap);
}
#endif

namespace {

struct StatPlan { int cluster, resident; unsigned smem; long long grid; };

// Two forms of the same kernel.  Default: one CTA per cutout, the pixels stay in global memory and every pass reads
// them through L2 — with two CTAs per SM the cutouts in flight (296 x 256 KB at 256^2) fit the 126 MB L2, so DRAM still
// sees each cutout about once (profiles/r2t_*), and the passes are bound by instruction issue, not by memory.
// PM_STAT_CLUSTER = 1|2|4|8: the cutout is spread over a thread-block cluster of that many CTAs and held in their
// shared memory (read from HBM exactly once; partial sums and histogram counters exchanged through distributed
// shared memory).  Measured slower on B200 at every BASELINE size (256^2 float32, sky statistics: 844k cutouts/s
// global form, 588k / 392k / 236k for clusters of 2 / 4 / 8): the per-pass barriers and exchanges cost more than the
// L2 reads they save, so it is kept as an option (and as the tested multi-CTA form), not as the default.
int make_plan(int64_t B, int n, size_t elem, bool with_mask, StatPlan* P, char* err, size_t err_len) {
#ifndef PM_EMULATE
    int dev = 0, sm = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { snprintf(err, err_len, "no CUDA device"); return PM_ENODEV; }
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (sm <= 0) { snprintf(err, err_len, "no CUDA device"); return PM_ENODEV; }
    const unsigned per_sm = 228u * 1024u, cap = (unsigned)optin;
#else
    const int sm = 2;
    const unsigned per_sm = 228u * 1024u, cap = 227u * 1024u;
    (void)err; (void)err_len;
#endif
    int forced = -1;
    if (const char* env = getenv("PM_STAT_CLUSTER")) forced = atoi(env);
    int pick = 0;
#ifdef PM_EMULATE
    const int sizes[1] = {1};
#else
    const int sizes[4] = {1, 2, 4, 8};
#endif
    if (forced > 0) {
        for (int c : sizes) {      // two CTAs per SM
            if (forced > 0 && c != forced) continue;
            if (stat_layout(n, c, elem, with_mask, true).total + 1024u <= per_sm / 2) { pick = c; break; }
        }
        if (!pick)
            for (int c : sizes) {  // one CTA per SM
                if (forced > 0 && c != forced) continue;
                if (stat_layout(n, c, elem, with_mask, true).total <= cap) { pick = c; break; }
            }
    }
    P->resident = pick ? 1 : 0;
    P->cluster = pick ? pick : 1;
    P->smem = stat_layout(n, P->cluster, elem, with_mask, P->resident != 0).total;
    if (P->smem > cap) { snprintf(err, err_len, "cutout size needs more shared memory than the device has"); return PM_EINVAL; }
    long long per = (long long)(per_sm / (P->smem + 1024u));
    per = per < 1 ? 1 : per > PM_STAT_MIN_CTAS ? PM_STAT_MIN_CTAS : per;           // what the register bound allows
    long long clusters = per * sm / P->cluster;
    if (clusters < 1) clusters = 1;
    if (clusters > B) clusters = B;
    P->grid = clusters * P->cluster;
    return PM_OK;
}

#ifndef PM_EMULATE
template <typename Kern, typename... Args>
cudaError_t launch_clustered(Kern kern, const StatPlan& P, cudaStream_t st, Args... args) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)P.grid, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = P.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e;
}
#endif

bool bad_images(const void* images, int dtype, int64_t B, int n) {
    return !images || B < 0 || n < 1 || n > PM_MAX_N || (dtype != PM_F32 && dtype != PM_F64);
}

}  // namespace

extern "C" {

int pm_sky_stats_batch(const void* images, int dtype, int64_t B, int n, const double* ellipse, double* out, void* cuda_stream) {
    pm_set_error("");
    if (B == 0) return PM_OK;
    if (bad_images(images, dtype, B, n) || !ellipse || !out) { pm_set_error("bad argument"); return PM_EINVAL; }
    StatPlan P;
    char err[200];
    int rc = make_plan(B, n, dtype == PM_F32 ? 4 : 8, true, &P, err, sizeof err);
    if (rc != PM_OK) { pm_set_error(err); return rc; }
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaError_t e = dtype == PM_F32
        ? launch_clustered(pm_sky_stats_kernel<float>, P, st, (const float*)images, (long long)B, n, ellipse, out, P.resident)
        : launch_clustered(pm_sky_stats_kernel<double>, P, st, (const double*)images, (long long)B, n, ellipse, out, P.resident);
    if (e != cudaSuccess) { snprintf(err, sizeof err, "pm_sky_stats_batch: %s", cudaGetErrorString(e)); pm_set_error(err); return PM_ECUDA; }
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < P.grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)P.grid, kThreads, P.smem, [&]() { sky_body<float>((const float*)images, B, n, ellipse, out, P.resident); });
        else pm_emul::run_block((int)blk, (int)P.grid, kThreads, P.smem, [&]() { sky_body<double>((const double*)images, B, n, ellipse, out, P.resident); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}

static int clipped_stats_impl(const void* images, int dtype, int64_t B, int n, double sigma, int maxiters, double* out, int snap_at,
                              double* out_snap, void* cuda_stream, const char* what) {
    pm_set_error("");
    if (B == 0) return PM_OK;
    if (bad_images(images, dtype, B, n) || n < 8 || !out || !(sigma > 0.0) || (out_snap && snap_at < 1)) { pm_set_error("bad argument"); return PM_EINVAL; }
    StatPlan P;
    char err[200];
    int rc = make_plan(B, n, dtype == PM_F32 ? 4 : 8, false, &P, err, sizeof err);
    if (rc != PM_OK) { pm_set_error(err); return rc; }
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaError_t e = dtype == PM_F32
        ? launch_clustered(pm_clipstats_kernel<float>, P, st, (const float*)images, (long long)B, n, sigma, maxiters, out, P.resident, snap_at, out_snap)
        : launch_clustered(pm_clipstats_kernel<double>, P, st, (const double*)images, (long long)B, n, sigma, maxiters, out, P.resident, snap_at, out_snap);
    if (e != cudaSuccess) { snprintf(err, sizeof err, "%s: %s", what, cudaGetErrorString(e)); pm_set_error(err); return PM_ECUDA; }
#else
    (void)cuda_stream; (void)what;
    for (long long blk = 0; blk < P.grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)P.grid, kThreads, P.smem, [&]() { clipstats_body<float>((const float*)images, B, n, sigma, maxiters, out, P.resident, snap_at, out_snap); });
        else pm_emul::run_block((int)blk, (int)P.grid, kThreads, P.smem, [&]() { clipstats_body<double>((const double*)images, B, n, sigma, maxiters, out, P.resident, snap_at, out_snap); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}

int pm_clipped_stats_batch(const void* images, int dtype, int64_t B, int n, double sigma, int maxiters, double* out, void* cuda_stream) {
    return clipped_stats_impl(images, dtype, B, n, sigma, maxiters, out, 0, nullptr, cuda_stream, "pm_clipped_stats_batch");
}

int pm_clipped_stats2_batch(const void* images, int dtype, int64_t B, int n, double sigma, int maxiters_first, double* out_first,
                            double* out_converged, void* cuda_stream) {
    if (!out_first) { pm_set_error("bad argument"); return PM_EINVAL; }
    return clipped_stats_impl(images, dtype, B, n, sigma, 0, out_converged, maxiters_first, out_first, cuda_stream, "pm_clipped_stats2_batch");
}

// what a call would use (tests, bench): cluster size, whether the cutout is resident in shared memory, bytes per CTA
int pm_stats_plan(int n, int dtype, int with_mask, int* cluster, int* resident, int* smem_bytes) {
    StatPlan P;
    char err[200];
    int rc = make_plan(1 << 20, n, dtype == PM_F32 ? 4 : 8, with_mask != 0, &P, err, sizeof err);
    if (rc != PM_OK) { pm_set_error(err); return rc; }
    if (cluster) *cluster = P.cluster;
    if (resident) *resident = P.resident;
    if (smem_bytes) *smem_bytes = (int)P.smem;
    return PM_OK;
}

}  // extern "C"
