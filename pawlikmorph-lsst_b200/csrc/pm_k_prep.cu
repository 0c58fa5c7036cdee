// pm_k_prep.cu — C ABI and launches of the pre-processing kernels (pm_prep.cuh): the Gaussian fit of
// skyBackground._gauss2dfit, source detection + the cleaning of imageutils.maskstarsSEG.  One CTA per cutout.
#include <stdio.h>
#include <string.h>

#include "pm_prep.cuh"

using namespace pmk;

extern "C" const char* pm_last_cuda_error(void);
extern "C" void pm_set_error(const char* msg);
extern "C" void pm_count_launch(void);

// Gaussian fit: 4 CTAs/SM (64 registers; the 7 x 7 algebra of warp 0 spills, the pixel loop does not).  Measured on 8 192 x 256^2:
// 84 k / 98 k / 105 k / 108 k cutouts/s at 1 / 2 / 3 / 4 CTAs per SM (profiles/r2af_gauss_dmma.txt).
#ifndef PM_GAUSS_MIN_CTAS
#define PM_GAUSS_MIN_CTAS 4
#endif
#ifndef PM_EMULATE
template <typename T> __global__ void __launch_bounds__(kThreads, PM_GAUSS_MIN_CTAS)
pm_gaussfit_kernel(const T* images, long long B, int n, const double* median, const double* vmax, const double* p0, double* out) {
    gaussfit_body<T>(images, B, n, median, vmax, p0, out);
}
template <typename T> __global__ void __launch_bounds__(kThreads, PM_GAUSS_MIN_CTAS)
pm_sersicfit_kernel(const T* images, long long B, int n, const double* p0, double* out) {
    sersicfit_body<T>(images, B, n, p0, out);
}
template <typename T> __global__ void __launch_bounds__(kThreads, 4)
pm_ellipse_sum_kernel(const T* images, int n, long long M, const long long* idx, const double* ell, double* out) {
    ellipse_sum_body<T>(images, n, M, idx, ell, out);
}
template <typename T> __global__ void __launch_bounds__(kThreads, 2)
pm_maskstars_kernel(const T* images, long long B, int n, const double* threshold, int npixels, const double* mean, const double* std_,
                    unsigned long long seed, int* labels_out, T* clean_out, double* out, unsigned char* scratch,
                    unsigned long long scratch_per_cta, int max_runs) {
    maskstars_body<T>(images, B, n, threshold, npixels, mean, std_, seed, labels_out, clean_out, out, scratch, scratch_per_cta, max_runs);
}
static int sm_count() {
    int dev = 0, sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    return sm;
}
#define PREP_CHECK(what)                                                                     \
    do {                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess) {                                                             \
            char m_[200];                                                                    \
            snprintf(m_, sizeof m_, what ": %s", cudaGetErrorString(e_));                    \
            pm_set_error(m_);                                                                \
            return PM_ECUDA;                                                                 \
        }                                                                                    \
    } while (0)
#else
static int sm_count() { return 2; }
#endif

static bool bad_images(const void* images, int dtype, int64_t B, int n) {
    return !images || B < 0 || n < 8 || n > PM_MAX_N || (dtype != PM_F32 && dtype != PM_F64);
}

extern "C" {

int pm_gaussfit_batch(const void* images, int dtype, int64_t B, int n, const double* median, const double* vmax, const double* p0,
                      double* out, void* cuda_stream) {
    pm_set_error("");
    if (B == 0) return PM_OK;
    if (bad_images(images, dtype, B, n) || !out || (!p0 && (!median || !vmax))) { pm_set_error("bad argument"); return PM_EINVAL; }
    const unsigned smem = align16((unsigned)sizeof(Sh)) + align16((unsigned)sizeof(GaussSmem));
    const int sm = sm_count();
    if (sm <= 0) { pm_set_error("no CUDA device"); return PM_ENODEV; }
    // iteration counts differ a lot between cutouts: eight short static slices per resident CTA balance the SMs (the
    // scheduler hands out the later CTAs as the earlier ones finish)
    const long long grid = B < 8LL * PM_GAUSS_MIN_CTAS * sm ? B : 8LL * PM_GAUSS_MIN_CTAS * sm;
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == PM_F32) pm_gaussfit_kernel<float><<<(unsigned)grid, kThreads, smem, st>>>((const float*)images, B, n, median, vmax, p0, out);
    else pm_gaussfit_kernel<double><<<(unsigned)grid, kThreads, smem, st>>>((const double*)images, B, n, median, vmax, p0, out);
    PREP_CHECK("pm_gaussfit_batch");
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { gaussfit_body<float>((const float*)images, B, n, median, vmax, p0, out); });
        else pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { gaussfit_body<double>((const double*)images, B, n, median, vmax, p0, out); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}

size_t pm_maskstars_workspace_bytes(int64_t B, int n) {
    if (B <= 0 || n < 8 || n > PM_MAX_N) return 0;
    const long long grid = B < PM_MAX_GRID ? B : PM_MAX_GRID;
    return (size_t)grid * (size_t)maskstars_scratch_bytes(n);
}

int pm_maskstars_batch(const void* images, int dtype, int64_t B, int n, const double* threshold, int npixels, const double* mean,
                       const double* std_, uint64_t seed, int32_t* labels_out, void* clean_out, double* out, void* workspace,
                       size_t workspace_bytes, void* cuda_stream) {
    pm_set_error("");
    if (B == 0) return PM_OK;
    if (bad_images(images, dtype, B, n) || !threshold || !out || !workspace || npixels < 1 || (clean_out && (!mean || !std_))) {
        pm_set_error("bad argument");
        return PM_EINVAL;
    }
    if (workspace_bytes < pm_maskstars_workspace_bytes(B, n) || ((uintptr_t)workspace & 15)) { pm_set_error("workspace too small / misaligned"); return PM_ENOSPC; }
    const unsigned smem = maskstars_smem_bytes(n);
    const int sm = sm_count();
    if (sm <= 0) { pm_set_error("no CUDA device"); return PM_ENODEV; }
    long long grid = B < 2LL * sm ? B : 2LL * sm;
    if (grid > PM_MAX_GRID) grid = PM_MAX_GRID;
    const unsigned long long per = maskstars_scratch_bytes(n);
    const int max_runs = maskstars_max_runs(n);
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == PM_F32) {
        cudaFuncSetAttribute(pm_maskstars_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pm_maskstars_kernel<float><<<(unsigned)grid, kThreads, smem, st>>>((const float*)images, B, n, threshold, npixels, mean, std_, seed, labels_out,
                                                                          (float*)clean_out, out, (unsigned char*)workspace, per, max_runs);
    } else {
        cudaFuncSetAttribute(pm_maskstars_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pm_maskstars_kernel<double><<<(unsigned)grid, kThreads, smem, st>>>((const double*)images, B, n, threshold, npixels, mean, std_, seed, labels_out,
                                                                           (double*)clean_out, out, (unsigned char*)workspace, per, max_runs);
    }
    PREP_CHECK("pm_maskstars_batch");
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { maskstars_body<float>((const float*)images, B, n, threshold, npixels, mean, std_, seed, labels_out, (float*)clean_out, out, (unsigned char*)workspace, per, max_runs); });
        else pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { maskstars_body<double>((const double*)images, B, n, threshold, npixels, mean, std_, seed, labels_out, (double*)clean_out, out, (unsigned char*)workspace, per, max_runs); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}

int pm_sersicfit_batch(const void* images, int dtype, int64_t B, int n, const double* p0, double* out, void* cuda_stream) {
    pm_set_error("");
    if (B == 0) return PM_OK;
    if (bad_images(images, dtype, B, n) || !out || !p0) { pm_set_error("bad argument"); return PM_EINVAL; }
    const unsigned smem = align16((unsigned)sizeof(Sh)) + align16((unsigned)sizeof(GaussSmem));
    const int sm = sm_count();
    if (sm <= 0) { pm_set_error("no CUDA device"); return PM_ENODEV; }
    const long long grid = B < 8LL * PM_GAUSS_MIN_CTAS * sm ? B : 8LL * PM_GAUSS_MIN_CTAS * sm;
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == PM_F32) pm_sersicfit_kernel<float><<<(unsigned)grid, kThreads, smem, st>>>((const float*)images, B, n, p0, out);
    else pm_sersicfit_kernel<double><<<(unsigned)grid, kThreads, smem, st>>>((const double*)images, B, n, p0, out);
    PREP_CHECK("pm_sersicfit_batch");
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { sersicfit_body<float>((const float*)images, B, n, p0, out); });
        else pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { sersicfit_body<double>((const double*)images, B, n, p0, out); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}

int pm_ellipse_sum_batch(const void* images, int dtype, int n, int64_t M, const int64_t* idx, const double* ell, double* out,
                         void* cuda_stream) {
    pm_set_error("");
    if (M == 0) return PM_OK;
    if (bad_images(images, dtype, M, n) || !ell || !out) { pm_set_error("bad argument"); return PM_EINVAL; }
    const unsigned smem = align16((unsigned)sizeof(Sh));
    const int sm = sm_count();
    if (sm <= 0) { pm_set_error("no CUDA device"); return PM_ENODEV; }
    const long long grid = M < 8LL * sm ? M : 8LL * sm;
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == PM_F32) pm_ellipse_sum_kernel<float><<<(unsigned)grid, kThreads, smem, st>>>((const float*)images, n, M, (const long long*)idx, ell, out);
    else pm_ellipse_sum_kernel<double><<<(unsigned)grid, kThreads, smem, st>>>((const double*)images, n, M, (const long long*)idx, ell, out);
    PREP_CHECK("pm_ellipse_sum_batch");
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < grid; blk++) {
        if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { ellipse_sum_body<float>((const float*)images, n, M, (const long long*)idx, ell, out); });
        else pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { ellipse_sum_body<double>((const double*)images, n, M, (const long long*)idx, ell, out); });
    }
#endif
    pm_count_launch();
    return PM_OK;
}


}  // extern "C"
