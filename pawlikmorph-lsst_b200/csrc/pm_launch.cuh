// pm_launch.cuh — one kernel of the per-cutout path per translation unit (they compile side by side): the macro below
// defines the __global__ entry and its C launcher `NAME_launch`; pm_capi.cu declares and calls the launchers.  With
// PM_EMULATE (CPU test-suite) the launcher runs the groups on the fibre emulator instead.
#pragma once
#include <stdio.h>
#include <string.h>

#include "pm_kernels.cuh"

#define PM_LAUNCHER_PROTO(NAME)                                                                                       \
    extern "C" int NAME##_launch(const void* params, size_t params_bytes, int dtype, unsigned grid, unsigned threads, \
                                 unsigned smem, void* cuda_stream, char* err, size_t err_len)

#ifndef PM_EMULATE
#define PM_DEFINE_KERNEL(NAME, BODY, THREADS, MIN_CTAS)                                                                \
    template <typename T> __global__ void __launch_bounds__(THREADS, MIN_CTAS) NAME##_kernel(const __grid_constant__ PM_NS::Params p) { \
        PM_NS::BODY<T>(p);                                                                                             \
    }                                                                                                                  \
    PM_LAUNCHER_PROTO(NAME) {                                                                                          \
        PM_NS::Params p;                                                                                               \
        if (params_bytes != sizeof p) { snprintf(err, err_len, "parameter block mismatch"); return PM_EINVAL; }        \
        memcpy(&p, params, sizeof p);                                                                                  \
        cudaStream_t st = (cudaStream_t)cuda_stream;                                                                   \
        cudaError_t e;                                                                                                 \
        if (dtype == PM_F32) {                                                                                         \
            e = cudaFuncSetAttribute(NAME##_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e == cudaSuccess) NAME##_kernel<float><<<grid, threads, smem, st>>>(p);                                \
        } else {                                                                                                       \
            e = cudaFuncSetAttribute(NAME##_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
            if (e == cudaSuccess) NAME##_kernel<double><<<grid, threads, smem, st>>>(p);                               \
        }                                                                                                              \
        if (e == cudaSuccess) e = cudaGetLastError();                                                                  \
        if (e != cudaSuccess) { snprintf(err, err_len, #NAME ": %s", cudaGetErrorString(e)); return PM_ECUDA; }        \
        return PM_OK;                                                                                                  \
    }
#else
#define PM_DEFINE_KERNEL(NAME, BODY, THREADS, MIN_CTAS)                                                                \
    PM_LAUNCHER_PROTO(NAME) {                                                                                          \
        PM_NS::Params p;                                                                                               \
        if (params_bytes != sizeof p) { snprintf(err, err_len, "parameter block mismatch"); return PM_EINVAL; }        \
        memcpy(&p, params, sizeof p);                                                                                  \
        (void)cuda_stream;                                                                                             \
        for (unsigned blk = 0; blk < grid; blk++) {                                                                    \
            if (dtype == PM_F32) pm_emul::run_block((int)blk, (int)grid, (int)threads, smem, [&]() { PM_NS::BODY<float>(p); });   \
            else pm_emul::run_block((int)blk, (int)grid, (int)threads, smem, [&]() { PM_NS::BODY<double>(p); });       \
        }                                                                                                              \
        return PM_OK;                                                                                                  \
    }
#endif
