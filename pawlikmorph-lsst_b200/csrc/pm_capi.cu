// pm_capi.cu — the C ABI of libpawlik_b200.so (include/pawlik_b200.h) and the kernel launches.
//
// Built by nvcc for sm_100a (the product).  The same file is also compiled by g++ with
// -DPM_EMULATE against tests/emul (CPU test-suite only): then "launching" runs every CTA on the
// fibre emulator and all pointers are host pointers.  The product library contains no CPU path.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "pm_kernels.cuh"
#include "pm_fits.cuh"

using namespace pmk;

static std::atomic<unsigned long long> g_launches{0};
static thread_local char g_err[256] = "";

// the kernels of the per-cutout path live in their own translation units (pm_k_*.cu, pm_tail.cu; pm_launch.cuh)
#define PM_LAUNCHER_DECL(NAME)                                                                                        \
    extern "C" int NAME##_launch(const void* params, size_t params_bytes, int dtype, unsigned grid, unsigned threads, \
                                 unsigned smem, void* cuda_stream, char* err, size_t err_len)
PM_LAUNCHER_DECL(pm_analyse);
PM_LAUNCHER_DECL(pm_big);
PM_LAUNCHER_DECL(pm_front);
PM_LAUNCHER_DECL(pm_minapix);
PM_LAUNCHER_DECL(pm_tail_cta);
PM_LAUNCHER_DECL(pm_tail_warp);
extern "C" unsigned pm_tail_warp_fixed_smem(int n);
extern "C" int pm_tail_warp_groups_per_cta(void);
extern "C" int pm_tail_warp_ctas_per_sm(void);
extern "C" int pm_minapix_min_ctas(void);
extern "C" int pm_front_min_ctas(void);
extern "C" int pm_tail_cta_min_ctas(void);
extern "C" int pm_tail_cta_threads(void);
extern "C" int pm_front_threads(void);
extern "C" int pm_mono_threads(void);
extern "C" int pm_mono_min_ctas(void);
extern "C" int pm_minapix_threads(void);

#ifndef PM_EMULATE
__global__ void __launch_bounds__(kThreads, 1)
pm_aperpixmap_kernel(int n, const double* rad, long long B, int ns, double frac, uint8_t* mask_out, unsigned smem_bytes) {
    aperpixmap_body(n, rad, B, ns, frac, mask_out, smem_bytes);
}
static thread_local int g_sm_count = 0, g_cc_major = 0, g_cc_minor = 0, g_dev = -1;
static thread_local size_t g_smem_optin = 0;
static int device_props() {      // properties of the CURRENT device of the calling thread (re-queried when it changes)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaGetDevice: %s", cudaGetErrorString(e)); return PM_ENODEV; }
    if (dev == g_dev && g_sm_count) return PM_OK;
    int sm = 0, maj = 0, mnr = 0, optin = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (sm <= 0) { snprintf(g_err, sizeof g_err, "no CUDA device"); return PM_ENODEV; }
    if (maj < 9) { snprintf(g_err, sizeof g_err, "device compute capability %d.%d: this library is built for sm_100a", maj, mnr); return PM_ENODEV; }
    g_cc_major = maj; g_cc_minor = mnr; g_smem_optin = (size_t)optin; g_sm_count = sm; g_dev = dev;
    return PM_OK;
}
#else
static int g_sm_count = 2, g_cc_major = 10, g_cc_minor = 0;
static size_t g_smem_optin = 232448;
static int device_props() { return PM_OK; }
#endif

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// shared-memory size and CTAs per SM of a kernel that keeps every fixed piece in shared memory (monolith, front, CTA
// tail) for (n, fs); `built`: the CTAs per SM its launch bounds were compiled for
static void launch_shape(int n, int fs, unsigned* smem, int* ctas_per_sm, int built = PM_MIN_CTAS, const char* env = "PM_CTAS_PER_SM") {
    const unsigned fixed = smem_fixed_bytes(n, fs);
    const unsigned cap1 = (unsigned)(g_smem_optin ? g_smem_optin : 232448);   // 227 KB
    int ctas = built;
    while (ctas > 1 && fixed + 24 * 1024 > (cap1 + 1024) / (unsigned)ctas - 1024) ctas--;   // large n: fewer, fatter CTAs
    ctas = env_int(env, ctas);
    if (ctas < 1) ctas = 1;
    if (ctas > built) ctas = built;
    unsigned per = (cap1 + 1024) / (unsigned)ctas - 1024;
    per &= ~15u;
    const int forced = env_int("PM_SMEM_BYTES", 0);
    if (forced > 0) per = (unsigned)forced & ~15u;
    if (per > cap1) per = cap1;
    *smem = per;
    *ctas_per_sm = ctas;
}
// the minapix kernel: aperture plane + window + list; as many CTAs per SM as the launch bounds allow
static void minapix_shape(int n, unsigned* smem, int* ctas_per_sm) {
    const unsigned fixed = smem_fixed_bytes(n, 1, SM_APER);
    const unsigned cap1 = (unsigned)(g_smem_optin ? g_smem_optin : 232448);
    const int built = pm_minapix_min_ctas();
    int ctas = env_int("PM_MINAPIX_CTAS", built > 3 ? 3 : built);
    if (ctas > built) ctas = built;
    if (ctas < 1) ctas = 1;
    while (ctas > 1 && fixed + 16 * 1024 > (cap1 + 1024) / (unsigned)ctas - 1024) ctas--;
    unsigned per = ((cap1 + 1024) / (unsigned)ctas - 1024) & ~15u;
    if (per > cap1) per = cap1;
    *smem = per;
    *ctas_per_sm = ctas;
}

// ---- workspace plan of the split path -------------------------------------------------------------------------------
// [0, 256) work counters | big list (int per cutout of a chunk) | per-group scratch (the largest of the kernels') |
// hand-over slots of one chunk.  Upper bounds on the grids are used so that the size does not depend on the device.
#define PM_WS_HEADER 256
#define PM_MAX_TAIL_GROUPS 5120        /* 32 warps x 160 SMs */
#define PM_MAX_MINAPIX_GRID 640
struct Plan {
    long long chunk;
    size_t off_big, off_scratch, scratch_bytes, off_state, total;
};
static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
static Plan make_plan(int64_t B, int n) {
    Plan pl;
    long long chunk = 49152;                                               // >= 20 waves of the tail kernel's groups
    const unsigned long long stride = state_stride(n);
    const long long by_mem = (long long)((4ull << 30) / stride);             // <= 4 GiB of hand-over slots
    if (chunk > by_mem) chunk = by_mem < 1024 ? 1024 : by_mem;
    const int forced = env_int("PM_CHUNK", 0);
    if (forced > 0) chunk = forced;
    if (chunk > B) chunk = B;
    pl.chunk = chunk;
    auto cap = [&](long long g) { return (unsigned long long)(B < g ? B : g); };
    unsigned long long sc = cap(PM_MAX_GRID) * scratch_bytes_per_cta(n);                            // front, monolith pass, CTA tail
    const unsigned long long sm_ = cap(PM_MAX_MINAPIX_GRID) * scratch_bytes_minapix(n);
    const unsigned long long st_ = cap(PM_MAX_TAIL_GROUPS) * scratch_bytes_tail(n);
    if (sm_ > sc) sc = sm_;
    if (st_ > sc) sc = st_;
    pl.off_big = PM_WS_HEADER;
    pl.off_scratch = up256(pl.off_big + 4 * (size_t)chunk);
    pl.scratch_bytes = up256((size_t)sc);
    pl.off_state = pl.off_scratch + pl.scratch_bytes;
    pl.total = pl.off_state + (size_t)chunk * (size_t)stride;
    return pl;
}

extern "C" {

int pm_abi_version(void) { return PM_ABI_VERSION; }
const char* pm_last_cuda_error(void) { return g_err; }
void pm_set_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); }     // for the other translation units
void pm_count_launch(void) { g_launches.fetch_add(1); }
uint64_t pm_launch_count(void) { return g_launches.load(); }

int pm_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    int rc = device_props();
    if (rc != PM_OK) return rc;
    if (sm_count) *sm_count = g_sm_count;
    if (cc_major) *cc_major = g_cc_major;
    if (cc_minor) *cc_minor = g_cc_minor;
    if (smem_optin) *smem_optin = g_smem_optin;
    return PM_OK;
}

size_t pm_workspace_bytes(int64_t B, int n, int dtype) {
    (void)dtype;
    if (B <= 0 || n <= 0 || n > PM_MAX_N) return 0;
    return make_plan(B, n).total;
}

// one kernel launch of the per-cutout path (the first failure sticks; later launches are skipped)
#define PM_LAUNCH(NAME, GRID, THREADS, SMEM)                                                                          \
    do {                                                                                                              \
        if (lrc == PM_OK) {                                                                                           \
            lrc = NAME##_launch(&p, sizeof p, dtype, (unsigned)(GRID), (unsigned)(THREADS), (SMEM), cuda_stream, g_err, sizeof g_err); \
            g_launches.fetch_add(1);                                                                                  \
        }                                                                                                             \
    } while (0)

int pm_analyse_batch(const void* images, int dtype, int64_t B, int n, const double* sky, const double* sky_err,
                     int filter_size, uint32_t flags, const pm_overrides_t* ov, const pm_outputs_t* out,
                     pm_row_t* rows, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    g_err[0] = 0;
    if (B == 0) return PM_OK;
    if (!images || !rows || !workspace || B < 0) { snprintf(g_err, sizeof g_err, "null pointer / negative batch"); return PM_EINVAL; }
    if (dtype != PM_F32 && dtype != PM_F64) { snprintf(g_err, sizeof g_err, "dtype must be PM_F32 or PM_F64"); return PM_EINVAL; }
    if (n < 8 || n > PM_MAX_N) { snprintf(g_err, sizeof g_err, "n must be in [8, %d]", PM_MAX_N); return PM_EINVAL; }
    if (filter_size < 1 || filter_size > 15 || (filter_size & 1) == 0 || n / filter_size < 1) {
        snprintf(g_err, sizeof g_err, "filter_size must be odd, in 1..15 and <= n");   // pixmap.py:160-162
        return PM_EINVAL;
    }
    if (ov && ov->starmask) { snprintf(g_err, sizeof g_err, "star masks are not supported yet"); return PM_EINVAL; }
    if (B > 0x7fffffffLL) { snprintf(g_err, sizeof g_err, "batch too large"); return PM_EINVAL; }
    const Plan pl = make_plan(B, n);
    if (workspace_bytes < pl.total) { snprintf(g_err, sizeof g_err, "workspace too small"); return PM_ENOSPC; }
    if (((uintptr_t)workspace & 15) != 0) { snprintf(g_err, sizeof g_err, "workspace must be 16-byte aligned"); return PM_EINVAL; }
    int rc = device_props();
    if (rc != PM_OK) return rc;

    Params p;
    memset(&p, 0, sizeof p);
    p.images = images; p.B = B; p.n = n; p.fs = filter_size; p.flags = flags;
    p.sky = sky; p.sky_err = sky_err;
    if (ov) p.ov = *ov;
    if (out) p.out = *out;
    p.rows = rows;
    p.counter = (unsigned long long*)workspace;
    p.scratch = (unsigned char*)workspace + pl.off_scratch;
    p.scratch_per_cta = scratch_bytes_per_cta(n);
    p.phase_cycles = (unsigned long long*)p.out.phase_cycles_out;
    p.big_list = (int*)((unsigned char*)workspace + pl.off_big);
    p.state = (unsigned char*)workspace + pl.off_state;
    p.state_stride = state_stride(n);
    p.list_cap = state_list_cap(n);
    p.cand_cap = state_cand_cap();
    {   // test hooks: smaller hand-over slots push more cutouts through the monolith pass
        const int lc = env_int("PM_TEST_LIST_CAP", 0), cc = env_int("PM_TEST_CAND_CAP", 0);
        if (cc > 0 && cc < p.cand_cap) p.cand_cap = cc;
        if (lc > 0 && lc < p.list_cap) p.list_cap = lc;
    }
    unsigned smem = 0;
    int ctas = 1;
    launch_shape(n, filter_size, &smem, &ctas, pm_mono_min_ctas());
    if (smem < smem_fixed_bytes(n, filter_size) + 4096) { snprintf(g_err, sizeof g_err, "shared memory too small for n"); return PM_EINVAL; }
    long long grid_cap = (long long)g_sm_count * ctas;
    if (grid_cap > PM_MAX_GRID) grid_cap = PM_MAX_GRID;
    const int g_env = env_int("PM_GRID", 0);
    if (g_env > 0 && g_env < grid_cap) grid_cap = g_env;
    const bool mono = (flags & PM_PATH_SPLIT) == 0 && env_int("PM_PATH_SPLIT", 0) == 0;
    const bool tail_cta = (flags & PM_TAIL_WARP) == 0;

    int lrc = PM_OK;
    auto zero_counters = [&]() -> int {
#ifndef PM_EMULATE
        cudaError_t e = cudaMemsetAsync(p.counter, 0, PM_WS_HEADER, (cudaStream_t)cuda_stream);
        if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "pm_analyse_batch: %s", cudaGetErrorString(e)); return PM_ECUDA; }
#else
        memset(p.counter, 0, PM_WS_HEADER);
#endif
        return PM_OK;
    };
    if (mono) {
        p.smem_bytes = smem;
        p.b0 = 0; p.nb = B;
        const long long grid = grid_cap < B ? grid_cap : B;
        lrc = zero_counters();
        PM_LAUNCH(pm_analyse, grid, pm_mono_threads(), smem);
        return lrc;
    }
    // ---- the split path, chunk by chunk
    unsigned smem_f = 0, smem_t = 0;
    int ctas_f = 1, ctas_t = 1;
    launch_shape(n, filter_size, &smem_f, &ctas_f, pm_front_min_ctas(), "PM_FRONT_CTAS");
    launch_shape(n, filter_size, &smem_t, &ctas_t, pm_tail_cta_min_ctas(), "PM_TAILCTA_CTAS");
    long long grid_f_cap = (long long)g_sm_count * ctas_f, grid_t_cap = (long long)g_sm_count * ctas_t;
    if (grid_f_cap > PM_MAX_GRID) grid_f_cap = PM_MAX_GRID;
    if (grid_t_cap > PM_MAX_GRID) grid_t_cap = PM_MAX_GRID;
    unsigned smem_m = 0;
    int ctas_m = 1;
    minapix_shape(n, &smem_m, &ctas_m);
    long long grid_m_cap = (long long)g_sm_count * ctas_m;
    if (grid_m_cap > PM_MAX_MINAPIX_GRID) grid_m_cap = PM_MAX_MINAPIX_GRID;
    // tail, one warp per cutout: G groups per SM, each with an equal share of the shared memory
    const int gpc = pm_tail_warp_groups_per_cta();
    int tail_ctas = env_int("PM_TAIL_CTAS", pm_tail_warp_ctas_per_sm());
    if (tail_ctas < 1) tail_ctas = 1;
    if (tail_ctas > pm_tail_warp_ctas_per_sm()) tail_ctas = pm_tail_warp_ctas_per_sm();
    const unsigned cap1 = (unsigned)(g_smem_optin ? g_smem_optin : 232448);
    unsigned smem_g = (((cap1 + 1024) / (unsigned)tail_ctas - 1024) / (unsigned)gpc) & ~15u;
    if (smem_g > 65536u) smem_g = 65536u;
    if (smem_g < pm_tail_warp_fixed_smem(n) + 1024u) { snprintf(g_err, sizeof g_err, "shared memory too small for the tail kernel"); return PM_EINVAL; }
    long long groups_cap = (long long)g_sm_count * tail_ctas * gpc;
    if (groups_cap > PM_MAX_TAIL_GROUPS) groups_cap = PM_MAX_TAIL_GROUPS;
    for (long long b0 = 0; b0 < B && lrc == PM_OK; b0 += pl.chunk) {
        const long long nb = B - b0 < pl.chunk ? B - b0 : pl.chunk;
        p.b0 = b0; p.nb = nb;
        const long long grid = grid_cap < nb ? grid_cap : nb;
        lrc = zero_counters();
        // front
        p.smem_bytes = smem_f; p.scratch_per_cta = scratch_bytes_per_cta(n);
        PM_LAUNCH(pm_front, grid_f_cap < nb ? grid_f_cap : nb, pm_front_threads(), smem_f);
        // minapix (skipped when the centre is supplied or the path stops before it)
        if (!p.ov.centroid_in && !(flags & (PM_STOP_AFTER_PIXELMAP | PM_STOP_AFTER_RMAX))) {
            const long long grid_m = grid_m_cap < nb ? grid_m_cap : nb;
            p.smem_bytes = smem_m; p.scratch_per_cta = scratch_bytes_minapix(n);
            PM_LAUNCH(pm_minapix, grid_m, pm_minapix_threads(), smem_m);
        }
        // tail (also writes the rows of the cutouts that ended early)
        if (tail_cta) {
            p.smem_bytes = smem_t; p.scratch_per_cta = scratch_bytes_per_cta(n);
            PM_LAUNCH(pm_tail_cta, grid_t_cap < nb ? grid_t_cap : nb, pm_tail_cta_threads(), smem_t);
        } else {
            p.smem_bytes = smem_g; p.scratch_per_cta = scratch_bytes_tail(n);
            const long long groups = groups_cap < nb ? groups_cap : nb;
            PM_LAUNCH(pm_tail_warp, (groups + gpc - 1) / gpc, 32 * gpc, smem_g * (unsigned)gpc);
        }
        // the monolith pass over whatever the front kernel handed over (normally nothing)
        p.smem_bytes = smem; p.scratch_per_cta = scratch_bytes_per_cta(n);
        PM_LAUNCH(pm_big, grid, pm_mono_threads(), smem);
    }
    return lrc;
}

int pm_aperpixmap_batch(int n, const double* rad, int64_t B, int nsubpix, double frac, uint8_t* mask_out,
                        void* cuda_stream) {
    g_err[0] = 0;
    if (B == 0) return PM_OK;
    if (!rad || !mask_out || B < 0 || n < 1 || n > PM_MAX_N || nsubpix < 1 || nsubpix > 32) {
        snprintf(g_err, sizeof g_err, "bad argument");
        return PM_EINVAL;
    }
    int rc = device_props();
    if (rc != PM_OK) return rc;
    const unsigned smem = aperpixmap_smem_bytes(n, nsubpix);
    long long grid = B < 4LL * g_sm_count ? B : 4LL * g_sm_count;
#ifndef PM_EMULATE
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaError_t e = cudaFuncSetAttribute(pm_aperpixmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        pm_aperpixmap_kernel<<<(unsigned)grid, kThreads, smem, st>>>(n, rad, B, nsubpix, frac, mask_out, smem);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "pm_aperpixmap_batch: %s", cudaGetErrorString(e)); return PM_ECUDA; }
#else
    (void)cuda_stream;
    for (long long blk = 0; blk < grid; blk++)
        pm_emul::run_block((int)blk, (int)grid, kThreads, smem, [&]() { aperpixmap_body(n, rad, B, nsubpix, frac, mask_out, smem); });
#endif
    g_launches.fetch_add(1);
    return PM_OK;
}

int pm_fits_decode_batch(const void* payload, int bitpix, double bzero, double bscale, int64_t count, void* out,
                         int out_dtype, void* cuda_stream) {
    g_err[0] = 0;
    if (count == 0) return PM_OK;
    if (!payload || !out || count < 0 || fits_bytes_per_pixel(bitpix) == 0 || (out_dtype != PM_F32 && out_dtype != PM_F64)) {
        snprintf(g_err, sizeof g_err, "bad argument");
        return PM_EINVAL;
    }
#ifndef PM_EMULATE
    int rc = device_props();
    if (rc != PM_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const unsigned char* raw = (const unsigned char*)payload;
    // HBM-bound byte work: up to 8 CTAs of 256 threads per SM (a multiple of the SM count), grid-stride beyond that
    const unsigned max_ctas = 8u * (unsigned)g_sm_count;
    if (out_dtype == PM_F32) fits_decode_launch<float>(raw, bitpix, bzero, bscale, (size_t)count, (float*)out, max_ctas, st);
    else fits_decode_launch<double>(raw, bitpix, bzero, bscale, (size_t)count, (double*)out, max_ctas, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "pm_fits_decode_batch: %s", cudaGetErrorString(e)); return PM_ECUDA; }
#else
    (void)cuda_stream;
    for (int64_t i = 0; i < count; i++) {
        const double v = fits_pixel((const unsigned char*)payload, (size_t)i, bitpix, bzero, bscale);
        if (out_dtype == PM_F32) ((float*)out)[i] = (float)v; else ((double*)out)[i] = v;
    }
#endif
    g_launches.fetch_add(1);
    return PM_OK;
}

int pm_fits_decode_window_batch(const void* payload, int bitpix, double bzero, double bscale, int64_t B, int H, int W,
                                const int32_t* origin, int npix, void* out, int out_dtype, void* cuda_stream) {
    g_err[0] = 0;
    if (B == 0) return PM_OK;
    const int bpp = fits_bytes_per_pixel(bitpix);
    if (!payload || !out || !origin || B < 0 || bpp == 0 || H <= 0 || W <= 0 || npix <= 0 || npix > H || npix > W ||
        (out_dtype != PM_F32 && out_dtype != PM_F64) || ((uintptr_t)payload % (uintptr_t)bpp) != 0) {
        snprintf(g_err, sizeof g_err, "bad argument");
        return PM_EINVAL;
    }
#ifndef PM_EMULATE
    int rc = device_props();
    if (rc != PM_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const unsigned char* raw = (const unsigned char*)payload;
    const unsigned max_ctas = 8u * (unsigned)g_sm_count;
    const bool vec = npix % 4 == 0 && (((uintptr_t)out) & 15) == 0;
    if (out_dtype == PM_F32) {
        if (vec) fits_window_launch<float, 4>(raw, bitpix, bzero, bscale, B, H, W, origin, npix, (float*)out, max_ctas, st);
        else fits_window_launch<float, 1>(raw, bitpix, bzero, bscale, B, H, W, origin, npix, (float*)out, max_ctas, st);
    } else {
        if (vec) fits_window_launch<double, 4>(raw, bitpix, bzero, bscale, B, H, W, origin, npix, (double*)out, max_ctas, st);
        else fits_window_launch<double, 1>(raw, bitpix, bzero, bscale, B, H, W, origin, npix, (double*)out, max_ctas, st);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "pm_fits_decode_window_batch: %s", cudaGetErrorString(e)); return PM_ECUDA; }
#else
    (void)cuda_stream;
    for (int64_t b = 0; b < B; b++)
        for (int r = 0; r < npix; r++)
            for (int c = 0; c < npix; c++) {
                const size_t src = ((size_t)b * H + (size_t)(origin[2 * b] + r)) * W + (size_t)(origin[2 * b + 1] + c);
                const double v = fits_pixel((const unsigned char*)payload, src, bitpix, bzero, bscale);
                const size_t dst = ((size_t)b * npix + r) * npix + c;
                if (out_dtype == PM_F32) ((float*)out)[dst] = (float)v; else ((double*)out)[dst] = v;
            }
#endif
    g_launches.fetch_add(1);
    return PM_OK;
}

}  // extern "C"
