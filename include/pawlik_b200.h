/* pawlik_b200.h — C ABI of the B200-native PawlikMorph-LSST morphology hot path.
 *
 * One shared object (libpawlik_b200.so), extern "C", plain pointers and sizes.
 * The reference (lewisfish/PawlikMorph-LSST) is pure Python and has no FFI layer;
 * its drop-in boundary is the Python module API of
 *   pawlikMorphLSST/pixmap.py, apertures.py, asymmetry.py, casgm.py
 * as called from main.py:421-495.  Each entry point below names the reference
 * function(s) it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - All data pointers are CALLER-OWNED DEVICE memory (e.g. torch tensors) unless
 *    marked HOST.  The library allocates nothing persistent and frees nothing.
 *  - Calls are asynchronous on `cuda_stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 on success, a negative PM_E* code for argument / launch
 *    errors.  Per-cutout failures never abort a batch: they are reported in
 *    pm_row_t.status and the statistics of that row stay at the reference's
 *    sentinel -99 (result.py:26-54).
 *  - Segmaps, star masks and aperture masks are binary {0,1} uint8 [B, n, n].
 *  - Images are [B, n, n] row-major, float32 (PM_F32) or float64 (PM_F64), WITH the
 *    sky still in them; `sky[b]` is subtracted on the fly in float64 exactly where
 *    main.py:442 does `img -= sky`.  Pass sky = 0 for already-subtracted images.
 *  - Re-entrant, no global state (a launch counter aside).  n <= PM_MAX_N.
 */
#ifndef PAWLIK_B200_H
#define PAWLIK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_ABI_VERSION 2
#define PM_MAX_N 640

/* dtype of `images` */
#define PM_F32 0
#define PM_F64 1

/* error codes */
#define PM_OK 0
#define PM_EINVAL (-1)     /* bad argument (null pointer, n out of range, even filter size, ...) */
#define PM_ENOSPC (-2)     /* workspace too small: call pm_workspace_bytes() */
#define PM_ECUDA (-3)      /* CUDA launch / runtime error; see pm_last_cuda_error() */
#define PM_ENODEV (-4)     /* no CUDA device / not an sm_100 device */

/* which statistics to compute (flags).  Mirrors the CLI switches of imganalysis.py:8-70
 * as main.py:470-495 uses them: A is computed for -A/-Aall/-cas, As+As90 for -As/-Aall,
 * r20/r80/C/gini/m20 for -cas.  PM_DO_S adds the smoothness S that main.py:494 leaves
 * commented out (casgm.py:276-406). */
#define PM_DO_A 1u
#define PM_DO_AS 2u
#define PM_DO_CAS 4u
#define PM_DO_S 8u
#define PM_DO_ALL 15u
/* debugging / test switches */
#define PM_FORCE_EXACT_FILTER 0x100u /* pixelmap: take the sequential scipy-recurrence path for every pixel */
/* stage control, used by the per-cutout Python functions (pixmap.pixelmap, pixmap.calcRmax, ...) */
#define PM_STOP_AFTER_PIXELMAP 0x200u /* only pixelmap + checkPixelmapEdges */
#define PM_STOP_AFTER_RMAX 0x400u     /* ... + calcRmax + aperpixmap */
#define PM_A_ANGLE90 0x800u           /* calcA on the image with angle = 90 instead of 180 (asymmetry.py:132) */
#define PM_NO_NOISECORRECT 0x1000u    /* calcA(..., noisecorrect=False): Abgr = 0, A not corrected */
#define PM_NO_PRUNE 0x4000u           /* tuning/test: evaluate every minapix candidate completely (no exact pruning) */
#define PM_L2_PREFETCH 0x2000u        /* tuning: TMA-prefetch the next cutout into L2 while the current one runs (off) */
#define PM_RADIX_SORT 0x20000u        /* tuning/test: always the LSD radix sort (the fallback of the bucket sort) */
#define PM_PATH_SPLIT 0x8000u         /* three kernels per chunk split by resource class (front / minapix / tail)
                                         instead of the one fused kernel (the default: measured faster, profiles/) */
#define PM_TAIL_WARP 0x10000u         /* with PM_PATH_SPLIT: the tail kernel (r20/r80, calcA, S) with one warp per cutout
                                         instead of one CTA per cutout */

/* pm_row_t.status */
#define PM_ST_OK 0
#define PM_ST_FAINT_CENTRE 1 /* pixmap.py:178-179 "Central pixel too faint" -> default Result (main.py:420-424) */
#define PM_ST_NO_CANDIDATE 2 /* no positive pixel in image*segmap: np.argmin([]) upstream (asymmetry.py:127) */
#define PM_ST_BAD_BRACKET 3  /* casgm.py:65-76 would raise (radius walked to <= 0 / no sign change) */
#define PM_ST_EMPTY_SEGMAP 4 /* supplied segmap has no pixel set: calcRmax would raise (pixmap.py:51-53) */
#define PM_ST_WORKSPACE 5    /* per-CTA scratch exhausted (cannot happen with pm_workspace_bytes()) */

/* One result row per cutout: the fields of result.Result (result.py:26-76) this path fills.
 * 16 x float64 = 128 bytes.  Unset statistics hold -99. */
typedef struct {
    double apix_col, apix_row; /* Result.apix = [col, row]            asymmetry.py:129 */
    double rmax;               /* Result.rmax                         pixmap.py:26-55 */
    double r20, r80, C;        /* casgm.py:124-184 */
    double A, Abgr;            /* Result.A = [A, Abgr]                asymmetry.py:132-257 */
    double S;                  /* casgm.py:276-338 */
    double As, As90;           /* Result.As[0], Result.As90[0] */
    double gini, m20;          /* casgm.py:187-273 */
    double object_edge;        /* Result.objectEdge (0/1)             pixmap.py:91-127 */
    double status;             /* PM_ST_* */
    double n_candidates;       /* number of centroid candidates tried by minapix (asymmetry.py:101-108) */
} pm_row_t;

/* Optional per-cutout overrides that turn the fused pipeline into the reference's individual
 * functions (used by the per-cutout Python API; all nullable, all DEVICE pointers). */
typedef struct {
    const uint8_t *segmap_in;   /* [B,n,n]  skip pixelmap, use this map (main.py:428-439 -segmap path) */
    const uint8_t *starmask;    /* [B,n,n]  star mask, 1 = keep (pixmap.py:156-158, asymmetry.py:82-86,:182-191) */
    const uint8_t *apermask_in; /* [B,n,n]  aperture mask instead of aperpixmap(n, rmax, 9, 0.1) */
    const double *centroid_in;  /* [B,2]    (col,row): skip minapix, use this centre (calcA / calcR20_R80 / smoothness) */
    const double *radius_in;    /* [B]      use this radius instead of calcRmax (calcR20_R80 / smoothness / aperpixmap) */
    const double *r20_in;       /* [B]      use this r20 in smoothness */
    const double *threshold_in; /* [B]      pixelmap threshold instead of sky + sky_err (pixmap.py:130) */
    const double *smooth_sky_in; /* [B]     the `sky` argument of casgm.smoothness (casgm.py:276) when it differs from
                                            the value subtracted from the image (already-subtracted images: sky = 0) */
} pm_overrides_t;

/* Optional outputs beyond the rows (all nullable, DEVICE). */
typedef struct {
    uint8_t *segmap_out;     /* [B,n,n]  pixmap.pixelmap result (0/1) */
    int32_t *sorted_idx_out; /* [B,n*n]  parity aid: flat indices of image*segmap in the canonical descending order
                                (value desc, index desc) for the POSITIVE prefix; the rest is -1.  asymmetry.py:94-95 */
    int32_t *n_positive_out; /* [B]      length of that prefix */
    uint8_t *apermask_out;   /* [B,n,n]  the aperture mask used (apertures.aperpixmap) */
    double *cand_a_out;      /* [B,cand_cap] parity aid: asymmetry of each candidate in order (asymmetry.py:125) */
    int32_t cand_cap;
    uint64_t *phase_cycles_out; /* tuning aid: zeroed DEVICE buffer of PM_PHASE_ROWS x 48 uint64; the launches of this call
                                   accumulate per-group clock64() cycles per phase (0 pixelmap, 1 list, 2 rmax/aperture,
                                   3 sort, 4 gini/m20/candidates, 5 minapix, 6 calcA, 7 r20/r80, 8 smoothness, 9.. parts) */
} pm_outputs_t;
#define PM_PHASE_ROWS 8192

/* Bytes of device scratch pm_analyse_batch needs for (B, n, dtype).  HOST call, no CUDA work. */
size_t pm_workspace_bytes(int64_t B, int n, int dtype);

/* The whole hot path for a batch of cutouts:
 *   pixmap.pixelmap + checkPixelmapEdges  (pixmap.py:130-213, :91-127)       main.py:421,427
 *   pixmap.calcRmax, apertures.aperpixmap (pixmap.py:26-55, apertures.py:42-128) main.py:463-464
 *   asymmetry.minapix                     (asymmetry.py:50-129)               main.py:467
 *   asymmetry.calcA x3                    (asymmetry.py:132-257)              main.py:470-478
 *   casgm.calcR20_R80, concentration, gini, m20, smoothness (casgm.py)        main.py:490-495
 * One fused kernel, one CTA per cutout, the image streamed from HBM once (default); with PM_PATH_SPLIT three kernels per
 * chunk of the batch, split by resource class: front (pixelmap .. candidates), minapix, tail (r20/r80,
 * calcA x3, smoothness), their hand-over state in `workspace`.
 * images  [B,n,n] dtype;  sky, sky_err [B] float64;  filter_size odd in 1..15;
 * ov / out may be NULL;  rows [B];  workspace >= pm_workspace_bytes(B,n,dtype). */
int pm_analyse_batch(const void *images, int dtype, int64_t B, int n, const double *sky,
                     const double *sky_err, int filter_size, uint32_t flags,
                     const pm_overrides_t *ov, const pm_outputs_t *out, pm_row_t *rows,
                     void *workspace, size_t workspace_bytes, void *cuda_stream);

/* apertures.aperpixmap(npix, rad, nsubpix, frac) for B radii at once (apertures.py:42-128):
 * mask_out [B,n,n] uint8.  General float radius / nsubpix / frac (the float64 sub-sample test). */
int pm_aperpixmap_batch(int n, const double *rad, int64_t B, int nsubpix, double frac,
                        uint8_t *mask_out, void *cuda_stream);

/* FITS image payload -> device floats: the input format of the path.  Replaces, for a batch of cutouts,
 * `fits.getdata(...)` + `byteswap().newbyteorder()` + `astype(float64)` of Image.py:132-134, :148.
 * payload: DEVICE pointer to `count` big-endian pixels of type BITPIX (8, 16, 32, -32, -64), headers stripped;
 * out[i] = BZERO + BSCALE * stored[i] (float64 arithmetic), written as out_dtype (PM_F32 / PM_F64).
 * BITPIX 16 with BZERO = 32768, BSCALE = 1 is the unsigned-16 convention of the bundled SDSS frames. */
int pm_fits_decode_batch(const void *payload, int bitpix, double bzero, double bscale, int64_t count,
                         void *out, int out_dtype, void *cuda_stream);

/* The same decode restricted to one npix x npix window per frame: `Cutout2D(image, position, (npix, npix), mode="strict").data`
 * of Image.py:175-186 fused with the decode.  payload: DEVICE pointer to B frames of H x W big-endian pixels (aligned to the
 * pixel size); origin: DEVICE int32 [B, 2] = (first row, first column) of each window, 0-based, the window inside the frame
 * (the caller validates: a window that leaves the frame is the reference's PartialOverlapError, raised on the host);
 * out: DEVICE [B, npix, npix] of out_dtype. */
int pm_fits_decode_window_batch(const void *payload, int bitpix, double bzero, double bscale, int64_t B, int H, int W,
                                const int32_t *origin, int npix, void *out, int out_dtype, void *cuda_stream);

/* Sky-region statistics of skyBackground._calcSkybgr (skyBackground.py:134-164, `_getSkyRegion` :265-303) for a batch:
 * pixels whose centre lies outside the ellipse of semi-axes a, b (= 2 min / 2 max of the fitted FWHMs) and angle theta
 * about (n/2 + 1, n/2 + 1); count, mean, median, population std; when mean > median, sigma_clipped_stats(sigma=3,
 * maxiters=5) and sky = 3 median - 2 mean, else sky = mean.  images: DEVICE [B, n, n] of dtype; ellipse: DEVICE double
 * [B, 4] = (a, b, cos theta, sin theta); out: DEVICE double [B, PM_SKY_COLS] = count, mean, median, std, clip iterations,
 * clipped count, mean, median, std, sky, sky_err, status (1 = fewer than 100 sky pixels, the reference's _SkyError).
 * The Gaussian fit that yields (fwhms, theta) is the caller's (skyBackground.py:306-395). */
#define PM_SKY_COLS 12
int pm_sky_stats_batch(const void *images, int dtype, int64_t B, int n, const double *ellipse, double *out, void *cuda_stream);

/* astropy.stats.sigma_clipped_stats(image, sigma=sigma, maxiters=maxiters, cenfunc=median, stdfunc=std) over every pixel of
 * each cutout (imageutils.py:62; photutils.detect_threshold, skyBackground.py:109 / imageutils.py:67, is mean + nsigma * std
 * of the call with maxiters = None, passed here as maxiters <= 0), plus the unclipped statistics and the maximum
 * (gaussfitter.moments needs the median and the maximum, gaussfitter.py:71-72).
 * out: DEVICE double [B, PM_CLIP_COLS] = count, mean, median, std, clip iterations, clipped count, mean, median, std, max. */
#define PM_CLIP_COLS 10
int pm_clipped_stats_batch(const void *images, int dtype, int64_t B, int n, double sigma, int maxiters, double *out, void *cuda_stream);
/* The same for two stopping rules from ONE sequence of clipping rounds: out_first = the rows of maxiters = maxiters_first (>= 1),
 * out_converged = the rows of maxiters = None (imageutils.py:62 and :67 ask for both of the same image). */
int pm_clipped_stats2_batch(const void *images, int dtype, int64_t B, int n, double sigma, int maxiters_first, double *out_first,
                            double *out_converged, void *cuda_stream);

/* How pm_sky_stats_batch (with_mask = 1) / pm_clipped_stats_batch (with_mask = 0) run a cutout size: CTAs per thread-block
 * cluster sharing one cutout (1, 2, 4 or 8), whether the cutout is held in the cluster's shared memory (read from HBM once) or
 * left in global memory, and the dynamic shared memory per CTA.  Host-side query for tests and benchmarks. */
int pm_stats_plan(int n, int dtype, int with_mask, int *cluster, int *resident, int *smem_bytes);

/* skyBackground._gauss2dfit (skyBackground.py:306-331): the least-squares fit of gaussfitter.twodgaussian (gaussfitter.py:87-148)
 * to each cutout that scipy.optimize.curve_fit performs from the start values of gaussfitter.moments (:40-84) — here a
 * Levenberg-Marquardt iteration with the analytic Jacobian run to the optimum (curve_fit stops at ftol = xtol = 1.5e-8, so the
 * two agree to that tolerance, not bit for bit).  median, vmax: DEVICE double [B] (columns 2 and 9 of pm_clipped_stats_batch),
 * used for the start values unless p0 (DEVICE double [B, 7], nullable) supplies them.
 * out: DEVICE double [B, PM_GAUSS_COLS] = height, amplitude, xo, yo, sigma_x, sigma_y, theta (raw optimum, before the np.abs of
 * skyBackground.py:331), cost, iterations, status (0 converged; 1 iteration limit: curve_fit's RuntimeError, the caller falls
 * back like skyBackground.py:117-118; 2 NaN start values: gaussfitter.py:76-77; 3 singular system). */
#define PM_GAUSS_COLS 10
int pm_gaussfit_batch(const void *images, int dtype, int64_t B, int n, const double *median, const double *vmax, const double *p0,
                      double *out, void *cuda_stream);

/* sersic.fitSersic (sersic.py:53-136), the two pieces that touch pixels.
 * pm_ellipse_sum_batch: photutils EllipticalAperture((xc, yc), a, b, theta).do_photometry(image, method="exact")[0][0]
 * (sersic.py:46-47, :101-121) — the sum of image x (exact area of ellipse ∩ pixel; pixel (row r, column c) is the unit square
 * about (x, y) = (c, r)).  M entries; ell: DEVICE double [M, 5] = xc, yc, a, b, theta; idx: DEVICE int64 [M] cutout of each entry
 * (NULL: entry i is cutout i) so that the root search of sersic.py:103-116 can probe the same cutouts again; out: DEVICE double [M].
 * pm_sersicfit_batch: the least-squares fit of astropy's Sersic2D (amplitude, r_eff, n, x_0, y_0, ellip, theta; x = column) that
 * LevMarLSQFitter performs at sersic.py:134 (scipy leastsq, stopped at 1e-8) — here the same trust-region iteration with the analytic
 * Jacobian run to 1e-10.  p0: DEVICE double [B, 7] start values (sersic.py:98-129); out: DEVICE double [B, PM_GAUSS_COLS] = the 7
 * parameters, cost, iterations, status (as pm_gaussfit_batch). */
int pm_ellipse_sum_batch(const void *images, int dtype, int n, int64_t M, const int64_t *idx, const double *ell, double *out,
                         void *cuda_stream);
int pm_sersicfit_batch(const void *images, int dtype, int64_t B, int n, const double *p0, double *out, void *cuda_stream);

/* photutils.detect_sources(image, threshold, npixels, kernel = 3 x 3 Gaussian of FWHM 3) as skyBackground.py:104-110 and
 * imageutils.py:66-71 call it — smooth, `> threshold[b]`, 8-connected labels, segments of fewer than npixels pixels dropped —
 * and, when clean_out is given, the cleaning of imageutils.maskstarsSEG (imageutils.py:73-89): every kept segment whose bounding
 * box does not contain (n/2 + 1, n/2 + 1) is a "star", and inside a star's bounding box the pixels of kept segments are
 * replaced by mean[b] + std[b] * N(0, 1).  The reference draws from the unseeded np.random.normal; here the normal deviate of
 * (seed, cutout index, pixel index) comes from Philox4x32-10 + Box-Muller, so that runs are reproducible and checkable.
 * labels_out: DEVICE int32 [B, n, n] (nullable; 0 background, 1.. in raster order of the first pixel); clean_out: DEVICE
 * [B, n, n] of dtype (nullable); out: DEVICE double [B, PM_STAR_COLS] = segments kept, segments cleaned, pixels replaced, status;
 * workspace >= pm_maskstars_workspace_bytes(B, n), 16-byte aligned. */
#define PM_STAR_COLS 4
size_t pm_maskstars_workspace_bytes(int64_t B, int n);
int pm_maskstars_batch(const void *images, int dtype, int64_t B, int n, const double *threshold, int npixels, const double *mean,
                       const double *std_, uint64_t seed, int32_t *labels_out, void *clean_out, double *out, void *workspace,
                       size_t workspace_bytes, void *cuda_stream);

/* Library / device introspection (HOST). */
int pm_abi_version(void);
int pm_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin);
const char *pm_last_cuda_error(void);
#define PM_MAX_GRID 592
/* number of kernel launches issued by this library since load (bench.py reports it) */
uint64_t pm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PAWLIK_B200_H */
