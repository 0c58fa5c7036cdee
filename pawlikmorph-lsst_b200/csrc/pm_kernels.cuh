// pm_kernels.cuh — the per-cutout morphology kernel: one CTA per galaxy, persistent CTAs
// pulling galaxy indices from a work counter.  Phases follow the dataflow of the reference's
// main.py:421-495 (pixelmap -> rmax/aperture -> minapix -> calcA x3 -> r20/r80/C -> gini/m20 -> S).
//
// Data layout in shared memory (per CTA): three n x n bit planes (segmap, its 9x9 dilation,
// the aperture disc), the (n, filterSize) index tables, and an arena that holds the per-cutout
// working set (raster list of segmap pixels, sort buffers, the masked window of the image over
// the segmap bounding box, the folded radial table).  Anything that does not fit the arena
// spills to a per-CTA slice of the caller's workspace in HBM (L2-resident in practice).
// The image itself stays in global memory: it is streamed once from HBM by the pixelmap phase
// and later phases re-touch only small parts of it through L1/L2.
#pragma once
#include "../../include/pawlik_b200.h"
#include "pm_block.cuh"

namespace PM_NS {

struct Params {
    const void* images;
    long long B;
    int n, fs;
    unsigned flags;
    const double* sky;
    const double* sky_err;
    pm_overrides_t ov;
    pm_outputs_t out;
    pm_row_t* rows;
    unsigned char* scratch;       // per-CTA global scratch, scratch_per_cta bytes each
    unsigned long long scratch_per_cta;
    unsigned long long* counter;  // work counters (zeroed by the host before the launches): [0] monolith, [1] front,
                                  // [2] minapix, [3] tail, [4] number of cutouts handed to the monolith pass
    unsigned smem_bytes;          // dynamic shared memory given to each group (CTA, or warp of the tail kernel)
    unsigned long long* phase_cycles;   // optional [PM_PHASE_ROWS, 48] per-group cycle counters (tuning aid), or null
    // split path: the cutouts [b0, b0 + nb) of the batch are in flight, their hand-over state lives at state + i * stride
    long long b0, nb;
    unsigned char* state;
    unsigned long long state_stride;
    int* big_list;                // cutouts of the chunk that take the monolith pass (local indices), counter[4] of them
    int list_cap, cand_cap;       // hand-over capacities (entries), see state_stride()
};

struct Geo {
    int n, m, fs, h, nw, mw, half, N;
};
PM_HOSTDEV Geo make_geo(int n, int fs) {
    Geo g;
    g.n = n; g.fs = fs; g.h = fs / 2;
    g.m = n / fs;                    // int(n / filterSize), pixmap.py:171
    g.nw = (n + 31) / 32;
    g.mw = (g.m + 31) / 32;
    g.half = n / 2;
    g.N = n * n;
    return g;
}
struct SmemMap {
    unsigned sh, seg, dil, aper, rowoff, i0, tw, ri, rstart, arena, arena_bytes;
};
constexpr unsigned kNoSmem = 0xffffffffu;
enum { SM_SEG = 1, SM_DIL = 2, SM_APER = 4, SM_ROWOFF = 8, SM_TABLES = 16, SM_ALL = 31 };
PM_HOSTDEV unsigned align16(unsigned v) { return (v + 15u) & ~15u; }
// Fixed part of a group's shared memory.  `mask`: the pieces this kernel keeps there; with `optional` a piece is
// placed only while it leaves >= 1 KB of arena (the body then takes it from the arena / global scratch instead).
PM_HOSTDEV SmemMap make_smem_map(const Geo& g, unsigned total, unsigned mask = SM_ALL, bool optional = false) {
    SmemMap s;
    unsigned o = 0;
    s.sh = o; o = align16(o + (unsigned)sizeof(Sh));
    unsigned plane = (unsigned)(g.n * g.nw * 4);
    unsigned zplane = (unsigned)(g.m * g.mw * 4);
    if (zplane > plane) plane = zplane;   // cannot happen (m <= n) but keep the aliasing safe
    plane = align16(plane);               // every piece (and the arena behind them) starts 16-byte aligned
    auto fits = [&](unsigned bytes) { return !optional || o + bytes + 1024u <= total; };
    s.seg = s.dil = s.aper = s.rowoff = s.tw = s.i0 = s.ri = s.rstart = kNoSmem;
    const unsigned rowoff_bytes = align16((unsigned)(g.n + 1) * 4);
    if ((mask & SM_ROWOFF) && fits(rowoff_bytes)) { s.rowoff = o; o += rowoff_bytes; }
    if ((mask & SM_SEG) && fits(plane)) { s.seg = o; o += plane; }
    if ((mask & SM_APER) && fits(plane)) { s.aper = o; o += plane; }
    if ((mask & SM_DIL) && fits(plane)) { s.dil = o; o += plane; }
    if (mask & SM_TABLES) {
        s.tw = o; o = align16(o + (unsigned)g.m * 8);
        s.i0 = o; o = align16(o + (unsigned)g.m * 2);
        s.ri = o; o = align16(o + (unsigned)g.n * 2);
        s.rstart = o; o = align16(o + (unsigned)(g.m + 1) * 2);
    }
    s.arena = o;
    s.arena_bytes = total > o ? total - o : 0;
    return s;
}
PM_HOSTDEV unsigned smem_fixed_bytes(int n, int fs, unsigned mask = SM_ALL) {
    Geo g = make_geo(n, fs);
    return make_smem_map(g, 0, mask).arena;
}
// per-CTA global scratch: exact-filter planes (2 x 8 n^2), or list/sort/window spill, or the folded table
PM_HOSTDEV unsigned long long scratch_bytes_per_cta(int n) {
    return 64ull * (unsigned long long)n * (unsigned long long)n + 65536ull;
}
// ---- split path: hand-over state of one cutout between the front, minapix and tail kernels (HBM; one slot per cutout
// of the chunk in flight).  256-byte header, the segmap and aperture bit planes, then list_cap entries holding the
// raster list of map pixels followed by the centroid candidates.  Cutouts whose lists do not fit (or whose radius /
// candidate count would break the scratch bounds of the later kernels) take the monolith pass instead.
struct GalState {
    pm_row_t row;                  // the result row as far as it is known
    double sky, thr, rmax, abs_sum, gini, m20;
    int status, done;              // done: 1 = nothing left for the later kernels, 2 = handed to the monolith pass
    int n_g, r0, r1, c0, c1;       // map pixel count, bounding box (inclusive)
    int n_cand, cx, cy;
};
static_assert(sizeof(GalState) <= 256, "hand-over header");
PM_HOSTDEV int state_list_cap(int n) { return n * n / 2 + 2048; }
PM_HOSTDEV int state_cand_cap() { return 1024; }
PM_HOSTDEV unsigned long long state_stride(int n) {
    const unsigned long long plane = (unsigned long long)n * ((n + 31) / 32) * 4;
    const unsigned long long b = 256 + 2 * plane + 4ull * (unsigned long long)state_list_cap(n);
    return (b + 255) & ~255ull;
}
// per-group global scratch of the minapix kernel (window of the image over the bounding box when it does not fit
// shared memory + the candidate bookkeeping) and of the tail kernel (folded radial table <= 8 (n/2 + 4)^2 bytes x 2,
// bit planes, band buffers)
PM_HOSTDEV unsigned long long scratch_bytes_minapix(int n) {
    return ((8ull * (unsigned long long)n * (unsigned long long)n + 255ull) & ~255ull) + 131072ull;
}
PM_HOSTDEV unsigned long long scratch_bytes_tail(int n) {
    return ((6ull * (unsigned long long)n * (unsigned long long)n + 255ull) & ~255ull) + 131072ull;
}

PM_DEV int refl(int i, int L) {  // scipy "reflect" (d c b a | a b c d | d c b a), any distance
    if (i >= 0 && i < L) return i;
    int p = 2 * L;
    i %= p;
    if (i < 0) i += p;
    return i < L ? i : p - 1 - i;
}

// order-preserving integer keys -------------------------------------------------------------------
PM_DEV uint32_t key_of(float x, double) {
    if (x == 0.0f) x = 0.0f;  // -0.0 -> +0.0 (numpy compares them equal; asymmetry.py:94-95)
    uint32_t b = pm::f2bits(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
PM_DEV uint64_t key_of(double x, double sky) {
    double v = x - sky;
    if (v == 0.0) v = 0.0;
    uint64_t b = (uint64_t)pm::d2bits(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
PM_DEV double val_of_key(uint32_t k, double sky) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return (double)pm::bits2f(b) - sky;
}
PM_DEV double val_of_key(uint64_t k, double) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return pm::bits2d((long long)b);
}
template <typename T> struct KeyT;
template <> struct KeyT<float> { typedef uint32_t type; };
template <> struct KeyT<double> { typedef uint64_t type; };
PM_DEV float nan_of(float) { return pm::bits2f(0x7fc00000u); }
PM_DEV double nan_of(double) { return pm::bits2d(0x7ff8000000000000ll); }

constexpr int kPhaseSlots = 48;      // per CTA; 0..15 phases of analyse_one, 16.. sub-phases and counters (bench.py --phases)
struct PhaseClock {
    unsigned long long* slot;
    long long last;
    PM_DEV void start(unsigned long long* s) { slot = s; last = slot ? pm::clock() : 0; }
    PM_DEV void lap(int phase) {
        if (!slot) return;
        const long long now = pm::clock();
        if (pm::tid() == 0) slot[phase] += (unsigned long long)(now - last);
        last = now;
    }
    PM_DEV void count(int phase, int v) { if (slot && pm::tid() == 0) slot[phase] += (unsigned long long)v; }
};
// Per-cutout context (registers / local; identical in every thread) ---------------------------------
template <typename T> struct Ctx {
    Geo g;
    Sh* sh;
    unsigned* seg; unsigned* dil; unsigned* aper; unsigned* rowoff;
    double* tw; uint16_t* i0; uint16_t* ri; uint16_t* rstart;
    Arena ar;
    const T* img;        // this cutout, [n, n]
    double sky, thr;
    unsigned flags;
    // segmap statistics
    int n_g, r0, r1, c0, c1;   // pixel count, bounding box (inclusive)
    unsigned* posR;      // raster list of segmap pixels, (row << 16) | col
    const unsigned* plist;   // the list minapix iterates over: posR, or the brightness-sorted list (pruned path)
    T* mw;               // masked window over the bounding box (NaN outside the segmap)
    PhaseClock pc;       // tuning aid
    const T* crop;       // copy of the central part of the image in shared memory (late phases), or null
    int crop_r0, crop_c0, crop_h, crop_w;
    unsigned mbar_phase[2];   // parity of the two staging mbarriers (uniform across the CTA)
    int bw;              // window width
};

// image value at (r, col): from the shared-memory crop when it covers the pixel, else from global memory
template <typename T> PM_DEV T ld_raw(const Ctx<T>& c, int r, int col) {
    const unsigned rr = (unsigned)(r - c.crop_r0), cc = (unsigned)(col - c.crop_c0);
    if (rr < (unsigned)c.crop_h && cc < (unsigned)c.crop_w) return c.crop[rr * (unsigned)c.crop_w + cc];
    return pm::ldg(c.img + (size_t)r * c.g.n + col);
}
template <typename T> PM_DEV double ld(const Ctx<T>& c, int r, int col) { return (double)ld_raw(c, r, col); }
PM_DEV bool bit_at(const unsigned* plane, int nw, int r, int c) {
    return (plane[r * nw + (c >> 5)] >> (c & 31)) & 1u;
}
// number of set bits of row r of a bit plane, read by ONE thread: the words are visited in an order rotated by r / 4 so
// that the 32 rows a warp counts (lane stride nw words) fall on different banks (nw = 8: conflict-free instead of 8-way)
PM_DEV unsigned row_popcount(const unsigned* plane, int nw, int r) {
    unsigned cnt = 0;
    int w = (r >> 2) % nw;
    for (int k = 0; k < nw; k++) {
        cnt += (unsigned)pm::popc(plane[r * nw + w]);
        if (++w == nw) w = 0;
    }
    return cnt;
}
template <typename T> PM_DEV T mw_raw(const Ctx<T>& c, int r, int col) {
    if (r < c.r0 || r > c.r1 || col < c.c0 || col > c.c1) return nan_of(T());
    return c.mw[(r - c.r0) * c.bw + (col - c.c0)];
}
// masked, sky-subtracted value M[r, col] = (img - sky) * segmap (asymmetry.py:82-86, casgm.py:151)
template <typename T> PM_DEV double mval(const Ctx<T>& c, int r, int col) {
    T w = mw_raw(c, r, col);
    return (w != w) ? 0.0 : (double)w - c.sky;
}

// ====================================================================================================
// Phase 0 — tables that depend only on (n, filterSize): built once per CTA
// ====================================================================================================
template <typename T> PM_DEV void build_tables(Ctx<T>& c) {
    const Geo& g = c.g;
    // skimage.transform.resize (pixmap.py:171-173) == scipy.ndimage.zoom(order=1, grid_mode=True):
    // source coordinate of output i:  cc = (i + 0.5) * (n / m) - 0.5 ; taps floor(cc), floor(cc) + 1
    const double zoom = pm::ddiv((double)g.n, (double)g.m);
    for (int i = pm::tid(); i < g.m; i += kThreads) {
        double cc = pm::dsub(pm::dmul((double)i + 0.5, zoom), 0.5);
        double f = floor(cc);
        int a = (int)f;
        if (a < 0) a = 0;
        if (a > g.n - 1) a = g.n - 1;
        c.i0[i] = (uint16_t)a;
        c.tw[i] = pm::dsub(cc, f);
    }
    // skimage.transform.resize(order=0, mode="edge") (pixmap.py:211-212): nearest source index
    const double zin = pm::ddiv((double)g.m, (double)g.n);
    for (int i = pm::tid(); i < g.n; i += kThreads) {
        double cc = pm::dsub(pm::dmul((double)i + 0.5, zin), 0.5);
        int a = (int)floor(pm::dadd(cc, 0.5));
        if (a < 0) a = 0;
        if (a > g.m - 1) a = g.m - 1;
        c.ri[i] = (uint16_t)a;
    }
    pm::sync();
    // rstart[k] = first image row whose source row is >= k   (ri is non-decreasing)
    for (int k = pm::tid(); k <= g.m; k += kThreads) {
        int lo = 0, hi = g.n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if ((int)c.ri[mid] >= k) hi = mid; else lo = mid + 1;
        }
        c.rstart[k] = (uint16_t)lo;
    }
    pm::sync();
}

// ====================================================================================================
// Phase 1 — pixmap.pixelmap (pixmap.py:130-213)
// ====================================================================================================
// Bilinear combination exactly as scipy's NI_ZoomShift accumulates it (tap order, product order).
PM_DEV double zoom_taps(double f00, double f01, double f10, double f11, double ta, double tb) {
    const double wa0 = pm::dsub(1.0, ta), wb0 = pm::dsub(1.0, tb);
    double s = 0.0;
    s = pm::dadd(s, pm::dmul(pm::dmul(f00, wa0), wb0));
    s = pm::dadd(s, pm::dmul(pm::dmul(f01, wa0), tb));
    s = pm::dadd(s, pm::dmul(pm::dmul(f10, ta), wb0));
    s = pm::dadd(s, pm::dmul(pm::dmul(f11, ta), tb));
    return s;
}

// (fs+1) x (fs+1) patch of staged rows -> the four fs x fs window sums around (a0, b0).  FS > 0: compile-time
// filter size (fully unrolled); kInterior: no reflection needed.
template <typename T, int FS, bool kInterior>
PM_DEV float patch_sums(const T* xb, int rlo, int n, int fs_rt, int a0, int b0, double& faa, double& fab, double& fba, double& fbb) {
    const int fs = FS ? FS : fs_rt, h = fs / 2;
    const bool ra = kInterior || a0 + 1 <= n - 1, rb = kInterior || b0 + 1 <= n - 1;   // second tap = first + 1 (else clamped)
    faa = fab = fba = fbb = 0.0;
    float amax = 0.0f;
#pragma unroll
    for (int k = 0; k <= fs; k++) {
        const int rr = kInterior ? a0 - h + k : refl(a0 - h + k, n);
        const T* row = xb + (rr - rlo) * n;
        double sa = 0.0, sb = 0.0;       // sums over the b0 / b1 column windows
#pragma unroll
        for (int q = 0; q <= fs; q++) {
            const T xr = row[kInterior ? b0 - h + q : refl(b0 - h + q, n)];
            amax = fmaxf(amax, fabsf((float)xr));     // NaN-propagating max is not needed: NaN shows up in Z
            const double x = (double)xr;
            if (q < fs) sa += x;
            if (rb ? (q >= 1) : (q < fs)) sb += x;
        }
        if (k < fs) { faa += sa; fab += sb; }
        if (ra ? (k >= 1) : (k < fs)) { fba += sa; fbb += sb; }
    }
    return amax;
}

// fast path: the image rows a band of shrunk rows needs are staged in shared memory — by a TMA bulk copy
// (cp.async.bulk, completion on an mbarrier, double-buffered so that the next band lands while the current
// one is consumed) when the band is 16-byte aligned, by coalesced independent loads otherwise.  This is the
// only pass that streams the image from HBM.  Each thread then forms the (fs+1) x (fs+1) patch sums of one
// shrunk pixel from the staged rows.  Reports min |Z - thr| and max |x| for the band test.
// xb0 / xb1: the two staging buffers; R: shrunk rows per band.
template <typename T, int FS>
PM_DEV void pixelmap_fast(Ctx<T>& c, unsigned* zb, int* centre_faint, double* minabs, double* maxabs, T* xb0, T* xb1, int R) {
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    const int fs = g.fs, h = g.h, n = g.n, m = g.m;
    const double inv = 1.0 / ((double)fs * (double)fs);
    double mn = 1e300;
    float mx = 0.0f;
    int faint = 0, bad = 0;
    const int cm = m / 2;
    for (int i = pm::tid(); i < m * g.mw; i += kThreads) zb[i] = 0;
    pm::sync();      // the bit plane is filled with atomics below
    const int nbands = (m + R - 1) / R;
    auto band_rows = [&](int k, int* rlo, int* cnt) {
        const int band = k * R, rows = min(R, m - band);
        *rlo = max(0, (int)c.i0[band] - h);
        const int rhi = min(n - 1, (int)c.i0[band + rows - 1] + h + 1);
        *cnt = (rhi - *rlo + 1) * n;
    };
    // TMA needs 16-byte aligned source / size and BOTH staging buffers in shared memory (large n with a large
    // filter can push them to the global scratch: then the cooperative-load path is used)
    const bool staging_in_smem = xb0 != xb1 && c.ar.in_smem(xb0) && c.ar.in_smem(xb1);
    auto tma_ok = [&](int rlo, int cnt) {
        return staging_in_smem && (((uintptr_t)(c.img + (size_t)rlo * n)) & 15) == 0 && (((size_t)cnt * sizeof(T)) & 15) == 0;
    };
    auto issue = [&](int k) {       // thread 0: start the bulk copy of band k if it qualifies
        int rlo, cnt;
        band_rows(k, &rlo, &cnt);
        if (tma_ok(rlo, cnt) && pm::tid() == 0)
            pm::tma_load_1d((k & 1) ? xb1 : xb0, c.img + (size_t)rlo * n, (unsigned)((size_t)cnt * sizeof(T)), &sh.mbar[k & 1]);
    };
    issue(0);
    for (int k = 0; k < nbands; k++) {
        const int band = k * R, rows = min(R, m - band);
        int rlo, cnt;
        band_rows(k, &rlo, &cnt);
        T* xb = (k & 1) ? xb1 : xb0;
        if (k + 1 < nbands) issue(k + 1);
        if (tma_ok(rlo, cnt)) {
            pm::mbar_wait(&sh.mbar[k & 1], c.mbar_phase[k & 1]);
            c.mbar_phase[k & 1] ^= 1u;
        } else {
            const T* src = c.img + (size_t)rlo * n;
            int i = pm::tid();
            for (; i + 3 * kThreads < cnt; i += 4 * kThreads) {     // four independent loads in flight per thread
                const T x0 = pm::ldg(src + i), x1 = pm::ldg(src + i + kThreads), x2 = pm::ldg(src + i + 2 * kThreads),
                        x3 = pm::ldg(src + i + 3 * kThreads);
                xb[i] = x0; xb[i + kThreads] = x1; xb[i + 2 * kThreads] = x2; xb[i + 3 * kThreads] = x3;
                const float sm = fabsf((float)x0) + fabsf((float)x1) + fabsf((float)x2) + fabsf((float)x3);
                if (!(sm <= 3.0e38f)) bad = 1;
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf((float)x0), fabsf((float)x1)), fmaxf(fabsf((float)x2), fabsf((float)x3))));
            }
            for (; i < cnt; i += kThreads) {
                const T x = pm::ldg(src + i);
                xb[i] = x;
                const float ax = fabsf((float)x);
                if (!(ax <= 3.0e38f)) bad = 1;
                mx = fmaxf(mx, ax);
            }
            pm::sync();
        }
        for (int item = pm::tid(); item < rows * m; item += kThreads) {
            const int rb = item / m, j = item - rb * m;
            const int ii = band + rb;
            const int a0 = c.i0[ii], b0 = c.i0[j];
            double faa, fab, fba, fbb;
            float am;
            if (a0 - h >= 0 && a0 + h + 1 <= n - 1 && b0 - h >= 0 && b0 + h + 1 <= n - 1)
                am = patch_sums<T, FS, true>(xb, rlo, n, fs, a0, b0, faa, fab, fba, fbb);
            else
                am = patch_sums<T, FS, false>(xb, rlo, n, fs, a0, b0, faa, fab, fba, fbb);
            if (!(am <= 3.0e38f)) bad = 1;           // inf (or beyond float range): let the exact path decide
            mx = fmaxf(mx, am);
            const double z = zoom_taps(faa * inv, fab * inv, fba * inv, fbb * inv, c.tw[ii], c.tw[j]);
            const double d = z - c.thr;
            if (!(fabs(d) >= 0.0)) bad = 1;          // NaN anywhere in the patch
            mn = fmin(mn, fabs(d));
            if (z > c.thr) pm::atomic_or(&zb[ii * g.mw + (j >> 5)], 1u << (j & 31));
            if (ii == cm && j == cm) faint = (z < c.thr) ? 1 : 0;
        }
        pm::sync();      // everybody is done with this buffer before band k + 2 is copied into it
    }
    if (bad) mn = 0.0;
    *minabs = block_min(sh, mn);
    *maxabs = block_max(sh, (double)mx * 1.0000002);   // float rounding of |x| never under-estimates the band
    *centre_faint = pm::sync_or(faint);
}

// exact path: scipy.ndimage.uniform_filter's running-sum recurrence, axis 0 then axis 1,
// bit-for-bit.  G and F are n x n float64 planes in the per-CTA global scratch.
template <typename T> PM_DEV void pixelmap_exact(Ctx<T>& c, unsigned* zb, int* centre_faint, double* G, double* F) {
    const Geo& g = c.g;
    const int fs = g.fs, h = g.h, n = g.n, m = g.m;
    const double dfs = (double)fs;
    for (int col = pm::tid(); col < n; col += kThreads) {       // pass 1: along axis 0
        double s = 0.0;
        for (int l = 0; l < fs; l++) s = pm::dadd(s, ld(c, refl(l - h, n), col));
        G[col] = pm::ddiv(s, dfs);
        for (int k = 1; k < n; k++) {
            const double pn = ld(c, refl(k + fs - 1 - h, n), col), po = ld(c, refl(k - 1 - h, n), col);
            s = pm::dadd(s, pm::dsub(pn, po));
            G[(size_t)k * n + col] = pm::ddiv(s, dfs);
        }
    }
    pm::sync();
    for (int row = pm::tid(); row < n; row += kThreads) {       // pass 2: along axis 1
        const double* gr = G + (size_t)row * n;
        double* fr = F + (size_t)row * n;
        double s = 0.0;
        for (int l = 0; l < fs; l++) s = pm::dadd(s, gr[refl(l - h, n)]);
        fr[0] = pm::ddiv(s, dfs);
        for (int k = 1; k < n; k++) {
            s = pm::dadd(s, pm::dsub(gr[refl(k + fs - 1 - h, n)], gr[refl(k - 1 - h, n)]));
            fr[k] = pm::ddiv(s, dfs);
        }
    }
    pm::sync();
    int faint = 0;
    const int cm = m / 2;
    for (int task = pm::warp(); task < m * g.mw; task += kWarps) {
        const int i = task / g.mw, w = task - i * g.mw;
        const int j = w * 32 + pm::lane();
        bool above = false;
        if (j < m) {
            const int a0 = c.i0[i], b0 = c.i0[j];
            const int a1 = min(a0 + 1, n - 1), b1 = min(b0 + 1, n - 1);
            const double z = zoom_taps(F[(size_t)a0 * n + b0], F[(size_t)a0 * n + b1], F[(size_t)a1 * n + b0],
                                       F[(size_t)a1 * n + b1], c.tw[i], c.tw[j]);
            above = z > c.thr;
            if (i == cm && j == cm) faint = (z < c.thr) ? 1 : 0;
        }
        unsigned bits = pm::ballot(above);
        if (pm::lane() == 0) zb[i * g.mw + w] = bits;
    }
    *centre_faint = pm::sync_or(faint);
}

// 8-connected grow from the centre of the m x m grid (pixmap.py:182-208), bit-parallel
PM_DEV unsigned fill_runs(unsigned A, unsigned S) {
    unsigned up = (((A + S) ^ A) & A) | S;
    unsigned rA = pm::brev(A), rS = pm::brev(S);
    unsigned dn = pm::brev((((rA + rS) ^ rA) & rA) | rS);
    return up | dn;
}
PM_DEV unsigned spread_row(const unsigned* gb, int mw, int i, int w) {
    const unsigned* row = gb + i * mw;
    unsigned v = row[w];
    unsigned s = v | (v << 1) | (v >> 1);
    if (w > 0) s |= row[w - 1] >> 31;
    if (w + 1 < mw) s |= row[w + 1] << 31;
    return s;
}
PM_DEV void region_grow(const Geo& g, unsigned* zb, unsigned* gb) {
    const int m = g.m, mw = g.mw, cm = m / 2;
    for (int i = pm::tid(); i < m * mw; i += kThreads) gb[i] = 0;
    pm::sync();
    if (pm::tid() == 0) {
        zb[cm * mw + (cm >> 5)] |= 1u << (cm & 31);   // the centre is always set (pixmap.py:184)
        gb[cm * mw + (cm >> 5)] = 1u << (cm & 31);
    }
    pm::sync();
    for (;;) {
        int changed = 0;
        for (int task = pm::tid(); task < m * mw; task += kThreads) {
            const int i = task / mw, w = task - i * mw;
            const unsigned A = zb[task];
            if (A == 0) continue;
            const unsigned cur = gb[task];
            unsigned s = spread_row(gb, mw, i, w);
            if (i > 0) s |= spread_row(gb, mw, i - 1, w);
            if (i + 1 < m) s |= spread_row(gb, mw, i + 1, w);
            s = (s & A) | cur;
            if (s == 0) continue;
            const unsigned nw_ = fill_runs(A, s);
            if (nw_ != cur) { gb[task] = nw_; changed = 1; }
        }
        if (!pm::sync_or(changed)) break;
    }
}
// nearest-neighbour upsample of the grown mask to n x n bits (pixmap.py:211-212)
template <typename T> PM_DEV void upsample_mask(Ctx<T>& c, const unsigned* gb) {
    const Geo& g = c.g;
    for (int task = pm::tid(); task < g.m * g.nw; task += kThreads) {
        const int mi = task / g.nw, w = task - mi * g.nw;
        const unsigned* row = gb + mi * g.mw;
        unsigned bits = 0;
        const int cend = min(32, g.n - w * 32);
        for (int b = 0; b < cend; b++) {
            const int mj = c.ri[w * 32 + b];
            bits |= ((row[mj >> 5] >> (mj & 31)) & 1u) << b;
        }
        for (int r = c.rstart[mi]; r < (int)c.rstart[mi + 1]; r++) c.seg[r * g.nw + w] = bits;
    }
    pm::sync();
}
// uint8 [n, n] plane <-> bit plane
PM_DEV void load_plane_u8(const Geo& g, const uint8_t* src, unsigned* plane) {
    for (int task = pm::warp(); task < g.n * g.nw; task += kWarps) {
        const int r = task / g.nw, w = task - r * g.nw;
        const int col = w * 32 + pm::lane();
        int v = 0;
        if (col < g.n) v = pm::ldg(src + (size_t)r * g.n + col) != 0;
        unsigned bits = pm::ballot(v);
        if (pm::lane() == 0) plane[task] = bits;
    }
    pm::sync();
}
PM_DEV void store_plane_u8(const Geo& g, const unsigned* plane, uint8_t* dst) {
    if ((g.n & 3) == 0 && ((uintptr_t)dst & 3) == 0) {
        // aligned rows: one (row, 32-pixel word) per thread-step; 4 bits -> 4 bytes with one multiply
        // ((b & 0xF) * 0x00204081) & 0x01010101 puts bit k into byte k
        for (int task = pm::tid(); task < g.n * g.nw; task += kThreads) {
            const int r = task / g.nw, w = task - r * g.nw;
            const unsigned bits = plane[task];
            uint32_t* d32 = (uint32_t*)(dst + (size_t)r * g.n + w * 32);
            const int nq = min(8, (g.n - w * 32) >> 2);
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < nq) d32[q] = (((bits >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u;
        }
        return;
    }
    // general case: 4 pixels per thread-step; word stores when the destination is 4-byte aligned (N < 2^20)
    const int total = g.N;
    const int head = min(total, (int)((4 - ((uintptr_t)dst & 3)) & 3));
    for (int i = pm::tid(); i < head; i += kThreads) {
        const int r = i / g.n, col = i - r * g.n;
        dst[i] = bit_at(plane, g.nw, r, col) ? 1 : 0;
    }
    const int nwords = (total - head) / 4;
    uint32_t* d32 = (uint32_t*)(dst + head);
    for (int k = pm::tid(); k < nwords; k += kThreads) {
        const int f = head + 4 * k;
        int r = f / g.n, col = f - r * g.n;
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            v |= (bit_at(plane, g.nw, r, col) ? 1u : 0u) << (8 * b);
            if (++col == g.n) { col = 0; r++; }
        }
        d32[k] = v;
    }
    for (int f = head + 4 * nwords + pm::tid(); f < total; f += kThreads) {
        const int r = f / g.n, col = f - r * g.n;
        dst[f] = bit_at(plane, g.nw, r, col) ? 1 : 0;
    }
}

// returns status (PM_ST_OK or PM_ST_FAINT_CENTRE); fills c.seg
template <typename T> PM_DEV int phase_pixelmap(Ctx<T>& c) {
    const Geo& g = c.g;
    unsigned* zb = c.dil;   // the two m x m planes alias the dilation / aperture planes (not yet in use)
    unsigned* gb = c.aper;
    int faint = 0;
    bool exact = (c.flags & PM_FORCE_EXACT_FILTER) != 0;
    if (!exact) {
        double mn, mx;
        Arena::Mark mk = c.ar.mark();
        int R = max(1, 3 * kThreads / g.m);                        // about three shrunk pixels per thread per band
        for (;;) {                                                  // two staging buffers of <= 32 KB each
            const int nr = (int)(((long long)(R - 1) * g.n) / g.m) + g.fs + 3;
            if (R == 1 || (size_t)nr * g.n * sizeof(T) <= 32768) break;
            R--;
        }
        const int nrows = min(g.n, (int)(((long long)(R - 1) * g.n) / g.m) + g.fs + 3);
        const size_t xbytes = (size_t)nrows * g.n * sizeof(T);
        T* xb0 = (T*)c.ar.alloc(xbytes);
        T* xb1 = xb0;                                               // one buffer only when two do not fit shared memory
        if (2 * xbytes + 16 <= (size_t)(c.ar.s_cap - mk.s)) xb1 = (T*)c.ar.alloc(xbytes);
        if (g.fs == 3) pixelmap_fast<T, 3>(c, zb, &faint, &mn, &mx, xb0, xb1, R);
        else pixelmap_fast<T, 0>(c, zb, &faint, &mn, &mx, xb0, xb1, R);
        c.ar.release(mk);
        // |Z_fast - Z_scipy| <= (2 n + fs^2 + 8) * max|x| * 2^-52 (running-sum error growth); x4 margin
        const double band = 4.0 * (2.0 * g.n + (double)(g.fs * g.fs) + 8.0) * mx * 2.220446049250313e-16;
        if (!(mn > band)) exact = true;
    }
    if (exact) {
        pm::sync();
        double* G = (double*)c.ar.g_base;
        double* F = G + (size_t)g.N;
        pixelmap_exact(c, zb, &faint, G, F);
    }
    c.pc.lap(0);
    if (faint) return PM_ST_FAINT_CENTRE;
    region_grow(g, zb, gb);
    upsample_mask(c, gb);
    c.pc.lap(0);
    return PM_ST_OK;
}

// ====================================================================================================
// Phase 2 — segmap statistics, raster list, keys; pixmap.checkPixelmapEdges, pixmap.calcRmax
// ====================================================================================================
struct SegStats {
    double sum, m10, m01;   // Σ v, Σ row·v, Σ col·v  over segmap pixels (v = img - sky)
    double abs_sum;         // Σ |v|
    double vmax;            // max v over segmap pixels
    int argmax;             // flat index of the first pixel attaining it
    int edge;
};

template <typename T>
PM_DEV void phase_list(Ctx<T>& c, typename KeyT<T>::type* keys, SegStats* st) {
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    // c.rowoff holds the per-row counts (the caller needed their sum for the allocations) -> exclusive offsets
    block_scan_excl_array(sh, c.rowoff, g.n);
    c.pc.lap(37);
    double s = 0.0, sa = 0.0, m10 = 0.0, m01 = 0.0, vmax = -1.7976931348623157e308;
    int amax = 0x7fffffff;
    const int wlo = c.c0 >> 5, whi = c.c1 >> 5;          // words of the bounding box
    for (int r = pm::warp(); r < g.n; r += kWarps) {
        unsigned off = c.rowoff[r];
        if (c.rowoff[r + 1] == off) continue;
        // four words of the row per step: their image loads are issued together (the row is otherwise a chain of
        // dependent L2 round trips)
        for (int w0 = wlo; w0 <= whi; w0 += 4) {
            unsigned bits[4];
            T x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                bits[u] = w0 + u <= whi ? c.seg[r * g.nw + w0 + u] : 0u;
                const bool mine = (bits[u] >> pm::lane()) & 1u;
                x[u] = mine ? pm::ldg(c.img + (size_t)r * g.n + (w0 + u) * 32 + pm::lane()) : T(0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if ((bits[u] >> pm::lane()) & 1u) {
                    const int col = (w0 + u) * 32 + pm::lane();
                    const unsigned o = off + (unsigned)pm::popc(bits[u] & lt_mask());
                    c.posR[o] = ((unsigned)r << 16) | (unsigned)col;
                    keys[o] = key_of(x[u], c.sky);
                    const double v = (double)x[u] - c.sky;
                    s += v; sa += fabs(v); m10 += (double)r * v; m01 += (double)col * v;
                    const int flat = r * g.n + col;
                    if (v > vmax || (v == vmax && flat < amax)) { vmax = v; amax = flat; }
                }
                off += (unsigned)pm::popc(bits[u]);
            }
        }
    }
    c.pc.lap(38);
    double red[4] = {s, m10, m01, sa};
    block_sum<4>(sh, red);
    st->sum = red[0]; st->m10 = red[1]; st->m01 = red[2]; st->abs_sum = red[3];
    const double bm = block_max(sh, vmax);
    st->vmax = bm;
    st->argmax = block_min_i(sh, (vmax == bm) ? amax : 0x7fffffff);
}

// bounding box + edge flag from the bit plane
template <typename T> PM_DEV void phase_bbox(Ctx<T>& c, int* edge) {
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    for (int w = pm::tid(); w < g.nw; w += kThreads) sh.colmask[w] = 0;
    pm::sync();
    int rmin = 0x7fffffff, rmaxv = -1;
    for (int task = pm::tid(); task < g.n * g.nw; task += kThreads) {
        const unsigned bits = c.seg[task];
        if (bits) {
            const int r = task / g.nw, w = task - r * g.nw;
            pm::atomic_or(&sh.colmask[w], bits);
            rmin = min(rmin, r); rmaxv = max(rmaxv, r);
        }
    }
    c.r0 = block_min_i(sh, rmin);
    c.r1 = block_max_i(sh, rmaxv);   // the syncs inside also publish colmask
    int c0 = 0x7fffffff, c1 = -1;
    for (int w = 0; w < g.nw; w++) {
        const unsigned bits = sh.colmask[w];
        if (bits) {
            if (c0 == 0x7fffffff) c0 = w * 32 + pm::ffs(bits) - 1;
            c1 = w * 32 + 31 - pm::clz(bits);
        }
    }
    c.c0 = c0; c.c1 = c1;
    *edge = (c.r1 >= 0) && (c.r0 == 0 || c.r1 == g.n - 1 || c0 == 0 || c1 == g.n - 1);  // pixmap.py:91-127
    pm::sync();
}

// first pixel (raster order) that is NOT in the segmap; N if none
template <typename T> PM_DEV int first_nonseg(const Ctx<T>& c) {
    const Geo& g = c.g;
    for (int r = 0; r < g.n; r++) {
        if (c.rowoff[r + 1] - c.rowoff[r] == (unsigned)g.n) continue;
        for (int w = 0; w < g.nw; w++) {
            unsigned z = ~c.seg[r * g.nw + w];
            if (w == g.nw - 1 && (g.n & 31)) z &= (1u << (g.n & 31)) - 1u;
            if (z) return r * g.n + w * 32 + pm::ffs(z) - 1;
        }
    }
    return g.N;
}

// pixmap.calcRmax (pixmap.py:26-55): returns rmax^2 (integer) and the brightest pixel
template <typename T> PM_DEV int phase_rmax2(Ctx<T>& c, const SegStats& st, int* cen_flat) {
    const Geo& g = c.g;
    int cen = st.argmax;
    if (!(st.vmax > 0.0) && c.n_g < g.N) {   // np.argmax(image * segmap): zeros outside the map compete
        const int fz = first_nonseg(c);
        if (st.vmax < 0.0) cen = fz;
        else cen = min(cen, fz);              // vmax == 0: first of {zero-valued map pixel, outside pixel}
    }
    const int cr = cen / g.n, cc = cen - cr * g.n;
    int d2 = 0;
    for (int i = pm::tid(); i < c.n_g; i += kThreads) {
        const unsigned p = c.posR[i];
        const int dr = (int)(p >> 16) - cr, dc = (int)(p & 0xffff) - cc;
        d2 = max(d2, dr * dr + dc * dc);
    }
    *cen_flat = cen;
    return block_max_i(*c.sh, d2);
}

// ====================================================================================================
// apertures.aperpixmap (apertures.py:42-128) — float64 sub-sample test, as the reference evaluates it
// ====================================================================================================
// |coordinate| of sub-sample k (0..ns-1, counted outward) of the pixel at distance d >= 1 from index n//2:
// the reference's vector is  q / ns - (n//2 + 1)  for q < (n//2) * ns,  zeros(ns),  and the mirrored copy.
PM_DEV double sub_coord_sq(int d, int k, int ns, int half) {
    if (d == 0) return 0.0;
    const int i = half - d;                  // pixel index on the negative side
    const int q = i * ns + (ns - 1 - k);     // k = 0 is the sample nearest the centre
    const double v = pm::dsub(pm::ddiv((double)q, (double)ns), (double)(half + 1));
    return pm::dmul(v, v);
}
// count of sub-sample pairs (k, l) with sqrt(x_k^2 + y_l^2) <= rad, via the squared threshold `a_star`
PM_DEV int sub_count(const double* sq, int d, int e, int ns, double a_star) {
    int cnt = 0;      // (a two-pointer sweep over the sorted coordinates measured slower: divergent, dependent loads)
    for (int k = 0; k < ns; k++) {
        const double a = sq[d * ns + k];
        for (int l = 0; l < ns; l++) cnt += (pm::dadd(a, sq[e * ns + l]) <= a_star) ? 1 : 0;
    }
    return cnt;
}
// largest double `a` with sqrt(a) <= rad (sqrt is monotone and correctly rounded)
PM_DEV double sqrt_threshold(double rad) {
    if (!(rad >= 0.0)) return -1.0;
    double a = pm::dmul(rad, rad);
    if (!(a < 1.7e308)) return 1.7e308;
    for (int it = 0; it < 8 && pm::dsqrt(a) > rad; it++) a = pm::bits2d(pm::d2bits(a) - 1);
    for (int it = 0; it < 8; it++) {
        const double nx = pm::bits2d(pm::d2bits(a) + 1);
        if (pm::dsqrt(nx) <= rad) a = nx; else break;
    }
    return a;
}
// Builds the n x n aperture plane for radius `rad`.  `sq` ((dmax+1)*ns doubles) and `hw` (dmax+1 ints) are scratch.
PM_DEV void build_aperture(const Geo& g, double rad, int ns, double frac, unsigned* plane, double* sq, int* hw) {
    const int half = g.half;
    const int dmax = half;   // |i - n//2| <= n//2
    for (int i = pm::tid(); i < (dmax + 1) * ns; i += kThreads) sq[i] = sub_coord_sq(i / ns, i % ns, ns, half);
    pm::sync();
    const double a_star = sqrt_threshold(rad);
    int need = 0;   // smallest count with count / ns^2 >= frac (apertures.py:122-123)
    {
        const double den = (double)(ns * ns);
        while (need <= ns * ns && !(pm::ddiv((double)need, den) >= frac)) need++;
    }
    for (int d = pm::tid(); d <= dmax; d += kThreads) {
        // the count is non-increasing in e: binary search for the largest e that still qualifies
        int lo = -1, hi = dmax;   // invariant: qualifies(lo) (or lo == -1), !qualifies(hi + 1)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sub_count(sq, d, mid, ns, a_star) >= need) lo = mid; else hi = mid - 1;
        }
        hw[d] = lo;
    }
    pm::sync();
    for (int task = pm::tid(); task < g.n * g.nw; task += kThreads) {
        const int r = task / g.nw, w = task - r * g.nw;
        const int d = r > half ? r - half : half - r;
        const int e = hw[d];
        unsigned bits = 0;
        if (e >= 0) {
            // columns [half - e, half + e] ∩ [0, n)
            const int lo = max(half - e, 0), hi = min(half + e, g.n - 1);
            const int a = max(lo - w * 32, 0), b = min(hi - w * 32, 31);
            if (a <= b) bits = (b - a == 31) ? 0xffffffffu : (((1u << (b - a + 1)) - 1u) << a);
        }
        plane[task] = bits;
    }
    pm::sync();
}

// ====================================================================================================
// Phase 3 — sort, gini / m20 (casgm.py:187-273), centroid candidates (asymmetry.py:89-108)
// ====================================================================================================
struct SortOut {
    double gini, m20;
    int n_cand, n_pos;
};

// sorted keys/positions ascending in (sk, sp); st = segmap statistics
template <typename T>
PM_DEV void phase_gini_m20_cands(Ctx<T>& c, const typename KeyT<T>::type* sk, const unsigned* sp, const SegStats& st,
                                 SortOut* out) {
    Sh& sh = *c.sh;
    const Geo& g = c.g;
    const int n = c.n_g;
    const int chunk = (n + kThreads - 1) / kThreads;
    const int lo = min(n, pm::tid() * chunk), hi = min(n, lo + chunk);
    // ---- gini: Σ (2i - n - 1)|x_i| / (|mean| n (n - 1)), i = 1..n ascending (photutils.morphology.gini)
    // one pass, one reduction: gini sum, Σ negatives and the counts of positive / negative values (exact in float64)
    double gs = 0.0, loc = 0.0, negsum = 0.0;
    int npos = 0, nneg = 0;
    for (int i = lo; i < hi; i++) {
        const double v = val_of_key(sk[i], c.sky);
        gs += (double)(2 * (i + 1) - n - 1) * fabs(v);
        loc += v;
        npos += v > 0.0;
        if (v < 0.0) { nneg++; negsum += v; }
    }
    {
        double red4[4] = {gs, negsum, (double)npos, (double)nneg};
        block_sum<4>(sh, red4);
        gs = red4[0]; negsum = red4[1]; out->n_pos = (int)red4[2]; nneg = (int)red4[3];
    }
    const double mean = st.sum / (double)n;
    out->gini = gs / (fabs(mean) * (double)n * (double)(n - 1));
    c.pc.lap(32);
    // ---- m20 threshold: first element of the ascending order of ALL N values (zeros outside the map
    //      included) whose cumulative fraction exceeds 0.8  (casgm.py:261-263)
    double tot_seq;
    const double pre = block_excl_prefix_d(sh, loc, &tot_seq);
    const int nzero = g.N - n;                 // zeros contributed by pixels outside the map
    int first_rank = 0x7fffffff;
    {
        double cum = pre;
        for (int i = lo; i < hi; i++) {
            const double v = val_of_key(sk[i], c.sky);
            cum += v;
            if (cum / st.sum > 0.8) {
                const int rank = (v < 0.0) ? i : i + nzero;
                first_rank = min(first_rank, rank);
                break;
            }
        }
    }
    // the block of outside zeros sits after the negatives: cumulative sum there = Σ negatives
    if (nzero > 0 && (negsum / st.sum > 0.8)) first_rank = min(first_rank, nneg);
    first_rank = block_min_i(sh, first_rank);
    double thresh;
    bool have_thresh = first_rank != 0x7fffffff;
    if (!have_thresh) thresh = 0.0;
    else if (nzero > 0 && first_rank >= nneg && first_rank < nneg + nzero) thresh = 0.0;
    else thresh = val_of_key(sk[first_rank < nneg ? first_rank : first_rank - nzero], c.sky);
    c.pc.lap(33);
    // ---- second moments about the flux centroid (casgm.py:253-271)
    const double rc = st.m10 / st.sum, cc = st.m01 / st.sum;
    double mt = 0.0, m2 = 0.0;
    for (int i = pm::tid(); i < n; i += kThreads) {
        const double v = val_of_key(sk[i], c.sky);
        const unsigned p = sp[i];
        const double dr = (double)(p >> 16) - rc, dc = (double)(p & 0xffff) - cc;
        const double q = v * (dr * dr + dc * dc);
        mt += q;
        if (v >= thresh) m2 += q;
    }
    double red[2] = {mt, m2};
    block_sum<2>(sh, red);
    out->m20 = have_thresh ? log10(red[1] / red[0]) : nan_of(double());
    c.pc.lap(34);
    // ---- candidates: walk the DESCENDING order while the running flux is below 20 % (asymmetry.py:101-108)
    const double t20 = 0.2 * st.sum;
    double dloc = 0.0;
    for (int i = lo; i < hi; i++) {          // descending element i  ==  ascending element n - 1 - i
        const double v = val_of_key(sk[n - 1 - i], c.sky);
        if (v > 0.0) dloc += v;
    }
    double dtot;
    const double dpre = block_excl_prefix_d(sh, dloc, &dtot);
    unsigned nc = 0;
    {
        double run = dpre;   // flux accumulated BEFORE descending element i
        for (int i = lo; i < hi; i++) {
            const double v = val_of_key(sk[n - 1 - i], c.sky);
            if (!(v > 0.0)) break;
            if (i == 0 || run < t20) nc++;
            run += v;
        }
    }
    out->n_cand = (int)block_sum_u(sh, nc);
    c.pc.lap(35);
}

// ====================================================================================================
// Phase 4 — asymmetry.minapix (asymmetry.py:50-129)
// ====================================================================================================
// disc test for the re-centred aperture of candidate (px, py): apertures.apercentre rolls the mask by
// pix - (n//2 + 1) on both axes (cyclic), with pix = [col, row]  (asymmetry.py:106,118; apertures.py:231-238)
template <bool kWrap> PM_DEV void wrap_rc(const Geo& g, int& rr, int& cc) {
    if (kWrap) {
        if (rr < 0) rr += g.n; else if (rr >= g.n) rr -= g.n;
        if (cc < 0) cc += g.n; else if (cc >= g.n) cc -= g.n;
    }
}
PM_DEV float lds_t(unsigned a, float) { return pm::lds_f32(a); }
PM_DEV double lds_t(unsigned a, double) { return pm::lds_f64(a); }

constexpr int kCandBlock = 4;   // candidates evaluated together per pass over the pixel list

// One work item = (group of kCandBlock candidates, slice of the raster pixel list), evaluated by one warp:
// every lane loads a map pixel once and pairs it with its mirror image about each candidate of the group.
// kSmem: the pixel list and the masked window live in shared memory (explicit 32-bit addressed loads).
template <typename T, bool kWrap, bool kSmem>
PM_DEV void minapix_item(const Ctx<T>& c, const int (&px)[kCandBlock], const int (&py)[kCandBlock], int first, int end, int stride,
                         double (&num)[kCandBlock], double (&den)[kCandBlock]) {
    const Geo& g = c.g;
    const int n = g.n, bw = c.bw, bh = c.r1 - c.r0 + 1, r0 = c.r0, c0 = c.c0, nw = g.nw;
    const unsigned ap32 = pm::smem_addr(c.aper);
    const unsigned mw32 = kSmem ? pm::smem_addr(c.mw) : 0u, pos32 = kSmem ? pm::smem_addr(c.plist) : 0u;
    int d0[kCandBlock], d1[kCandBlock];
#pragma unroll
    for (int j = 0; j < kCandBlock; j++) { d0[j] = px[j] - (g.half + 1); d1[j] = py[j] - (g.half + 1); num[j] = 0.0; den[j] = 0.0; }
    for (int i = first; i < end; i += stride) {
        const unsigned p = kSmem ? pm::lds_u32(pos32 + 4u * (unsigned)i) : c.plist[i];
        const int r = (int)(p >> 16), col = (int)(p & 0xffff);
        const unsigned vi = (unsigned)((r - r0) * bw + (col - c0));
        const T vraw = kSmem ? lds_t(mw32 + (unsigned)sizeof(T) * vi, T()) : c.mw[vi];
        const double v = (double)vraw - c.sky;
        const double av = fabs(v);
        // all loads of the kCandBlock candidates first (independent, in flight together) ...
        T w[kCandBlock];
        unsigned b1[kCandBlock], b2[kCandBlock];
        bool inwin[kCandBlock], inb[kCandBlock];
#pragma unroll
        for (int j = 0; j < kCandBlock; j++) {
            const int tr = 2 * py[j] - r, tc = 2 * px[j] - col;      // mirror image of the pixel about the candidate
            const unsigned wr = (unsigned)(tr - r0), wc = (unsigned)(tc - c0);
            inwin[j] = wr < (unsigned)bh && wc < (unsigned)bw;
            const unsigned wi = inwin[j] ? wr * (unsigned)bw + wc : 0u;
            w[j] = kSmem ? lds_t(mw32 + (unsigned)sizeof(T) * wi, T()) : c.mw[wi];
            inb[j] = (unsigned)tr < (unsigned)n && (unsigned)tc < (unsigned)n;
            int rr = r - d0[j], cc = col - d1[j];
            wrap_rc<kWrap>(g, rr, cc);
            b1[j] = pm::lds_u32(ap32 + 4u * (unsigned)(rr * nw + (cc >> 5))) >> (cc & 31);
            int r2 = inb[j] ? tr - d0[j] : g.half, c2 = inb[j] ? tc - d1[j] : g.half;
            wrap_rc<kWrap>(g, r2, c2);
            b2[j] = pm::lds_u32(ap32 + 4u * (unsigned)(r2 * nw + (c2 >> 5))) >> (c2 & 31);
        }
        // ... then the (predicated) float64 accumulation
#pragma unroll
        for (int j = 0; j < kCandBlock; j++) {
            const bool t_in_seg = inwin[j] && (w[j] == w[j]);
            double mt = 0.0;
            if (t_in_seg) mt = (double)w[j] - c.sky;
            const bool a1 = (b1[j] & 1u) != 0;
            pm::dadd_if(a1, num[j], fabs(v - mt));
            pm::dadd_if(a1, den[j], av);
            // mirror position inside the image but outside the map: there M = 0 and M180 = v
            pm::dadd_if((b2[j] & 1u) && inb[j] && !t_in_seg, num[j], av);
        }
    }
}

// One warp: candidates cidx[0..3] (indices into `cand`) against the list elements first, first + stride, ... < end;
// warp-reduced sums to out[2 j], out[2 j + 1] (lane 0).
template <typename T>
PM_DEV void minapix_group(const Ctx<T>& c, const unsigned* cand, const int (&cidx)[kCandBlock], int first, int end, int stride,
                          double* out) {
    const Geo& g = c.g;
    int px[kCandBlock], py[kCandBlock];
    bool nowrap = true;
#pragma unroll
    for (int j = 0; j < kCandBlock; j++) {
        const unsigned pp = cand[cidx[j]];
        py[j] = (int)(pp >> 16); px[j] = (int)(pp & 0xffff);
        // rows / columns the aperture is probed at: the map's bounding box and its mirror image (clipped)
        const int d0 = px[j] - (g.half + 1), d1 = py[j] - (g.half + 1);
        const int rlo = min(c.r0, max(0, 2 * py[j] - c.r1)), rhi = max(c.r1, min(g.n - 1, 2 * py[j] - c.r0));
        const int clo = min(c.c0, max(0, 2 * px[j] - c.c1)), chi = max(c.c1, min(g.n - 1, 2 * px[j] - c.c0));
        nowrap = nowrap && (rlo - d0 >= 0) && (rhi - d0 < g.n) && (clo - d1 >= 0) && (chi - d1 < g.n);
    }
    double num[kCandBlock], den[kCandBlock];
    const bool in_smem = c.ar.in_smem(c.mw) && c.ar.in_smem(c.plist);
    if (in_smem) {
        if (nowrap) minapix_item<T, false, true>(c, px, py, first, end, stride, num, den);
        else minapix_item<T, true, true>(c, px, py, first, end, stride, num, den);
    } else {
        if (nowrap) minapix_item<T, false, false>(c, px, py, first, end, stride, num, den);
        else minapix_item<T, true, false>(c, px, py, first, end, stride, num, den);
    }
#pragma unroll
    for (int j = 0; j < kCandBlock; j++) {
        const double a = pm::warp_sum(num[j]), b = pm::warp_sum(den[j]);
        if (pm::lane() == 0) { out[2 * j] = a; out[2 * j + 1] = b; }
    }
}

template <typename T>
PM_DEV void phase_minapix(Ctx<T>& c, const unsigned* cand, int n_cand, double* a_out, double* parts, int slices,
                          int* best) {
    Sh& sh = *c.sh;
    const int groups = (n_cand + kCandBlock - 1) / kCandBlock;
    const int items = groups * slices;
    for (int item = pm::warp(); item < items; item += kWarps) {
        const int grp = item / slices, sl = item - grp * slices;
        int cidx[kCandBlock];
#pragma unroll
        for (int j = 0; j < kCandBlock; j++) cidx[j] = min(grp * kCandBlock + j, n_cand - 1);
        minapix_group(c, cand, cidx, sl * 32 + pm::lane(), c.n_g, 32 * slices, parts + 2 * item * kCandBlock);
    }
    pm::sync();
    double amin = 1.7976931348623157e308 * 2.0;   // +inf
    int nan_idx = 0x7fffffff;
    for (int k = pm::tid(); k < n_cand; k += kThreads) {
        const int grp = k / kCandBlock, j = k - grp * kCandBlock;
        double num = 0.0, den = 0.0;
        for (int s = 0; s < slices; s++) {
            const int item = grp * slices + s;
            num += parts[2 * (item * kCandBlock + j)];
            den += parts[2 * (item * kCandBlock + j) + 1];
        }
        const double a = num / (2.0 * den);
        a_out[k] = a;
        if (a != a) nan_idx = min(nan_idx, k); else amin = fmin(amin, a);
    }
    pm::sync();
    nan_idx = block_min_i(sh, nan_idx);
    if (nan_idx != 0x7fffffff) { *best = nan_idx; return; }   // np.argmin returns the first NaN
    amin = block_min(sh, amin);
    int idx = 0x7fffffff;
    for (int k = pm::tid(); k < n_cand; k += kThreads)
        if (a_out[k] == amin) { idx = k; break; }
    *best = block_min_i(sh, idx);
}

// minapix with exact pruning (used when the per-candidate asymmetries are not requested as an output).
// Every term of the numerator Σ|M - M180| is non-negative and the denominator is at most D = Σ_map |M|, so a
// candidate whose PARTIAL numerator already exceeds 2 D a* — a* being the asymmetry of some fully evaluated
// candidate — cannot be the minimum and is dropped.  The first kCandBlock candidates (the brightest pixels; on
// the bench set the minimum is among them 99 % of the time) are evaluated completely first; the others advance
// through kStages contiguous chunks of the raster list — the middle rows, which hold most of the flux, first —
// and are re-grouped after every chunk.  The value of a surviving candidate does not depend on which others
// were dropped (fixed chunk order, one warp per group and chunk), so results are reproducible.
constexpr int kStages = 8;
template <typename T>
PM_DEV void phase_minapix_pruned(Ctx<T>& c, const unsigned* cand, int n_cand, double dmax, double* acc /*[2 n_cand]*/,
                                 int* alive /*[n_cand]*/, int* alive2 /*[n_cand]*/, unsigned* flag /*[n_cand + 1]*/,
                                 double* parts /*[kWarps*8 + 8*groups]*/,
                                 int* best) {
    Sh& sh = *c.sh;
    const int n = c.n_g;
    // ---- the first group, completely: all warps, interleaved slices
    {
        int cidx[kCandBlock];
#pragma unroll
        for (int j = 0; j < kCandBlock; j++) cidx[j] = min(j, n_cand - 1);
        minapix_group(c, cand, cidx, pm::warp() * 32 + pm::lane(), n, 32 * kWarps, parts + 8 * pm::warp());
        pm::sync();
        for (int k = pm::tid(); k < 2 * n_cand; k += kThreads) acc[k] = 0.0;
        pm::sync();
        if (pm::tid() < min(kCandBlock, n_cand)) {
            double num = 0.0, den = 0.0;
            for (int w = 0; w < kWarps; w++) { num += parts[8 * w + 2 * pm::tid()]; den += parts[8 * w + 2 * pm::tid() + 1]; }
            acc[2 * pm::tid()] = num; acc[2 * pm::tid() + 1] = den;
        }
        pm::sync();
    }
    c.pc.lap(20);   // (minapix: first group)
    double astar = 1.7976931348623157e308 * 2.0;
    bool any_nan = false;
    for (int k = 0; k < min(kCandBlock, n_cand); k++) {
        const double a = acc[2 * k] / (2.0 * acc[2 * k + 1]);
        if (a != a) any_nan = true; else astar = fmin(astar, a);
    }
    // ---- the rest, chunk by chunk
    int n_alive = max(0, n_cand - kCandBlock);
    for (int k = pm::tid(); k < n_alive; k += kThreads) alive[k] = kCandBlock + k;
    pm::sync();
    const int order[kStages] = {3, 4, 2, 5, 1, 6, 0, 7};   // raster list: the middle rows hold most of the flux
    const double cut = any_nan ? 1.7976931348623157e308 * 2.0 : 2.0 * dmax * (1.0 + 1e-9) * astar * (1.0 + 1e-9);
    for (int st = 0; st < kStages && n_alive > 0; st++) {
        const int ch = order[st];
        const int lo = (int)(((long long)n * ch) / kStages), hi = (int)(((long long)n * (ch + 1)) / kStages);
        const int groups = (n_alive + kCandBlock - 1) / kCandBlock;
        for (int grp = pm::warp(); grp < groups; grp += kWarps) {
            int cidx[kCandBlock];
#pragma unroll
            for (int j = 0; j < kCandBlock; j++) cidx[j] = alive[min(grp * kCandBlock + j, n_alive - 1)];
            minapix_group(c, cand, cidx, lo + pm::lane(), hi, 32, parts + 8 * kWarps + 8 * grp);
        }
        pm::sync();
        c.pc.lap(21);   // (minapix: stage sums)
        c.pc.count(24 + st, n_alive);
        // accumulate, decide, compact the survivors (order preserved)
        for (int k = pm::tid(); k < n_alive; k += kThreads) {
            const int ci = alive[k];
            const double num = acc[2 * ci] + parts[8 * kWarps + 8 * (k / kCandBlock) + 2 * (k % kCandBlock)];
            const double den = acc[2 * ci + 1] + parts[8 * kWarps + 8 * (k / kCandBlock) + 2 * (k % kCandBlock) + 1];
            acc[2 * ci] = num; acc[2 * ci + 1] = den;
            flag[k] = (st == kStages - 1 || !(num > cut)) ? 1u : 0u;     // NaN partial sums are never dropped
        }
        pm::sync();
        if (st == kStages - 1) break;
        const int keep = (int)block_scan_excl_array(sh, flag, n_alive);
        for (int k = pm::tid(); k < n_alive; k += kThreads)
            if (flag[k + 1] != flag[k]) alive2[flag[k]] = alive[k];
        pm::sync();
        { int* t = alive; alive = alive2; alive2 = t; }
        n_alive = keep;
        c.pc.lap(22);   // (minapix: decide + compact)
    }
    // ---- first minimum over the completely evaluated candidates (dropped ones are strictly larger than a*)
    double amin = 1.7976931348623157e308 * 2.0;
    int nan_idx = 0x7fffffff;
    for (int k = pm::tid(); k < n_alive + min(kCandBlock, n_cand); k += kThreads) {
        const int ci = k < min(kCandBlock, n_cand) ? k : alive[k - min(kCandBlock, n_cand)];
        const double a = acc[2 * ci] / (2.0 * acc[2 * ci + 1]);
        if (a != a) nan_idx = min(nan_idx, ci); else amin = fmin(amin, a);
    }
    nan_idx = block_min_i(sh, nan_idx);
    if (nan_idx != 0x7fffffff) { *best = nan_idx; return; }
    amin = block_min(sh, amin);
    int idx = 0x7fffffff;
    for (int k = pm::tid(); k < n_alive + min(kCandBlock, n_cand); k += kThreads) {
        const int ci = k < min(kCandBlock, n_cand) ? k : alive[k - min(kCandBlock, n_cand)];
        if (acc[2 * ci] / (2.0 * acc[2 * ci + 1]) == amin) idx = min(idx, ci);
    }
    *best = block_min_i(sh, idx);
}

// ====================================================================================================
// Phase 5 — asymmetry.calcA x3 (asymmetry.py:132-257): A + Abgr (image), As (180°) and As90 (90°) on the map
// ====================================================================================================
PM_DEV void dilate9(const Geo& g, const unsigned* seg, unsigned* dil, unsigned* tmp) {
    // horizontal OR over offsets -4..+4 (zero border), then vertical OR over 9 rows   (asymmetry.py:220)
    for (int task = pm::tid(); task < g.n * g.nw; task += kThreads) {
        const int r = task / g.nw, w = task - r * g.nw;
        const unsigned v = seg[task];
        const unsigned lft = w > 0 ? seg[task - 1] : 0u, rgt = w + 1 < g.nw ? seg[task + 1] : 0u;
        unsigned o = v;
#pragma unroll
        for (int s = 1; s <= 4; s++) o |= (v << s) | (lft >> (32 - s)) | (v >> s) | (rgt << (32 - s));
        if (w == g.nw - 1 && (g.n & 31)) o &= (1u << (g.n & 31)) - 1u;
        tmp[task] = o;
        (void)r;
    }
    pm::sync();
    for (int task = pm::tid(); task < g.n * g.nw; task += kThreads) {
        const int r = task / g.nw, w = task - r * g.nw;
        unsigned o = 0;
        for (int dr = -4; dr <= 4; dr++) {
            const int rr = r + dr;
            if (rr >= 0 && rr < g.n) o |= tmp[rr * g.nw + w];
        }
        dil[task] = o;
    }
    pm::sync();
}

struct BgrCtx {
    const unsigned* dil;
    const unsigned* rowoff;   // exclusive prefix of dilated-mask pixels per row, [n + 1]
    const unsigned* wordoff;  // ... and per (row, word), [n * nw]
    int nb, nm;
    int nfree;                // pixels before the first row that holds a mask pixel: there, select(j) == j
};
// rank of a mask pixel among mask pixels in raster order (wordoff: exclusive prefix per (row, word))
PM_DEV int mask_rank(const Geo& g, const BgrCtx& b, int r, int col) {
    const int w = col >> 5;
    return (int)b.wordoff[r * g.nw + w] + pm::popc(b.dil[r * g.nw + w] & ((1u << (col & 31)) - 1u));
}
// flat index of the j-th (0-based) background pixel in raster order
PM_DEV int bgr_select(const Geo& g, const BgrCtx& b, int j) {
    if (j < b.nfree) return j;     // the leading all-background rows: the common case
    int lo = 0, hi = g.n - 1;      // largest row r with  r*n - rowoff[r] <= j
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (mid * g.n - (int)b.rowoff[mid] <= j) lo = mid; else hi = mid - 1;
    }
    int k = j - (lo * g.n - (int)b.rowoff[lo]);
    const unsigned* row = b.dil + lo * g.nw;
    for (int w = 0; w < g.nw; w++) {
        unsigned z = ~row[w];
        if (w == g.nw - 1 && (g.n & 31)) z &= (1u << (g.n & 31)) - 1u;
        const int cnt = pm::popc(z);
        if (k < cnt) return lo * g.n + w * 32 + pm::nth_set_bit(z, k);
        k -= cnt;
    }
    return lo * g.n;   // unreachable
}
// background image B of asymmetry.py:214-244: flat index of the image pixel whose value B takes at (r, col)
PM_DEV int bgr_source(const Geo& g, const BgrCtx& b, int r, int col) {
    if (!bit_at(b.dil, g.nw, r, col)) return r * g.n + col;
    int k = mask_rank(g, b, r, col);
    if (k >= b.nb) k %= b.nb;                          // :227-239 truncate or tile
    return bgr_select(g, b, k);
}

struct AsymOut { double A, Abgr, As, As90; };

template <typename T>
PM_DEV void phase_calcA(Ctx<T>& c, int cx, int cy, AsymOut* out, unsigned* tmp_plane) {
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    const bool do_img = (c.flags & (PM_DO_A | PM_DO_CAS)) != 0;
    const bool do_shape = (c.flags & PM_DO_AS) != 0;
    const bool noise = do_img && !(c.flags & PM_NO_NOISECORRECT);
    const bool rot90 = (c.flags & PM_A_ANGLE90) != 0;
    BgrCtx b;
    b.dil = c.dil; b.rowoff = c.rowoff; b.wordoff = tmp_plane; b.nb = 0; b.nm = 0; b.nfree = 0;
    if (noise) {
        dilate9(g, c.seg, c.dil, tmp_plane);
        for (int r = pm::tid(); r < g.n; r += kThreads) {
            c.rowoff[r] = row_popcount(c.dil, g.nw, r);
        }
        pm::sync();
        b.nm = (int)block_scan_excl_array(sh, c.rowoff, g.n);
        b.nb = g.N - b.nm;
        int fm = 0x7fffffff;
        for (int r = pm::tid(); r < g.n; r += kThreads)
            if (c.rowoff[r + 1] != c.rowoff[r]) fm = min(fm, r);
        fm = block_min_i(sh, fm);
        b.nfree = (fm == 0x7fffffff) ? g.N : fm * g.n;
        for (int r = pm::tid(); r < g.n; r += kThreads) {        // per-word prefix (the dilation scratch is free again)
            unsigned run = c.rowoff[r];
            for (int w = 0; w < g.nw; w++) { tmp_plane[r * g.nw + w] = run; run += (unsigned)pm::popc(c.dil[r * g.nw + w]); }
        }
        pm::sync();
    }
    const bool bgr_ok = noise && b.nb > 0 && b.nm > 1;      // asymmetry.py:225-226
    c.pc.lap(9);    // (calcA: dilation + row scan)
    // compact list of the non-empty aperture words (deterministic order)
    Arena::Mark mk_t = c.ar.mark();
    const int nwords = g.n * g.nw;
    unsigned* flag = (unsigned*)c.ar.alloc((size_t)(nwords + 1) * 4);
    unsigned* tlist = (unsigned*)c.ar.alloc((size_t)nwords * 4);      // (row << 16) | word
    for (int t = pm::tid(); t < nwords; t += kThreads) flag[t] = c.aper[t] != 0u;
    pm::sync();
    const int ntask = (int)block_scan_excl_array(sh, flag, nwords);
    for (int t = pm::tid(); t < nwords; t += kThreads)
        if (c.aper[t] != 0u) { const int r = t / g.nw; tlist[flag[t]] = ((unsigned)r << 16) | (unsigned)(t - r * g.nw); }
    pm::sync();
    double s[3] = {0.0, 0.0, 0.0};           // numA denA numB
    unsigned u180 = 0, uden = 0, u90 = 0;    // shape asymmetry: integer pixel counts
    // kTasks aperture words per warp-step, in three stages so that the image loads of all of them are in flight
    // together: (1) source indices (shared-memory work only, branchy), (2) all loads back to back, (3) arithmetic
    constexpr int kTasks = 2;
    for (int k0 = pm::warp(); k0 < ntask; k0 += kTasks * kWarps) {
        int r_[kTasks], col_[kTasks], fR[kTasks], fB[kTasks], fBr[kTasks], rR[kTasks], cR[kTasks];
        bool on[kTasks], in180[kTasks], in90[kTasks];
#pragma unroll
        for (int u = 0; u < kTasks; u++) {
            const int k = k0 + u * kWarps;
            const unsigned task = tlist[min(k, ntask - 1)];
            const int r = (int)(task >> 16), col = (int)(task & 0xffffu) * 32 + pm::lane();
            const unsigned bits = c.aper[r * g.nw + (int)(task & 0xffffu)];
            on[u] = k < ntask && ((bits >> pm::lane()) & 1u);
            r_[u] = r; col_[u] = col;
            // 180°: (r, c) -> (2cy - r, 2cx - c);  90° (skimage, counter-clockwise): (cy + (c - cx), cx - (r - cy))
            const int r180 = 2 * cy - r, c180 = 2 * cx - col;
            const int r90 = cy + (col - cx), c90 = cx - (r - cy);
            in180[u] = r180 >= 0 && r180 < g.n && c180 >= 0 && c180 < g.n;
            in90[u] = r90 >= 0 && r90 < g.n && c90 >= 0 && c90 < g.n;
            fR[u] = -1; fB[u] = -1; fBr[u] = -1; rR[u] = r; cR[u] = col;
            if (do_img && on[u]) {
                if (rot90) { if (in90[u]) { fR[u] = 1; rR[u] = r90; cR[u] = c90; } }
                else if (in180[u]) { fR[u] = 1; rR[u] = r180; cR[u] = c180; }
                if (bgr_ok) {
                    // the background image is rotated about the SWAPPED centre (asymmetry.py:245)
                    const int rb = 2 * cx - r, cb = 2 * cy - col;
                    fB[u] = bgr_source(g, b, r, col);
                    if (rb >= 0 && rb < g.n && cb >= 0 && cb < g.n) fBr[u] = bgr_source(g, b, rb, cb);
                }
            }
        }
        if (do_img) {
            T xI[kTasks], xR[kTasks], xB[kTasks], xBr[kTasks];
#pragma unroll
            for (int u = 0; u < kTasks; u++) {
                const int fI = r_[u] * g.n + col_[u];
                xI[u] = on[u] ? ld_raw(c, r_[u], col_[u]) : T(0);
                xR[u] = on[u] ? ld_raw(c, rR[u], cR[u]) : T(0);
                xB[u] = (on[u] && bgr_ok) ? pm::ldg(c.img + (fB[u] >= 0 ? fB[u] : fI)) : T(0);
                xBr[u] = (on[u] && bgr_ok) ? pm::ldg(c.img + (fBr[u] >= 0 ? fBr[u] : fI)) : T(0);
            }
#pragma unroll
            for (int u = 0; u < kTasks; u++) {
                if (!on[u]) continue;
                const double I = (double)xI[u] - c.sky;
                const double Ir = fR[u] >= 0 ? (double)xR[u] - c.sky : 0.0;
                s[0] += fabs(I - Ir);
                s[1] += fabs(I);
                if (bgr_ok) {
                    const double B = (double)xB[u] - c.sky;
                    const double Br = fBr[u] >= 0 ? (double)xBr[u] - c.sky : 0.0;
                    s[2] += fabs(B - Br);
                }
            }
        }
        if (do_shape) {
#pragma unroll
            for (int u = 0; u < kTasks; u++) {
                if (!on[u]) continue;
                const int r = r_[u], col = col_[u];
                const unsigned v = bit_at(c.seg, g.nw, r, col) ? 1u : 0u;
                const unsigned v180 = in180[u] && bit_at(c.seg, g.nw, 2 * cy - r, 2 * cx - col) ? 1u : 0u;
                const unsigned v90 = in90[u] && bit_at(c.seg, g.nw, cy + (col - cx), cx - (r - cy)) ? 1u : 0u;
                u180 += v ^ v180;
                uden += v;
                u90 += v ^ v90;
            }
        }
    }
    block_sum<3>(sh, s);
    const double n180 = (double)block_sum_u(sh, u180), nden = (double)block_sum_u(sh, uden), n90 = (double)block_sum_u(sh, u90);
    c.ar.release(mk_t);
    out->A = -99.0; out->Abgr = -99.0; out->As = -99.0; out->As90 = -99.0;
    if (do_img) {
        const double den = 2.0 * s[1];
        double A = s[0] / den, Ab = 0.0;
        if (noise) {
            if (bgr_ok) { Ab = s[2] / den; A = A - Ab; }
            else Ab = -99.0;
        }
        out->A = A; out->Abgr = Ab;
    }
    if (do_shape) {
        out->As = n180 / (2.0 * nden);
        out->As90 = n90 / (2.0 * nden);
    }
}

// ====================================================================================================
// Phase 6 — casgm.calcR20_R80 / concentration (casgm.py:22-184)
// ====================================================================================================
// Folded radial table: T[a(a+1)/2 + b] = Σ of M over the (up to 8) pixels at offsets (±a, ±b), (±b, ±a)
// from the integer centre, 0 <= b <= a <= A;  P = prefix of T along b within a column a.
struct Fold {
    double* T;
    double* P;
    int A;
    double abs_sum;   // Σ |T|
};
PM_DEV int tri(int a, int b) { return a * (a + 1) / 2 + b; }

template <typename T_> PM_DEV void build_fold(Ctx<T_>& c, int cx, int cy, Fold& f) {
    const Geo& g = c.g;
    const int A = f.A;
    const int entries = (A + 1) * (A + 2) / 2;
    // masked value at offset (dr, dc) from the centre; an unconditional (clamped) load so that the 8 mirror-image
    // loads of an entry — and those of a second entry — are all in flight together
    auto fetch = [&](int dr, int dc, bool valid) -> double {
        const int r = cy + dr, col = cx + dc;
        const bool ok = valid && r >= c.r0 && r <= c.r1 && col >= c.c0 && col <= c.c1;    // the window lies inside the image
        if (c.mw) {
            const T_ w = c.mw[ok ? (r - c.r0) * c.bw + (col - c.c0) : 0];
            return (ok && w == w) ? (double)w - c.sky : 0.0;
        }
        // no masked window (split path): the map's bit plane and the image itself
        const bool in = ok && bit_at(c.seg, c.g.nw, ok ? r : c.r0, ok ? col : c.c0);
        const T_ x = pm::ldg(c.img + (size_t)(in ? r : c.r0) * c.g.n + (in ? col : c.c0));
        return in ? (double)x - c.sky : 0.0;
    };
    auto entry_ab = [&](int e, int& a, int& b) {
        a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);     // invert the triangular index
        while (tri(a + 1, 0) <= e) a++;
        while (tri(a, 0) > e) a--;
        b = e - tri(a, 0);
    };
    // the (up to 8) mirror images of (a, b), duplicates masked out: a == b, b == 0, a == 0
    auto entry_sum = [&](int a, int b) -> double {
        const bool nz_a = a > 0, nz_b = b > 0, off = a != b;
        const double v0 = fetch(a, b, true), v1 = fetch(a, -b, nz_b), v2 = fetch(-a, b, nz_a), v3 = fetch(-a, -b, nz_a && nz_b);
        const double v4 = fetch(b, a, off), v5 = fetch(b, -a, off && nz_a), v6 = fetch(-b, a, off && nz_b), v7 = fetch(-b, -a, off && nz_b);
        double sum = v0;
        sum += v1; sum += v2; sum += v3; sum += v4; sum += v5; sum += v6; sum += v7;
        return sum;
    };
    for (int e0 = pm::tid(); e0 < entries; e0 += 2 * kThreads) {
        const int e1 = e0 + kThreads;
        int a0, b0, a1 = 0, b1 = 0;
        entry_ab(e0, a0, b0);
        if (e1 < entries) entry_ab(e1, a1, b1);
        const double s0 = entry_sum(a0, b0), s1 = entry_sum(a1, b1);
        f.T[e0] = s0;
        if (e1 < entries) f.T[e1] = s1;
    }
    pm::sync();
    double asum = 0.0;
    for (int a = pm::tid(); a <= A; a += kThreads) {
        double run = 0.0;
        for (int b = 0; b <= a; b++) { const double t = f.T[tri(a, b)]; run += t; asum += fabs(t); f.P[tri(a, b)] = run; }
    }
    f.abs_sum = block_sum1(*c.sh, asum);
}
// Column a of the folded table for a circle of radius r: entries b < b_lo lie fully inside (prefix sum P),
// entries b_lo <= b < b_hi are cut by the circle, the rest is outside.  The predicates are photutils'
// (d < r - pr, d < r + pr with d = sqrt(dx^2 + dy^2) in float64); a float estimate is corrected with them.
// smallest integer d2 >= 0 with !(sqrt((double)d2) < lim): photutils' float64 predicate as an integer threshold
PM_DEV int d2_threshold(double lim) {
    if (!(lim > 0.0)) return 0;
    if (!(lim < 46000.0)) return 0x7fffffff;
    int d = (int)(lim * lim);
    while (d > 0 && !(sqrt((double)(d - 1)) < lim)) d--;
    while (sqrt((double)d) < lim) d++;
    return d;
}
PM_DEV int count_below(int a, int lim);
// cut range of column a for integer thresholds (d2 < d_in: fully inside; d2 >= d_out: outside)
PM_DEV void fold_range_i(int a, int d_in, int d_out, int* b_lo_out, int* b_hi_out) {
    const int b_lo = count_below(a, d_in - 1), b_hi = count_below(a, d_out - 1);
    *b_lo_out = min(b_lo, a + 1);
    *b_hi_out = min(max(b_hi, b_lo), a + 1);
}
// cheap, conservative bounds of column a's contribution: squared-distance tests with a safety margin
// (no float64 square roots); every possibly-cut pixel is counted with weight 0 or 1, whichever is worse
PM_DEV int count_below(int a, int lim) {              // number of b >= 0 with a^2 + b^2 <= lim  (integers only)
    const int t = lim - a * a;
    if (t < 0) return 0;
    int b = (int)sqrtf((float)t);                      // exact to +-1 for t < 2^24 * 4 ...
    b -= (b * b > t) ? 1 : 0;                          // ... so one correction step each way is enough,
    b += ((b + 1) * (b + 1) <= t) ? 1 : 0;             //     and a second one covers the large-t rounding
    b -= (b * b > t) ? 1 : 0;
    b += ((b + 1) * (b + 1) <= t) ? 1 : 0;
    return b + 1;
}
// exact F(r) for nr <= kRedSlots radii; results in acc[0..nr) of every thread.  Pass 1: one (radius, column) item per
// thread finds the cut range and adds the fully-inside prefix; pass 2: the cut pixels of all columns are
// dealt out evenly, one exact circle/pixel overlap per thread-step.  off / blo: nr * (A + 1) + 1 entries each.
PM_DEV void eval_flux(Sh& sh, const Fold& f, const double* radii, int nr, unsigned* off, int* blo, double (&acc)[kRedSlots]) {
#pragma unroll
    for (int q = 0; q < kRedSlots; q++) acc[q] = 0.0;
    const int ncol = f.A + 1, items = nr * ncol;
    if (pm::tid() < nr) {                      // integer distance thresholds of each radius, once
        const double pr = 0.5 * sqrt(2.0);
        sh.ia[8 + 2 * pm::tid()] = d2_threshold(radii[pm::tid()] - pr);
        sh.ia[9 + 2 * pm::tid()] = d2_threshold(radii[pm::tid()] + pr);
    }
    pm::sync();
    for (int item = pm::tid(); item < items; item += kThreads) {
        const int q = item / ncol, a = item - q * ncol;
        const double r = radii[q];
        int b_lo = 0, b_hi = 0;
        if (!((double)a > r + 1.0)) fold_range_i(a, sh.ia[8 + 2 * q], sh.ia[9 + 2 * q], &b_lo, &b_hi);
        const double v = b_lo > 0 ? f.P[tri(a, b_lo - 1)] : 0.0;
#pragma unroll
        for (int k = 0; k < kRedSlots; k++)
            if (k == q) acc[k] += v;
        blo[item] = b_lo;
        off[item] = (unsigned)(b_hi - b_lo);
    }
    pm::sync();
    const int total = (int)block_scan_excl_array(sh, off, items);
    for (int e = pm::tid(); e < total; e += kThreads) {
        int lo = 0, hi = items - 1;          // last item with off[item] <= e
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)off[mid] <= e) lo = mid; else hi = mid - 1;
        }
        const int q = lo / ncol, a = lo - q * ncol, b = blo[lo] + (e - (int)off[lo]);
        const double v = f.T[tri(a, b)] * overlap_pixel(a, b, radii[q]);
#pragma unroll
        for (int k = 0; k < kRedSlots; k++)
            if (k == q) acc[k] += v;
    }
    // every thread ends up with F(radii[q]) in acc[q]; the reduction is as wide as the batch
    if (nr <= 1) { double t[1] = {acc[0]}; block_sum<1>(sh, t); acc[0] = t[0]; }
    else if (nr <= 2) { double t[2] = {acc[0], acc[1]}; block_sum<2>(sh, t); acc[0] = t[0]; acc[1] = t[1]; }
    else if (nr <= 4) { double t[4] = {acc[0], acc[1], acc[2], acc[3]}; block_sum<4>(sh, t); acc[0] = t[0]; acc[1] = t[1]; acc[2] = t[2]; acc[3] = t[3]; }
    else block_sum<kRedSlots>(sh, acc);
}
// bounds lo <= F(r_k) <= hi for the walk radii rk[1..K]: one warp per radius.  A pixel at squared distance
// d2 is certainly inside when d2 <= floor((r - pr - 1e-9)^2) - 1 and certainly outside when d2 > ceil((r + pr + 1e-9)^2).
PM_DEV void eval_flux_bounds(Sh& sh, const Fold& f, int K) {
    const double pr = 0.5 * sqrt(2.0);
    // two radii per warp-step: the two column scans are independent chains (ILP)
    for (int k0 = 1 + pm::warp(); k0 <= K; k0 += 2 * kWarps) {
        const int k1 = k0 + kWarps;
        const bool two = k1 <= K;
        const double r0 = sh.rk[k0], r1 = sh.rk[two ? k1 : k0];
        const double in0 = r0 - pr - 1e-9, out0 = r0 + pr + 1e-9, in1 = r1 - pr - 1e-9, out1 = r1 + pr + 1e-9;
        const int li0 = in0 > 0.0 ? (int)floor(in0 * in0) - 1 : -1, lo0_ = (int)ceil(out0 * out0);
        const int li1 = in1 > 0.0 ? (int)floor(in1 * in1) - 1 : -1, lo1_ = (int)ceil(out1 * out1);
        double lo0 = 0.0, hi0 = 0.0, lo1 = 0.0, hi1 = 0.0;
        const int amax = min(f.A, (int)(fmax(r0, r1) + 2.0));
        for (int a = pm::lane(); a <= amax; a += 32) {
            const int bi0 = min(count_below(a, li0), a + 1), bo0 = min(max(count_below(a, lo0_), bi0), a + 1);
            const int bi1 = min(count_below(a, li1), a + 1), bo1 = min(max(count_below(a, lo1_), bi1), a + 1);
            const double s0 = bi0 > 0 ? f.P[tri(a, bi0 - 1)] : 0.0, s1 = bi1 > 0 ? f.P[tri(a, bi1 - 1)] : 0.0;
            double l0 = s0, h0 = s0, l1 = s1, h1 = s1;
            const int steps = max(bo0 - bi0, bo1 - bi1);
            for (int i = 0; i < steps; i++) {
                const bool p0 = bi0 + i < bo0, p1 = bi1 + i < bo1;
                const double t0 = p0 ? f.T[tri(a, bi0 + i)] : 0.0, t1 = p1 ? f.T[tri(a, bi1 + i)] : 0.0;
                if (t0 < 0.0) l0 += t0; else h0 += t0;
                if (t1 < 0.0) l1 += t1; else h1 += t1;
            }
            lo0 += l0; hi0 += h0; lo1 += l1; hi1 += h1;
        }
        lo0 = pm::warp_sum(lo0); hi0 = pm::warp_sum(hi0); lo1 = pm::warp_sum(lo1); hi1 = pm::warp_sum(hi1);
        if (pm::lane() == 0) {
            sh.fk[k0] = lo0; sh.fh[k0] = hi0;
            if (two) { sh.fk[k1] = lo1; sh.fh[k1] = hi1; }
        }
    }
    pm::sync();
}

// scipy.optimize.brentq (Zeros/brentq.c), xtol = 2e-12, rtol = 4 eps, maxiter = 100, as a resumable state machine
struct Brent {
    double xpre, xcur, xblk, fpre, fcur, fblk, spre, scur;
    double root;
    int state;   // 0 = needs f(xcur), 1 = converged, 2 = sign error, 3 = no convergence
    int iter;
};
PM_DEV void brent_init(Brent& b, double xa, double xb, double fa, double fb) {
    b.xpre = xa; b.xcur = xb; b.fpre = fa; b.fcur = fb;
    b.xblk = 0.0; b.fblk = 0.0; b.spre = 0.0; b.scur = 0.0; b.iter = 0; b.state = 0; b.root = 0.0;
    if (fa == 0.0) { b.state = 1; b.root = xa; return; }
    if (fb == 0.0) { b.state = 1; b.root = xb; return; }
    if (signbit(fa) == signbit(fb)) { b.state = 2; return; }
}
// one iteration of the loop body up to (not including) the function evaluation; sets xcur to the next abscissa
PM_DEV void brent_step(Brent& b) {
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    if (b.iter >= 100) { b.state = 3; return; }
    b.iter++;
    if (b.fpre != 0.0 && b.fcur != 0.0 && (signbit(b.fpre) != signbit(b.fcur))) {
        b.xblk = b.xpre; b.fblk = b.fpre;
        b.spre = b.scur = b.xcur - b.xpre;
    }
    if (fabs(b.fblk) < fabs(b.fcur)) {
        b.xpre = b.xcur; b.xcur = b.xblk; b.xblk = b.xpre;
        b.fpre = b.fcur; b.fcur = b.fblk; b.fblk = b.fpre;
    }
    const double delta = (xtol + rtol * fabs(b.xcur)) / 2.0;
    const double sbis = (b.xblk - b.xcur) / 2.0;
    if (b.fcur == 0.0 || fabs(sbis) < delta) { b.state = 1; b.root = b.xcur; return; }
    if (fabs(b.spre) > delta && fabs(b.fcur) < fabs(b.fpre)) {
        double stry;
        if (b.xpre == b.xblk) {
            stry = -b.fcur * (b.xcur - b.xpre) / (b.fcur - b.fpre);
        } else {
            const double dpre = (b.fpre - b.fcur) / (b.xpre - b.xcur);
            const double dblk = (b.fblk - b.fcur) / (b.xblk - b.xcur);
            stry = -b.fcur * (b.fblk * dblk - b.fpre * dpre) / (dblk * dpre * (b.fblk - b.fpre));
        }
        if (2.0 * fabs(stry) < fmin(fabs(b.spre), 3.0 * fabs(sbis) - delta)) { b.spre = b.scur; b.scur = stry; }
        else { b.spre = sbis; b.scur = sbis; }
    } else {
        b.spre = sbis; b.scur = sbis;
    }
    b.xpre = b.xcur; b.fpre = b.fcur;
    if (fabs(b.scur) > delta) b.xcur += b.scur;
    else b.xcur += (sbis > 0.0 ? delta : -delta);
}

struct RadiiOut { double r20, r80; int ok; };

template <typename T_>
PM_DEV void phase_r20_r80(Ctx<T_>& c, int cx, int cy, double radius, RadiiOut* out) {
    Sh& sh = *c.sh;
    out->ok = 0; out->r20 = -99.0; out->r80 = -99.0;
    if (!(radius > 0.0) || !(radius < 1e6)) return;   // photutils: "'r' must be a positive scalar"
    Arena::Mark mk = c.ar.mark();
    Fold f;
    f.A = (int)(radius + 0.5 * sqrt(2.0)) + 1;
    const int entries = (f.A + 1) * (f.A + 2) / 2;
    f.T = (double*)c.ar.alloc((size_t)entries * 8);
    f.P = (double*)c.ar.alloc((size_t)entries * 8);
    unsigned* eoff = (unsigned*)c.ar.alloc((size_t)(kRedSlots * (f.A + 1) + 1) * 4);
    int* eblo = (int*)c.ar.alloc((size_t)(kRedSlots * (f.A + 1) + 1) * 4);
    int* eblo2 = eblo;
    unsigned char* vscratch = (unsigned char*)c.ar.alloc(2 * 132 + 16);
    if (c.ar.overflow) { c.ar.release(mk); return; }
    build_fold(c, cx, cy, f);
    c.pc.lap(10);   // (r20/r80: folded table)
    // total flux inside `radius` (casgm.py:51-52)
    if (pm::tid() == 0) sh.rk[0] = radius;
    // the walk radii (casgm.py:55-73): r_1 = R - δ, r_{k+1} = r_k - δ, δ = (R / 200) * 2
    const double delta = (radius / 200.0) * 2.0;
    int K = 0;   // number of walk radii that are > 0
    {
        // the reference subtracts delta repeatedly (rounding accumulates): every thread runs the short serial chain in
        // registers and thread k keeps r_k — no stores or branches on the chain
        if (kThreads > 130) {
            double r = radius, mine = 0.0;
#pragma unroll 10
            for (int k = 1; k <= 130; k++) {
                r -= delta;
                K += (r > 0.0) ? 1 : 0;               // r decreases monotonically: counts the leading positive radii
                if (k == pm::tid()) mine = r;
            }
            if (pm::tid() >= 1 && pm::tid() <= K) sh.rk[pm::tid()] = mine;
        } else {                                      // small groups: thread t keeps the radii k = t (mod kThreads)
            double r = radius;
            for (int k = 1; k <= 130; k++) {
                r -= delta;
                const bool pos = r > 0.0;
                K += pos ? 1 : 0;
                if (pos && (k % kThreads) == pm::tid()) sh.rk[k] = r;
            }
        }
    }
    pm::sync();
    double fv[kRedSlots];
    eval_flux(sh, f, sh.rk, 1, eoff, eblo, fv);
    const double total = fv[0];
    c.pc.lap(39);
    // Walk down (casgm.py:61-73): first k with F(r_k) / total <= fraction.  Cheap bounds decide most radii;
    // only the undecided ones near the crossing are evaluated exactly.
    eval_flux_bounds(sh, f, K);
    c.pc.lap(11);   // (r20/r80: total + walk bounds)
    const double slack = 1e-9 * f.abs_sum;          // >> rounding of any partial sum, << one pixel of flux
    const double fr[2] = {0.2, 0.8};
    // verdict per radius and fraction: 1 = certainly F/total <= fraction, 0 = certainly not, 2 = undecided
    unsigned char* verdict = vscratch;              // [2][K + 1]
    for (int item = pm::tid(); item < 2 * K; item += kThreads) {
        const int q = item / K, k = 1 + item - q * K;
        const double lo = sh.fk[k] - slack, hi = sh.fh[k] + slack;
        int v;
        if (total > 0.0) v = (hi / total <= fr[q] - 1e-12) ? 1 : ((lo / total > fr[q] + 1e-12) ? 0 : 2);
        else if (total < 0.0) v = (lo / total <= fr[q] - 1e-12) ? 1 : ((hi / total > fr[q] + 1e-12) ? 0 : 2);
        else v = 2;
        verdict[q * (K + 1) + k] = (unsigned char)v;
    }
    pm::sync();
    int kfound[2] = {-1, -1};
    int pos[2] = {1, 1};
    // Both fractions advance together: per round warp 0 scans the verdicts of each unfinished fraction 32 at a time
    // (ballots) and collects its undecided radii up to the first certain one — at most kRedSlots in total, half each
    // while both are active — then ONE batched exact evaluation serves both; the outcome is broadcast through
    // shared memory.  Per fraction the sequence of decisions is the reference's walk (casgm.py:61-73).
    for (;;) {
        const bool act[2] = {kfound[0] < 0 && pos[0] <= K, kfound[1] < 0 && pos[1] <= K};
        if (!act[0] && !act[1]) break;
        const int cap = (act[0] && act[1]) ? kRedSlots / 2 : kRedSlots;
        if (pm::warp() == 0) {
            for (int q = 0; q < 2; q++) {
                int nl_ = 0, sure_ = -1, k_ = pos[q];
                while (act[q] && k_ <= K && nl_ < cap && sure_ < 0) {
                    const int kk_ = k_ + pm::lane();
                    const int v = kk_ <= K ? (int)verdict[q * (K + 1) + kk_] : 0;
                    const unsigned m1 = pm::ballot(v == 1);
                    unsigned m2 = pm::ballot(v == 2);
                    const int f1 = m1 ? pm::ffs(m1) - 1 : 32;             // first certain radius of this group
                    if (f1 < 32) m2 &= (1u << f1) - 1u;
                    int consumed = 32;                                     // radii of this group dealt with
                    while (m2 && nl_ < cap) {
                        const int bpos = pm::ffs(m2) - 1;
                        if (pm::lane() == 0) sh.ia[8 + (act[0] && act[1] ? q * (kRedSlots / 2) : 0) + nl_] = k_ + bpos;
                        nl_++;
                        m2 &= m2 - 1u;
                        if (nl_ == cap) consumed = bpos + 1;
                    }
                    if (nl_ == cap && consumed <= f1) { k_ += consumed; break; }
                    if (f1 < 32) { sure_ = k_ + f1; break; }
                    k_ += 32;
                }
                if (pm::lane() == 0) { sh.ia[1 + 3 * q] = nl_; sh.ia[2 + 3 * q] = sure_; sh.ia[3 + 3 * q] = min(k_, K + 1); }
            }
        }
        pm::sync();
        int list[kRedSlots], nl[2], sure[2], nxt[2];
        for (int q = 0; q < 2; q++) { nl[q] = sh.ia[1 + 3 * q]; sure[q] = sh.ia[2 + 3 * q]; nxt[q] = sh.ia[3 + 3 * q]; }
        const int ntot = nl[0] + nl[1];
        for (int i = 0; i < kRedSlots; i++) {
            // both active: fraction 0's entries sit at ia[8..], fraction 1's at ia[8 + kRedSlots/2 ..]; one active: at ia[8..]
            const int src = (act[0] && act[1] && i >= nl[0]) ? kRedSlots / 2 + (i - nl[0]) : i;
            list[i] = i < ntot ? sh.ia[8 + src] : 0;
        }
        pm::sync();
        if (ntot > 0) {
            if (pm::tid() == 0)
                for (int i = 0; i < ntot; i++) sh.bc[8 + i] = sh.rk[list[i]];
            pm::sync();
            eval_flux(sh, f, sh.bc + 8, ntot, eoff, eblo2, fv);
            for (int i = 0; i < ntot; i++) {
                const int q = (act[0] && i < nl[0]) ? 0 : 1;
                if (kfound[q] < 0 && fv[i] / total <= fr[q]) kfound[q] = list[i];
            }
            pm::sync();      // sh.bc[8..] is rewritten by the next batch
        }
        for (int q = 0; q < 2; q++) {
            if (!act[q]) continue;
            if (kfound[q] < 0 && sure[q] > 0) kfound[q] = sure[q];
            pos[q] = (sure[q] > 0) ? K + 1 : nxt[q];
        }
    }
    const int k20 = kfound[0], k80 = kfound[1];
    c.pc.lap(12);   // (r20/r80: exact walk evaluations)
    if (k20 < 0 || k80 < 0) { c.ar.release(mk); return; }   // the reference walks to r <= 0 and raises
    // brentq on [r_k, r_k + δ] for both fractions, in lock-step.  Inside a bracket only the pixels with
    // lo - pr <= d < hi + pr can change weight: cache them once (entry + table value), everything closer to
    // the centre is a constant base sum.  One cached pixel per thread-step and evaluation.
    Brent b[2];
    const int kk[2] = {k20, k80};
    double blo_r[2], bhi_r[2];
    for (int q = 0; q < 2; q++) { blo_r[q] = sh.rk[kk[q]]; bhi_r[q] = sh.rk[kk[q]] + delta; }
    const double pr = 0.5 * sqrt(2.0);
    const int ncol = f.A + 1;
    const int dlim[4] = {d2_threshold(blo_r[0] - pr), d2_threshold(bhi_r[0] + pr), d2_threshold(blo_r[1] - pr), d2_threshold(bhi_r[1] + pr)};
    double base[2] = {0.0, 0.0};
    for (int item = pm::tid(); item < 2 * ncol; item += kThreads) {
        const int q = item / ncol, a = item - q * ncol;
        int bl = 0, bh = 0;
        if (!((double)a > bhi_r[q] + 1.0)) fold_range_i(a, dlim[2 * q], dlim[2 * q + 1], &bl, &bh);
        if (bl > 0) base[q] += f.P[tri(a, bl - 1)];
        eblo[item] = bl;
        eoff[item] = (unsigned)(bh - bl);
    }
    pm::sync();
    const int ncache = (int)block_scan_excl_array(sh, eoff, 2 * ncol);
    block_sum<2>(sh, base);
    unsigned* cent = (unsigned*)c.ar.alloc((size_t)max(ncache, 1) * 4);
    double* ctv = (double*)c.ar.alloc((size_t)max(ncache, 1) * 8);
    if (c.ar.overflow) { c.ar.release(mk); return; }
    for (int e = pm::tid(); e < ncache; e += kThreads) {
        int lo = 0, hi = 2 * ncol - 1;          // last item with eoff[item] <= e
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)eoff[mid] <= e) lo = mid; else hi = mid - 1;
        }
        const int q = lo / ncol, a = lo - q * ncol, bb = eblo[lo] + (e - (int)eoff[lo]);
        cent[e] = ((unsigned)q << 30) | ((unsigned)a << 15) | (unsigned)bb;
        ctv[e] = f.T[tri(a, bb)];
    }
    pm::sync();
    // F_q(r_q) for the active fractions -> fq[q] (in every thread).  (A variant that kept the few dozen cached pixels
    // in the registers of warp 0 and reduced with shuffles only was slower: two overlap chains per lane, 62k vs 52k cycles.)
    double fq[2] = {0.0, 0.0};
    auto eval_block = [&](const double (&rq)[2], const bool (&act)[2]) {
        double acc[2] = {0.0, 0.0};
        for (int e = pm::tid(); e < ncache; e += kThreads) {
            const unsigned ce = cent[e];
            const int q = (int)(ce >> 30);
            if (!act[q]) continue;
            const double v = ctv[e] * circle_weight((int)((ce >> 15) & 0x7fff), (int)(ce & 0x7fff), rq[q]);
            if (q == 0) acc[0] += v; else acc[1] += v;
        }
        block_sum<2>(sh, acc);
        fq[0] = base[0] + acc[0]; fq[1] = base[1] + acc[1];
    };
    const double fr2[2] = {0.2, 0.8};
    auto drive = [&](auto&& evalf) {       // brentq on both brackets in lock-step (uniform control flow)
        {
            const bool both[2] = {true, true};
            double flo[2], fhi[2];
            evalf(blo_r, both);
            flo[0] = fq[0]; flo[1] = fq[1];
            evalf(bhi_r, both);
            fhi[0] = fq[0]; fhi[1] = fq[1];
            for (int q = 0; q < 2; q++) brent_init(b[q], blo_r[q], bhi_r[q], flo[q] / total - fr2[q], fhi[q] / total - fr2[q]);
        }
        for (;;) {
            bool act[2];
            double rq[2];
            for (int q = 0; q < 2; q++) {
                if (b[q].state == 0) brent_step(b[q]);
                act[q] = b[q].state == 0;
                rq[q] = b[q].xcur;
            }
            if (!act[0] && !act[1]) break;
            evalf(rq, act);
            for (int q = 0; q < 2; q++)
                if (act[q]) b[q].fcur = fq[q] / total - fr2[q];
        }
    };
    drive(eval_block);
    c.ar.release(mk);
    if (b[0].state != 1 || b[1].state != 1) return;
    if (!(fabs(b[0].root) < 1.7e308) || !(fabs(b[1].root) < 1.7e308)) return;
    out->r20 = b[0].root; out->r80 = b[1].root; out->ok = 1;
}

// ====================================================================================================
// Phase 7 — casgm.smoothness (casgm.py:276-406)
// ====================================================================================================
// k x k box means with scipy.ndimage.uniform_filter geometry (window [i - k/2, i - k/2 + k - 1], "reflect"
// boundary of the sub-image [rlo, rlo + rlen) x [clo, clo + clen)) for the pixels of rows [ra, rb) x cols [ca, cb).
// Separable: vertical window sums of a band of rows are staged in `vbuf`, then summed horizontally.
// fn(r, col, mean) is called once per pixel by the thread that owns it.
// Σ_{i<K} (p[i * stride] - sky), loads issued together (compile-time K)
template <int K, typename T_> PM_DEV double strided_sum(const T_* p, int stride, double sky) {
    T_ x[K];
#pragma unroll
    for (int i = 0; i < K; i++) x[i] = p[(size_t)i * stride];      // generic load: global image or shared-memory crop
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) s += (double)x[i] - sky;
    return s;
}

// One band of rows of box_means() without any reflection (the common case), compile-time window size K.  Every warp
// works on TWO rows of the band at once so that two independent load / add chains are in flight per lane.
template <int K, typename T_, typename W_, typename F>
PM_DEV void box_band_fast(const Ctx<T_>& c, int band, int rows, int ca, int W, int E, double* vbuf, double inv, W_& want, F& fn) {
    const int h = K / 2, n = c.g.n;
    const int c_first = ca - h;                                   // image column of vbuf column 0
    const bool in_crop = c.crop != nullptr && c_first >= c.crop_c0 && c_first + E <= c.crop_c0 + c.crop_w &&
                         band - h >= c.crop_r0 && band + rows - 1 - h + K - 1 < c.crop_r0 + c.crop_h;
    const T_* base = in_crop ? c.crop + (size_t)(band - h - c.crop_r0) * c.crop_w + (c_first - c.crop_c0)
                             : c.img + (size_t)(band - h) * n + c_first;
    const int stride = in_crop ? c.crop_w : n;
    for (int q0 = pm::warp(); q0 < rows; q0 += 2 * kWarps) {
        const int q1 = q0 + kWarps;
        const bool two = q1 < rows;
        const T_* p0 = base + (size_t)q0 * stride;
        const T_* p1 = base + (size_t)(two ? q1 : q0) * stride;
        for (int e = pm::lane(); e < E; e += 32) {
            T_ x0[K], x1[K];
#pragma unroll
            for (int i = 0; i < K; i++) { x0[i] = p0[(size_t)i * stride + e]; x1[i] = p1[(size_t)i * stride + e]; }
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < K; i++) { s0 += (double)x0[i] - c.sky; s1 += (double)x1[i] - c.sky; }
            vbuf[q0 * E + e] = s0;
            if (two) vbuf[q1 * E + e] = s1;
        }
    }
    pm::sync();
    for (int q0 = pm::warp(); q0 < rows; q0 += 2 * kWarps) {
        const int q1 = q0 + kWarps;
        const bool two = q1 < rows;
        for (int x = pm::lane(); x < W; x += 32) {
            const bool w0 = want(band + q0, ca + x), w1 = two && want(band + q1, ca + x);
            if (!w0 && !w1) continue;
            const double* v0 = vbuf + q0 * E + x;
            const double* v1 = vbuf + (two ? q1 : q0) * E + x;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < K; i++) { s0 += v0[i]; s1 += v1[i]; }
            if (w0) fn(band + q0, ca + x, s0 * inv);
            if (w1) fn(band + q1, ca + x, s1 * inv);
        }
    }
    pm::sync();
}

template <typename T_, typename W_, typename F>
PM_DEV void box_means(const Ctx<T_>& c, int k, int rlo, int rlen, int clo, int clen, int ra, int rb, int ca, int cb,
                      double* vbuf, int vcap, W_ want, F fn) {
    const int h = k / 2;
    const int W = cb - ca, E = W + k - 1;
    if (W <= 0 || rb <= ra) return;
    const int R = max(1, min(4 * kWarps, vcap / E));
    const double inv = 1.0 / ((double)k * (double)k);
    const bool cols_inside = (ca - h >= clo) && (cb - 1 - h + k - 1 < clo + clen);     // no column reflection needed
    for (int band = ra; band < rb; band += R) {
        const int rows = min(R, rb - band);
        if (cols_inside && k <= 8 && band - h >= rlo && band + rows - 1 - h + k - 1 < rlo + rlen) {
            switch (k) {
                case 1: box_band_fast<1>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 2: box_band_fast<2>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 3: box_band_fast<3>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 4: box_band_fast<4>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 5: box_band_fast<5>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 6: box_band_fast<6>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                case 7: box_band_fast<7>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
                default: box_band_fast<8>(c, band, rows, ca, W, E, vbuf, inv, want, fn); break;
            }
            continue;
        }
        // general band: reflection at the borders of the sub-image and / or a large window
        for (int q = pm::warp(); q < rows; q += kWarps) {
            const int r = band + q;
            for (int e = pm::lane(); e < E; e += 32) {
                const int cc = cols_inside ? ca - h + e : clo + refl(ca - clo - h + e, clen);
                double sum = 0.0;
                int dr = 0;
                for (; dr + 3 < k; dr += 4) {          // four independent loads in flight
                    const double x0 = ld(c, rlo + refl(r - rlo - h + dr, rlen), cc), x1 = ld(c, rlo + refl(r - rlo - h + dr + 1, rlen), cc),
                                 x2 = ld(c, rlo + refl(r - rlo - h + dr + 2, rlen), cc), x3 = ld(c, rlo + refl(r - rlo - h + dr + 3, rlen), cc);
                    sum += x0 - c.sky; sum += x1 - c.sky; sum += x2 - c.sky; sum += x3 - c.sky;
                }
                for (; dr < k; dr++) sum += ld(c, rlo + refl(r - rlo - h + dr, rlen), cc) - c.sky;
                vbuf[q * E + e] = sum;
            }
        }
        pm::sync();
        for (int q = pm::warp(); q < rows; q += kWarps) {
            const int r = band + q;
            for (int x = pm::lane(); x < W; x += 32) {
                if (!want(r, ca + x)) continue;
                const double* v = vbuf + q * E + x;
                double sum = 0.0;
                for (int dc = 0; dc < k; dc++) sum += v[dc];
                fn(r, ca + x, sum * inv);
            }
        }
        pm::sync();
    }
}

template <typename T_>
PM_DEV double phase_smoothness(Ctx<T_>& c, int cx, int cy, double rmax, double r20, double sky_test) {
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    const int k = (int)r20;
    if (k < 1 || !(rmax > r20)) return -99.0;   // uniform_filter(size=0) / CircularAnnulus(r_out <= r_in) raise upstream
    // ---- annulus sums (casgm.py:319-336): weights = overlap(r_out) - overlap(r_in), exact
    const int ixmin = (int)floor((double)cx - rmax + 0.5), ixmax = (int)ceil((double)cx + rmax + 0.5);
    const int iymin = (int)floor((double)cy - rmax + 0.5), iymax = (int)ceil((double)cy + rmax + 0.5);
    const int x0 = max(ixmin, 0), x1 = min(ixmax, g.n), y0 = max(iymin, 0), y1 = min(iymax, g.n);
    Arena::Mark mk = c.ar.mark();
    const int vcap = max(2 * (x1 - x0 + k), 4096);
    double* vbuf = (double*)c.ar.alloc((size_t)vcap * 8);
    const int A = (int)(rmax + 0.5 * sqrt(2.0)) + 1;
    unsigned* eoff = (unsigned*)c.ar.alloc((size_t)(2 * (A + 1) + 1) * 4);
    int* eblo = (int*)c.ar.alloc((size_t)(2 * (A + 1) + 1) * 4);
    if (c.ar.overflow) { c.ar.release(mk); return -99.0; }
    // Pixels cut by neither circle have weight exactly 1 (inside the annulus) or 0: integer thresholds on the squared
    // distance reproduce photutils' float64 predicates  d < r_out - pr  and  d < r_in + pr  (d = sqrt(dx^2 + dy^2)).
    const double pr = 0.5 * sqrt(2.0);
    const int d2_full_out = d2_threshold(rmax - pr), d2_zero_out = d2_threshold(rmax + pr);   // first d2 not fully inside / first outside
    const int d2_full_in = d2_threshold(r20 - pr), d2_zero_in = d2_threshold(r20 + pr);
    double s[2] = {0.0, 0.0};   // Σ w I ,  Σ w max(I - Is, 0)
    // ---- (A) interior of the annulus, weight 1: banded separable box means over the bounding box
    box_means(c, k, 0, g.n, 0, g.n, y0, y1, x0, x1, vbuf, vcap,
              [&](int r, int col) {
                  const int dx = col - cx, dy = r - cy, d2 = dx * dx + dy * dy;
                  return d2 < d2_full_out && d2 >= d2_zero_in;
              },
              [&](int r, int col, double Is) {
                  const double I = ld(c, r, col) - c.sky;
                  const double d = I - Is;
                  s[0] += I;
                  s[1] += d < 0.0 ? 0.0 : d;
              });
    c.pc.lap(13);   // (smoothness: interior pass)
    // ---- (B) pixels cut by either circle: enumerated from the folded octant (centre is an integer pixel), one
    // (entry, mirror image) per thread-step; the box mean is summed directly (k x k loads)
    {
        if (pm::tid() == 0) { sh.bc[8] = rmax; sh.bc[9] = r20; }
        pm::sync();
        const int ncol = A + 1, items = 2 * ncol;
        for (int item = pm::tid(); item < items; item += kThreads) {
            const int q = item / ncol, a = item - q * ncol;
            const double r = sh.bc[8 + q];
            int b_lo = 0, b_hi = 0;
            if (!((double)a > r + 1.0)) fold_range_i(a, q == 0 ? d2_full_out : d2_full_in, q == 0 ? d2_zero_out : d2_zero_in, &b_lo, &b_hi);
            eblo[item] = b_lo;
            eoff[item] = (unsigned)(b_hi - b_lo);
        }
        pm::sync();
        const int total = (int)block_scan_excl_array(sh, eoff, items);
        // B1: one exact weight per cut entry (shared by its mirror images)
        double* wcut = (double*)c.ar.alloc((size_t)max(total, 1) * 8);
        unsigned* ecut = (unsigned*)c.ar.alloc((size_t)max(total, 1) * 4);
        if (c.ar.overflow) { c.ar.release(mk); return -99.0; }
        for (int e = pm::tid(); e < total; e += kThreads) {
            int lo = 0, hi = items - 1;          // last item with eoff[item] <= e
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((int)eoff[mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int q = lo / ncol, a = lo - q * ncol, b = eblo[lo] + (e - (int)eoff[lo]);
            const int d2 = a * a + b * b;
            const bool cut_out = !(d2 < d2_full_out) && d2 < d2_zero_out;
            double w = 0.0;
            if (!(q == 1 && cut_out))                                 // cut by both circles: counted with the outer one
                w = circle_weight(a, b, rmax) - circle_weight(a, b, r20);
            wcut[e] = w;
            ecut[e] = ((unsigned)a << 16) | (unsigned)b;
        }
        pm::sync();
        // B2: the (up to 8) mirror images of every entry, one per thread-step
        for (int e8 = pm::tid(); e8 < total * 8; e8 += kThreads) {
            const int e = e8 >> 3, sym = e8 & 7;
            const double w = wcut[e];
            if (w == 0.0) continue;
            const int a = (int)(ecut[e] >> 16), b = (int)(ecut[e] & 0xffff);
            int dx = (sym & 4) ? b : a, dy = (sym & 4) ? a : b;
            if ((sym & 4) && a == b) continue;
            if ((sym & 1) && dx == 0) continue;
            if ((sym & 2) && dy == 0) continue;
            if (sym & 1) dx = -dx;
            if (sym & 2) dy = -dy;
            const int col = cx + dx, r = cy + dy;
            if (r < y0 || r >= y1 || col < x0 || col >= x1) continue;
            const double I = ld(c, r, col) - c.sky;
            const int h = k / 2;
            double bs = 0.0;
            if (k <= 8 && r - h >= 0 && r - h + k - 1 < g.n && col - h >= 0 && col - h + k - 1 < g.n) {
                const bool in_crop = c.crop != nullptr && r - h >= c.crop_r0 && r - h + k - 1 < c.crop_r0 + c.crop_h &&
                                     col - h >= c.crop_c0 && col - h + k - 1 < c.crop_c0 + c.crop_w;
                for (int dr = 0; dr < k; dr++) {
                    const T_* p = in_crop ? c.crop + (size_t)(r - h + dr - c.crop_r0) * c.crop_w + (col - h - c.crop_c0)
                                          : c.img + (size_t)(r - h + dr) * g.n + (col - h);
                    double rs;
                    switch (k) {
                        case 1: rs = strided_sum<1>(p, 1, c.sky); break;
                        case 2: rs = strided_sum<2>(p, 1, c.sky); break;
                        case 3: rs = strided_sum<3>(p, 1, c.sky); break;
                        case 4: rs = strided_sum<4>(p, 1, c.sky); break;
                        case 5: rs = strided_sum<5>(p, 1, c.sky); break;
                        case 6: rs = strided_sum<6>(p, 1, c.sky); break;
                        case 7: rs = strided_sum<7>(p, 1, c.sky); break;
                        default: rs = strided_sum<8>(p, 1, c.sky); break;
                    }
                    bs += rs;
                }
            } else {
                for (int dr = 0; dr < k; dr++) {
                    const int rr = refl(r - h + dr, g.n);
                    double rs = 0.0;
                    for (int dc = 0; dc < k; dc++) rs += ld(c, rr, refl(col - h + dc, g.n)) - c.sky;
                    bs += rs;
                }
            }
            const double d = I - bs / ((double)k * (double)k);
            s[0] += w * I;
            s[1] += w * (d < 0.0 ? 0.0 : d);
        }
    }
    block_sum<2>(sh, s);
    c.pc.lap(14);   // (smoothness: cut pixels)
    // ---- background box search (casgm.py:367-398); one candidate position per warp per round
    // positions are visited in the order (size 32: y = 0, 2, ..; x = 0, 2, ..), then size 16, then 8.
    int bx = 0, by = 0, bsz = 32;
    // the very first box (size 32 at the origin) is accepted almost always: test it with the whole block first
    bool first_ok = false;
    {
        double bs = 0.0;
        for (int i = pm::tid(); i < 32 * 32; i += kThreads) {
            const int r = i >> 5, col = i & 31;
            if (r < g.n && col < g.n) {
                double v = bit_at(c.seg, g.nw, r, col) ? 0.0 : (ld(c, r, col) - c.sky);
                if (v == 0.0) v = -99.0;
                bs += v;
            }
        }
        bs = block_sum1(sh, bs);
        first_ok = fabs(sky_test) > fabs(bs / 1024.0);
    }
    if (!first_ok) {
        int size = 32;
        for (;;) {
            // candidate grid for this size: x in {0, 2, ..} with x < W - size, y likewise; (0, 0) always
            const int nx = max(1, (g.n - size + 1) / 2), ny = max(1, (g.n - size + 1) / 2);
            const int total = nx * ny;
            int first = 0x7fffffff;
            for (int base = 0; base < total && first == 0x7fffffff; base += kWarps) {
                const int idx = base + pm::warp();
                int ok = 0;
                if (idx < total) {
                    const int y = (idx / nx) * 2, x = (idx % nx) * 2;
                    double bs = 0.0;
                    // size is 32, 16 or 8: a warp covers 32 / size rows per step; loads are independent
                    const int sh_ = size == 32 ? 5 : (size == 16 ? 4 : 3);
                    for (int i0 = 0; i0 < size * size; i0 += 128) {
                        double vv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = i0 + u * 32 + pm::lane();
                            const int r = y + (i >> sh_), col = x + (i & (size - 1));
                            const bool ok_ = i < size * size && r < g.n && col < g.n;
                            const bool inseg = ok_ && bit_at(c.seg, g.nw, r, col);
                            const double raw = (ok_ && !inseg) ? ld(c, r, col) : 0.0;
                            double v = inseg ? 0.0 : raw - c.sky;
                            if (v == 0.0) v = -99.0;
                            vv[u] = ok_ ? v : 0.0;
                        }
                        bs += vv[0]; bs += vv[1]; bs += vv[2]; bs += vv[3];
                    }
                    bs = pm::warp_sum(bs);
                    ok = fabs(sky_test) > fabs(bs / (double)(size * size));
                }
                first = block_min_i(sh, ok ? idx : 0x7fffffff);
            }
            if (first != 0x7fffffff) { by = (first / nx) * 2; bx = (first % nx) * 2; bsz = size; break; }
            if (size > 8) size /= 2;
            else {
                // gave up (casgm.py:397-398): the walk stops with x = 0 and the first y >= H - size
                bsz = size; bx = 0;
                by = 2 * ny;
                break;
            }
        }
    }
    c.pc.lap(15);   // (smoothness: background box search)
    const int rows = max(0, min(bsz, g.n - by)), cols = max(0, min(bsz, g.n - bx));
    double bd = 0.0;
    box_means(c, k, by, rows, bx, cols, by, by + rows, bx, bx + cols, vbuf, vcap,
              [&](int, int) { return true; },
              [&](int r, int col, double Is) {
                  const double d = (ld(c, r, col) - c.sky) - Is;
                  bd += d < 0.0 ? 0.0 : d;
              });
    c.ar.release(mk);
    bd = block_sum1(sh, bd);
    const double bs = bd / (double)(rows * cols);
    const double area = 3.141592653589793 * (rmax * rmax - r20 * r20);
    return (s[1] - area * bs) / s[0];
}

// ====================================================================================================
// The kernel
// ====================================================================================================
PM_DEV void row_defaults(pm_row_t& row) {
    row.apix_col = row.apix_row = row.rmax = row.r20 = row.r80 = row.C = -99.0;
    row.A = row.Abgr = row.S = row.As = row.As90 = row.gini = row.m20 = -99.0;
    row.object_edge = 0.0; row.status = 0.0; row.n_candidates = 0.0;
}

template <typename T> PM_DEV void analyse_one(Ctx<T>& c, const Params& p, long long b) {
    typedef typename KeyT<T>::type K;
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    const size_t plane_elems = (size_t)g.N;
    c.img = (const T*)p.images + (size_t)b * plane_elems;
    c.sky = p.sky ? p.sky[b] : 0.0;
    c.thr = p.ov.threshold_in ? p.ov.threshold_in[b] : c.sky + (p.sky_err ? p.sky_err[b] : 0.0);
    c.flags = p.flags;
    c.ar.reset();
    c.crop = nullptr; c.crop_r0 = c.crop_c0 = 0; c.crop_h = c.crop_w = 0;
    pm_row_t row;
    row_defaults(row);
    PhaseClock& pc = c.pc;
    pc.start(p.phase_cycles ? p.phase_cycles + (size_t)pm::block() * kPhaseSlots : nullptr);
    uint8_t* seg_out = p.out.segmap_out ? p.out.segmap_out + (size_t)b * plane_elems : nullptr;
    int32_t* sidx_out = p.out.sorted_idx_out ? p.out.sorted_idx_out + (size_t)b * plane_elems : nullptr;

    // ---- pixelmap (main.py:421) or supplied segmap (main.py:428-439)
    int status = PM_ST_OK;
    int edge = 0;
    if (p.ov.segmap_in) {
        load_plane_u8(g, p.ov.segmap_in + (size_t)b * plane_elems, c.seg);
    } else {
        status = phase_pixelmap(c);
        if (status != PM_ST_OK) {
            for (int i = pm::tid(); i < g.n * g.nw; i += kThreads) c.seg[i] = 0;
            pm::sync();
        }
    }
    if (seg_out) store_plane_u8(g, c.seg, seg_out);
    pc.lap(0);   // pixelmap (+ segmap store)
    if (status == PM_ST_OK) phase_bbox(c, &edge);
    if (status == PM_ST_OK && !p.ov.segmap_in) row.object_edge = (double)edge;   // main.py:427 (only on the pixelmap path)
    if (p.ov.segmap_in && (p.flags & PM_STOP_AFTER_PIXELMAP)) row.object_edge = (double)edge;
    bool done = status != PM_ST_OK || (p.flags & PM_STOP_AFTER_PIXELMAP);
    if (!done && c.r1 < 0) { status = PM_ST_EMPTY_SEGMAP; done = true; }

    if (!done) {
        // ---- raster list + keys, statistics
        unsigned ng = 0;
        for (int r = pm::tid(); r < g.n; r += kThreads) {   // n_g first (for the allocations)
            const unsigned cnt = row_popcount(c.seg, g.nw, r);
            c.rowoff[r] = cnt;            // stays there for phase_list
            ng += cnt;
        }
        c.n_g = (int)block_sum_u(sh, ng);
        const int n_g = c.n_g;
        // arena plan: [raster list][sorted positions -> candidates][sort temporaries], then — once the sort
        // temporaries are released — the masked window (so that both the sort and minapix run out of shared memory)
        c.bw = c.c1 - c.c0 + 1;
        const int bh = c.r1 - c.r0 + 1;
        c.mw = nullptr;
        c.posR = (unsigned*)c.ar.alloc((size_t)n_g * 4);
        Arena::Mark mk_list = c.ar.mark();
        unsigned* spA = (unsigned*)c.ar.alloc((size_t)n_g * 4);     // sorted positions end up here
        Arena::Mark mk_sorted = c.ar.mark();
        K* skA = (K*)c.ar.alloc((size_t)n_g * sizeof(K));
        Arena::Mark mk_keys = c.ar.mark();
        K* skB = (K*)c.ar.alloc((size_t)n_g * sizeof(K));
        unsigned* spB = (unsigned*)c.ar.alloc((size_t)n_g * 4);
        // 9-bit digits (3 passes for typical float keys) when their histogram fits shared memory, else 8-bit digits
        int sort_db = kMaxDigitBits;
        Arena::Mark mk_hist = c.ar.mark();
        unsigned* hist = (unsigned*)c.ar.alloc((size_t)max(kWarps * kMaxDigits, kBuckets + 2) * 4);
        if (!c.ar.in_smem(hist)) {
            c.ar.release(mk_hist);
            sort_db = 8;
            hist = (unsigned*)c.ar.alloc((size_t)max(kWarps * 256, kBuckets + 2) * 4);
        }
        const unsigned* cand = spA;
        if (c.ar.overflow) { status = PM_ST_WORKSPACE; done = true; }   // workspace contract violated
        SegStats st;
        st.sum = st.m10 = st.m01 = st.vmax = st.abs_sum = 0.0; st.argmax = 0; st.edge = edge;
        int cen_flat = 0, rmax2 = 0;
        double rmax = 0.0;
        SortOut so;
        so.gini = so.m20 = -99.0; so.n_cand = 0; so.n_pos = 0;
        if (!done) {
            phase_list(c, skA, &st);
            pc.lap(1);   // bbox, counts, window, raster list
            rmax2 = phase_rmax2(c, st, &cen_flat);
            pc.lap(40);
            rmax = p.ov.radius_in ? p.ov.radius_in[b] : pm::dsqrt((double)rmax2);
            row.rmax = rmax;
            // ---- aperture plane (main.py:464) — uses the key buffers' neighbours as scratch before the sort
            if (p.ov.apermask_in) {
                load_plane_u8(g, p.ov.apermask_in + (size_t)b * plane_elems, c.aper);
            } else {
                Arena::Mark mk = c.ar.mark();
                double* sq = (double*)c.ar.alloc((size_t)(g.half + 1) * 9 * 8);
                int* hw = (int*)c.ar.alloc((size_t)(g.half + 1) * 4);
                build_aperture(g, rmax, 9, 0.1, c.aper, sq, hw);
                c.ar.release(mk);
            }
            if (p.out.apermask_out) store_plane_u8(g, c.aper, p.out.apermask_out + (size_t)b * plane_elems);
            pc.lap(2);   // rmax + aperture
            if (p.flags & PM_STOP_AFTER_RMAX) done = true;
        }
        if (!done) {
            // ---- sort ascending by (value, flat index); the reverse is the canonical descending order
            for (int i = pm::tid(); i < n_g; i += kThreads) spA[i] = c.posR[i];
            pm::sync();
            int where = (p.flags & PM_RADIX_SORT) ? -1 : bucket_sort_pairs<K>(sh, skA, spA, skB, spB, hist, n_g, [&](int ph) { pc.lap(ph); });
            if (where < 0) where = radix_sort_pairs<K>(sh, skA, spA, skB, spB, hist, n_g, sort_db, [&](int ph) { pc.lap(ph); });
            if (where) {
                for (int i = pm::tid(); i < n_g; i += kThreads) { spA[i] = spB[i]; skA[i] = skB[i]; }
                pm::sync();
            }
            c.ar.release(mk_keys);    // drop skB, spB, hist
            pc.lap(3);   // sort
            phase_gini_m20_cands(c, skA, spA, st, &so);
            row.n_candidates = (double)so.n_cand;
            // reverse in place: the canonical DESCENDING order (value desc, flat index desc), asymmetry.py:94-95
            for (int i = pm::tid(); i < n_g / 2; i += kThreads) {
                const unsigned a = spA[i], z = spA[n_g - 1 - i];
                spA[i] = z; spA[n_g - 1 - i] = a;
            }
            pm::sync();
            if (sidx_out) {
                for (int i = pm::tid(); i < g.N; i += kThreads) {
                    int v = -1;
                    if (i < so.n_pos) { const unsigned q = spA[i]; v = (int)(q >> 16) * g.n + (int)(q & 0xffff); }
                    sidx_out[i] = v;
                }
            }
            if (p.out.n_positive_out && pm::tid() == 0) p.out.n_positive_out[b] = so.n_pos;
            pm::sync();
            pc.lap(36);
            // keep only the candidates (the first n_cand entries of the descending order)
            c.ar.release(mk_list);
            c.ar.keep(spA, (size_t)max(so.n_cand, 1) * 4);
            // ---- masked window over the bounding box (NaN outside the map)
            c.mw = (T*)c.ar.alloc((size_t)c.bw * bh * sizeof(T));
            if (c.ar.overflow) { status = PM_ST_WORKSPACE; done = true; }
            else {
                for (int wr = pm::warp(); wr < bh; wr += kWarps) {        // one window row per warp, loads independent
                    const int r = c.r0 + wr;
                    const T* src = c.img + (size_t)r * g.n + c.c0;
                    for (int x = pm::lane(); x < c.bw; x += 32)
                        c.mw[wr * c.bw + x] = bit_at(c.seg, g.nw, r, c.c0 + x) ? pm::ldg(src + x) : nan_of(T());
                }
                pm::sync();
            }
            pc.lap(4);   // gini, m20, candidates
            (void)mk_sorted;
        }
        int cx = 0, cy = 0;
        if (!done) {
            // ---- minapix (main.py:467) or supplied centroid
            if (p.ov.centroid_in) {
                cx = (int)p.ov.centroid_in[2 * b]; cy = (int)p.ov.centroid_in[2 * b + 1];
            } else if (so.n_cand == 0) {
                status = PM_ST_NO_CANDIDATE; done = true;
            } else {
                Arena::Mark mk = c.ar.mark();
                const int groups = (so.n_cand + kCandBlock - 1) / kCandBlock;
                int best = 0;
                const bool prune = p.out.cand_a_out == nullptr && so.n_cand > 2 * kCandBlock && !(p.flags & PM_NO_PRUNE);
                c.plist = c.posR;      // (a brightness-ordered list prunes earlier but loses locality: measured slower)
                if (prune) {
                    double* acc = (double*)c.ar.alloc((size_t)so.n_cand * 16);
                    int* alive = (int*)c.ar.alloc((size_t)so.n_cand * 4);
                    int* alive2 = (int*)c.ar.alloc((size_t)so.n_cand * 4);
                    unsigned* flag = (unsigned*)c.ar.alloc((size_t)(so.n_cand + 1) * 4);
                    double* parts = (double*)c.ar.alloc((size_t)(kWarps + groups) * 64);
                    if (c.ar.overflow) { status = PM_ST_WORKSPACE; done = true; }
                    else phase_minapix_pruned(c, cand, so.n_cand, st.abs_sum, acc, alive, alive2, flag, parts, &best);
                } else {
                    int slices = (3 * kWarps + groups - 1) / groups;       // aim at >= 3 work items per warp
                    slices = max(1, min(slices, (n_g + 63) / 64));
                    double* a_arr = (double*)c.ar.alloc((size_t)so.n_cand * 8);
                    double* parts = (double*)c.ar.alloc((size_t)groups * slices * kCandBlock * 16);
                    if (c.ar.overflow) { status = PM_ST_WORKSPACE; done = true; }
                    else {
                        phase_minapix(c, cand, so.n_cand, a_arr, parts, slices, &best);
                        if (p.out.cand_a_out) {
                            for (int i = pm::tid(); i < p.out.cand_cap; i += kThreads)
                                p.out.cand_a_out[(size_t)b * p.out.cand_cap + i] = i < so.n_cand ? a_arr[i] : -99.0;
                        }
                    }
                }
                if (!done) {
                    const unsigned q = cand[best];
                    cy = (int)(q >> 16); cx = (int)(q & 0xffff);
                }
                pm::sync();
                c.ar.release(mk);
            }
        }
        if (!done) {
            pc.lap(5);   // minapix
            row.apix_col = (double)cx; row.apix_row = (double)cy;
            if (p.flags & PM_DO_CAS) { row.gini = so.gini; row.m20 = so.m20; }   // (not reached without a candidate)
            // ---- r20, r80, C (main.py:491-492) — before calcA: it is the last user of the masked window
            if (p.flags & PM_DO_CAS) {
                RadiiOut ro;
                phase_r20_r80(c, cx, cy, rmax, &ro);
                if (ro.ok) {
                    row.r20 = ro.r20; row.r80 = ro.r80;
                    row.C = 5.0 * log10(ro.r80 / ro.r20);
                } else {
                    status = PM_ST_BAD_BRACKET;
                }
            }
            pc.lap(7);   // r20 / r80
            // ---- the arena is free now: stage the central part of the image (what calcA and smoothness touch:
            // the aperture about the image centre, its mirror image about apix, the annulus about apix) in shared memory
            c.ar.reset();
            c.posR = nullptr; c.mw = nullptr;
            if (p.flags & (PM_DO_A | PM_DO_CAS | PM_DO_S)) {
                // bounding box of {image centre, apix, mirror of the centre about apix} grown by rmax + margin
                const int reach = (int)fmin(rmax, 4096.0) + 12;
                int lo_r = max(0, min(min(g.half, cy), 2 * cy - g.half) - reach), hi_r = min(g.n - 1, max(max(g.half, cy), 2 * cy - g.half) + reach);
                int lo_c = max(0, min(min(g.half, cx), 2 * cx - g.half) - reach), hi_c = min(g.n - 1, max(max(g.half, cx), 2 * cx - g.half) + reach);
                // leave room for the scratch of calcA / smoothness (bit plane, task list, band buffers)
                const long long room = (long long)c.ar.s_cap - (long long)(g.n * g.nw * 12 + 24 * 1024);
                while ((long long)(hi_r - lo_r + 1) * (hi_c - lo_c + 1) * (long long)sizeof(T) > room && hi_r - lo_r > 16 && hi_c - lo_c > 16) {
                    lo_r++; hi_r--; lo_c++; hi_c--;
                }
                const int ch = hi_r - lo_r + 1, cw = hi_c - lo_c + 1;
                if (ch > 0 && cw > 0 && (long long)ch * cw * (long long)sizeof(T) <= room) {
                    T* cb = (T*)c.ar.alloc((size_t)ch * cw * sizeof(T));
                    for (int wr = pm::warp(); wr < ch; wr += kWarps) {
                        const T* src = c.img + (size_t)(lo_r + wr) * g.n + lo_c;
                        int x = pm::lane();
                        for (; x + 96 < cw; x += 128) {       // four independent loads in flight
                            const T a0 = pm::ldg(src + x), a1 = pm::ldg(src + x + 32), a2 = pm::ldg(src + x + 64), a3 = pm::ldg(src + x + 96);
                            cb[wr * cw + x] = a0; cb[wr * cw + x + 32] = a1; cb[wr * cw + x + 64] = a2; cb[wr * cw + x + 96] = a3;
                        }
                        for (; x < cw; x += 32) cb[wr * cw + x] = pm::ldg(src + x);
                    }
                    pm::sync();
                    c.crop = cb; c.crop_r0 = lo_r; c.crop_c0 = lo_c; c.crop_h = ch; c.crop_w = cw;
                }
            }
            // ---- calcA x3 (main.py:470-478)
            if (p.flags & (PM_DO_A | PM_DO_AS | PM_DO_CAS)) {
                Arena::Mark mk = c.ar.mark();
                unsigned* tmp = (unsigned*)c.ar.alloc((size_t)g.n * g.nw * 4);
                AsymOut ao;
                phase_calcA(c, cx, cy, &ao, tmp);
                c.ar.release(mk);
                if (p.flags & (PM_DO_A | PM_DO_CAS)) { row.A = ao.A; row.Abgr = ao.Abgr; }
                if (p.flags & PM_DO_AS) { row.As = ao.As; row.As90 = ao.As90; }
            }
            pc.lap(6);   // calcA x3
            // ---- smoothness (casgm.py:276; the call is commented out at main.py:494)
            if (p.flags & PM_DO_S) {
                const double r20 = p.ov.r20_in ? p.ov.r20_in[b] : row.r20;
                const double skyt = p.ov.smooth_sky_in ? p.ov.smooth_sky_in[b] : c.sky;
                if (status == PM_ST_OK && (p.ov.r20_in || (p.flags & PM_DO_CAS)))
                    row.S = phase_smoothness(c, cx, cy, rmax, r20, skyt);
            }
        }
    }
    pc.lap(8);   // smoothness (or whatever ended the cutout early)
    row.status = (double)status;
    if (pm::tid() == 0) p.rows[b] = row;
    pm::sync();
}

// group setup shared by all kernels: the fixed shared-memory pieces named by `mask` (the others come out of
// the arena, i.e. global scratch when shared memory is short), the arena, the tables
template <typename T> PM_DEV void setup_ctx(Ctx<T>& c, const Params& p, unsigned mask, bool optional, unsigned long long scratch_per_group) {
    unsigned char* smem = pm::dyn_smem(p.smem_bytes);
    c.g = make_geo(p.n, p.fs);
    const SmemMap sm = make_smem_map(c.g, p.smem_bytes, mask, optional);
    c.sh = (Sh*)(smem + sm.sh);
    c.ar.init(smem + sm.arena, sm.arena_bytes, p.scratch + (size_t)pm::block() * scratch_per_group, (size_t)scratch_per_group);
    const size_t plane = (size_t)c.g.n * c.g.nw * 4;
    auto piece = [&](unsigned off, bool wanted, size_t bytes) -> void* {
        if (!wanted) return nullptr;
        return off != kNoSmem ? (void*)(smem + off) : c.ar.alloc(bytes);
    };
    c.rowoff = (unsigned*)piece(sm.rowoff, mask & SM_ROWOFF, (size_t)(c.g.n + 1) * 4);
    c.seg = (unsigned*)piece(sm.seg, mask & SM_SEG, plane);
    c.aper = (unsigned*)piece(sm.aper, mask & SM_APER, plane);
    c.dil = (unsigned*)piece(sm.dil, mask & SM_DIL, plane);
    c.tw = (mask & SM_TABLES) ? (double*)(smem + sm.tw) : nullptr;
    c.i0 = (mask & SM_TABLES) ? (uint16_t*)(smem + sm.i0) : nullptr;
    c.ri = (mask & SM_TABLES) ? (uint16_t*)(smem + sm.ri) : nullptr;
    c.rstart = (mask & SM_TABLES) ? (uint16_t*)(smem + sm.rstart) : nullptr;
    c.ar.set_floor();
    c.n_g = 0; c.r0 = c.c0 = 0; c.r1 = c.c1 = -1; c.posR = nullptr; c.plist = nullptr; c.mw = nullptr; c.bw = 0;
    c.crop = nullptr; c.crop_r0 = c.crop_c0 = 0; c.crop_h = c.crop_w = 0;
    c.mbar_phase[0] = c.mbar_phase[1] = 0u;
    c.sh->par[pm::tid()] = 0;
    if (mask & SM_TABLES) {
        if (pm::tid() == 0) { pm::mbar_init(&c.sh->mbar[0], 1); pm::mbar_init(&c.sh->mbar[1], 1); pm::mbar_fence_init(); }
        build_tables(c);
    }
    pm::sync();
}
template <typename T> PM_DEV void analyse_body(const Params& p) {
    Ctx<T> c;
    setup_ctx(c, p, SM_ALL, false, p.scratch_per_cta);
    // Persistent CTA: galaxy indices come from a global work counter, fetched ONE AHEAD so that the next cutout
    // can optionally be pulled towards L2 (TMA bulk prefetch, PM_L2_PREFETCH) while the current one is analysed.
    // Off by default: with the TMA-staged pixelmap it no longer pays (measured -1 %: it competes for L2 capacity).
    const size_t img_bytes = (size_t)c.g.N * sizeof(T);
    if (pm::tid() == 0) c.sh->ia[0] = (int)pm::atomic_add(p.counter, 1ull);
    pm::sync();
    long long cur = (long long)c.sh->ia[0];
    pm::sync();
    while (cur < p.B) {
        if (pm::tid() == 0) c.sh->ia[0] = (int)pm::atomic_add(p.counter, 1ull);
        pm::sync();
        const long long nxt = (long long)c.sh->ia[0];
        pm::sync();
        if (nxt < p.B && (p.flags & PM_L2_PREFETCH)) {
            const uintptr_t base = (uintptr_t)p.images + (size_t)nxt * img_bytes;
            const uintptr_t lo = (base + 15) & ~(uintptr_t)15, hi = (base + img_bytes) & ~(uintptr_t)15;
            for (uintptr_t a = lo + (uintptr_t)pm::tid() * 4096; a < hi; a += (uintptr_t)kThreads * 4096)
                pm::prefetch_l2((const void*)a, (unsigned)min((uintptr_t)4096, hi - a));
        }
        analyse_one(c, p, cur);
        cur = nxt;
    }
}

// ====================================================================================================
// The split path: the same phases as analyse_one(), cut into three kernels by resource class so that every kernel
// runs at the occupancy its phases allow (and fits the instruction cache):
//   front   (one CTA per cutout)   pixelmap -> bbox -> raster list / keys -> rmax, aperture -> sort -> gini, m20, candidates
//   minapix (one CTA per cutout, more CTAs per SM: the loop needs ~64 registers and window + list of shared memory)
//   tail    (one WARP per cutout in pm_tail.cu, or one CTA)   r20/r80, calcA x3, smoothness -> result row
// The hand-over goes through GalState slots in the workspace.  Cutouts that do not fit a slot take analyse_one().
// ====================================================================================================
PM_DEV GalState* state_of(const Params& p, long long i) { return (GalState*)(p.state + (size_t)i * p.state_stride); }
PM_DEV unsigned* state_seg(GalState* gs) { return (unsigned*)((unsigned char*)gs + 256); }
PM_DEV unsigned* state_aper(GalState* gs, const Geo& g) { return state_seg(gs) + g.n * g.nw; }
PM_DEV unsigned* state_list(GalState* gs, const Geo& g) { return state_seg(gs) + 2 * g.n * g.nw; }

// hand a cutout to the monolith pass
PM_DEV void push_big(const Params& p, long long i, GalState* gs) {
    if (pm::tid() == 0) {
        const unsigned long long k = pm::atomic_add(p.counter + 4, 1ull);
        p.big_list[k] = (int)i;
        gs->done = 2;
    }
}

template <typename T> PM_DEV void front_one(Ctx<T>& c, const Params& p, long long i) {
    typedef typename KeyT<T>::type K;
    const Geo& g = c.g;
    Sh& sh = *c.sh;
    const long long b = p.b0 + i;
    GalState* gs = state_of(p, i);
    const size_t plane_elems = (size_t)g.N;
    c.img = (const T*)p.images + (size_t)b * plane_elems;
    c.sky = p.sky ? p.sky[b] : 0.0;
    c.thr = p.ov.threshold_in ? p.ov.threshold_in[b] : c.sky + (p.sky_err ? p.sky_err[b] : 0.0);
    c.flags = p.flags;
    c.ar.reset();
    c.crop = nullptr; c.crop_r0 = c.crop_c0 = 0; c.crop_h = c.crop_w = 0;
    c.mw = nullptr;
    pm_row_t row;
    row_defaults(row);
    PhaseClock& pc = c.pc;
    pc.start(p.phase_cycles ? p.phase_cycles + (size_t)pm::block() * kPhaseSlots : nullptr);
    uint8_t* seg_out = p.out.segmap_out ? p.out.segmap_out + (size_t)b * plane_elems : nullptr;
    int32_t* sidx_out = p.out.sorted_idx_out ? p.out.sorted_idx_out + (size_t)b * plane_elems : nullptr;

    // ---- pixelmap (main.py:421) or supplied segmap (main.py:428-439)
    int status = PM_ST_OK;
    int edge = 0;
    if (p.ov.segmap_in) {
        load_plane_u8(g, p.ov.segmap_in + (size_t)b * plane_elems, c.seg);
    } else {
        status = phase_pixelmap(c);
        if (status != PM_ST_OK) {
            for (int k = pm::tid(); k < g.n * g.nw; k += kThreads) c.seg[k] = 0;
            pm::sync();
        }
    }
    if (seg_out) store_plane_u8(g, c.seg, seg_out);
    pc.lap(0);   // pixelmap (+ segmap store)
    if (status == PM_ST_OK) phase_bbox(c, &edge);
    if (status == PM_ST_OK && !p.ov.segmap_in) row.object_edge = (double)edge;   // main.py:427 (only on the pixelmap path)
    if (p.ov.segmap_in && (p.flags & PM_STOP_AFTER_PIXELMAP)) row.object_edge = (double)edge;
    bool done = status != PM_ST_OK || (p.flags & PM_STOP_AFTER_PIXELMAP);
    if (!done && c.r1 < 0) { status = PM_ST_EMPTY_SEGMAP; done = true; }
    double rmax = 0.0, gini = -99.0, m20 = -99.0, abs_sum = 0.0;
    int n_cand = 0;
    if (!done) {
        unsigned ng = 0;
        for (int r = pm::tid(); r < g.n; r += kThreads) {   // n_g first (for the allocations)
            const unsigned cnt = row_popcount(c.seg, g.nw, r);
            c.rowoff[r] = cnt;            // stays there for phase_list
            ng += cnt;
        }
        c.n_g = (int)block_sum_u(sh, ng);
        const int n_g = c.n_g;
        if (n_g + p.cand_cap > p.list_cap) { push_big(p, i, gs); return; }     // lists would not fit the slot
        c.bw = c.c1 - c.c0 + 1;
        c.posR = (unsigned*)c.ar.alloc((size_t)n_g * 4);
        unsigned* spA = (unsigned*)c.ar.alloc((size_t)n_g * 4);     // sorted positions end up here
        K* skA = (K*)c.ar.alloc((size_t)n_g * sizeof(K));
        Arena::Mark mk_keys = c.ar.mark();
        K* skB = (K*)c.ar.alloc((size_t)n_g * sizeof(K));
        unsigned* spB = (unsigned*)c.ar.alloc((size_t)n_g * 4);
        int sort_db = kMaxDigitBits;
        Arena::Mark mk_hist = c.ar.mark();
        unsigned* hist = (unsigned*)c.ar.alloc((size_t)max(kWarps * kMaxDigits, kBuckets + 2) * 4);
        if (!c.ar.in_smem(hist)) {
            c.ar.release(mk_hist);
            sort_db = 8;
            hist = (unsigned*)c.ar.alloc((size_t)max(kWarps * 256, kBuckets + 2) * 4);
        }
        if (c.ar.overflow) { status = PM_ST_WORKSPACE; done = true; }
        SegStats st;
        st.sum = st.m10 = st.m01 = st.vmax = st.abs_sum = 0.0; st.argmax = 0; st.edge = edge;
        SortOut so;
        so.gini = so.m20 = -99.0; so.n_cand = 0; so.n_pos = 0;
        if (!done) {
            phase_list(c, skA, &st);
            pc.lap(1);   // bbox, counts, raster list
            int cen_flat = 0;
            const int rmax2 = phase_rmax2(c, st, &cen_flat);
            pc.lap(40);
            rmax = p.ov.radius_in ? p.ov.radius_in[b] : pm::dsqrt((double)rmax2);
            // the warp-group tail sizes its folded radial table by the radius: very large ones take the monolith pass
            if ((p.flags & PM_TAIL_WARP) && !(rmax <= 0.5 * (double)g.n)) { push_big(p, i, gs); return; }
            row.rmax = rmax;
            abs_sum = st.abs_sum;
            if (p.ov.apermask_in) {
                load_plane_u8(g, p.ov.apermask_in + (size_t)b * plane_elems, c.aper);
            } else {
                Arena::Mark mk = c.ar.mark();
                double* sq = (double*)c.ar.alloc((size_t)(g.half + 1) * 9 * 8);
                int* hw = (int*)c.ar.alloc((size_t)(g.half + 1) * 4);
                build_aperture(g, rmax, 9, 0.1, c.aper, sq, hw);
                c.ar.release(mk);
            }
            if (p.out.apermask_out) store_plane_u8(g, c.aper, p.out.apermask_out + (size_t)b * plane_elems);
            pc.lap(2);   // rmax + aperture
            if (p.flags & PM_STOP_AFTER_RMAX) done = true;
        }
        if (!done) {
            unsigned* st_list = state_list(gs, g);
            for (int k = pm::tid(); k < n_g; k += kThreads) { const unsigned q = c.posR[k]; spA[k] = q; st_list[k] = q; }
            pm::sync();
            int where = (p.flags & PM_RADIX_SORT) ? -1 : bucket_sort_pairs<K>(sh, skA, spA, skB, spB, hist, n_g, [&](int ph) { pc.lap(ph); });
            if (where < 0) where = radix_sort_pairs<K>(sh, skA, spA, skB, spB, hist, n_g, sort_db, [&](int ph) { pc.lap(ph); });
            if (where) {
                for (int k = pm::tid(); k < n_g; k += kThreads) { spA[k] = spB[k]; skA[k] = skB[k]; }
                pm::sync();
            }
            c.ar.release(mk_keys);    // drop skB, spB, hist
            pc.lap(3);   // sort
            phase_gini_m20_cands(c, skA, spA, st, &so);
            row.n_candidates = (double)so.n_cand;
            n_cand = so.n_cand;
            gini = so.gini; m20 = so.m20;
            if (n_cand > p.cand_cap) { push_big(p, i, gs); return; }
            // the canonical DESCENDING order (value desc, flat index desc), asymmetry.py:94-95: the candidates are its head
            for (int k = pm::tid(); k < n_cand; k += kThreads) st_list[n_g + k] = spA[n_g - 1 - k];
            if (sidx_out) {
                for (int k = pm::tid(); k < g.N; k += kThreads) {
                    int v = -1;
                    if (k < so.n_pos) { const unsigned q = spA[n_g - 1 - k]; v = (int)(q >> 16) * g.n + (int)(q & 0xffff); }
                    sidx_out[k] = v;
                }
            }
            if (p.out.n_positive_out && pm::tid() == 0) p.out.n_positive_out[b] = so.n_pos;
            pc.lap(4);   // gini, m20, candidates
        }
    }
    // ---- hand-over
    if (!done) {
        unsigned* ss = state_seg(gs);
        unsigned* sa = state_aper(gs, g);
        for (int k = pm::tid(); k < g.n * g.nw; k += kThreads) { ss[k] = c.seg[k]; sa[k] = c.aper[k]; }
    }
    if (pm::tid() == 0) {
        gs->row = row;
        gs->sky = c.sky; gs->thr = c.thr; gs->rmax = rmax; gs->abs_sum = abs_sum; gs->gini = gini; gs->m20 = m20;
        gs->status = status; gs->done = done ? 1 : 0;
        gs->n_g = c.n_g; gs->r0 = c.r0; gs->r1 = c.r1; gs->c0 = c.c0; gs->c1 = c.c1;
        gs->n_cand = n_cand; gs->cx = 0; gs->cy = 0;
    }
    pm::sync();
}

// image over the map's bounding box with NaN outside the map (the window minapix pairs pixels through), from the image
// and the map's bit plane `seg` (any address space)
template <typename T> PM_DEV void build_window(const Ctx<T>& c, const unsigned* seg, T* mw) {
    const Geo& g = c.g;
    const int bh = c.r1 - c.r0 + 1;
    for (int wr = pm::warp(); wr < bh; wr += kWarps) {        // one window row per warp, loads independent
        const int r = c.r0 + wr;
        const T* src = c.img + (size_t)r * g.n + c.c0;
        for (int x = pm::lane(); x < c.bw; x += 32)
            mw[wr * c.bw + x] = bit_at(seg, g.nw, r, c.c0 + x) ? pm::ldg(src + x) : nan_of(T());
    }
}

template <typename T> PM_DEV void minapix_one(Ctx<T>& c, const Params& p, long long i) {
    const Geo& g = c.g;
    const long long b = p.b0 + i;
    GalState* gs = state_of(p, i);
    if (gs->done) return;
    PhaseClock& pc = c.pc;
    pc.start(p.phase_cycles ? p.phase_cycles + (size_t)pm::block() * kPhaseSlots : nullptr);
    const int n_cand = gs->n_cand;
    if (n_cand == 0) {                                   // np.argmin([]) upstream (asymmetry.py:127)
        pm::sync();
        if (pm::tid() == 0) { gs->status = PM_ST_NO_CANDIDATE; gs->done = 1; }
        pm::sync();
        return;
    }
    c.img = (const T*)p.images + (size_t)b * (size_t)g.N;
    c.sky = gs->sky; c.flags = p.flags;
    c.n_g = gs->n_g; c.r0 = gs->r0; c.r1 = gs->r1; c.c0 = gs->c0; c.c1 = gs->c1;
    c.bw = c.c1 - c.c0 + 1;
    const int bh = c.r1 - c.r0 + 1, n_g = c.n_g;
    const double abs_sum = gs->abs_sum;
    c.ar.reset();
    c.crop = nullptr;
    const unsigned* st_seg = state_seg(gs);
    const unsigned* st_aper = state_aper(gs, g);
    const unsigned* st_list = state_list(gs, g);
    for (int k = pm::tid(); k < g.n * g.nw; k += kThreads) c.aper[k] = st_aper[k];
    // the window and the list in shared memory when they fit (else the window in global scratch, the list in place)
    T* mw = (T*)c.ar.alloc((size_t)c.bw * bh * sizeof(T));
    unsigned* plist = (unsigned*)c.ar.alloc_smem((size_t)n_g * 4);
    unsigned* cand = (unsigned*)c.ar.alloc((size_t)n_cand * 4);
    int status = PM_ST_OK;
    if (c.ar.overflow) status = PM_ST_WORKSPACE;
    int best = 0;
    if (status == PM_ST_OK) {
        build_window(c, st_seg, mw);
        if (plist) { for (int k = pm::tid(); k < n_g; k += kThreads) plist[k] = st_list[k]; }
        for (int k = pm::tid(); k < n_cand; k += kThreads) cand[k] = st_list[n_g + k];
        pm::sync();
        c.mw = mw;
        c.posR = plist ? plist : const_cast<unsigned*>(st_list);
        c.plist = c.posR;
        pc.lap(41);   // (minapix kernel: window + lists)
        const int groups = (n_cand + kCandBlock - 1) / kCandBlock;
        const bool prune = p.out.cand_a_out == nullptr && n_cand > 2 * kCandBlock && !(p.flags & PM_NO_PRUNE);
        if (prune) {
            double* acc = (double*)c.ar.alloc((size_t)n_cand * 16);
            int* alive = (int*)c.ar.alloc((size_t)n_cand * 4);
            int* alive2 = (int*)c.ar.alloc((size_t)n_cand * 4);
            unsigned* flag = (unsigned*)c.ar.alloc((size_t)(n_cand + 1) * 4);
            double* parts = (double*)c.ar.alloc((size_t)(kWarps + groups) * 64);
            if (c.ar.overflow) status = PM_ST_WORKSPACE;
            else phase_minapix_pruned(c, cand, n_cand, abs_sum, acc, alive, alive2, flag, parts, &best);
        } else {
            int slices = (3 * kWarps + groups - 1) / groups;       // aim at >= 3 work items per warp
            slices = max(1, min(slices, (n_g + 63) / 64));
            double* a_arr = (double*)c.ar.alloc((size_t)n_cand * 8);
            double* parts = (double*)c.ar.alloc((size_t)groups * slices * kCandBlock * 16);
            if (c.ar.overflow) status = PM_ST_WORKSPACE;
            else {
                phase_minapix(c, cand, n_cand, a_arr, parts, slices, &best);
                if (p.out.cand_a_out) {
                    for (int k = pm::tid(); k < p.out.cand_cap; k += kThreads)
                        p.out.cand_a_out[(size_t)b * p.out.cand_cap + k] = k < n_cand ? a_arr[k] : -99.0;
                }
            }
        }
    }
    pm::sync();
    if (pm::tid() == 0) {
        if (status != PM_ST_OK) { gs->status = status; gs->done = 1; }
        else { const unsigned q = cand[best]; gs->cy = (int)(q >> 16); gs->cx = (int)(q & 0xffff); }
    }
    pc.lap(5);   // minapix
    pm::sync();
    c.mw = nullptr; c.posR = nullptr; c.plist = nullptr;
}

template <typename T> PM_DEV void tail_one(Ctx<T>& c, const Params& p, long long i) {
    const Geo& g = c.g;
    const long long b = p.b0 + i;
    GalState* gs = state_of(p, i);
    const int done_in = gs->done;
    if (done_in == 2) return;                            // the monolith pass writes this row
    pm_row_t row = gs->row;
    int status = gs->status;
    PhaseClock& pc = c.pc;
    pc.start(p.phase_cycles ? p.phase_cycles + (size_t)pm::block() * kPhaseSlots : nullptr);
    if (!done_in) {
        c.img = (const T*)p.images + (size_t)b * (size_t)g.N;
        c.sky = gs->sky; c.flags = p.flags;
        c.n_g = gs->n_g; c.r0 = gs->r0; c.r1 = gs->r1; c.c0 = gs->c0; c.c1 = gs->c1;
        c.bw = c.c1 - c.c0 + 1;
        c.mw = nullptr; c.posR = nullptr; c.plist = nullptr;
        c.crop = nullptr; c.crop_r0 = c.crop_c0 = 0; c.crop_h = c.crop_w = 0;
        c.ar.reset();
        const double rmax = gs->rmax;
        int cx, cy;
        if (p.ov.centroid_in) { cx = (int)p.ov.centroid_in[2 * b]; cy = (int)p.ov.centroid_in[2 * b + 1]; }
        else { cx = gs->cx; cy = gs->cy; }
        const unsigned* st_seg = state_seg(gs);
        const unsigned* st_aper = state_aper(gs, g);
        for (int k = pm::tid(); k < g.n * g.nw; k += kThreads) { c.seg[k] = st_seg[k]; c.aper[k] = st_aper[k]; }
        pm::sync();
        pc.lap(42);  // (tail kernel: planes)
        row.apix_col = (double)cx; row.apix_row = (double)cy;
        if (p.flags & PM_DO_CAS) { row.gini = gs->gini; row.m20 = gs->m20; }
        // ---- r20, r80, C (main.py:491-492)
        if (p.flags & PM_DO_CAS) {
            RadiiOut ro;
            phase_r20_r80(c, cx, cy, rmax, &ro);
            if (ro.ok) {
                row.r20 = ro.r20; row.r80 = ro.r80;
                row.C = 5.0 * log10(ro.r80 / ro.r20);
            } else {
                status = PM_ST_BAD_BRACKET;
            }
        }
        pc.lap(7);   // r20 / r80
        c.ar.reset();
        // ---- CTA groups: stage the central part of the image (what calcA and smoothness touch) in shared memory
        if (kWarps > 1 && (p.flags & (PM_DO_A | PM_DO_CAS | PM_DO_S))) {
            const int reach = (int)fmin(rmax, 4096.0) + 12;
            int lo_r = max(0, min(min(g.half, cy), 2 * cy - g.half) - reach), hi_r = min(g.n - 1, max(max(g.half, cy), 2 * cy - g.half) + reach);
            int lo_c = max(0, min(min(g.half, cx), 2 * cx - g.half) - reach), hi_c = min(g.n - 1, max(max(g.half, cx), 2 * cx - g.half) + reach);
            const long long room = (long long)(c.ar.s_cap - c.ar.s_top) - (long long)(g.n * g.nw * 12 + 24 * 1024);
            while ((long long)(hi_r - lo_r + 1) * (hi_c - lo_c + 1) * (long long)sizeof(T) > room && hi_r - lo_r > 16 && hi_c - lo_c > 16) {
                lo_r++; hi_r--; lo_c++; hi_c--;
            }
            const int ch = hi_r - lo_r + 1, cw = hi_c - lo_c + 1;
            if (ch > 0 && cw > 0 && (long long)ch * cw * (long long)sizeof(T) <= room) {
                T* cb = (T*)c.ar.alloc((size_t)ch * cw * sizeof(T));
                for (int wr = pm::warp(); wr < ch; wr += kWarps) {
                    const T* src = c.img + (size_t)(lo_r + wr) * g.n + lo_c;
                    int x = pm::lane();
                    for (; x + 96 < cw; x += 128) {       // four independent loads in flight
                        const T a0 = pm::ldg(src + x), a1 = pm::ldg(src + x + 32), a2 = pm::ldg(src + x + 64), a3 = pm::ldg(src + x + 96);
                        cb[wr * cw + x] = a0; cb[wr * cw + x + 32] = a1; cb[wr * cw + x + 64] = a2; cb[wr * cw + x + 96] = a3;
                    }
                    for (; x < cw; x += 32) cb[wr * cw + x] = pm::ldg(src + x);
                }
                pm::sync();
                c.crop = cb; c.crop_r0 = lo_r; c.crop_c0 = lo_c; c.crop_h = ch; c.crop_w = cw;
            }
        }
        // ---- calcA x3 (main.py:470-478)
        if (p.flags & (PM_DO_A | PM_DO_AS | PM_DO_CAS)) {
            Arena::Mark mk = c.ar.mark();
            unsigned* tmp = (unsigned*)c.ar.alloc((size_t)g.n * g.nw * 4);
            AsymOut ao;
            phase_calcA(c, cx, cy, &ao, tmp);
            c.ar.release(mk);
            if (p.flags & (PM_DO_A | PM_DO_CAS)) { row.A = ao.A; row.Abgr = ao.Abgr; }
            if (p.flags & PM_DO_AS) { row.As = ao.As; row.As90 = ao.As90; }
        }
        pc.lap(6);   // calcA x3
        // ---- smoothness (casgm.py:276; the call is commented out at main.py:494)
        if (p.flags & PM_DO_S) {
            const double r20 = p.ov.r20_in ? p.ov.r20_in[b] : row.r20;
            const double skyt = p.ov.smooth_sky_in ? p.ov.smooth_sky_in[b] : c.sky;
            if (status == PM_ST_OK && (p.ov.r20_in || (p.flags & PM_DO_CAS)))
                row.S = phase_smoothness(c, cx, cy, rmax, r20, skyt);
        }
        c.crop = nullptr;
    }
    pc.lap(8);   // smoothness (or whatever ended the cutout early)
    row.status = (double)status;
    if (pm::tid() == 0) p.rows[b] = row;
    pm::sync();
}

// persistent groups: cutout indices of the chunk come from a work counter
template <typename F> PM_DEV void work_loop(Sh& sh, unsigned long long* counter, long long nb, F&& fn) {
    for (;;) {
        pm::sync();
        if (pm::tid() == 0) sh.ia[0] = (int)pm::atomic_add(counter, 1ull);
        pm::sync();
        const long long cur = (long long)sh.ia[0];
        if (cur >= nb) break;
        fn(cur);
    }
}
template <typename T> PM_DEV void front_body(const Params& p) {
    Ctx<T> c;
    setup_ctx(c, p, SM_ALL, false, p.scratch_per_cta);
    work_loop(*c.sh, p.counter + 1, p.nb, [&](long long i) { front_one(c, p, i); });
}
template <typename T> PM_DEV void minapix_body(const Params& p) {
    Ctx<T> c;
    setup_ctx(c, p, SM_APER, false, p.scratch_per_cta);
    work_loop(*c.sh, p.counter + 2, p.nb, [&](long long i) { minapix_one(c, p, i); });
}
template <typename T> PM_DEV void tail_body(const Params& p) {
    Ctx<T> c;
    setup_ctx(c, p, SM_ROWOFF | SM_SEG | SM_APER | SM_DIL, kWarps == 1, p.scratch_per_cta);
    work_loop(*c.sh, p.counter + 3, p.nb, [&](long long i) { tail_one(c, p, i); });
}
// the monolith pass over the cutouts the front kernel handed over (rare: lists / radii beyond the slot bounds)
template <typename T> PM_DEV void big_body(const Params& p) {
    Ctx<T> c;
    setup_ctx(c, p, SM_ALL, false, p.scratch_per_cta);
    const long long nbig = (long long)p.counter[4];
    work_loop(*c.sh, p.counter + 0, nbig, [&](long long k) { analyse_one(c, p, p.b0 + (long long)p.big_list[k]); });
}

// apertures.aperpixmap for a batch of radii (one CTA per mask)
PM_DEV void aperpixmap_body(int n, const double* rad, long long B, int ns, double frac, uint8_t* mask_out,
                            unsigned smem_bytes) {
    unsigned char* smem = pm::dyn_smem(smem_bytes);
    const Geo g = make_geo(n, 1);
    unsigned* plane = (unsigned*)smem;
    double* sq = (double*)(smem + align16((unsigned)(g.n * g.nw * 4)));
    int* hw = (int*)((unsigned char*)sq + align16((unsigned)((g.half + 1) * ns * 8)));
    (void)smem_bytes;
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        build_aperture(g, rad[b], ns, frac, plane, sq, hw);
        store_plane_u8(g, plane, mask_out + (size_t)b * g.N);
        pm::sync();
    }
}
PM_HOSTDEV unsigned aperpixmap_smem_bytes(int n, int ns) {
    const Geo g = make_geo(n, 1);
    return align16((unsigned)(g.n * g.nw * 4)) + align16((unsigned)((g.half + 1) * ns * 8)) + align16((unsigned)((g.half + 1) * 4));
}

}  // namespace PM_NS
