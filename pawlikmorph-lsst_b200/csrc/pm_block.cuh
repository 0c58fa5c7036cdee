// pm_block.cuh — block-wide building blocks of the per-cutout kernel: deterministic
// reductions and scans, the stable LSD radix sort, the per-cutout scratch arena and
// the exact circle/pixel overlap geometry.  All reductions use a fixed combination
// order so that results do not depend on scheduling (and therefore not on the number
// of GPUs a batch is sharded over).
#pragma once
#include "pm_platform.cuh"

// PM_NS: the namespace of this translation unit's copy of the kernels.  The library compiles the kernel sources in
// two units — pm_capi.cu (CTA groups, namespace pmk) and pm_tail.cu (PM_GROUP_WARP: warp groups, namespace pmk_w) —
// and the sizes of the block-shared structures differ between them.
#ifndef PM_NS
#define PM_NS pmk
#endif
namespace PM_NS {

// Group shape.  CTA groups: 256 threads x 2 CTAs per SM for the front kernel (two independent CTAs per SM overlap each
// other's barrier waits; measured +12 % over one 512-thread CTA, profiles/).  Warp groups: 32 threads.
#if defined(PM_GROUP_WARP)
#undef PM_THREADS
#define PM_THREADS 32
#endif
#ifndef PM_THREADS
#define PM_THREADS 256
#endif
#ifndef PM_MIN_CTAS
#define PM_MIN_CTAS 2
#endif
constexpr int kThreads = PM_THREADS;
static_assert(kThreads % 32 == 0, "whole warps");
constexpr int kWarps = kThreads / 32;
constexpr int kRedSlots = 8;

// Block-shared scalars.  Lives at offset 0 of the dynamic shared memory.
struct Sh {
    double red[2][kWarps * kRedSlots];  // per-warp partials of block reductions, two alternating halves: a reduction
                                        // needs ONE barrier (the previous half is still being read while this one fills)
    double bc[16];                   // broadcast slots
    double rk[132];                  // walk radii of casgm.py:55-73
    double fk[132];                  // F(rk): lower bounds during the walk
    double fh[132];                  // upper bounds
    unsigned scan_tmp[2][kWarps + 2];
    unsigned char par[kThreads];     // per-thread parity of the reduction scratch (each thread touches only its own)
    unsigned colmask[24];            // OR of all segmap rows (PM_MAX_N / 32 words)
    unsigned long long mbar[2];      // mbarriers of the two TMA staging buffers (pixelmap)
    int ia[8 + 2 * kRedSlots];       // work index; integer thresholds of the radii being evaluated
    int status;
};

PM_DEV unsigned lt_mask() { return (1u << pm::lane()) - 1u; }

// ---- reductions ------------------------------------------------------------------------
// All block reductions: per-warp partials go to one of two alternating scratch halves, ONE barrier, then every
// thread combines the partials in a fixed order.  The half written by reduction k is read until the barrier of
// reduction k + 1 at the latest, and written again only by reduction k + 2: no trailing barrier is needed.
PM_DEV unsigned red_half(Sh& sh) {
    const unsigned h = sh.par[pm::tid()];
    sh.par[pm::tid()] = (unsigned char)(h ^ 1u);
    return h;
}
template <int K> PM_DEV void block_sum(Sh& sh, double (&v)[K]) {
    static_assert(K <= kRedSlots, "too many reduction slots");
    if (kWarps == 1) {                       // warp group: the butterfly leaves the total in every lane
#pragma unroll
        for (int k = 0; k < K; k++) v[k] = pm::warp_sum(v[k]);
        return;
    }
    double* red = sh.red[red_half(sh)];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = pm::warp_sum(v[k]);
    if (pm::lane() == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) red[pm::warp() * K + k] = v[k];
    }
    pm::sync();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = 0.0;
        for (int w = 0; w < kWarps; w++) s += red[w * K + k];
        v[k] = s;
    }
}
PM_DEV double block_sum1(Sh& sh, double v) {
    double a[1] = {v};
    block_sum<1>(sh, a);
    return a[0];
}
PM_DEV double block_max(Sh& sh, double v) {
    double* red = sh.red[red_half(sh)];
    v = pm::warp_max(v);
    if (pm::lane() == 0) red[pm::warp()] = v;
    pm::sync();
    double s = red[0];
    for (int w = 1; w < kWarps; w++) s = fmax(s, red[w]);
    return s;
}
PM_DEV double block_min(Sh& sh, double v) {
    double* red = sh.red[red_half(sh)];
    v = pm::warp_min(v);
    if (pm::lane() == 0) red[pm::warp()] = v;
    pm::sync();
    double s = red[0];
    for (int w = 1; w < kWarps; w++) s = fmin(s, red[w]);
    return s;
}
PM_DEV int block_min_i(Sh& sh, int v) {
    unsigned* tmp = sh.scan_tmp[red_half(sh)];
    v = pm::warp_min(v);
    if (pm::lane() == 0) tmp[pm::warp()] = (unsigned)v;
    pm::sync();
    int s = (int)tmp[0];
    for (int w = 1; w < kWarps; w++) s = min(s, (int)tmp[w]);
    return s;
}
PM_DEV int block_max_i(Sh& sh, int v) {
    unsigned* tmp = sh.scan_tmp[red_half(sh)];
    v = pm::warp_max(v);
    if (pm::lane() == 0) tmp[pm::warp()] = (unsigned)v;
    pm::sync();
    int s = (int)tmp[0];
    for (int w = 1; w < kWarps; w++) s = max(s, (int)tmp[w]);
    return s;
}
PM_DEV unsigned block_sum_u(Sh& sh, unsigned v) {
    unsigned* tmp = sh.scan_tmp[red_half(sh)];
    v = pm::warp_sum(v);
    if (pm::lane() == 0) tmp[pm::warp()] = v;
    pm::sync();
    unsigned s = 0;
    for (int w = 0; w < kWarps; w++) s += tmp[w];
    return s;
}
// exclusive prefix of one value per thread (thread order); *total gets the block total
PM_DEV unsigned block_excl_prefix_u(Sh& sh, unsigned v, unsigned* total) {
    unsigned* tmp = sh.scan_tmp[red_half(sh)];
    unsigned inc = pm::warp_incl_scan(v);
    if (pm::lane() == 31) tmp[pm::warp()] = inc;
    pm::sync();
    unsigned base = 0, tot = 0;
    for (int w = 0; w < kWarps; w++) {
        unsigned t = tmp[w];
        if (w < pm::warp()) base += t;
        tot += t;
    }
    *total = tot;
    return base + inc - v;
}
PM_DEV double block_excl_prefix_d(Sh& sh, double v, double* total) {
    double* red = sh.red[red_half(sh)];
    double inc = pm::warp_incl_scan(v);
    if (pm::lane() == 31) red[pm::warp()] = inc;
    pm::sync();
    double base = 0.0, tot = 0.0;
    for (int w = 0; w < kWarps; w++) {
        double t = red[w];
        if (w < pm::warp()) base += t;
        tot += t;
    }
    *total = tot;
    return base + (inc - v);
}
// in-place exclusive scan of arr[0..count) ; arr[count] = total.  Any count: every thread owns a
// contiguous chunk.
PM_DEV unsigned block_scan_excl_array(Sh& sh, unsigned* arr, int count) {
    const int per = (count + kThreads - 1) / kThreads;
    const int lo = min(count, pm::tid() * per), hi = min(count, lo + per);
    unsigned sum = 0;
    for (int i = lo; i < hi; i++) sum += arr[i];
    unsigned tot;
    unsigned ex = block_excl_prefix_u(sh, sum, &tot);
    for (int i = lo; i < hi; i++) { const unsigned v = arr[i]; arr[i] = ex; ex += v; }
    if (pm::tid() == 0) arr[count] = tot;
    pm::sync();
    return tot;
}

// ---- scratch arena: shared memory first, per-CTA global scratch when it does not fit ------
struct Arena {
    unsigned char* s_base;
    unsigned s_cap, s_top;
    unsigned char* g_base;
    size_t g_cap, g_top;
    int overflow;
    unsigned s_floor;       // per-cutout resets return to these marks (what lies below lives as long as the kernel)
    size_t g_floor;
    PM_DEV void init(unsigned char* sb, unsigned sc, unsigned char* gb, size_t gc) {
        s_base = sb; s_cap = sc; s_top = 0; g_base = gb; g_cap = gc; g_top = 0; overflow = 0; s_floor = 0; g_floor = 0;
#if defined(PM_EMULATE)
        if (((uintptr_t)sb & 15) || ((uintptr_t)gb & 15)) { fprintf(stderr, "pm_emul: misaligned arena base\n"); abort(); }   // x86 would not notice
#endif
    }
    PM_DEV void set_floor() { s_floor = s_top; g_floor = g_top; }
    PM_DEV void reset() { s_top = s_floor; g_top = g_floor; overflow = 0; }
    // global scratch even when shared memory has room (streamed data that should not crowd out random-access data)
    PM_DEV void* alloc_global(size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        if (bytes <= g_cap - g_top) {
            void* p = g_base + g_top;
            g_top += bytes;
            return p;
        }
        overflow = 1;
        return g_base;
    }
    // shared memory or nothing
    PM_DEV void* alloc_smem(size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        if (bytes > (size_t)(s_cap - s_top)) return nullptr;
        void* p = s_base + s_top;
        s_top += (unsigned)bytes;
        return p;
    }
    PM_DEV void* alloc(size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        if (bytes <= (size_t)(s_cap - s_top)) {
            void* p = s_base + s_top;
            s_top += (unsigned)bytes;
            return p;
        }
        if (bytes <= g_cap - g_top) {
            void* p = g_base + g_top;
            g_top += bytes;
            return p;
        }
        overflow = 1;
        return g_base;  // never dereferenced meaningfully: the caller checks `overflow`
    }
    PM_DEV bool in_smem(const void* p) const {
        const unsigned char* q = (const unsigned char*)p;
        return q >= s_base && q < s_base + s_cap;
    }
    struct Mark { unsigned s; size_t g; };
    PM_DEV Mark mark() const { Mark m; m.s = s_top; m.g = g_top; return m; }
    PM_DEV void release(Mark m) { s_top = m.s; g_top = m.g; }
    // after release(mark): re-reserve the first `bytes` of `p`, which was the first allocation made after `mark`
    PM_DEV void keep(const void* p, size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        const unsigned char* q = (const unsigned char*)p;
        if (q >= s_base && q < s_base + s_cap) s_top = (unsigned)(q - s_base) + (unsigned)bytes;
        else g_top = (size_t)(q - g_base) + bytes;
    }
};

// lanes of the warp that hold the same digit as this lane (invalid lanes match nobody).  One MATCH.ANY: ~11 cycles
// per DISTINCT value (370-510 cycles for the typical all-different digits, profiles/microbench), partly hidden by
// batching four chunks per step.  Measured alternatives for the scatter sweep: nine ballots per chunk (79k vs 50k
// cycles), ranks taken from the counting sweep's atomics with a MATCH among colliding lanes only (68k + 27k).
PM_DEV unsigned same_digit_mask(unsigned d, bool valid) {
    return pm::match_any(valid ? d : 0x10000u + (unsigned)pm::lane());
}

// ---- stable LSD radix sort of (key, value) pairs, ascending ---------------------------------------------
// Only the bits that differ somewhere in the batch are sorted (found with one OR-reduction), in the fewest
// passes of at most max_digit_bits (<= kMaxDigitBits) bits.  hist: kWarps << max_digit_bits counters.
// Returns 0 when the sorted data ends in (keyA, valA), 1 for B.
constexpr int kMaxDigitBits = 9;
constexpr int kMaxDigits = 1 << kMaxDigitBits;
PM_DEV int high_bit(unsigned long long v) {          // position of the highest set bit, -1 for 0
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    return hi ? 63 - pm::clz(hi) : (lo ? 31 - pm::clz(lo) : -1);
}
struct NoLap { PM_DEV void operator()(int) const {} };     // lap(phase): optional tuning probe
template <typename K, typename Lap = NoLap>
PM_DEV int radix_sort_pairs(Sh& sh, K* keyA, unsigned* valA, K* keyB, unsigned* valB, unsigned* hist, int n,
                            int max_digit_bits = kMaxDigitBits, Lap lap = Lap()) {
    const int w = pm::warp(), l = pm::lane(), t = pm::tid();
    if (n <= 1) return 0;
    // bits that differ from the first key anywhere
    unsigned dlo = 0, dhi = 0;
    {
        const K k0 = keyA[0];
        for (int i = t; i < n; i += kThreads) {
            const unsigned long long x = (unsigned long long)(keyA[i] ^ k0);
            dlo |= (unsigned)x; dhi |= (unsigned)(x >> 32);
        }
        dlo = pm::warp_or(dlo); dhi = pm::warp_or(dhi);
        unsigned* tmp = sh.scan_tmp[red_half(sh)];
        if (l == 0) { tmp[w] = dlo; }
        pm::sync();
        unsigned a = 0;
        for (int ww = 0; ww < kWarps; ww++) a |= tmp[ww];
        dlo = a;
        tmp = sh.scan_tmp[red_half(sh)];
        if (l == 0) { tmp[w] = dhi; }
        pm::sync();
        a = 0;
        for (int ww = 0; ww < kWarps; ww++) a |= tmp[ww];
        dhi = a;
    }
    lap(16);
    const int nbits = high_bit(((unsigned long long)dhi << 32) | dlo) + 1;
    if (nbits == 0) return 0;                            // all keys equal: already sorted (stable)
    const int passes = (nbits + max_digit_bits - 1) / max_digit_bits;
    const int db = (nbits + passes - 1) / passes;       // bits per digit, <= max_digit_bits
    const int D = 1 << db;
    const unsigned mask = (unsigned)D - 1u;
    int tile = (n + kWarps - 1) / kWarps;
    tile = (tile + 31) & ~31;
    const int lo = min(n, w * tile), hi = min(n, lo + tile);
    K* src = keyA; unsigned* sv = valA; K* dst = keyB; unsigned* dv = valB;
    int where = 0;
    for (int pass = 0; pass < passes; pass++) {
        const int shift = pass * db;
        for (int i = t; i < kWarps * D; i += kThreads) hist[i] = 0;
        pm::sync();
        // sweep 1: per-warp digit counts, one shared-memory atomic per key (the digits of a warp's 32 keys are mostly
        // different: little same-address serialisation, and no warp vote or MATCH is needed: 37k -> 16k cycles)
        for (int base = lo; base < hi; base += 128) {
            unsigned d[4];
            bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + 32 * u + l;
                valid[u] = i < hi;
                d[u] = valid[u] ? ((unsigned)(src[valid[u] ? i : lo] >> shift) & mask) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (valid[u]) pm::atomic_add(&hist[w * D + d[u]], 1u);
        }
        pm::sync();
        lap(17);
        // exclusive offsets in (digit-major, warp-minor) order; a thread owns digits t, t + kThreads, ...
        int skip = 0;
        unsigned carry = 0;
        for (int dbase = 0; dbase < D; dbase += kThreads) {
            const int dg = dbase + t;
            unsigned total = 0;
            if (dg < D) {
                for (int ww = 0; ww < kWarps; ww++) {
                    const unsigned c = hist[ww * D + dg];
                    hist[ww * D + dg] = total;
                    total += c;
                }
            }
            skip |= (dg < D && total == (unsigned)n) ? 1 : 0;
            unsigned grand;
            const unsigned basep = block_excl_prefix_u(sh, total, &grand) + carry;
            if (dg < D) {
                for (int ww = 0; ww < kWarps; ww++) hist[ww * D + dg] += basep;
            }
            carry += grand;
        }
        skip = pm::sync_or(skip);
        lap(18);
        if (skip) continue;  // every key has the same digit: the pass is the identity
        // sweep 2: stable scatter.  Same four-chunk batching; only the short offset read / update is ordered.
        for (int base = lo; base < hi; base += 128) {
            K k[4];
            unsigned v[4], d[4], m[4];
            bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + 32 * u + l;
                valid[u] = i < hi;
                k[u] = src[valid[u] ? i : lo];
                v[u] = sv[valid[u] ? i : lo];
                d[u] = valid[u] ? ((unsigned)(k[u] >> shift) & mask) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) m[u] = same_digit_mask(d[u], valid[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                unsigned pos = 0;
                if (valid[u]) pos = hist[w * D + d[u]] + (unsigned)pm::popc(m[u] & lt_mask());
                pm::syncwarp();
                if (valid[u]) {
                    dst[pos] = k[u];
                    dv[pos] = v[u];
                    if (l == pm::ffs(m[u]) - 1) hist[w * D + d[u]] += (unsigned)pm::popc(m[u]);
                }
                pm::syncwarp();
            }
        }
        pm::sync();
        lap(19);
        K* tk = src; src = dst; dst = tk;
        unsigned* tv = sv; sv = dv; dv = tv;
        where ^= 1;
    }
    return where;
}

// ---- bucket sort of (key, value) pairs, ascending by (key, value) ---------------------------------------------
// The values are unique and ascending in the input (raster positions), so ordering by (key, value) IS the stable
// ascending order of np.argsort(kind="stable") that the canonical tie rule needs.  One counting pass
// over up to 4096 buckets that are linear in the key bits (order-preserving float bits: logarithmic in the pixel
// value, which matches the skewed brightness distribution of a galaxy), one scatter in arbitrary order inside a
// bucket (shared-memory atomics only: no warp votes, no MATCH), then every element finds its rank inside its
// bucket by comparing itself with the bucket's other elements and moves to its final place.  Two data movements
// instead of the 2 x 3 of the LSD radix sort; the compare step costs sum(bucket size^2) — typically < 20 per
// element.  Returns 0 (sorted data in keyA / valA) or -1 when a bucket holds more than kMaxBucket elements (many
// equal keys: integer-valued frames, flat maps): the caller then falls back to radix_sort_pairs.
constexpr int kBucketBits = 12;
constexpr int kBuckets = 1 << kBucketBits;
constexpr int kMaxBucket = 192;
template <typename K> PM_DEV K block_min_max_key(Sh& sh, K kmin, K kmax, K* out_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = pm::shfl_xor((unsigned long long)kmin, o), b = pm::shfl_xor((unsigned long long)kmax, o);
        kmin = (K)a < kmin ? (K)a : kmin;
        kmax = (K)b > kmax ? (K)b : kmax;
    }
    unsigned long long* red = (unsigned long long*)sh.red[red_half(sh)];
    if (pm::lane() == 0) { red[2 * pm::warp()] = (unsigned long long)kmin; red[2 * pm::warp() + 1] = (unsigned long long)kmax; }
    pm::sync();
    K lo = (K)red[0], hi = (K)red[1];
    for (int w = 1; w < kWarps; w++) {
        const K a = (K)red[2 * w], b = (K)red[2 * w + 1];
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    *out_max = hi;
    return lo;
}
template <typename K, typename Lap = NoLap>
PM_DEV int bucket_sort_pairs(Sh& sh, K* keyA, unsigned* valA, K* keyB, unsigned* valB, unsigned* hist /* kBuckets + 2 */, int n,
                             Lap lap = Lap()) {
    const int t = pm::tid();
    if (n <= 1) return 0;
    K kmin = ~(K)0, kmax = 0;
    for (int i = t; i < n; i += kThreads) {
        const K k = keyA[i];
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
    kmin = block_min_max_key<K>(sh, kmin, kmax, &kmax);
    lap(16);
    if (kmax == kmin) return 0;                          // all keys equal: the input order is the answer
    const unsigned long long range = (unsigned long long)(kmax - kmin);
    const int hb = high_bit(range) + 1;                   // range < 2^hb
    const int shift = hb > kBucketBits ? hb - kBucketBits : 0;
    const int nb = (int)(range >> shift) + 1;             // <= kBuckets
    for (int i = t; i <= nb; i += kThreads) hist[i] = 0;
    pm::sync();
    for (int i = t; i < n; i += kThreads) pm::atomic_add(&hist[(unsigned)((keyA[i] - kmin) >> shift)], 1u);
    pm::sync();
    lap(17);
    int cmax = 0;
    for (int i = t; i < nb; i += kThreads) cmax = max(cmax, (int)hist[i]);
    cmax = block_max_i(sh, cmax);
    if (cmax > kMaxBucket) return -1;
    block_scan_excl_array(sh, hist, nb);                  // hist[b] = first slot of bucket b
    lap(18);
    // scatter (arbitrary order inside a bucket); afterwards hist[b] = END of bucket b
    for (int i0 = t; i0 < n; i0 += 4 * kThreads) {
        K k[4];
        unsigned v[4], slot[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * kThreads;
            k[u] = keyA[i < n ? i : 0];
            v[u] = valA[i < n ? i : 0];
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            slot[u] = (i0 + u * kThreads < n) ? pm::atomic_add(&hist[(unsigned)((k[u] - kmin) >> shift)], 1u) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * kThreads < n) { keyB[slot[u]] = k[u]; valB[slot[u]] = v[u]; }
    }
    pm::sync();
    // rank inside the bucket, final placement
    for (int i = t; i < n; i += kThreads) {
        const K k = keyB[i];
        const unsigned v = valB[i];
        const unsigned b = (unsigned)((k - kmin) >> shift);
        const int lo = b ? (int)hist[b - 1] : 0, hi = (int)hist[b];
        int r = 0, j = lo;
        for (; j + 3 < hi; j += 4) {                          // four independent compares in flight
            const K k0 = keyB[j], k1 = keyB[j + 1], k2 = keyB[j + 2], k3 = keyB[j + 3];
            const unsigned v0 = valB[j], v1 = valB[j + 1], v2 = valB[j + 2], v3 = valB[j + 3];
            r += ((k0 < k || (k0 == k && v0 < v)) ? 1 : 0) + ((k1 < k || (k1 == k && v1 < v)) ? 1 : 0) +
                 ((k2 < k || (k2 == k && v2 < v)) ? 1 : 0) + ((k3 < k || (k3 == k && v3 < v)) ? 1 : 0);
        }
        for (; j < hi; j++) {
            const K kj = keyB[j];
            const unsigned vj = valB[j];
            r += (kj < k || (kj == k && vj < v)) ? 1 : 0;
        }
        keyA[lo + r] = k;
        valA[lo + r] = v;
    }
    pm::sync();
    lap(19);
    return 0;
}

// ---- exact area of circle(0, r) ∩ rectangle — photutils geometry -------------------
// Restates photutils/geometry/circular_overlap.pyx + core.pyx (photutils 0.7.x; the reference calls
// CircularAperture(...).do_photometry(method="exact") at casgm.py:51-52, :66-67, :116-117, :333-336).
PM_DEV double fsqrt0(double v) { return v > 0.0 ? sqrt(v) : 0.0; }
// (theta - sin theta) / h^3 for theta = 2 asin(h), as a series in x = h^2 <= 1/4:
//   theta - sin theta = 2 asin(h) - 2 h sqrt(1 - h^2) = sum_{k>=1} g_k h^(2k+1),  g_k = 2 C(2k,k)/4^k * 4k/((2k+1)(2k-1)).
// 30 terms (truncation < 1e-18 at x = 1/4, all coefficients positive), Estrin's scheme: five FMA levels.
PM_DEV double seg_series(double x) {
    const double g[32] = {
        1.3333333333333333, 0.4, 0.21428571428571427, 0.1388888888888889, 0.09943181818181818, 0.07572115384615384,
        0.06015625, 0.04928768382352941, 0.04134328741776316, 0.035327729724702384, 0.030642965565557064, 0.02691009521484375,
        0.023878556710702402, 0.02137669201554923, 0.019283352359648672, 0.017510842193256725, 0.015994278181876456,
        0.014684730763169559, 0.013544676879134316, 0.012544909330289357, 0.011662389546007373, 0.010878726333127512,
        0.010179079039942812, 0.00955135411245743, 0.008985608055959748, 0.008473597936544683, 0.008008438888979117,
        0.007584340273350919, 0.007196400350168018, 0.006840444990846536, 0.0, 0.0};
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
    double p[16], q[8], o[4];
#pragma unroll
    for (int i = 0; i < 16; i++) p[i] = fma(g[2 * i + 1], x, g[2 * i]);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = fma(p[2 * i + 1], x2, p[2 * i]);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = fma(q[2 * i + 1], x4, q[2 * i]);
    const double s0 = fma(o[1], x8, o[0]), s1 = fma(o[3], x8, o[2]);
    return fma(s1, x16, s0);
}
// circular segment cut off by the chord (x1, y1)-(x2, y2): photutils takes a = chord, theta = 2 asin(a / 2r),
// area = r^2 (theta - sin theta) / 2.  Here theta - sin theta comes from the series in h^2 = (a / 2r)^2 when
// h <= 1/2 (every pixel-sized chord of a circle with r >= sqrt(2)): no asin, no division, no cancellation, and only
// one square root on the dependency chain (the polynomial runs beside it).  Agrees with the closed form to < 1e-15
// relative (the closed form itself loses ~1e-10 relative to cancellation for short chords).  inv4r2 = 1 / (4 r^2).
PM_DEV double area_arc(double x1, double y1, double x2, double y2, double r, double inv4r2) {
    const double c2 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);      // chord^2
    const double h2 = c2 * inv4r2;
    if (h2 <= 0.25) return 0.125 * c2 * (sqrt(h2) * seg_series(h2));      // r^2 h^3 g / 2 = (a^2 / 8) h g
    const double a = sqrt(c2);
    const double h = 0.5 * a / r;
    const double theta = 2.0 * asin(h);
    const double sin_theta = 2.0 * h * sqrt(fmax(1.0 - h * h, 0.0));
    return 0.5 * r * r * (theta - sin_theta);
}
PM_DEV double area_tri(double x1, double y1, double x2, double y2, double x3, double y3) {
    return 0.5 * fabs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));
}
// rectangle in the first quadrant (0 <= xmin, 0 <= ymin).  photutils (circular_overlap_core) branches on which
// corners lie inside the circle; here the four corner configurations share ONE instruction stream — the two
// intersection points, the arc segment and the (up to two) triangles are formed from selected operands — so that
// the lanes of a warp, which hold different pixels, do not serialise four copies of the sqrt / asin chain
// (measured: 1915 -> ~500 cycles per call, profiles/microbench).  Per configuration the arithmetic is photutils'.
PM_DEV double overlap_core(double xmin, double ymin, double xmax, double ymax, double r) {
    const double r2 = r * r, inv4r2 = 0.25 / r2;
    const double full = (xmax - xmin) * (ymax - ymin);
    const bool none = xmin * xmin + ymin * ymin > r2, all = xmax * xmax + ymax * ymax < r2;
    // corner tests on squared distances (photutils compares the square roots: same decisions up to rounding at
    // the exact boundary, where the area formulas agree anyway)
    const bool in1 = xmax * xmax + ymin * ymin < r2, in2 = xmin * xmin + ymax * ymax < r2;
    const bool cA = in1 && in2, cB = in1 && !in2, cC = !in1 && in2;       // else: neither far corner inside
    // the two edges the circle crosses: P1 on (A: top, B: left, C / D: bottom), P2 on (A / B: right, C: top, D: left)
    const double sa = cA ? ymax : (cB ? xmin : ymin);
    const double sb = (cA || cB) ? xmax : (cC ? ymax : xmin);
    const double qa = fsqrt0(r2 - sa * sa), qb = fsqrt0(r2 - sb * sb);
    const double x1 = cB ? sa : qa, y1 = cB ? qa : sa;
    const double x2 = cC ? qb : sb, y2 = cC ? sb : qb;
    const double arc = area_arc(x1, y1, x2, y2, r, inv4r2);
    // triangle 1: (P1, V2, V3); triangle 2 (configurations B and C only): (P1, W2, P2)
    const double v2x = cB ? x1 : (cC ? xmin : x2), v2y = cB ? ymin : (cC ? y1 : y2);
    const double v3x = (cA || cB) ? xmax : xmin, v3y = (cA || cC) ? ymax : ymin;
    const double w2x = cB ? x2 : xmin, w2y = cB ? ymin : y2;
    const double t1 = area_tri(x1, y1, v2x, v2y, v3x, v3y);
    const double t2 = (cB || cC) ? area_tri(x1, y1, w2x, w2y, x2, y2) : 0.0;
    const double cut = cA ? (full - t1) + arc : (arc + t1) + t2;
    return none ? 0.0 : (all ? full : cut);
}
// unit pixel centred at the INTEGER offset (a, b) from the circle centre, any quadrant.  A pixel that straddles an
// axis is split there by photutils and each part mapped to the first quadrant: for integer offsets the parts are
// mirror images of each other, so one evaluation is doubled (a == 0 or b == 0) or quadrupled (a == b == 0).
PM_DEV double overlap_pixel(int a, int b, double r) {
    const int ia = a < 0 ? -a : a, ib = b < 0 ? -b : b;
    const double px = (double)ia, py = (double)ib;
    const double xmin = ia == 0 ? 0.0 : px - 0.5, ymin = ib == 0 ? 0.0 : py - 0.5;
    const double part = overlap_core(xmin, ymin, px + 0.5, py + 0.5, r);
    const double mult = (ia == 0 ? 2.0 : 1.0) * (ib == 0 ? 2.0 : 1.0);
    return mult * part;
}
// photutils weight of the pixel at integer offset (dx, dy) from the circle centre
PM_DEV double circle_weight(int dx, int dy, double r) {
    const double pr = 0.5 * sqrt(2.0);
    double d = sqrt((double)(dx * dx + dy * dy));
    if (d < r - pr) return 1.0;
    if (d < r + pr) return overlap_pixel(dx, dy, r);
    return 0.0;
}

}  // namespace PM_NS
