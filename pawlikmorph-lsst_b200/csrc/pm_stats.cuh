// pm_stats.cuh — order statistics of one cutout, for the two pre-processing steps that are made of them:
//   * sky_body        the sky-region statistics of skyBackground._calcSkybgr (skyBackground.py:134-164): the pixels whose
//                     centre lies outside the ellipse of semi-axes a = 2 min(fwhm), b = 2 max(fwhm), angle theta about
//                     (int(n/2) + 1, int(n/2) + 1) (`_getSkyRegion`, :265-303: photutils EllipticalAperture.to_mask(
//                     method="center")), their count / mean / median / population std (np.ma.mean / median / std,
//                     :153-155), and, when mean > median (the crowded branch, :161-164), astropy's
//                     sigma_clipped_stats(sigma=3, maxiters=5, cenfunc=median, stdfunc=std), sky = 3 median - 2 mean
//   * clipstats_body  astropy.stats.sigma_clipped_stats(image, sigma, maxiters) over a whole cutout, plus the plain
//                     statistics and the maximum: imageutils.py:62, photutils.detect_threshold (skyBackground.py:109,
//                     imageutils.py:67) and gaussfitter.moments' median / max (gaussfitter.py:71-72)
// Every statistics set costs several passes over the pixels (sums, then a radix select of 12 bits per pass for the
// median), and a clipped statistic is up to six sets.  Passes are merged where the key range is known beforehand: a
// clipped set [lo, hi] builds the histogram of its first digit in the pass that takes its sums, the keys under a
// select prefix are one key range (a single unsigned compare per pixel), and the upper middle element of an even count
// is read off the last digit's histogram.
// The code serves two forms (pm_k_stats.cu picks): one CTA per cutout with the pixels in global memory (passes through
// L2), or a thread-block cluster of C CTAs per cutout, CTA q holding rows [q n/C, (q+1) n/C) in its shared memory.  In
// the cluster form the CTAs exchange three sums, two extremes and — per select pass — 64 coarse histogram sums plus the
// 64 counters under the coarse bin that holds the rank, read from each other's shared memory (distributed shared
// memory) after a cluster barrier; every CTA continues with the same totals, so all take the same decisions.
#pragma once
#include "pm_block.cuh"

namespace PM_NS {

PM_HOSTDEV unsigned st_align16(unsigned v) { return (v + 15u) & ~15u; }

constexpr int kSkyCols = 12;    // count, mean, median, std, iterations, clipped count, mean, median, std, sky, sky_err, status
constexpr int kClipCols = 10;   // count, mean, median, std (all pixels), iterations, clipped count, mean, median, std, max
constexpr int kStatDigit = 12;  // bits resolved per select pass
constexpr int kStatBins = 1 << kStatDigit;
constexpr int kStatRun = kStatBins / kThreads;     // histogram bins a thread owns when the rank is located
static_assert(kStatBins % kThreads == 0 && kStatRun >= 1, "bins per thread");
constexpr unsigned kNoBin = 0xffffffffu;
constexpr int kStatCollect = 512;      // a prefix holding at most this many keys is finished by gathering them (one pass)

struct StatScratch {
    unsigned hist[2][kStatBins];      // alternating: the other CTAs may still read the previous pass's counters
    unsigned coarse[2][64];           // sums of 64 counters each (what the other CTAs read first)
    double xd[2][8];                  // cluster exchange slots, alternating per exchange
    double xall[40];                  // the slots of all CTAs, gathered
    unsigned bin, left, val;          // result of locating a rank (written by the owner of the bin)
    unsigned nlist;
    unsigned long long found[2];
    unsigned long long list[kStatCollect];   // the keys under a sparsely populated prefix, gathered for a direct selection
};

// order-preserving keys.  -0.0 and +0.0 get adjacent keys (numpy compares them equal: a bound that is a zero is
// widened to take in both, set_bounds; as a median either one is the same number)
PM_DEV unsigned stat_key(float x) { const unsigned u = pm::f2bits(x); return u ^ ((unsigned)((int)u >> 31) | 0x80000000u); }
PM_DEV unsigned long long stat_key(double x) {
    const unsigned long long u = (unsigned long long)pm::d2bits(x);
    return u ^ ((unsigned long long)((long long)u >> 63) | 0x8000000000000000ull);
}
PM_DEV double stat_val(unsigned k, float) { return (double)pm::bits2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
PM_DEV double stat_val(unsigned long long k, double) {
    return pm::bits2d((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
PM_DEV int stat_clz(unsigned v) { return v ? pm::clz(v) : 32; }
PM_DEV int stat_clz(unsigned long long v) { return (v >> 32) ? pm::clz((unsigned)(v >> 32)) : 32 + stat_clz((unsigned)v); }

// the cluster as this thread sees it, and which of the alternating buffers the next pass / exchange uses
struct StatCl {
    unsigned rank, size;
    unsigned hsel, xsel;
};

template <typename T> struct StatSet {      // pixels of this CTA's rows that belong to the region and lie inside [lo, hi]
    typedef decltype(stat_key(T())) K;
    const T* data;           // first local row: the shared-memory copy, or the cutout in global memory
    const unsigned* mask;    // region bit plane of the local rows, or nullptr (every pixel)
    int n, nw, rows;
    double lo, hi;
    K klo, khi;              // [lo, hi] as a key range: membership is two integer compares
    PM_DEV void set_bounds(double lo_, double hi_) {
        lo = lo_; hi = hi_;
        klo = stat_key((T)lo_);
        if (stat_val(klo, T()) < lo_) klo += (K)1;        // (T)lo rounded down: the next key up is the first one inside
        if ((T)lo_ == (T)0 && lo_ <= 0.0) klo = stat_key(-(T)0);     // x >= 0 holds for -0.0 too
        khi = stat_key((T)hi_);
        if (stat_val(khi, T()) > hi_) khi -= (K)1;
        if ((T)hi_ == (T)0 && hi_ >= 0.0) khi = stat_key((T)0);      // x <= 0 holds for +0.0 too
    }
};

// f(x, key) for every member of the set: one row per warp, 32 columns per word, the loads of kStatGroup words issued
// together (unconditionally: the pixels outside the region are few).  Membership in [klo, khi] is one unsigned
// compare of key - klo against khi - klo.  kFull: n is a multiple of 32 * kStatGroup (no bounds checks at all).
constexpr int kStatGroup = 8;
template <typename T, bool kFull, bool kMask, bool kSparse, typename F> PM_DEV void stat_for_each_impl(const StatSet<T>& s, F&& f) {
    typedef typename StatSet<T>::K K;
    const K klo = s.klo, span = s.khi - s.klo;        // callers guarantee klo <= khi
    for (int r = pm::warp(); r < s.rows; r += kWarps) {
        const T* row = s.data + (size_t)r * s.n + pm::lane();
        const unsigned* mrow = kMask ? s.mask + r * s.nw : nullptr;
        for (int w0 = 0; w0 < s.nw; w0 += kStatGroup) {
            T xv[kStatGroup];
            unsigned bits[kStatGroup];
#pragma unroll
            for (int j = 0; j < kStatGroup; j++) {
                const int wi = w0 + j;
                const bool ok = kFull || (wi < s.nw && wi * 32 + pm::lane() < s.n);
                xv[j] = ok ? row[wi * 32] : (T)0;
                bits[j] = !ok ? 0u : kMask ? mrow[wi] >> pm::lane() : 1u;
            }
            if (kSparse || kMask) {
                // membership of the eight pixels first, ONE branch for the group: in the select passes hardly any pixel is a member;
                // with a region plane it also pays in the dense passes (measured: sky 0.92 -> 1.00 M cutouts/s), without one it costs 15 %
                K key[kStatGroup];
                unsigned member = 0u;
#pragma unroll
                for (int j = 0; j < kStatGroup; j++) {
                    key[j] = stat_key(xv[j]);
                    const bool in = (!kMask && kFull ? true : (bits[j] & 1u) != 0u) && (K)(key[j] - klo) <= span;
                    member |= (in ? 1u : 0u) << j;
                }
                if (member) {
#pragma unroll
                    for (int j = 0; j < kStatGroup; j++)
                        if ((member >> j) & 1u) f(xv[j], key[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kStatGroup; j++) {
                    const K key = stat_key(xv[j]);
                    if ((!kMask && kFull ? true : (bits[j] & 1u) != 0u) && (K)(key - klo) <= span) f(xv[j], key);     // NaNs sort outside every range
                }
            }
        }
    }
}
template <typename T, bool kSparse = false, typename F> PM_DEV void stat_for_each(const StatSet<T>& s, F&& f) {
    if (s.klo > s.khi) return;                        // empty range
    const bool full = s.n % (32 * kStatGroup) == 0;
    if (s.mask) {
        if (full) stat_for_each_impl<T, true, true, kSparse>(s, f); else stat_for_each_impl<T, false, true, kSparse>(s, f);
    } else {
        if (full) stat_for_each_impl<T, true, false, kSparse>(s, f); else stat_for_each_impl<T, false, false, kSparse>(s, f);
    }
}

// ---- cluster exchange: v[0..2] summed, v[3] minimum, v[4] maximum over the CTAs, combined in rank order by everyone
// (after a block reduction: v is the same in all threads of a CTA).  One cluster barrier; it also orders the histogram
// counters of the pass before it.
PM_DEV void stat_exchange(StatScratch& sc, StatCl& cl, double (&v)[5]) {
    if (cl.size == 1) return;
    double* slot = sc.xd[cl.xsel];
    if (pm::tid() == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) slot[k] = v[k];
    }
    pm::cluster_sync();
    // one remote read per thread (all in flight together), then everybody combines the local copy in rank order
    if (pm::tid() < 5 * (int)cl.size) {
        const unsigned r = (unsigned)pm::tid() / 5u, k = (unsigned)pm::tid() % 5u;
        sc.xall[pm::tid()] = pm::dsmem_ld_f64(pm::dsmem_addr(pm::smem_addr(slot + k), r));
    }
    pm::sync();
#pragma unroll
    for (int k = 0; k < 5; k++) v[k] = sc.xall[k];
    for (unsigned r = 1; r < cl.size; r++) {
        const double* o = sc.xall + 5 * r;
        v[0] += o[0]; v[1] += o[1]; v[2] += o[2];
        v[3] = fmin(v[3], o[3]); v[4] = fmax(v[4], o[4]);
    }
    cl.xsel ^= 1u;
}
// barrier after a histogram pass that has no exchange of its own
PM_DEV void stat_pass_barrier(const StatCl& cl) {
    if (cl.size == 1) pm::sync(); else pm::cluster_sync();
}

// ---- locating a rank in the histogram of the whole cluster.  The counters stay where they were counted; every CTA
// adds up 64 coarse sums (64 bins each) of its own histogram before the barrier that ends the pass, and afterwards
// reads, from every CTA, the coarse sums and then the 64 counters under the coarse bin that holds the rank:
// 128 words per CTA instead of the whole histogram.
constexpr int kStatCoarse = 64;
constexpr int kStatFine = kStatBins / kStatCoarse;
static_assert(kStatFine == 64 && kThreads >= 64, "two levels of 64");
constexpr int kStatLanesPerCoarse = kStatFine / kStatRun;       // threads that share one coarse bin (consecutive lanes)
static_assert(kStatLanesPerCoarse >= 1 && kStatLanesPerCoarse <= 32 && (kStatLanesPerCoarse & (kStatLanesPerCoarse - 1)) == 0, "lanes per coarse bin");

// after the local counters are complete (block barrier): coarse sums of this CTA's histogram h
PM_DEV void stat_coarse_sums(StatScratch& sc, unsigned h) {
    const unsigned* mine = sc.hist[h] + pm::tid() * kStatRun;
    unsigned v = 0;
#pragma unroll
    for (int j = 0; j < kStatRun; j++) v += mine[j];
#pragma unroll
    for (int o = 1; o < kStatLanesPerCoarse; o <<= 1) v += pm::shfl_xor(v, o);
    if ((pm::tid() & (kStatLanesPerCoarse - 1)) == 0) sc.coarse[h][pm::tid() / kStatLanesPerCoarse] = v;
}
// word `idx` of the array at shared address `base`, summed over the CTAs of the cluster (threads idx < 64)
PM_DEV unsigned stat_cluster_word(const StatCl& cl, unsigned base, unsigned idx) {
    unsigned v[8];
#pragma unroll
    for (unsigned r = 0; r < 8; r++) v[r] = r < cl.size ? pm::dsmem_ld_u32(pm::dsmem_addr(base + 4u * idx, r)) : 0u;
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}
// first index > after among the threads' 64 values that is non-zero (kNoBin if none); val is this thread's value
PM_DEV unsigned stat_next_nonzero(Sh& sh, unsigned val, int after) {
    int cand = 0x7fffffff;
    if (pm::tid() < 64 && pm::tid() > after && val) cand = pm::tid();
    cand = block_min_i(sh, cand);
    return cand == 0x7fffffff ? kNoBin : (unsigned)cand;
}
// rank kk among the threads' 64 values (threads >= 64 hold 0): index holding it and the rank inside, to sc.bin / sc.left
PM_DEV void stat_locate64(Sh& sh, StatScratch& sc, unsigned val, unsigned kk) {
    unsigned total;
    const unsigned before = block_excl_prefix_u(sh, val, &total);
    if (pm::tid() < 64 && kk >= before && kk < before + val) { sc.bin = (unsigned)pm::tid(); sc.left = kk - before; sc.val = val; }
    pm::sync();
}
// The bin holding rank kk of the cluster's histogram h and the rank inside it; with want_next also the bin of rank
// kk + 1 (kNoBin when the histogram ends first).  Call after the barrier that ends the pass.  The results are the
// same in every thread of every CTA.
struct StatLoc { unsigned bin, left, next, count; };
PM_DEV StatLoc stat_locate(Sh& sh, StatScratch& sc, const StatCl& cl, unsigned h, unsigned kk, bool want_next) {
    StatLoc L;
    const unsigned cbase = pm::smem_addr(sc.coarse[h]), hbase = pm::smem_addr(sc.hist[h]);
    const unsigned t = (unsigned)pm::tid();
    const unsigned cval = t < 64 ? (cl.size > 1 ? stat_cluster_word(cl, cbase, t) : sc.coarse[h][t]) : 0u;
    stat_locate64(sh, sc, cval, kk);
    const unsigned cbin = sc.bin, cleft = sc.left;
    pm::sync();
    const unsigned fval = t < 64 ? (cl.size > 1 ? stat_cluster_word(cl, hbase, cbin * kStatFine + t) : sc.hist[h][cbin * kStatFine + t]) : 0u;
    stat_locate64(sh, sc, fval, cleft);
    const unsigned fbin = sc.bin, fval_at = sc.val;
    L.bin = cbin * kStatFine + fbin; L.left = sc.left; L.next = kNoBin; L.count = fval_at;
    pm::sync();
    if (want_next) {
        if (L.left + 1u < fval_at) L.next = L.bin;
        else {
            unsigned nf = stat_next_nonzero(sh, fval, (int)fbin);
            if (nf != kNoBin) L.next = cbin * kStatFine + nf;
            else {
                const unsigned nc = stat_next_nonzero(sh, cval, (int)cbin);
                if (nc != kNoBin) {
                    const unsigned gval = t < 64 ? (cl.size > 1 ? stat_cluster_word(cl, hbase, nc * kStatFine + t) : sc.hist[h][nc * kStatFine + t]) : 0u;
                    nf = stat_next_nonzero(sh, gval, -1);
                    L.next = nc * kStatFine + nf;
                }
            }
        }
    }
    return L;
}

struct StatOut { double cnt, mean, median, std, vmax; };

// Statistics of the set.  bounded: the key range [klo, khi] is the a-priori range (a clipped set); otherwise the
// range is measured (first set: every finite pixel of the region).  The sums run on x - shift (any value near the
// data: the squared deviations then come from one pass, sum (x-c)^2 - (sum (x-c))^2 / cnt, without cancellation worth
// speaking of).
template <typename T> PM_DEV StatOut stat_set(Sh& sh, StatScratch& sc, StatCl& cl, const StatSet<T>& s, double shift, bool bounded) {
    typedef typename StatSet<T>::K K;
    constexpr int KB = (int)(8 * sizeof(K));
    const double inf = 1.7976931348623157e308 * 2.0;
    StatOut st;
    double a[3] = {0.0, 0.0, 0.0};
    K kmin = ~(K)0, kmax = (K)0;
    unsigned mine = 0;
    int rem = 0;
    bool first_hist = false;
    const bool top = !bounded;      // the first digit of the first set: the top 12 bits of any key
    if (bounded && s.klo <= s.khi) {
        rem = KB - stat_clz((K)(s.klo ^ s.khi));
        first_hist = rem > 0;
    } else if (top) {
        rem = KB;
        first_hist = true;
    }
    if (first_hist) {
        const int bits = rem < kStatDigit ? rem : kStatDigit, shr = rem - bits;
        unsigned* hist = sc.hist[cl.hsel];
        for (int t = pm::tid(); t < kStatBins; t += kThreads) hist[t] = 0u;
        pm::sync();
        const unsigned bmask = (1u << bits) - 1u;
        if (top) {
            stat_for_each(s, [&](T x, K key) {
                const double d = (double)x - shift;
                mine++; a[1] += d; a[2] += d * d;
                kmax = key > kmax ? key : kmax;
                pm::atomic_add(&hist[(unsigned)(key >> shr) & bmask], 1u);
            });
        } else {
            stat_for_each(s, [&](T x, K key) {
                const double d = (double)x - shift;
                mine++; a[1] += d; a[2] += d * d;
                pm::atomic_add(&hist[(unsigned)(key >> shr) & bmask], 1u);
            });
        }
    } else {
        stat_for_each(s, [&](T x, K key) {
            const double d = (double)x - shift;
            mine++; a[1] += d; a[2] += d * d;
            kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
        });
    }
    a[0] = (double)mine;
    block_sum<3>(sh, a);
    if (first_hist) {            // (the reduction's barrier has ordered the counters of this CTA)
        stat_coarse_sums(sc, cl.hsel);
        if (cl.size == 1) pm::sync();
    }
    double v[5] = {a[0], a[1], a[2], inf, -inf};
    if (!first_hist) v[3] = block_min(sh, mine ? stat_val(kmin, T()) : inf);
    if (!first_hist || top) v[4] = block_max(sh, mine ? stat_val(kmax, T()) : -inf);
    stat_exchange(sc, cl, v);
    st.cnt = v[0];
    st.vmax = v[4];
    if (v[0] == 0.0) {
        if (first_hist) cl.hsel ^= 1u;
        st.mean = st.median = st.std = pm::bits2d(0x7ff8000000000000ll);
        return st;
    }
    const double md = v[1] / v[0];
    st.mean = shift + md;
    st.std = pm::dsqrt(fmax(v[2] / v[0] - md * md, 0.0));
    // ---- median: element (cnt - 1) / 2 of the ascending order, averaged with the next one when the count is even
    const unsigned cnt = (unsigned)v[0];
    const bool even = (cnt & 1u) == 0u;
    unsigned k = (cnt - 1u) / 2u;
    K common;
    if (first_hist) common = s.klo;
    else {
        common = stat_key((T)v[3]);
        rem = KB - stat_clz((K)(common ^ stat_key((T)v[4])));
    }
    unsigned long long prefix = rem >= KB ? 0ull : (unsigned long long)(common >> rem);
    K k1 = common, upper = common;
    bool upper_known = true;
    bool have = first_hist;
    while (rem > 0) {
        const int bits = rem < kStatDigit ? rem : kStatDigit, shr = rem - bits;
        if (!have) {
            unsigned* hist = sc.hist[cl.hsel];
            for (int t = pm::tid(); t < kStatBins; t += kThreads) hist[t] = 0u;
            pm::sync();
            // keys under the prefix form one key range: intersect it with the set's own
            StatSet<T> sub = s;
            if (rem < KB) {
                const K plo = (K)(prefix << rem), phi = (K)(plo | (K)((((K)1) << rem) - (K)1));
                sub.klo = s.klo > plo ? s.klo : plo;
                sub.khi = s.khi < phi ? s.khi : phi;
            }
            const unsigned bmask = (1u << bits) - 1u;
            stat_for_each<T, true>(sub, [&](T, K key) { pm::atomic_add(&hist[(unsigned)(key >> shr) & bmask], 1u); });
            pm::sync();
            stat_coarse_sums(sc, cl.hsel);
            stat_pass_barrier(cl);
        }
        have = false;
        const bool last = shr == 0;
        const StatLoc loc = stat_locate(sh, sc, cl, cl.hsel, k, even && last);
        cl.hsel ^= 1u;
        const unsigned bin = loc.bin, nxt = loc.next;
        k = loc.left;
        if (last) {
            k1 = (K)((prefix << bits) | bin);
            if (even) {
                if (nxt != kNoBin) upper = (K)((prefix << bits) | nxt);
                else upper_known = false;
            } else upper = k1;
        }
        prefix = (prefix << bits) | bin;
        rem = shr;
        if (rem > 0 && loc.count <= (unsigned)kStatCollect && cl.size == 1) {
            // few keys are left under the prefix (64-bit keys of sparse data: after two or three digits): gather them and
            // rank them against each other instead of resolving the remaining digits pass by pass
            if (pm::tid() == 0) sc.nlist = 0u;
            pm::sync();
            StatSet<T> sub = s;
            const K plo = (K)(prefix << rem), phi = (K)(plo | (K)((((K)1) << rem) - (K)1));
            sub.klo = s.klo > plo ? s.klo : plo;
            sub.khi = s.khi < phi ? s.khi : phi;
            stat_for_each<T, true>(sub, [&](T, K key) { sc.list[pm::atomic_add(&sc.nlist, 1u)] = (unsigned long long)key; });
            pm::sync();
            const unsigned m = sc.nlist;                       // == loc.count
            for (unsigned i = (unsigned)pm::tid(); i < m; i += (unsigned)kThreads) {
                const unsigned long long mine_k = sc.list[i];
                unsigned rank = 0;
                for (unsigned j = 0; j < m; j++) {
                    const unsigned long long o = sc.list[j];
                    rank += (o < mine_k || (o == mine_k && j < i)) ? 1u : 0u;
                }
                if (rank == k) sc.found[0] = mine_k;
                if (rank == k + 1u) sc.found[1] = mine_k;
            }
            pm::sync();
            k1 = (K)sc.found[0];
            if (even) {
                if (k + 1u < m) upper = (K)sc.found[1];
                else upper_known = false;
            } else upper = k1;
            pm::sync();
            rem = 0;
        }
    }
    double med = stat_val(k1, T());
    double up = stat_val(upper, T());
    if (!upper_known) {      // (rare) the upper middle element is the smallest key above k1, outside the last prefix
        double nx = inf;
        stat_for_each(s, [&](T x, K key) { if (key > k1) nx = fmin(nx, (double)x); });
        double w[5] = {0.0, 0.0, 0.0, block_min(sh, nx), -inf};
        stat_exchange(sc, cl, w);
        up = w[3];
    }
    if (even) med = (med + up) / 2.0;       // np.median: mean of the two middle elements
    st.median = med;
    return st;
}

// ---- the cutout's rows of this CTA: copied to shared memory (16-byte vectors when the alignment allows)
struct alignas(16) StatVec4 { unsigned x, y, z, w; };
template <typename T> PM_DEV void stat_stage_rows(T* dst, const T* src, size_t count) {
    const size_t bytes = count * sizeof(T);
    if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0 && (bytes & 15u) == 0) {
        const StatVec4* s4 = (const StatVec4*)src;
        StatVec4* d4 = (StatVec4*)dst;
        const size_t n4 = bytes / 16;
        size_t i = pm::tid();
        for (; i + 3 * (size_t)kThreads < n4; i += 4 * (size_t)kThreads) {       // four vectors in flight per thread
            const StatVec4 v0 = s4[i], v1 = s4[i + kThreads], v2 = s4[i + 2 * kThreads], v3 = s4[i + 3 * kThreads];
            d4[i] = v0; d4[i + kThreads] = v1; d4[i + 2 * kThreads] = v2; d4[i + 3 * kThreads] = v3;
        }
        for (; i < n4; i += kThreads) d4[i] = s4[i];
    } else {
        for (size_t i = pm::tid(); i < count; i += kThreads) dst[i] = src[i];
    }
}

struct StatLayout { unsigned sc, mask, data, total; int rows_cap; };
// shared memory of one CTA: Sh | StatScratch | region plane of its rows (with_mask) | its rows (resident)
PM_HOSTDEV StatLayout stat_layout(int n, int cluster, size_t elem, bool with_mask, bool resident) {
    StatLayout L;
    const int nw = (n + 31) / 32;
    L.rows_cap = (n + cluster - 1) / cluster;
    unsigned o = st_align16((unsigned)sizeof(Sh));
    L.sc = o; o += st_align16((unsigned)sizeof(StatScratch));
    L.mask = o; if (with_mask) o += st_align16((unsigned)(L.rows_cap * nw * 4));
    L.data = o; if (resident) o += st_align16((unsigned)((size_t)L.rows_cap * n * elem));
    L.total = o;
    return L;
}

struct StatCtx {
    Sh* sh; StatScratch* sc; unsigned* mask; void* data;
    StatCl cl;
    int row_lo, rows;
    unsigned cid, ncl;
};
PM_DEV StatCtx stat_begin(int n, size_t elem, bool with_mask, bool resident) {
    unsigned char* smem = pm::dyn_smem(0);
    StatCtx c;
    c.cl.rank = pm::cluster_rank(); c.cl.size = pm::cluster_size(); c.cl.hsel = 0; c.cl.xsel = 0;
    c.cid = pm::cluster_id(); c.ncl = pm::cluster_count();
    const StatLayout L = stat_layout(n, (int)c.cl.size, elem, with_mask, resident);
    c.sh = (Sh*)smem;
    c.sc = (StatScratch*)(smem + L.sc);
    c.mask = (unsigned*)(smem + L.mask);
    c.data = smem + L.data;
    c.row_lo = (int)c.cl.rank * L.rows_cap;
    const int hi = c.row_lo + L.rows_cap < n ? c.row_lo + L.rows_cap : n;
    c.rows = hi > c.row_lo ? hi - c.row_lo : 0;
    c.sh->par[pm::tid()] = 0;
    pm::sync();
    return c;
}

// sky region = pixels whose centre is NOT inside the ellipse (photutils elliptical_overlap, subpixels = 1); local rows
PM_DEV void sky_build_mask(unsigned* mask, int n, int nw, int row_lo, int rows, double a, double b, double ca, double sa) {
    const int cen = n / 2 + 1;
    const double inv_a2 = 1.0 / (a * a), inv_b2 = 1.0 / (b * b);
    for (int lr = pm::warp(); lr < rows; lr += kWarps) {
        const double dy = (double)(row_lo + lr - cen);
        for (int wi = 0; wi < nw; wi++) {
            const int c = wi * 32 + pm::lane();
            const double dx = (double)(c - cen);
            const double xt = pm::dadd(pm::dmul(dy, sa), pm::dmul(dx, ca)), yt = pm::dsub(pm::dmul(dy, ca), pm::dmul(dx, sa));
            const bool inside = pm::dadd(pm::dmul(pm::dmul(xt, xt), inv_a2), pm::dmul(pm::dmul(yt, yt), inv_b2)) < 1.0;
            const unsigned w = pm::ballot(c < n && !inside);
            if (pm::lane() == 0) mask[lr * nw + wi] = w;
        }
    }
}

template <typename T>
PM_DEV void sky_body(const T* images, long long B, int n, const double* ellipse, double* out, int resident) {
    StatCtx c = stat_begin(n, sizeof(T), true, resident != 0);
    Sh& sh = *c.sh;
    StatScratch& sc = *c.sc;
    const double inf = 1.7976931348623157e308 * 2.0;
    for (long long b = c.cid; b < B; b += c.ncl) {
        const T* img = images + (size_t)b * n * n;
        StatSet<T> s;
        s.n = n; s.nw = (n + 31) / 32; s.rows = c.rows; s.mask = c.mask;
        s.data = img + (size_t)c.row_lo * n;
        if (resident) {
            stat_stage_rows((T*)c.data, s.data, (size_t)c.rows * n);
            s.data = (const T*)c.data;
        }
        // cos / sin of theta come from the caller's libm, like photutils
        sky_build_mask(c.mask, n, s.nw, c.row_lo, c.rows, ellipse[4 * b], ellipse[4 * b + 1], ellipse[4 * b + 2], ellipse[4 * b + 3]);
        pm::sync();
        s.set_bounds(-inf, inf);
        StatOut cur = stat_set(sh, sc, c.cl, s, (double)pm::ldg(img), false);
        const StatOut all = cur;
        double iters = 0.0, sky, err;
        const double status = all.cnt < 100.0 ? 1.0 : 0.0;          // skyBackground.py:147-148: region too small
        const bool clipped = all.cnt > 0.0 && all.mean > all.median;
        if (!clipped) { sky = all.mean; err = all.std; }
        else {
            double changed = 1.0;
            while (changed != 0.0 && iters < 5.0) {
                iters += 1.0;
                s.set_bounds(fmax(s.lo, cur.median - cur.std * 3.0), fmin(s.hi, cur.median + cur.std * 3.0));
                const StatOut nxt = stat_set(sh, sc, c.cl, s, cur.median, true);
                changed = cur.cnt - nxt.cnt;
                cur = nxt;
            }
            sky = 3.0 * cur.median - 2.0 * cur.mean;
            err = cur.std;
        }
        if (c.cl.rank == 0 && pm::tid() == 0) {
            double* o = out + (size_t)b * kSkyCols;
            o[0] = all.cnt; o[1] = all.mean; o[2] = all.median; o[3] = all.std; o[4] = iters;
            o[5] = cur.cnt; o[6] = cur.mean; o[7] = cur.median; o[8] = cur.std; o[9] = sky; o[10] = err; o[11] = status;
        }
        pm::sync();      // the rows and the plane are overwritten by the next cutout
    }
    if (c.cl.size > 1) pm::cluster_sync();      // nobody leaves while its shared memory may still be read
}

// snap_at > 0 with out_snap: the row as it stands after snap_at rounds (or at convergence, if that comes first) goes to out_snap —
// sigma_clipped_stats(maxiters=snap_at) and (maxiters=maxiters) of the same image from one sequence of rounds
// (imageutils.py:62 and :67 need maxiters = 5 and the converged statistics of the same image).
template <typename T>
PM_DEV void clipstats_body(const T* images, long long B, int n, double sigma, int maxiters, double* out, int resident,
                           int snap_at = 0, double* out_snap = nullptr) {
    StatCtx c = stat_begin(n, sizeof(T), false, resident != 0);
    Sh& sh = *c.sh;
    StatScratch& sc = *c.sc;
    const double inf = 1.7976931348623157e308 * 2.0;
    for (long long b = c.cid; b < B; b += c.ncl) {
        const T* img = images + (size_t)b * n * n;
        StatSet<T> s;
        s.n = n; s.nw = (n + 31) / 32; s.rows = c.rows; s.mask = nullptr;
        s.data = img + (size_t)c.row_lo * n;
        if (resident) {
            stat_stage_rows((T*)c.data, s.data, (size_t)c.rows * n);
            s.data = (const T*)c.data;
            pm::sync();
        }
        s.set_bounds(-inf, inf);
        StatOut cur = stat_set(sh, sc, c.cl, s, (double)pm::ldg(img), false);
        const StatOut all = cur;
        // astropy SigmaClip: bounds median -/+ sigma * std (inclusive) from the surviving values, until a round removes
        // nothing or maxiters rounds are done (maxiters <= 0: until convergence)
        double iters = 0.0, changed = 1.0;
        bool snapped = false;
        auto write_row = [&](double* o) {
            o[0] = all.cnt; o[1] = all.mean; o[2] = all.median; o[3] = all.std; o[4] = iters;
            o[5] = cur.cnt; o[6] = cur.mean; o[7] = cur.median; o[8] = cur.std; o[9] = all.cnt > 0.0 ? all.vmax : -inf;
        };
        while (changed != 0.0 && (maxiters <= 0 || iters < (double)maxiters) && cur.cnt > 0.0) {
            if (out_snap && !snapped && iters >= (double)snap_at) {
                if (c.cl.rank == 0 && pm::tid() == 0) write_row(out_snap + (size_t)b * kClipCols);
                snapped = true;
            }
            iters += 1.0;
            s.set_bounds(fmax(s.lo, cur.median - cur.std * sigma), fmin(s.hi, cur.median + cur.std * sigma));
            const StatOut nxt = stat_set(sh, sc, c.cl, s, cur.median, true);
            changed = cur.cnt - nxt.cnt;
            cur = nxt;
        }
        if (c.cl.rank == 0 && pm::tid() == 0) {
            if (out_snap && !snapped) write_row(out_snap + (size_t)b * kClipCols);
            write_row(out + (size_t)b * kClipCols);
        }
        pm::sync();
    }
    if (c.cl.size > 1) pm::cluster_sync();
}

}  // namespace PM_NS
