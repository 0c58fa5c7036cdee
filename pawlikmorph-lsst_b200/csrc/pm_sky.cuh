// pm_sky.cuh — sky-region statistics of skyBackground._calcSkybgr (skyBackground.py:134-164) for a batch of cutouts:
// the pixels whose centre lies outside the ellipse of semi-axes a = 2 min(fwhm), b = 2 max(fwhm), angle theta about
// (int(n/2) + 1, int(n/2) + 1) (`_getSkyRegion`, :265-303: photutils EllipticalAperture.to_mask(method="center")),
// their count / mean / median / population std (np.ma.mean / median / std, :153-155), and, when mean > median (the
// crowded branch, :161-164), astropy's sigma_clipped_stats(sigma=3, maxiters=5, cenfunc=median, stdfunc=std) with
// sky = 3 median - 2 mean.  The Gaussian fit that supplies (fwhms, theta) (:306-395) is not part of this kernel.
// One CTA per cutout.  The sky region is a bit plane in shared memory (built once); the image is re-read from L2 for
// every pass (it does not fit shared memory at 256^2).  Per statistics set: one pass for count / sums / key range, then
// a radix select of 12 bits per pass over the bits the smallest and largest key do NOT share (two passes for typical
// sky pixels), plus one pass for the upper middle element of an even count.
#pragma once
#include "pm_kernels.cuh"

namespace pmk {

constexpr int kSkyCols = 12;   // count, mean, median, std, iterations, clipped count, mean, median, std, sky, sky_err, status
constexpr int kSkyDigit = 12;  // bits resolved per select pass
constexpr int kSkyBins = 1 << kSkyDigit;

struct SkyScratch {
    unsigned hist[kSkyBins];
    unsigned long long prefix;   // key bits decided so far
    unsigned k;                  // rank still to find inside the current prefix
    unsigned pad;
};

// order-preserving keys; -0.0 is filed as +0.0 (numpy compares them equal)
PM_DEV unsigned sky_key(float x) { x = x == 0.0f ? 0.0f : x; const unsigned u = pm::f2bits(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
PM_DEV unsigned long long sky_key(double x) {
    x = x == 0.0 ? 0.0 : x;
    const unsigned long long u = (unsigned long long)pm::d2bits(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
PM_DEV double sky_val(unsigned k, float) { return (double)pm::bits2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
PM_DEV double sky_val(unsigned long long k, double) {
    return pm::bits2d((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
PM_DEV int sky_clz(unsigned v) { return v ? pm::clz(v) : 32; }
PM_DEV int sky_clz(unsigned long long v) { return (v >> 32) ? pm::clz((unsigned)(v >> 32)) : 32 + sky_clz((unsigned)v); }

template <typename T> struct SkySet {      // the statistics set: sky-region pixels (bit plane in shared memory) inside [lo, hi]
    const T* img;
    const unsigned* mask;
    int n, nw;
    double lo, hi;
    decltype(sky_key(T())) klo, khi;   // [lo, hi] as a key range: membership is two integer compares
    PM_DEV void set_bounds(double lo_, double hi_) {
        typedef decltype(sky_key(T())) K;
        lo = lo_; hi = hi_;
        klo = sky_key((T)lo_);
        if (sky_val(klo, T()) < lo_) klo += (K)1;        // (T)lo rounded down: the next key up is the first one inside
        khi = sky_key((T)hi_);
        if (sky_val(khi, T()) > hi_) khi -= (K)1;
    }
};

// f(x) for every member of the set: one row per warp, 32 columns per word; the loads of kSkyGroup words are issued
// together (unconditionally, the object pixels are few) so that a lane has that many L2 requests in flight
// (Measured alternatives at 256^2: the plane walked as one flat list of words, 8 per step: 417k cutouts/s; 16 per step:
// 127 registers, 2 CTAs per SM, 419k; this row form: 516k; one word per step with a predicated load: 180k.)
constexpr int kSkyGroup = 8;
template <typename T, typename F> PM_DEV void sky_for_each(const SkySet<T>& s, F&& f) {
    for (int r = pm::warp(); r < s.n; r += kWarps) {
        const T* row = s.img + (size_t)r * s.n;
        const unsigned* mrow = s.mask + r * s.nw;
        for (int w0 = 0; w0 < s.nw; w0 += kSkyGroup) {
            T xv[kSkyGroup];
            unsigned bits[kSkyGroup];
#pragma unroll
            for (int j = 0; j < kSkyGroup; j++) {
                const int wi = w0 + j, c = wi * 32 + pm::lane();
                const bool ok = wi < s.nw && c < s.n;
                xv[j] = ok ? pm::ldg(row + c) : (T)0;
                bits[j] = wi < s.nw ? mrow[wi] : 0u;
            }
#pragma unroll
            for (int j = 0; j < kSkyGroup; j++) {
                if ((bits[j] >> pm::lane()) & 1u) {
                    const auto key = sky_key(xv[j]);
                    if (key >= s.klo && key <= s.khi) f(xv[j], key);     // NaNs sort outside every range
                }
            }
        }
    }
}

// sky region = pixels whose centre is NOT inside the ellipse (photutils elliptical_overlap, subpixels = 1)
PM_DEV void sky_build_mask(unsigned* mask, int n, int nw, double a, double b, double ca, double sa) {
    const int cen = n / 2 + 1;
    const double inv_a2 = 1.0 / (a * a), inv_b2 = 1.0 / (b * b);
    for (int r = pm::warp(); r < n; r += kWarps) {
        const double dy = (double)(r - cen);
        for (int wi = 0; wi < nw; wi++) {
            const int c = wi * 32 + pm::lane();
            const double dx = (double)(c - cen);
            const double xt = pm::dadd(pm::dmul(dy, sa), pm::dmul(dx, ca)), yt = pm::dsub(pm::dmul(dy, ca), pm::dmul(dx, sa));
            const bool inside = pm::dadd(pm::dmul(pm::dmul(xt, xt), inv_a2), pm::dmul(pm::dmul(yt, yt), inv_b2)) < 1.0;
            const unsigned w = pm::ballot(c < n && !inside);
            if (pm::lane() == 0) mask[r * nw + wi] = w;
        }
    }
    pm::sync();
}

struct SkyStats { double cnt, mean, median, std; };

// k-th smallest (0-based) key of the set, all of whose keys share the bits above `rem` with `common`
template <typename T, typename K> PM_DEV K sky_select(Sh& sh, SkyScratch& sc, const SkySet<T>& s, unsigned k, K common, int rem) {
    if (pm::tid() == 0) { sc.prefix = rem >= (int)(8 * sizeof(K)) ? 0ull : (unsigned long long)(common >> rem); sc.k = k; }
    pm::sync();
    while (rem > 0) {
        const int bits = rem < kSkyDigit ? rem : kSkyDigit, shift = rem - bits;
        for (int t = pm::tid(); t < kSkyBins; t += kThreads) sc.hist[t] = 0u;
        pm::sync();
        const unsigned long long prefix = sc.prefix;
        const bool all = rem >= (int)(8 * sizeof(K));
        sky_for_each(s, [&](T, K key) {
            if (all || (unsigned long long)(key >> rem) == prefix) pm::atomic_add(&sc.hist[(unsigned)(key >> shift) & ((1u << bits) - 1u)], 1u);
        });
        pm::sync();
        // the bin holding rank k: every thread sums a run of bins, a block prefix finds the run, its owner walks it
        constexpr int kRun = kSkyBins / kThreads;
        static_assert(kSkyBins % kThreads == 0, "bins per thread");
        unsigned mine = 0;
        for (int j = 0; j < kRun; j++) mine += sc.hist[pm::tid() * kRun + j];
        unsigned total;
        const unsigned before = block_excl_prefix_u(sh, mine, &total);
        const unsigned kk = sc.k;
        pm::sync();
        if (kk >= before && kk < before + mine) {
            unsigned left = kk - before;
            int j = 0;
            for (; j < kRun - 1; j++) {
                const unsigned h = sc.hist[pm::tid() * kRun + j];
                if (left < h) break;
                left -= h;
            }
            sc.k = left;
            sc.prefix = (prefix << bits) | (unsigned)(pm::tid() * kRun + j);
        }
        pm::sync();
        rem = shift;
    }
    return (K)sc.prefix;
}

// count, mean, population std and median of the set.  The sums run on x - shift (any value near the data: the squared
// deviations then come from one pass, sum (x-c)^2 - (sum (x-c))^2 / cnt, without cancellation worth speaking of).
template <typename T> PM_DEV SkyStats sky_stats(Sh& sh, SkyScratch& sc, const SkySet<T>& s, double shift) {
    typedef decltype(sky_key(T())) K;
    double a[3] = {0.0, 0.0, 0.0};
    K kmin = ~(K)0, kmax = (K)0;
    unsigned mine = 0;
    sky_for_each(s, [&](T x, K key) {
        const double d = (double)x - shift;
        mine++; a[1] += d; a[2] += d * d;
        kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
    });
    a[0] = (double)mine;
    block_sum<3>(sh, a);
    SkyStats st;
    st.cnt = a[0];
    if (a[0] == 0.0) { st.mean = st.median = st.std = pm::bits2d(0x7ff8000000000000ll); return st; }
    const double md = a[1] / a[0];
    st.mean = shift + md;
    st.std = pm::dsqrt(fmax(a[2] / a[0] - md * md, 0.0));
    // smallest / largest key: (value, as double) min / max are order-equivalent
    const double vmin = block_min(sh, sky_val(kmin, T())), vmax = block_max(sh, sky_val(kmax, T()));
    const K gmin = sky_key((T)vmin), gmax = sky_key((T)vmax);
    const int rem = (int)(8 * sizeof(K)) - sky_clz((K)(gmin ^ gmax));
    const unsigned cnt = (unsigned)a[0];
    const K k1 = sky_select<T, K>(sh, sc, s, (cnt - 1) / 2, gmin, rem);
    double med = sky_val(k1, T());
    if ((cnt & 1u) == 0u) {       // even count: mean of the two middle values; the upper one is k1 again or the next larger key
        unsigned le = 0;
        double nxt = 1.7976931348623157e308 * 2.0;
        sky_for_each(s, [&](T x, K key) {
            if (key <= k1) le++; else nxt = fmin(nxt, (double)x);
        });
        le = block_sum_u(sh, le);
        nxt = block_min(sh, nxt);
        const double upper = le > cnt / 2 ? med : nxt;
        med = (med + upper) / 2.0;       // np.median: mean of the two middle elements
    }
    st.median = med;
    return st;
}

template <typename T>
PM_DEV void sky_body(const T* images, long long B, int n, const double* ellipse, double* out) {
    unsigned char* smem = pm::dyn_smem(0);
    Sh& sh = *(Sh*)smem;
    SkyScratch& sc = *(SkyScratch*)(smem + align16((unsigned)sizeof(Sh)));
    unsigned* mask = (unsigned*)(smem + align16((unsigned)sizeof(Sh)) + align16((unsigned)sizeof(SkyScratch)));
    sh.par[pm::tid()] = 0;
    pm::sync();
    const double inf = 1.7976931348623157e308 * 2.0;
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        SkySet<T> s;
        s.img = images + (size_t)b * n * n;
        s.mask = mask;
        s.n = n;
        s.nw = (n + 31) / 32;
        // cos / sin of theta come from the caller's libm, like photutils
        sky_build_mask(mask, n, s.nw, ellipse[4 * b], ellipse[4 * b + 1], ellipse[4 * b + 2], ellipse[4 * b + 3]);
        s.set_bounds(-inf, inf);
        SkyStats cur = sky_stats(sh, sc, s, (double)pm::ldg(s.img));
        const SkyStats all = cur;
        double iters = 0.0, sky, err;
        double status = all.cnt < 100.0 ? 1.0 : 0.0;          // skyBackground.py:147-148: region too small
        const bool clipped = all.cnt > 0.0 && all.mean > all.median;
        if (!clipped) { sky = all.mean; err = all.std; }
        else {
            double changed = 1.0;
            while (changed != 0.0 && iters < 5.0) {
                iters += 1.0;
                s.set_bounds(fmax(s.lo, cur.median - cur.std * 3.0), fmin(s.hi, cur.median + cur.std * 3.0));
                const SkyStats nxt = sky_stats(sh, sc, s, cur.median);
                changed = cur.cnt - nxt.cnt;
                cur = nxt;
            }
            sky = 3.0 * cur.median - 2.0 * cur.mean;
            err = cur.std;
        }
        if (pm::tid() == 0) {
            double* o = out + (size_t)b * kSkyCols;
            o[0] = all.cnt; o[1] = all.mean; o[2] = all.median; o[3] = all.std; o[4] = iters;
            o[5] = cur.cnt; o[6] = cur.mean; o[7] = cur.median; o[8] = cur.std; o[9] = sky; o[10] = err; o[11] = status;
        }
        pm::sync();
    }
}
PM_HOSTDEV unsigned sky_smem_bytes(int n) {
    return align16((unsigned)sizeof(Sh)) + align16((unsigned)sizeof(SkyScratch)) + align16((unsigned)(n * ((n + 31) / 32) * 4));
}

}  // namespace pmk
