// pm_sky.cuh — sky-region statistics of skyBackground._calcSkybgr (skyBackground.py:134-164) for a batch of cutouts:
// the pixels whose centre lies outside the ellipse of semi-axes a = 2 min(fwhm), b = 2 max(fwhm), angle theta about
// (int(n/2) + 1, int(n/2) + 1) (`_getSkyRegion`, :265-303: photutils EllipticalAperture.to_mask(method="center")),
// their count / mean / median / population std (np.ma.mean / median / std, :153-155), and, when mean > median (the
// crowded branch, :161-164), astropy's sigma_clipped_stats(sigma=3, maxiters=5, cenfunc=median, stdfunc=std) with
// sky = 3 median - 2 mean.  The Gaussian fit that supplies (fwhms, theta) (:306-395) is not part of this kernel.
// One CTA per cutout; the image is re-read from L2 for every pass (it does not fit shared memory at 256^2):
// two reduction passes (mean, then squared deviations, like np.std) and an 8-bit radix select per statistics set.
#pragma once
#include "pm_kernels.cuh"

namespace pmk {

constexpr int kSkyCols = 12;   // count, mean, median, std, iterations, clipped count, mean, median, std, sky, sky_err, status

struct SkyScratch {
    unsigned hist[256];
    unsigned long long prefix;   // key bits decided so far (high bytes)
    unsigned k;                  // rank still to find inside the current prefix
    unsigned pad;
};

PM_DEV unsigned sky_key(float x) { const unsigned u = pm::f2bits(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
PM_DEV unsigned long long sky_key(double x) {
    const unsigned long long u = (unsigned long long)pm::d2bits(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
PM_DEV double sky_val(unsigned k, float) { return (double)pm::bits2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
PM_DEV double sky_val(unsigned long long k, double) {
    return pm::bits2d((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

template <typename T> struct SkySet {      // the statistics set: outside the ellipse and inside [lo, hi]
    const T* img;
    int n, cen;
    double ca, sa, inv_a2, inv_b2, lo, hi;
    PM_DEV bool has(int i, T& x) const {
        const int r = i / n, c = i - r * n;
        const double dx = (double)(c - cen), dy = (double)(r - cen);
        const double xt = pm::dadd(pm::dmul(dx, ca), pm::dmul(dy, sa)), yt = pm::dsub(pm::dmul(dy, ca), pm::dmul(dx, sa));
        if (pm::dadd(pm::dmul(pm::dmul(xt, xt), inv_a2), pm::dmul(pm::dmul(yt, yt), inv_b2)) < 1.0) return false;
        x = pm::ldg(img + i);
        const double v = (double)x;
        return v >= lo && v <= hi;
    }
};

struct SkyStats { double cnt, mean, median, std; };

// k-th smallest (0-based) key of the set
template <typename T, typename K> PM_DEV K sky_select(Sh& sh, SkyScratch& sc, const SkySet<T>& s, unsigned k) {
    constexpr int kBytes = (int)sizeof(K);
    if (pm::tid() == 0) { sc.prefix = 0ull; sc.k = k; }
    for (int pass = 0; pass < kBytes; pass++) {
        const int shift = 8 * (kBytes - 1 - pass);
        for (int t = pm::tid(); t < 256; t += kThreads) sc.hist[t] = 0u;
        pm::sync();
        const unsigned long long prefix = sc.prefix;
        for (int i = pm::tid(); i < s.n * s.n; i += kThreads) {
            T x;
            if (!s.has(i, x)) continue;
            const K key = sky_key(x);
            if (pass == 0 || (unsigned long long)(key >> (shift + 8)) == prefix) pm::atomic_add(&sc.hist[(unsigned)(key >> shift) & 255u], 1u);
        }
        pm::sync();
        if (pm::tid() == 0) {
            unsigned kk = sc.k, bin = 0;
            for (; bin < 255; bin++) {
                if (kk < sc.hist[bin]) break;
                kk -= sc.hist[bin];
            }
            sc.k = kk;
            sc.prefix = (sc.prefix << 8) | bin;
        }
        pm::sync();
    }
    (void)sh;
    return (K)sc.prefix;
}

template <typename T> PM_DEV SkyStats sky_stats(Sh& sh, SkyScratch& sc, const SkySet<T>& s) {
    typedef decltype(sky_key(T())) K;
    double a[2] = {0.0, 0.0};
    for (int i = pm::tid(); i < s.n * s.n; i += kThreads) {
        T x;
        if (s.has(i, x)) { a[0] += 1.0; a[1] += (double)x; }
    }
    block_sum<2>(sh, a);
    SkyStats st;
    st.cnt = a[0];
    if (a[0] == 0.0) { st.mean = st.median = st.std = pm::bits2d(0x7ff8000000000000ll); return st; }
    st.mean = a[1] / a[0];
    double q = 0.0;
    for (int i = pm::tid(); i < s.n * s.n; i += kThreads) {
        T x;
        if (s.has(i, x)) { const double d = (double)x - st.mean; q += d * d; }
    }
    q = block_sum1(sh, q);
    st.std = pm::dsqrt(q / a[0]);
    const unsigned cnt = (unsigned)a[0];
    const K k1 = sky_select<T, K>(sh, sc, s, (cnt - 1) / 2);
    double med = sky_val(k1, T());
    if ((cnt & 1u) == 0u) {       // even count: mean of the two middle values; the upper one is k1 again or the next larger key
        unsigned le = 0;
        double nxt = 1.7976931348623157e308 * 2.0;
        for (int i = pm::tid(); i < s.n * s.n; i += kThreads) {
            T x;
            if (!s.has(i, x)) continue;
            const K key = sky_key(x);
            if (key <= k1) le++; else nxt = fmin(nxt, (double)x);
        }
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
    unsigned char* smem = pm::dyn_smem();
    Sh& sh = *(Sh*)smem;
    SkyScratch& sc = *(SkyScratch*)(smem + align16((unsigned)sizeof(Sh)));
    sh.par[pm::tid()] = 0;
    pm::sync();
    const double inf = 1.7976931348623157e308 * 2.0;
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        SkySet<T> s;
        s.img = images + (size_t)b * n * n;
        s.n = n;
        s.cen = n / 2 + 1;
        const double a = ellipse[4 * b], bb = ellipse[4 * b + 1];
        s.ca = ellipse[4 * b + 2]; s.sa = ellipse[4 * b + 3];       // cos / sin of theta from the caller's libm, like photutils
        s.inv_a2 = 1.0 / (a * a); s.inv_b2 = 1.0 / (bb * bb);
        s.lo = -inf; s.hi = inf;
        SkyStats cur = sky_stats(sh, sc, s);
        const SkyStats all = cur;
        double iters = 0.0, sky, err;
        double status = all.cnt < 100.0 ? 1.0 : 0.0;          // skyBackground.py:147-148: region too small
        const bool clipped = all.cnt > 0.0 && all.mean > all.median;
        if (!clipped) { sky = all.mean; err = all.std; }
        else {
            double changed = 1.0;
            while (changed != 0.0 && iters < 5.0) {
                iters += 1.0;
                s.lo = fmax(s.lo, cur.median - cur.std * 3.0);
                s.hi = fmin(s.hi, cur.median + cur.std * 3.0);
                const SkyStats nxt = sky_stats(sh, sc, s);
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
PM_HOSTDEV unsigned sky_smem_bytes() { return align16((unsigned)sizeof(Sh)) + align16((unsigned)sizeof(SkyScratch)); }

}  // namespace pmk
