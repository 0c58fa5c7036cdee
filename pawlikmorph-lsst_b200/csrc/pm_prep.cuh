// pm_prep.cuh — the pre-processing around the morphology path, one CTA per cutout:
//   (the clipped statistics that feed both — astropy.stats.sigma_clipped_stats — are pm_stats.cuh)
//   * gaussfit_body    skyBackground._gauss2dfit (skyBackground.py:306-331): gaussfitter.moments start values and the
//                      least-squares fit of gaussfitter.twodgaussian that scipy.optimize.curve_fit performs, as a
//                      Levenberg-Marquardt iteration with the analytic Jacobian
//   * maskstars_body   photutils.detect_sources as skyBackground.py:110 / imageutils.py:71 call it (3x3 Gaussian smooth,
//                      strict threshold, 8-connected labels of >= npixels pixels) and the cleaning step of
//                      imageutils.maskstarsSEG (:42-91) with a counter-based (Philox) normal generator in place of the
//                      reference's unseeded np.random.normal
#pragma once
#include "pm_kernels.cuh"

namespace pmk {

// ----------------------------------------------------------------------------------------------------
// Least-squares fits of 7-parameter surface models: the Gaussian of skybgr, the Sersic profile of fitSersic
// ----------------------------------------------------------------------------------------------------
// out: the 7 parameters (the raw optimum), final cost (sum of squared residuals), iterations, status (0 converged,
//      1 iteration limit — curve_fit's RuntimeError —, 2 NaN in the start values / first residuals — gaussfitter.py:76-77 —,
//      3 singular normal equations)
constexpr int kGaussCols = 10;
constexpr int kGaussMaxIter = 200;

// Cholesky solve of the 7x7 system (A + par * diag(dg)^2) d = g; returns false when it is not positive definite
PM_DEV bool lm_solve7(const double* A /* packed lower triangle, 28 */, const double* g, double par, const double* dg, double* d) {
    double L[28];
    for (int i = 0; i < 7; i++) {
        for (int j = 0; j <= i; j++) {
            double sum = A[i * (i + 1) / 2 + j];
            if (i == j) sum += par * dg[i] * dg[i];
            for (int k = 0; k < j; k++) sum -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            if (i == j) {
                if (!(sum > 0.0) || !(sum < 1.7e308)) return false;
                L[i * (i + 1) / 2 + i] = sqrt(sum);
            } else {
                L[i * (i + 1) / 2 + j] = sum / L[j * (j + 1) / 2 + j];
            }
        }
    }
    double y[7];
    for (int i = 0; i < 7; i++) {
        double sum = g[i];
        for (int k = 0; k < i; k++) sum -= L[i * (i + 1) / 2 + k] * y[k];
        y[i] = sum / L[i * (i + 1) / 2 + i];
    }
    for (int i = 6; i >= 0; i--) {
        double sum = y[i];
        for (int k = i + 1; k < 7; k++) sum -= L[k * (k + 1) / 2 + i] * d[k];
        d[i] = sum / L[i * (i + 1) / 2 + i];
    }
    return true;
}

// Shared memory of a fit behind Sh: per warp a staging tile of the 8 augmented Jacobian columns (J0..J6, r) of its 32
// pixels, [8][kGaussTileStride] doubles (stride 36: the tensor-core fragment reads of a warp then take the minimum of two
// wavefronts), the warps' 8 x 8 products, their sum, and the next trial point posted by warp 0.
constexpr int kGaussTileStride = 36;
struct GaussSmem {
    double tile[kWarps][8 * kGaussTileStride];
    double prod[kWarps][64];
    double total[64];
    double trial[8];
    int go;
};

// One pass over the cutout at parameters q: cost = sum r^2, J^T J, J^T r for r = model - data.  With the residual as an
// eighth column, all 36 sums are the upper triangle of ONE 8 x 8 product [J r]^T [J r] accumulated over the pixels — a
// rank update, done on the FP64 tensor cores (DMMA.8x8x4: four pixels per instruction, the 8 x 8 accumulator spread
// over the warp, two registers per lane instead of 36 accumulators each): a lane evaluates the model at its pixel,
// the warp transposes the 32 x 8 values through the staging tile, eight DMMAs add their outer products.
// Model: setup(q, n) fixes the per-pass constants and the box [r0, r1] x [c0, c1] of pixels to visit; columns(row, col,
// dat, J) fills J0..J6 and the residual J7; finish(acc, ...) adds what the pixels outside the box contribute.
template <typename T, typename Model>
PM_DEV void fit_pass(Sh& sh, GaussSmem& gs, const T* img, int n, Model& model, const double* q, double* cost, double* jtj, double* jtr) {
    model.setup(q, n);
    const int r0 = model.r0, c0 = model.c0;
    const int bw = model.c1 >= model.c0 ? model.c1 - model.c0 + 1 : 0, bh = model.r1 >= model.r0 ? model.r1 - model.r0 + 1 : 0;
    const int total = bw * bh;
    // the pixel values of the next two steps are requested before the current one is worked on (a step is ~70 FP64
    // instructions, an L2 read several hundred cycles).  Box coordinates advance by the thread count per step without
    // divisions: (row, col) += (kThreads / bw, kThreads % bw) with one carry.
    const int bw1 = bw > 0 ? bw : 1;
    const int step_r = kThreads / bw1, step_c = kThreads - step_r * bw1;
    struct Pos { int r, c; };
    auto advance = [&](Pos& p) { p.r += step_r; p.c += step_c; if (p.c >= bw1) { p.c -= bw1; p.r++; } };
    auto fetch = [&](const Pos& p) -> T { return p.r < bh ? pm::ldg(img + (size_t)(r0 + p.r) * n + (c0 + p.c)) : (T)0; };
    double* tile = gs.tile[pm::warp()];
    const int lane = pm::lane(), g = lane >> 2, t = lane & 3;
    double d0 = 0.0, d1 = 0.0;       // this lane's two entries of the 8 x 8 product
    Pos cur, nxt;
    cur.r = pm::tid() / bw1; cur.c = pm::tid() - cur.r * bw1;
    nxt = cur;
    T ahead1 = fetch(nxt);
    advance(nxt);
    T ahead2 = fetch(nxt);
    for (int i = pm::tid(); i - lane < total; i += kThreads) {          // (whole warps: the product is a warp operation)
        const double dat = (double)ahead1;
        ahead1 = ahead2;
        advance(nxt);
        ahead2 = fetch(nxt);
        double J[8];
        if (i < total) model.columns(r0 + cur.r, c0 + cur.c, dat, J);
        else {
#pragma unroll
            for (int k = 0; k < 8; k++) J[k] = 0.0;
        }
        advance(cur);
#pragma unroll
        for (int k = 0; k < 8; k++) tile[k * kGaussTileStride + lane] = J[k];
        pm::syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double a = tile[g * kGaussTileStride + 4 * k + t];      // column g of pixel 4 k + t: A[g][t] and B[t][g] alike
            pm::dmma_8x8x4(a, a, d0, d1);
        }
        pm::syncwarp();
    }
    // the warps' products, summed in warp order
    gs.prod[pm::warp()][g * 8 + 2 * t] = d0;
    gs.prod[pm::warp()][g * 8 + 2 * t + 1] = d1;
    double bs[2] = {model.box_s1, model.box_s2};
    block_sum<2>(sh, bs);                  // (its barrier also orders the products)
    if (pm::tid() < 64) {
        double v = 0.0;
        for (int w = 0; w < kWarps; w++) v += gs.prod[w][pm::tid()];
        gs.total[pm::tid()] = v;
    }
    pm::sync();
    double acc[36];
    for (int a = 0; a < 7; a++) {
        acc[28 + a] = gs.total[a * 8 + 7];
        for (int b2 = 0; b2 <= a; b2++) acc[a * (a + 1) / 2 + b2] = gs.total[a * 8 + b2];
    }
    acc[35] = gs.total[63];
    pm::sync();                            // gs.total is rewritten by the next pass
    model.finish(acc, n, bw, bh, bs);
    *cost = acc[35];
    for (int k = 0; k < 28; k++) jtj[k] = acc[k];
    for (int k = 0; k < 7; k++) jtr[k] = acc[28 + k];
}

// Levenberg-Marquardt with MINPACK's trust-region logic (lmder: scaling by the running maximum of the column norms,
// step bound delta, ratio of actual to predicted reduction steering delta and the damping parameter) — the algorithm
// behind scipy's curve_fit / leastsq with the analytic Jacobian, stopped at `tol` where they stop at 1.5e-8 (the
// iteration converges linearly on noisy frames: 1e-10 leaves the Gaussian widths within 2e-5 of the optimum, curve_fit's
// own stop within 2e-4; 1e-13 costs 20 % more steps).  An undamped or merely Marquardt-damped iteration walks off to
// needle-shaped Gaussians from the start values of gaussfitter.moments.
// The passes over the pixels are the whole CTA's; the 7 x 7 algebra between them is warp 0's alone (the other warps wait
// at the barrier and leave the issue slots to the CTAs sharing the SM): it posts the next trial point, or the end, in
// gs.trial / gs.go.  p: start values in, optimum out (meaningful in warp 0); returns the status (warp 0).
template <typename T, typename Model>
PM_DEV int lm_minimise(Sh& sh, GaussSmem& gs, const T* img, int n, Model& model, double* p, double tol, double* cost_out, int* iter_out) {
    double cost = 0.0, jtj[28], jtr[7];
    int iter = 0, status = 1;
    const bool lead = pm::warp() == 0;
    fit_pass(sh, gs, img, n, model, p, &cost, jtj, jtr);
    if (!(cost == cost) || !(cost < 1.7e308)) status = 2;
    double dg[7], d[7], xnorm = 0.0, par = 0.0, pnorm = 0.0, delta = 100.0;
    if (lead) {
        for (int k = 0; k < 7; k++) {
            const double a = jtj[k * (k + 1) / 2 + k];
            dg[k] = a > 0.0 ? sqrt(a) : 1.0;
            xnorm += dg[k] * p[k] * dg[k] * p[k];
            d[k] = 0.0;
        }
        xnorm = sqrt(xnorm);
        delta = xnorm > 0.0 ? 100.0 * xnorm : 100.0;
        if (status == 1 && cost == 0.0) status = 0;
    }
    for (;;) {
        if (lead) {
            bool go = status == 1 && iter < kGaussMaxIter;
            if (go) {
                double maxratio = 0.0;
                for (int k = 0; k < 7; k++) {
                    const double a = jtj[k * (k + 1) / 2 + k];
                    if (a > 0.0) dg[k] = fmax(dg[k], sqrt(a));
                    maxratio = fmax(maxratio, a / (dg[k] * dg[k]));
                }
                double g[7];
                for (int k = 0; k < 7; k++) g[k] = -jtr[k];
                bool solved = false;
                for (int inner = 0; inner < 80; inner++) {
                    if (!lm_solve7(jtj, g, par, dg, d)) { par = fmax(par * 4.0, 1e-12 * fmax(maxratio, 1e-300)); continue; }
                    pnorm = 0.0;
                    for (int k = 0; k < 7; k++) pnorm += dg[k] * d[k] * dg[k] * d[k];
                    pnorm = sqrt(pnorm);
                    solved = true;
                    if (pnorm <= 1.1 * delta) break;
                    par = fmax(par * 4.0, 1e-8 * fmax(maxratio, 1e-300));
                }
                if (!solved) { status = 3; go = false; }
            }
            if (pm::lane() == 0) {
                gs.go = go ? 1 : 0;
                for (int k = 0; k < 7; k++) gs.trial[k] = p[k] + d[k];
            }
        }
        pm::sync();
        if (!gs.go) break;
        double q[7], cost_q, jtj_q[28], jtr_q[7];
        for (int k = 0; k < 7; k++) q[k] = gs.trial[k];
        fit_pass(sh, gs, img, n, model, q, &cost_q, jtj_q, jtr_q);
        if (lead) {
            if (iter == 0) delta = fmin(delta, pnorm);
            const bool finite_q = cost_q == cost_q && cost_q < 1.7e308;
            const double actred = finite_q ? 1.0 - cost_q / cost : -1.0;
            double jd2 = 0.0;                             // |J d|^2 = d^T (J^T J) d
            for (int a = 0; a < 7; a++)
                for (int b2 = 0; b2 < 7; b2++) jd2 += d[a] * d[b2] * jtj[a >= b2 ? a * (a + 1) / 2 + b2 : b2 * (b2 + 1) / 2 + a];
            const double t1 = jd2 / cost, t2 = par * pnorm * pnorm / cost;
            const double prered = t1 + 2.0 * t2, dirder = -(t1 + t2);
            const double ratio = prered != 0.0 ? actred / prered : 0.0;
            if (ratio <= 0.25) {
                double temp = actred >= 0.0 ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (!finite_q || 0.1 * cost_q >= cost || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                for (int k = 0; k < 7; k++) p[k] = q[k];
                cost = cost_q;
                for (int k = 0; k < 28; k++) jtj[k] = jtj_q[k];
                for (int k = 0; k < 7; k++) jtr[k] = jtr_q[k];
                xnorm = 0.0;
                for (int k = 0; k < 7; k++) xnorm += dg[k] * p[k] * dg[k] * p[k];
                xnorm = sqrt(xnorm);
            }
            if ((fabs(actred) <= tol && prered <= tol && 0.5 * ratio <= 1.0) || delta <= tol * xnorm || cost == 0.0) status = 0;
            iter++;
        }
    }
    *cost_out = cost;
    *iter_out = iter;
    return status;
}

// ---- the Gaussian of skyBackground._gauss2dfit (skyBackground.py:306-331): gaussfitter.twodgaussian
//      parameters: height, amplitude, xo, yo, sigma_x, sigma_y, theta (raw optimum, before the np.abs of :331)
// Far from the centre the Gaussian term is below 1e-10 of its amplitude (exp(-q/2), q > 46: farther than 6.8 times the
// larger width), the residual is height - data and only the height column of the Jacobian is non-zero.  Those pixels
// enter through three sums that do not depend on the parameters — their number, sum (d - m0) and sum (d - m0)^2 about a
// fixed level m0 near the data (no cancellation) —, taken once per cutout; a pass then evaluates the model only inside
// the bounding box of that ellipse and corrects the sums by the box's own.  (curve_fit stops at 1.5e-8: the neglected
// terms are far below what decides its optimum.)
constexpr double kGaussReach = 6.8;
struct GaussModel {
    double m0, s1, s2;                     // level, sum (d - m0), sum (d - m0)^2 over the WHOLE cutout
    double h, A, xo, yo, sx, sy, c, s, isx, isy;
    int r0, r1, c0, c1;
    double box_s1, box_s2;                 // this thread's sums of (d - m0), (d - m0)^2 over the box pixels it visited
    template <typename T> PM_DEV void whole_frame_sums(Sh& sh, const T* img, int n, double level) {
        double a[2] = {0.0, 0.0};
        for (int i = pm::tid(); i < n * n; i += kThreads) {
            const double d = (double)pm::ldg(img + i) - level;
            a[0] += d; a[1] += d * d;
        }
        block_sum<2>(sh, a);
        m0 = level; s1 = a[0]; s2 = a[1];
    }
    PM_DEV void setup(const double* q, int n) {
        h = q[0]; A = q[1]; xo = q[2]; yo = q[3]; sx = q[4]; sy = q[5];
        c = cos(q[6]); s = sin(q[6]);
        isx = 1.0 / sx; isy = 1.0 / sy;
        box_s1 = box_s2 = 0.0;
        r0 = 0; r1 = n - 1; c0 = 0; c1 = n - 1;
        // bounding box of the ellipse beyond which the Gaussian is dropped: rows about xo, columns about yo
        const double reach = kGaussReach * fmax(fabs(sx), fabs(sy)) + 1.0;
        if (reach == reach && xo == xo && yo == yo && reach < 4.0 * (double)n) {
            const double a0 = floor(xo - reach), a1 = ceil(xo + reach), b0 = floor(yo - reach), b1 = ceil(yo + reach);
            r0 = a0 > 0.0 ? (a0 < (double)n ? (int)a0 : n) : 0;
            r1 = a1 < (double)(n - 1) ? (a1 >= 0.0 ? (int)a1 : -1) : n - 1;
            c0 = b0 > 0.0 ? (b0 < (double)n ? (int)b0 : n) : 0;
            c1 = b1 < (double)(n - 1) ? (b1 >= 0.0 ? (int)b1 : -1) : n - 1;
        }
    }
    PM_DEV void columns(int row, int col, double dat, double (&J)[8]) {
        // gaussfitter.twodgaussian: centre (column yo, row xo) — center_y, center_x = xo, yo (gaussfitter.py:132)
        const double dx = yo - (double)col, dy = xo - (double)row;
        const double u = (dx * c - dy * s) * isx, v = (dx * s + dy * c) * isy;
        const double E = exp(-0.5 * (u * u + v * v));
        const double AE = A * E;
        const double dm = dat - m0;
        box_s1 += dm; box_s2 += dm * dm;
        J[0] = 1.0;
        J[1] = E;
        J[2] = AE * (u * s * isx - v * c * isy);
        J[3] = -AE * (u * c * isx + v * s * isy);
        J[4] = AE * u * u * isx;
        J[5] = AE * v * v * isy;
        J[6] = AE * u * v * (sy * isx - sx * isy);
        J[7] = h + AE - dat;
    }
    // the pixels outside the box: r = (h - m0) - (d - m0), J = (1, 0, ..., 0)
    PM_DEV void finish(double* acc, int n, int bw, int bh, const double* bs) const {
        const double nf = (double)n * (double)n - (double)bw * (double)bh;
        const double f1 = s1 - bs[0], f2 = s2 - bs[1], hm = h - m0;
        if (nf > 0.0) {
            acc[35] += nf * hm * hm - 2.0 * hm * f1 + f2;
            acc[0] += nf;                    // J0 J0
            acc[28] += nf * hm - f1;         // J0 r
        }
    }
};

template <typename T>
PM_DEV void gaussfit_body(const T* images, long long B, int n, const double* median, const double* vmax, const double* p0_in, double* out) {
    unsigned char* smem = pm::dyn_smem(0);
    Sh& sh = *(Sh*)smem;
    GaussSmem& gs = *(GaussSmem*)(smem + align16((unsigned)sizeof(Sh)));
    sh.par[pm::tid()] = 0;
    pm::sync();
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        const T* img = images + (size_t)b * n * n;
        double p[7];
        int status = 0;
        if (p0_in) {
            for (int k = 0; k < 7; k++) p[k] = p0_in[7 * b + k];
        } else {
            // gaussfitter.moments (gaussfitter.py:61-84): widths from the central row / column, height = median, amplitude = max - median
            const int y0 = n / 2, x0 = n / 2;
            double a4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = pm::tid(); i < n; i += kThreads) {
                const double cv = (double)pm::ldg(img + (size_t)y0 * n + i), rv = (double)pm::ldg(img + (size_t)i * n + x0);
                const double dc = (double)(i - y0), dr = (double)(i - x0);
                a4[0] += fabs(dc * dc * cv); a4[1] += fabs(cv); a4[2] += fabs(dr * dr * rv); a4[3] += fabs(rv);
            }
            block_sum<4>(sh, a4);
            p[0] = median[b]; p[1] = vmax[b] - median[b]; p[2] = (double)x0; p[3] = (double)y0;
            p[4] = sqrt(a4[0] / a4[1]); p[5] = sqrt(a4[2] / a4[3]); p[6] = 90.0 * 3.141592653589793 / 180.0;
            if (p[4] != p[4] || p[5] != p[5] || p[0] != p[0] || p[1] != p[1]) status = 2;
        }
        double cost = 0.0;
        int iter = 0;
        GaussModel model;
        model.whole_frame_sums(sh, img, n, p[0] == p[0] && fabs(p[0]) < 1.7e308 ? p[0] : 0.0);
        if (status == 0) status = lm_minimise(sh, gs, img, n, model, p, 1e-10, &cost, &iter);
        if (pm::tid() == 0) {
            double* o = out + (size_t)b * kGaussCols;
            for (int k = 0; k < 7; k++) o[k] = p[k];
            o[7] = cost; o[8] = (double)iter; o[9] = (double)status;
        }
        pm::sync();
    }
}

// ---- the Sersic profile of sersic.fitSersic (sersic.py:125-134): astropy Sersic2D.evaluate
//      f = amplitude exp(-b_n (z^(1/n) - 1)),  z^2 = (x_maj / r_eff)^2 + (x_min / ((1 - ellip) r_eff))^2,  b_n = gammaincinv(2 n, 1/2)
//      parameters: amplitude, r_eff, n, x_0, y_0, ellip, theta; x = column, y = row (np.mgrid, sersic.py:131)
// regularised lower incomplete gamma P(a, x) for x < a + 1 (series) and its inverse at 1/2 by Newton steps from the
// asymptotic value a - 1/3 + 8 / (405 a) (Ciotti & Bertin 1999 for b_n with a = 2 n)
PM_DEV double gamma_p_series(double a, double x) {
    double term = 1.0 / a, sum = term;
    for (int k = 1; k < 4000; k++) {                 // (terms grow until k ~ x - a: enough for any index a fit visits)
        term *= x / (a + (double)k);
        sum += term;
        if (fabs(term) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - lgamma(a));
}
PM_DEV double gammaincinv_half(double a) {
    if (!(a > 0.0)) return pm::bits2d(0x7ff8000000000000ll);
    double x = a - 1.0 / 3.0 + 8.0 / (405.0 * a);
    if (!(x > 0.0)) x = 0.5 * a;
    for (int it = 0; it < 60; it++) {
        const double f = gamma_p_series(a, x) - 0.5;
        const double df = exp(-x + (a - 1.0) * log(x) - lgamma(a));
        double step = f / df;
        if (!(step == step)) break;
        if (x - step <= 0.0) step = 0.5 * x;             // stay positive
        x -= step;
        if (fabs(step) <= 1e-16 * x) break;
    }
    return x;
}
struct SersicModel {
    double A, re, nn, x0, y0, el, ct, st, bn, dbn, ia, ib, inv_n;
    int r0, r1, c0, c1;
    double box_s1, box_s2;
    PM_DEV void setup(const double* q, int n) {
        A = q[0]; re = q[1]; nn = q[2]; x0 = q[3]; y0 = q[4]; el = q[5];
        ct = cos(q[6]); st = sin(q[6]);
        bn = gammaincinv_half(2.0 * nn);
        const double hstep = 1e-6 * fmax(fabs(nn), 1e-3);
        dbn = (gammaincinv_half(2.0 * (nn + hstep)) - gammaincinv_half(2.0 * (nn - hstep))) / (2.0 * hstep);
        ia = 1.0 / re; ib = 1.0 / ((1.0 - el) * re); inv_n = 1.0 / nn;
        r0 = 0; r1 = n - 1; c0 = 0; c1 = n - 1;
        box_s1 = box_s2 = 0.0;
    }
    PM_DEV void columns(int row, int col, double dat, double (&J)[8]) {
        const double dx = (double)col - x0, dy = (double)row - y0;
        const double xmaj = dx * ct + dy * st, xmin = -dx * st + dy * ct;
        const double ua = xmaj * ia, ub = xmin * ib;
        const double z2 = ua * ua + ub * ub, z = sqrt(z2);
        const double lz = log(z);                                    // -inf at the centre: w = 0 there
        const double w = z > 0.0 ? exp(lz * inv_n) : 0.0;
        const double e = exp(-bn * (w - 1.0));
        const double f = A * e;
        // df/dz = -b_n f w / (n z); the cusp at z = 0 (one pixel at most) is given slope 0
        const double fz = z > 0.0 ? -bn * f * w * inv_n / z : 0.0;
        const double iz = z > 0.0 ? 1.0 / z : 0.0;
        J[0] = e;
        J[1] = fz * (-z * ia);                                                           // z ~ 1 / r_eff
        J[2] = z > 0.0 ? -f * (dbn * (w - 1.0) - bn * w * lz * inv_n * inv_n) : -f * dbn * (w - 1.0);
        J[3] = fz * iz * (-ua * ia * ct + ub * ib * st);
        J[4] = fz * iz * (-ua * ia * st - ub * ib * ct);
        J[5] = fz * iz * (ub * ub / (1.0 - el));                                         // d b / d ellip = -r_eff
        J[6] = fz * iz * (xmaj * xmin * (ia * ia - ib * ib));
        J[7] = f - dat;
    }
    PM_DEV void finish(double*, int, int, int, const double*) const {}
};

// p0: DEVICE [B, 7] start values (sersic.py:98-123, the caller's); out: as the Gaussian fit
template <typename T>
PM_DEV void sersicfit_body(const T* images, long long B, int n, const double* p0_in, double* out) {
    unsigned char* smem = pm::dyn_smem(0);
    Sh& sh = *(Sh*)smem;
    GaussSmem& gs = *(GaussSmem*)(smem + align16((unsigned)sizeof(Sh)));
    sh.par[pm::tid()] = 0;
    pm::sync();
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        const T* img = images + (size_t)b * n * n;
        double p[7];
        int status = 0;
        for (int k = 0; k < 7; k++) { p[k] = p0_in[7 * b + k]; if (p[k] != p[k]) status = 2; }
        double cost = 0.0;
        int iter = 0;
        SersicModel model;
        if (status == 0) status = lm_minimise(sh, gs, img, n, model, p, 1e-10, &cost, &iter);
        if (pm::tid() == 0) {
            double* o = out + (size_t)b * kGaussCols;
            for (int k = 0; k < 7; k++) o[k] = p[k];
            o[7] = cost; o[8] = (double)iter; o[9] = (double)status;
        }
        pm::sync();
    }
}

// ---- exact elliptical aperture sums (photutils EllipticalAperture.do_photometry(method="exact"), sersic.py:46-47, :101-121)
// Area of (pixel ∩ ellipse): the pixel's corners are mapped to the frame in which the ellipse is the unit circle; the
// area of (parallelogram ∩ unit disc) is the sum over the edges of a triangle about the origin for the part of the edge
// inside the disc and of circular sectors for the parts outside (signed; counter-clockwise corners).
PM_DEV double unit_disc_edge(double px, double py, double qx, double qy) {
    const double dx = qx - px, dy = qy - py;
    const double a = dx * dx + dy * dy, b = 2.0 * (px * dx + py * dy), c = px * px + py * py - 1.0;
    const double disc = b * b - 4.0 * a * c;
    double t1 = 0.0, t2 = 0.0;
    bool hit = false;
    if (disc > 0.0 && a > 0.0) {
        const double sq = sqrt(disc);
        t1 = fmin(fmax((-b - sq) / (2.0 * a), 0.0), 1.0);
        t2 = fmin(fmax((-b + sq) / (2.0 * a), 0.0), 1.0);
        hit = t2 > t1;
    }
    if (!hit) {
        const bool q_in = qx * qx + qy * qy - 1.0 <= 0.0;
        if (c <= 0.0 && q_in) return 0.5 * (px * qy - py * qx);
        return 0.5 * atan2(px * qy - py * qx, px * qx + py * qy);
    }
    const double ax = px + t1 * dx, ay = py + t1 * dy, bx = px + t2 * dx, by = py + t2 * dy;
    double v = 0.5 * (ax * by - ay * bx);
    if (t1 > 0.0) v += 0.5 * atan2(px * ay - py * ax, px * ax + py * ay);
    if (t2 < 1.0) v += 0.5 * atan2(bx * qy - by * qx, bx * qx + by * qy);
    return v;
}
// ell: (xc, yc, a, b, theta) per entry; idx: cutout of each entry (nullptr: entry i is cutout i); out: the sums
template <typename T>
PM_DEV void ellipse_sum_body(const T* images, int n, long long M, const long long* idx, const double* ell, double* out) {
    unsigned char* smem = pm::dyn_smem(0);
    Sh& sh = *(Sh*)smem;
    sh.par[pm::tid()] = 0;
    pm::sync();
    for (long long m = pm::block(); m < M; m += pm::nblocks()) {
        const T* img = images + (size_t)(idx ? idx[m] : m) * n * n;
        const double xc = ell[5 * m], yc = ell[5 * m + 1], a = ell[5 * m + 2], b = ell[5 * m + 3], th = ell[5 * m + 4];
        const double ct = cos(th), st = sin(th);
        const double ex = sqrt(a * ct * a * ct + b * st * b * st), ey = sqrt(a * st * a * st + b * ct * b * ct);
        double sum = 0.0;
        if (a > 0.0 && b > 0.0 && ex == ex && ey == ey && xc == xc && yc == yc && ex < 1e9 && ey < 1e9) {
            const double fc0 = floor(xc - ex - 0.5), fc1 = ceil(xc + ex + 0.5), fr0 = floor(yc - ey - 0.5), fr1 = ceil(yc + ey + 0.5);
            const int c0 = fc0 > 0.0 ? (fc0 < (double)n ? (int)fc0 : n) : 0, c1 = fc1 < (double)(n - 1) ? (fc1 >= 0.0 ? (int)fc1 : -1) : n - 1;
            const int r0 = fr0 > 0.0 ? (fr0 < (double)n ? (int)fr0 : n) : 0, r1 = fr1 < (double)(n - 1) ? (fr1 >= 0.0 ? (int)fr1 : -1) : n - 1;
            const int bw = c1 >= c0 ? c1 - c0 + 1 : 0, bh = r1 >= r0 ? r1 - r0 + 1 : 0;
            const double ia = 1.0 / a, ib = 1.0 / b;
            for (int i = pm::tid(); i < bw * bh; i += kThreads) {
                const int rr = i / bw, row = r0 + rr, col = c0 + (i - rr * bw);
                double u[4], v[4];
                bool all_in = true;
#pragma unroll
                for (int k = 0; k < 4; k++) {                       // counter-clockwise corners
                    const double x = (double)col + ((k == 1 || k == 2) ? 0.5 : -0.5) - xc, y = (double)row + (k >= 2 ? 0.5 : -0.5) - yc;
                    u[k] = (x * ct + y * st) * ia; v[k] = (-x * st + y * ct) * ib;
                    all_in = all_in && (u[k] * u[k] + v[k] * v[k] <= 1.0);
                }
                double wgt = 1.0;
                if (!all_in) {
                    double area = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; k++) area += unit_disc_edge(u[k], v[k], u[(k + 1) & 3], v[(k + 1) & 3]);
                    wgt = fmin(fmax(area * a * b, 0.0), 1.0);
                }
                if (wgt != 0.0) sum += wgt * (double)pm::ldg(img + (size_t)row * n + col);
            }
        } else sum = pm::bits2d(0x7ff8000000000000ll);
        sum = block_sum1(sh, sum);
        if (pm::tid() == 0) out[m] = sum;
        pm::sync();
    }
}

// ----------------------------------------------------------------------------------------------------
// source detection + cleaning
// ----------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> four 32-bit words
PM_HOSTDEV void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// standard normal of (seed, cutout, pixel): Box-Muller on two 32-bit uniforms in (0, 1]
PM_DEV double philox_normal(unsigned long long seed, unsigned long long cutout, unsigned pixel) {
    unsigned w[4];
    philox4x32(pixel, (unsigned)cutout, (unsigned)(cutout >> 32), 0x504d4253u, (unsigned)seed, (unsigned)(seed >> 32), w);
    const double u1 = ((double)w[0] + 1.0) * (1.0 / 4294967296.0), u2 = ((double)w[1] + 1.0) * (1.0 / 4294967296.0);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// find with path halving on the parent array of the run forest
PM_DEV int uf_find(int* parent, int i) {
    for (;;) {
        const int p = parent[i];
        if (p == i) return i;
        const int gp = parent[p];
        if (gp != p) parent[i] = gp;
        i = gp;
    }
}
PM_DEV void uf_union(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }       // link the larger root under the smaller: roots only decrease
        const int old = pm::atomic_cas(&parent[a], a, b);
        if (old == a) return;
    }
}

constexpr int kStarCols = 4;   // number of segments kept, number of segments cleaned, pixels replaced, status (1: run table overflow)

// det: threshold[B]; fill: mean[B], std[B] of the replacement noise (imageutils.py:62, :86), seed.
// labels_out [B, n, n] int32 (nullable): 0 background, 1.. segments in raster order of their first pixel (scipy.ndimage.label
// order kept by photutils after the small ones are dropped); clean_out [B, n, n] of T (nullable): the cleaned image.
// scratch (per CTA): runs as (row << 20 | c0 << 10 | c1) — n <= 640 fits 10 bits —, parents, sizes, boxes.
template <typename T>
PM_DEV void maskstars_body(const T* images, long long B, int n, const double* threshold, int npixels, const double* mean, const double* std_,
                           unsigned long long seed, int* labels_out, T* clean_out, double* out, unsigned char* scratch_all,
                           unsigned long long scratch_per_cta, int max_runs) {
    unsigned char* smem = pm::dyn_smem(0);
    Sh& sh = *(Sh*)smem;
    const int nw = (n + 31) / 32;
    unsigned* plane = (unsigned*)(smem + align16((unsigned)sizeof(Sh)));
    unsigned* rowoff = plane + n * nw;                       // [n + 1] runs before each row
    sh.par[pm::tid()] = 0;
    pm::sync();
    unsigned char* scratch = scratch_all + (size_t)pm::block() * scratch_per_cta;
    unsigned* runs = (unsigned*)scratch;                     // [max_runs]
    int* parent = (int*)(runs + max_runs);                   // [max_runs]
    int* size = parent + max_runs;                           // [max_runs] pixel count per root
    int* newlab = size + max_runs;                           // [max_runs] final label per root (0: dropped)
    int* box = newlab + max_runs;                            // [max_runs][4] rmin, rmax, cmin, cmax per root
    // astropy Gaussian2DKernel(sigma = 3 * gaussian_fwhm_to_sigma, x_size = 3, y_size = 3), normalised (imageutils.py:68-70)
    const double sig = 3.0 * 0.42466090014400953, e1 = exp(-1.0 / (2.0 * sig * sig)), e2 = exp(-2.0 / (2.0 * sig * sig));
    const double ksum = 1.0 + 4.0 * e1 + 4.0 * e2, k0 = 1.0 / ksum, k1 = e1 / ksum, k2 = e2 / ksum;
    for (long long b = pm::block(); b < B; b += pm::nblocks()) {
        const T* img = images + (size_t)b * n * n;
        const double thr = threshold[b];
        // ---- smoothed image > threshold -> bit plane (convolution with zero fill beyond the frame)
        for (int task = pm::warp(); task < n * nw; task += kWarps) {
            const int r = task / nw, w = task - r * nw, col = w * 32 + pm::lane();
            bool above = false;
            if (col < n) {
                auto at = [&](int rr, int cc) -> double {
                    return (rr >= 0 && rr < n && cc >= 0 && cc < n) ? (double)pm::ldg(img + (size_t)rr * n + cc) : 0.0;
                };
                // scipy.ndimage.convolve accumulation order is not reproduced: the decision `> threshold` is robust to it except
                // at exact ties, which real data do not produce (threshold = mean + 1.5 std of the clipped pixels)
                const double corner = at(r - 1, col - 1) + at(r - 1, col + 1) + at(r + 1, col - 1) + at(r + 1, col + 1);
                const double edge = at(r - 1, col) + at(r + 1, col) + at(r, col - 1) + at(r, col + 1);
                above = k0 * at(r, col) + k1 * edge + k2 * corner > thr;
            }
            const unsigned bits = pm::ballot(above);
            if (pm::lane() == 0) plane[task] = bits;
        }
        pm::sync();
        // ---- runs per row
        for (int r = pm::tid(); r < n; r += kThreads) {
            unsigned cnt = 0, carry = 0;
            for (int w = 0; w < nw; w++) {
                const unsigned v = plane[r * nw + w];
                cnt += (unsigned)pm::popc(v & ~((v << 1) | carry));      // run starts: set bits whose left neighbour is clear
                carry = v >> 31;
            }
            rowoff[r] = cnt;
        }
        pm::sync();
        const int nruns = (int)block_scan_excl_array(sh, rowoff, n);
        int status = 0;
        if (nruns > max_runs) status = 1;
        if (status == 0) {
            for (int r = pm::tid(); r < n; r += kThreads) {
                unsigned o = rowoff[r];
                int start = -1;
                for (int col = 0; col <= n; col++) {
                    const bool on = col < n && ((plane[r * nw + (col >> 5)] >> (col & 31)) & 1u);
                    if (on && start < 0) start = col;
                    if (!on && start >= 0) { runs[o] = ((unsigned)r << 20) | ((unsigned)start << 10) | (unsigned)(col - 1); parent[o] = (int)o; o++; start = -1; }
                }
            }
            pm::sync();
            // ---- union runs of adjacent rows that touch (8-connectivity: column ranges overlap after widening by one)
            for (int i = pm::tid(); i < nruns; i += kThreads) {
                const unsigned ru = runs[i];
                const int r = (int)(ru >> 20), c0 = (int)((ru >> 10) & 1023u), c1 = (int)(ru & 1023u);
                if (r == 0) continue;
                for (int j = (int)rowoff[r - 1]; j < (int)rowoff[r]; j++) {
                    const unsigned rj = runs[j];
                    const int d0 = (int)((rj >> 10) & 1023u), d1 = (int)(rj & 1023u);
                    if (d0 > c1 + 1) break;
                    if (d1 >= c0 - 1) uf_union(parent, i, j);
                }
            }
            pm::sync();
            // ---- roots, sizes, boxes
            for (int i = pm::tid(); i < nruns; i += kThreads) { size[i] = 0; box[4 * i] = 0x7fffffff; box[4 * i + 1] = -1; box[4 * i + 2] = 0x7fffffff; box[4 * i + 3] = -1; }
            pm::sync();
            for (int i = pm::tid(); i < nruns; i += kThreads) {
                const int root = uf_find(parent, i);
                parent[i] = root;
            }
            pm::sync();
            for (int i = pm::tid(); i < nruns; i += kThreads) {
                const unsigned ru = runs[i];
                const int r = (int)(ru >> 20), c0 = (int)((ru >> 10) & 1023u), c1 = (int)(ru & 1023u), root = parent[i];
                pm::atomic_add(&size[root], c1 - c0 + 1);
                pm::atomic_min(&box[4 * root], r); pm::atomic_max(&box[4 * root + 1], r);
                pm::atomic_min(&box[4 * root + 2], c0); pm::atomic_max(&box[4 * root + 3], c1);
            }
            pm::sync();
            // ---- labels of the kept roots in raster order of their first pixel (= order of the root run index)
            for (int i = pm::tid(); i < nruns; i += kThreads) newlab[i] = (parent[i] == i && size[i] >= npixels) ? 1 : 0;
            pm::sync();
        }
        int nkept = 0;
        if (status == 0 && nruns > 0) nkept = (int)block_scan_excl_array(sh, (unsigned*)newlab, nruns);
        // newlab[i] = number of kept roots before run i (the table has room for the total at [nruns]): a kept root i gets
        // label newlab[i] + 1
        if (labels_out) {
            int* lab = labels_out + (size_t)b * n * n;
            for (int i = pm::tid(); i < n * n; i += kThreads) lab[i] = 0;
            pm::sync();
            if (status == 0) {
                for (int i = pm::warp(); i < nruns; i += kWarps) {
                    const unsigned ru = runs[i];
                    const int r = (int)(ru >> 20), c0 = (int)((ru >> 10) & 1023u), c1 = (int)(ru & 1023u), root = parent[i];
                    if (size[root] < npixels) continue;
                    const int l = newlab[root] + 1;
                    for (int col = c0 + pm::lane(); col <= c1; col += 32) lab[r * n + col] = l;
                }
            }
        }
        // ---- cleaning (imageutils.py:73-89): a kept segment whose bounding box does not contain cenpix is a "star"; inside
        // the bounding box of every star every pixel that belongs to ANY kept segment is replaced by N(mean, std) noise
        // (np.where(data_ma > 0, 1, 0) reads the label cutout, masked or not), the others keep the image value
        int ncleaned = 0, nrepl = 0;
        if (clean_out) {
            T* dst = clean_out + (size_t)b * n * n;
            for (int i = pm::tid(); i < n * n; i += kThreads) dst[i] = pm::ldg(img + i);
            pm::sync();
            if (status == 0 && mean && std_) {
                const double cen = (double)(n / 2 + 1);
                const double mu = mean[b], sd = std_[b];
                // star flag per root: _inBbox(extent = (cmin - 0.5, cmax + 0.5, rmin - 0.5, rmax + 0.5), cenpix), strict
                for (int i = pm::tid(); i < nruns; i += kThreads) {
                    if (parent[i] == i && size[i] >= npixels) {
                        const bool inbox = cen > (double)box[4 * i + 2] - 0.5 && cen > (double)box[4 * i] - 0.5 &&
                                           cen < (double)box[4 * i + 3] + 0.5 && cen < (double)box[4 * i + 1] + 0.5;
                        size[i] = inbox ? size[i] : -size[i];          // negative size marks a star
                        if (!inbox) ncleaned++;
                    }
                }
                pm::sync();
                // the stars' boxes as a compact list (newlab is free again once the labels are written)
                if (pm::tid() == 0) sh.ia[1] = 0;
                pm::sync();
                for (int i = pm::tid(); i < nruns; i += kThreads)
                    if (parent[i] == i && size[i] < 0 && -size[i] >= npixels) newlab[pm::atomic_add(&sh.ia[1], 1)] = i;
                pm::sync();
                const int nstars = sh.ia[1];
                // a pixel of a kept segment is replaced when it lies inside the bounding box of at least one star
                for (int i = pm::warp(); i < nruns; i += kWarps) {
                    const unsigned ru = runs[i];
                    const int r = (int)(ru >> 20), c0 = (int)((ru >> 10) & 1023u), c1 = (int)(ru & 1023u), root = parent[i];
                    if ((size[root] < 0 ? -size[root] : size[root]) < npixels) continue;
                    for (int col = c0 + pm::lane(); col <= c1; col += 32) {
                        bool hit = size[root] < 0;                      // its own segment is a star
                        for (int k = 0; k < nstars && !hit; k++) {
                            const int j = newlab[k];
                            hit = r >= box[4 * j] && r <= box[4 * j + 1] && col >= box[4 * j + 2] && col <= box[4 * j + 3];
                        }
                        if (hit) {
                            double v = mu + sd * philox_normal(seed, (unsigned long long)b, (unsigned)(r * n + col));
                            if (v == 0.0) v = (double)pm::ldg(img + (size_t)r * n + col);    // imageutils.py:89
                            dst[r * n + col] = (T)v;
                            nrepl++;
                        }
                    }
                }
            }
        }
        ncleaned = (int)block_sum_u(sh, (unsigned)ncleaned);
        nrepl = (int)block_sum_u(sh, (unsigned)nrepl);
        if (pm::tid() == 0) {
            double* o = out + (size_t)b * kStarCols;
            o[0] = (double)nkept; o[1] = (double)ncleaned; o[2] = (double)nrepl; o[3] = (double)status;
        }
        pm::sync();
    }
}
PM_HOSTDEV unsigned maskstars_smem_bytes(int n) {
    return align16((unsigned)sizeof(Sh)) + align16((unsigned)(n * ((n + 31) / 32) * 4)) + align16((unsigned)(n + 2) * 4);
}
PM_HOSTDEV int maskstars_max_runs(int n) { return n * ((n + 1) / 2) + 2; }                 // a run needs a gap: at most ceil(n / 2) per row
PM_HOSTDEV unsigned long long maskstars_scratch_bytes(int n) { return ((unsigned long long)maskstars_max_runs(n) * 8 * 4 + 255) & ~255ull; }

}  // namespace pmk
