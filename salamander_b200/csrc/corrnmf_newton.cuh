// Device restatement of SciPy's Newton-CG (truncated Newton, CG inner loop, More'-Thuente DCSRCH line search with SciPy's
// defaults) -- the third-party algorithm the reference calls at models/_utils_corrnmf.py:400-407 -- shared by the sample- and
// the signature-embedding kernels (corrnmf.cu, corrnmf_sig.cu).  oracle/corrnmf.py::newton_cg is the same restatement in numpy,
// pinned against scipy itself.  All arithmetic is float64.
#pragma once

#include "sal_common.cuh"

namespace {

constexpr int MAXM = 16;  // embedding dimensions handled (SAL_EUNSUPPORTED above)

// ---------------------------------------------------------------------------------------------------------
// DCSRCH / dcstep (MINPACK-2, as shipped in scipy/optimize/_dcsrch.py): scalar state machine
// ---------------------------------------------------------------------------------------------------------
struct StepState {
    double stx, fx, dx, sty, fy, dy, stp;
    bool brackt;
};

__device__ inline double sgn(double v) { return (v > 0.0) - (v < 0.0); }

__device__ void dcstep(StepState& s, double fp, double dp, double stpmin, double stpmax) {
    const double sgnd = sgn(dp) * sgn(s.dx);
    double stpf, stpc, stpq;
    if (fp > s.fx) {
        const double theta = 3.0 * (s.fx - fp) / (s.stp - s.stx) + s.dx + dp;
        const double sc = fmax(fabs(theta), fmax(fabs(s.dx), fabs(dp)));
        double gamma = sc * sqrt((theta / sc) * (theta / sc) - (s.dx / sc) * (dp / sc));
        if (s.stp < s.stx) gamma = -gamma;
        const double p = (gamma - s.dx) + theta, q = ((gamma - s.dx) + gamma) + dp, r = p / q;
        stpc = s.stx + r * (s.stp - s.stx);
        stpq = s.stx + ((s.dx / ((s.fx - fp) / (s.stp - s.stx) + s.dx)) / 2.0) * (s.stp - s.stx);
        stpf = fabs(stpc - s.stx) <= fabs(stpq - s.stx) ? stpc : stpc + (stpq - stpc) / 2.0;
        s.brackt = true;
    } else if (sgnd < 0.0) {
        const double theta = 3.0 * (s.fx - fp) / (s.stp - s.stx) + s.dx + dp;
        const double sc = fmax(fabs(theta), fmax(fabs(s.dx), fabs(dp)));
        double gamma = sc * sqrt((theta / sc) * (theta / sc) - (s.dx / sc) * (dp / sc));
        if (s.stp > s.stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + s.dx, r = p / q;
        stpc = s.stp + r * (s.stx - s.stp);
        stpq = s.stp + (dp / (dp - s.dx)) * (s.stx - s.stp);
        stpf = fabs(stpc - s.stp) > fabs(stpq - s.stp) ? stpc : stpq;
        s.brackt = true;
    } else if (fabs(dp) < fabs(s.dx)) {
        const double theta = 3.0 * (s.fx - fp) / (s.stp - s.stx) + s.dx + dp;
        const double sc = fmax(fabs(theta), fmax(fabs(s.dx), fabs(dp)));
        double gamma = sc * sqrt(fmax(0.0, (theta / sc) * (theta / sc) - (s.dx / sc) * (dp / sc)));
        if (s.stp > s.stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = (gamma + (s.dx - dp)) + gamma, r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = s.stp + r * (s.stx - s.stp);
        else
            stpc = s.stp > s.stx ? stpmax : stpmin;
        stpq = s.stp + (dp / (dp - s.dx)) * (s.stx - s.stp);
        if (s.brackt) {
            stpf = fabs(stpc - s.stp) < fabs(stpq - s.stp) ? stpc : stpq;
            stpf = s.stp > s.stx ? fmin(s.stp + 0.66 * (s.sty - s.stp), stpf) : fmax(s.stp + 0.66 * (s.sty - s.stp), stpf);
        } else {
            stpf = fabs(stpc - s.stp) > fabs(stpq - s.stp) ? stpc : stpq;
            stpf = fmin(fmax(stpf, stpmin), stpmax);
        }
    } else {
        if (s.brackt) {
            const double theta = 3.0 * (fp - s.fy) / (s.sty - s.stp) + s.dy + dp;
            const double sc = fmax(fabs(theta), fmax(fabs(s.dy), fabs(dp)));
            double gamma = sc * sqrt((theta / sc) * (theta / sc) - (s.dy / sc) * (dp / sc));
            if (s.stp > s.sty) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + s.dy, r = p / q;
            stpf = s.stp + r * (s.sty - s.stp);
        } else {
            stpf = s.stp > s.stx ? stpmax : stpmin;
        }
    }
    if (fp > s.fx) {
        s.sty = s.stp, s.fy = fp, s.dy = dp;
    } else {
        if (sgnd < 0.0) s.sty = s.stx, s.fy = s.fx, s.dy = s.dx;
        s.stx = s.stp, s.fx = fp, s.dx = dp;
    }
    s.stp = stpf;
}

// ---------------------------------------------------------------------------------------------------------
// Newton-CG.  Problem P provides (collectively for the calling threads; every thread gets the same values):
//   double f_grad_hess(const double* x, double* g, double* A)   -- all three at the same point in ONE sweep over the
//                                                                  terms (one exp per term, one collective reduction)
// SciPy evaluates fprime(xk) and fhess(xk) again at the top of every Newton iteration; xk is the line-search point that
// was just accepted, where DCSRCH has already asked for f and f'.  Every trial therefore evaluates f, f' and the Hessian
// together and the accepted trial's values are kept: one sweep per trial, none at the top of the iteration (same
// functions at the same points, so the iterates are SciPy's).
// ---------------------------------------------------------------------------------------------------------
// M > 0: the embedding dimension is a compile-time constant -- every vector and both Hessians live in registers and the loops
// unroll (with a run-time m they are local-memory arrays with dynamic indexing: the sample-embedding kernel, one Newton-CG per
// THREAD, spent most of its time there); M = 0: run-time dimension up to MAXM.
template <int M, class P>
__device__ void newton_cg(P& prob, double* x, int m_rt, int maxiter) {
    constexpr int MM = M > 0 ? M : MAXM;
    const int m = M > 0 ? M : m_rt;
    const double ftol = 1e-4, gtol = 0.9, ls_xtol = 1e-14, stpmin = 1e-8, stpmax = 50.0, eps64 = 2.220446049250313e-16;
    const double xtol = m * 1e-5;
    const int cg_maxiter = 20 * m;
    double b[MM], xs[MM], ri[MM], ps[MM], Ap[MM], xt[MM], gt[MM], gl[MM];
    double A[MM * MM], Al[MM * MM];  // Hessian at the current point / at the trial point (copied on acceptance)
    double old_fval = prob.f_grad_hess(x, gt, A), old_old_fval = 0.0;
    bool have_old_old = false;
    double update_l1 = 1.7976931348623157e308;
    int k = 0;
    while (update_l1 > xtol) {
        if (k >= maxiter) break;
        double maggrad = 0.0;
        _Pragma("unroll") for (int i = 0; i < m; ++i) b[i] = -gt[i], maggrad += fabs(b[i]);
        const double termcond = fmin(0.5, sqrt(maggrad)) * maggrad;
        double dri0 = 0.0;
        _Pragma("unroll") for (int i = 0; i < m; ++i) xs[i] = 0.0, ri[i] = -b[i], ps[i] = b[i], dri0 += ri[i] * ri[i];
        int it = 0;
        bool failed = true;
        for (int k2 = 0; k2 < cg_maxiter; ++k2) {
            double rn = 0.0;
            _Pragma("unroll") for (int i = 0; i < m; ++i) rn += fabs(ri[i]);
            if (rn <= termcond) {
                failed = false;
                break;
            }
            double curv = 0.0;
            _Pragma("unroll") for (int i = 0; i < m; ++i) {
                double t = 0.0;
                _Pragma("unroll") for (int j = 0; j < m; ++j) t += A[i * m + j] * ps[j];
                Ap[i] = t;
            }
            _Pragma("unroll") for (int i = 0; i < m; ++i) curv += ps[i] * Ap[i];
            if (curv >= 0.0 && curv <= 3.0 * eps64) {
                failed = false;
                break;
            } else if (curv < 0.0) {
                if (it == 0)
                    _Pragma("unroll") for (int i = 0; i < m; ++i) xs[i] = dri0 / (-curv) * b[i];
                failed = false;
                break;
            }
            const double alphai = dri0 / curv;
            double dri1 = 0.0;
            _Pragma("unroll") for (int i = 0; i < m; ++i) xs[i] += alphai * ps[i], ri[i] += alphai * Ap[i], dri1 += ri[i] * ri[i];
            const double betai = dri1 / dri0;
            _Pragma("unroll") for (int i = 0; i < m; ++i) ps[i] = -ri[i] + betai * ps[i];
            ++it;
            dri0 = dri1;
        }
        if (failed) break;  // "CG iterations didn't converge"
        // ---- line search along pk = xs (DCSRCH)
        double derphi0 = 0.0;
        _Pragma("unroll") for (int i = 0; i < m; ++i) derphi0 += gt[i] * xs[i];
        double alpha1 = 1.0;
        if (have_old_old && derphi0 != 0.0) {
            alpha1 = fmin(1.0, 1.01 * 2.0 * (old_fval - old_old_fval) / derphi0);
            if (alpha1 < 0.0) alpha1 = 1.0;
        }
        bool ok = false;
        double fnew = old_fval, stp_ok = 0.0;
        if (!(alpha1 < stpmin || alpha1 > stpmax || derphi0 >= 0.0)) {
            StepState s;
            s.brackt = false;
            int stage = 1;
            const double finit = old_fval, ginit = derphi0, gtest = ftol * ginit;
            double width = stpmax - stpmin, width1 = width / 0.5;
            s.stx = s.sty = 0.0, s.fx = s.fy = finit, s.dx = s.dy = ginit, s.stp = alpha1;
            double stmin = 0.0, stmax = alpha1 + 4.0 * alpha1;
            for (int ls = 0; ls < 99; ++ls) {
                _Pragma("unroll") for (int i = 0; i < m; ++i) xt[i] = x[i] + s.stp * xs[i];
                const double f = prob.f_grad_hess(xt, gl, Al);
                double g = 0.0;
                _Pragma("unroll") for (int i = 0; i < m; ++i) g += gl[i] * xs[i];
                const double ftest = finit + s.stp * gtest;
                if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
                bool warn = false;
                if (s.brackt && (s.stp <= stmin || s.stp >= stmax)) warn = true;
                if (s.brackt && stmax - stmin <= ls_xtol * stmax) warn = true;
                if (s.stp == stpmax && f <= ftest && g <= gtest) warn = true;
                if (s.stp == stpmin && (f > ftest || g >= gtest)) warn = true;
                if (f <= ftest && fabs(g) <= gtol * -ginit) {
                    ok = true, fnew = f, stp_ok = s.stp;
                    break;
                }
                if (warn) break;
                if (stage == 1 && f <= s.fx && f > ftest) {
                    StepState t = s;
                    t.fx = s.fx - s.stx * gtest, t.fy = s.fy - s.sty * gtest, t.dx = s.dx - gtest, t.dy = s.dy - gtest;
                    dcstep(t, f - s.stp * gtest, g - gtest, stmin, stmax);
                    s = t;
                    s.fx = t.fx + t.stx * gtest, s.fy = t.fy + t.sty * gtest, s.dx = t.dx + gtest, s.dy = t.dy + gtest;
                } else {
                    dcstep(s, f, g, stmin, stmax);
                }
                if (s.brackt) {
                    if (fabs(s.sty - s.stx) >= 0.66 * width1) s.stp = s.stx + 0.5 * (s.sty - s.stx);
                    width1 = width;
                    width = fabs(s.sty - s.stx);
                    stmin = fmin(s.stx, s.sty), stmax = fmax(s.stx, s.sty);
                } else {
                    stmin = s.stp + 1.1 * (s.stp - s.stx), stmax = s.stp + 4.0 * (s.stp - s.stx);
                }
                s.stp = fmin(fmax(s.stp, stpmin), stpmax);
                if ((s.brackt && (s.stp <= stmin || s.stp >= stmax)) || (s.brackt && stmax - stmin <= ls_xtol * stmax)) s.stp = s.stx;
                if (!isfinite(s.stp)) break;
            }
        }
        if (!ok) break;  // line search failed: keep the current point
        old_old_fval = old_fval, have_old_old = true, old_fval = fnew;
        update_l1 = 0.0;
        _Pragma("unroll") for (int i = 0; i < m; ++i) {
            const double xn = x[i] + stp_ok * xs[i];  // the very expression the trial point was formed with
            update_l1 += fabs(stp_ok * xs[i]);
            x[i] = xn;
            gt[i] = gl[i];
        }
#pragma unroll
        for (int i = 0; i < MM * MM; ++i)
            if (i < m * m) A[i] = Al[i];
        ++k;
    }
}

__device__ inline double snap_eps(double e) {
    const double eps = SAL_EPS_F32;
    if (e > 0.0 && e < eps) return eps;
    if (e < 0.0 && e > -eps) return -eps;
    return e;
}

}  // namespace
