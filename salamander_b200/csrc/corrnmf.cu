// Correlated NMF kernels (reference models/_utils_corrnmf.py, models/corrnmf_det.py).
//
//   exposures            H_dk = exp(a_k + b_d + l_k.u_d)                          _utils_corrnmf.py:11-25
//   sample scalings      b_d  = ln sum_v x_dv - ln sum_k exp(a_k + l_k.u_d)       :141-179
//   signature scalings   a_k  = ln sum_d aux_kd - ln sum_d exp(b_d + l_k.u_d)     :103-138
//   embedding updates    argmin_e -[sum_i (o_i.e) aux_i - sum_i exp(s + s_i + o_i.e) - |e|^2/(2 var)]   :182-410
//                        sample embeddings: one THREAD per sample (others = the k signatures, 3 Newton iterations,
//                        corrnmf_det.py:115-141); signature embeddings: one CTA per signature, every objective /
//                        gradient / Hessian evaluation is a fixed-order block reduction over the samples (:88-113)
//   variance             clip(mean([L;U]^2))                                      corrnmf_det.py:60-69
//   aux and the W numerator come from the fused pass (klnmf_pass.cu, SAL_PASS_NOCLIP): aux_kd = H_dk (W^T A)_kd.
//
// The minimiser is a device restatement of SciPy's Newton-CG -- truncated Newton, CG inner loop, More'-Thuente DCSRCH
// line search with SciPy's defaults -- the third-party algorithm the reference calls at _utils_corrnmf.py:400-407;
// oracle/corrnmf.py::newton_cg is the same restatement in numpy and is pinned against scipy itself.
// All arithmetic is float64 whatever the handle's storage dtype.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "sal_common.cuh"

#include <string.h>

namespace cg = cooperative_groups;

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
template <class P>
__device__ void newton_cg(P& prob, double* x, int m, int maxiter) {
    const double ftol = 1e-4, gtol = 0.9, ls_xtol = 1e-14, stpmin = 1e-8, stpmax = 50.0, eps64 = 2.220446049250313e-16;
    const double xtol = m * 1e-5;
    const int cg_maxiter = 20 * m;
    double b[MAXM], xs[MAXM], ri[MAXM], ps[MAXM], Ap[MAXM], xt[MAXM], gt[MAXM], gl[MAXM];
    double Abuf0[MAXM * MAXM], Abuf1[MAXM * MAXM];
    double *A = Abuf0, *Al = Abuf1;  // Hessian at the current point / at the trial point (swapped on acceptance)
    double old_fval = prob.f_grad_hess(x, gt, A), old_old_fval = 0.0;
    bool have_old_old = false;
    double update_l1 = 1.7976931348623157e308;
    int k = 0;
    while (update_l1 > xtol) {
        if (k >= maxiter) break;
        double maggrad = 0.0;
        for (int i = 0; i < m; ++i) b[i] = -gt[i], maggrad += fabs(b[i]);
        const double termcond = fmin(0.5, sqrt(maggrad)) * maggrad;
        double dri0 = 0.0;
        for (int i = 0; i < m; ++i) xs[i] = 0.0, ri[i] = -b[i], ps[i] = b[i], dri0 += ri[i] * ri[i];
        int it = 0;
        bool failed = true;
        for (int k2 = 0; k2 < cg_maxiter; ++k2) {
            double rn = 0.0;
            for (int i = 0; i < m; ++i) rn += fabs(ri[i]);
            if (rn <= termcond) {
                failed = false;
                break;
            }
            double curv = 0.0;
            for (int i = 0; i < m; ++i) {
                double t = 0.0;
                for (int j = 0; j < m; ++j) t += A[i * m + j] * ps[j];
                Ap[i] = t;
            }
            for (int i = 0; i < m; ++i) curv += ps[i] * Ap[i];
            if (curv >= 0.0 && curv <= 3.0 * eps64) {
                failed = false;
                break;
            } else if (curv < 0.0) {
                if (it == 0)
                    for (int i = 0; i < m; ++i) xs[i] = dri0 / (-curv) * b[i];
                failed = false;
                break;
            }
            const double alphai = dri0 / curv;
            double dri1 = 0.0;
            for (int i = 0; i < m; ++i) xs[i] += alphai * ps[i], ri[i] += alphai * Ap[i], dri1 += ri[i] * ri[i];
            const double betai = dri1 / dri0;
            for (int i = 0; i < m; ++i) ps[i] = -ri[i] + betai * ps[i];
            ++it;
            dri0 = dri1;
        }
        if (failed) break;  // "CG iterations didn't converge"
        // ---- line search along pk = xs (DCSRCH)
        double derphi0 = 0.0;
        for (int i = 0; i < m; ++i) derphi0 += gt[i] * xs[i];
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
                for (int i = 0; i < m; ++i) xt[i] = x[i] + s.stp * xs[i];
                const double f = prob.f_grad_hess(xt, gl, Al);
                double g = 0.0;
                for (int i = 0; i < m; ++i) g += gl[i] * xs[i];
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
        for (int i = 0; i < m; ++i) {
            const double xn = x[i] + stp_ok * xs[i];  // the very expression the trial point was formed with
            update_l1 += fabs(stp_ok * xs[i]);
            x[i] = xn;
            gt[i] = gl[i];
        }
        {
            double* t = A;
            A = Al, Al = t;
        }
        ++k;
    }
}

__device__ inline double snap_eps(double e) {
    const double eps = SAL_EPS_F32;
    if (e > 0.0 && e < eps) return eps;
    if (e < 0.0 && e > -eps) return -eps;
    return e;
}

// ---------------------------------------------------------------------------------------------------------
// sample embeddings: one thread per sample, the k signature embeddings / scalings in shared memory
// ---------------------------------------------------------------------------------------------------------
struct SampleProblem {
    const double* others;    // [k][m] shared
    const double* s_others;  // [k]    shared
    const double* aux;       // [k]    local
    const double* s_vec;     // [k]    local: the sample's scaling for every "other" (one value repeated for a single
                             //        modality; per-modality values in multimodal CorrNMF, mmcorrnmf.py:413-419)
    double inv_var;
    int k, m;
    __device__ double f_grad_hess(const double* x, double* g, double* A) const {
        double acc = 0.0, nrm = 0.0;
        for (int j = 0; j < m * m; ++j) A[j] = 0.0;
        for (int j = 0; j < m; ++j) g[j] = x[j] * inv_var, nrm += x[j] * x[j], A[j * m + j] = inv_var;
        for (int i = 0; i < k; ++i) {
            double sp = 0.0;
            for (int j = 0; j < m; ++j) sp += others[i * m + j] * x[j];
            const double e = exp(s_vec[i] + s_others[i] + sp);
            acc += sp * aux[i] - e;
            const double w = e - aux[i];
            for (int j = 0; j < m; ++j) g[j] += w * others[i * m + j];
            for (int p = 0; p < m; ++p)
                for (int q = 0; q < m; ++q) A[p * m + q] += e * others[i * m + p] * others[i * m + q];
        }
        return -(acc - 0.5 * nrm * inv_var);
    }
};

// b: [D] (b_is_matrix = 0, one scaling per sample) or [D][k] (b_is_matrix = 1, one per sample and signature)
template <typename T>
__global__ void __launch_bounds__(128) sample_embeddings_kernel(const T* auxT, const T* a, const T* b, int b_is_matrix, const T* L, T* U,
                                                               int64_t D, int k, int m, double variance, int maxiter) {
    __shared__ double sL[SAL_KMAX * MAXM];
    __shared__ double sa[SAL_KMAX];
    for (int i = threadIdx.x; i < k * m; i += blockDim.x) sL[i] = (double)L[i];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sa[i] = (double)a[i];
    __syncthreads();
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double aux[SAL_KMAX], sv[SAL_KMAX], x[MAXM];
    for (int i = 0; i < k; ++i) aux[i] = (double)auxT[d * k + i], sv[i] = (double)(b_is_matrix ? b[d * k + i] : b[d]);
    for (int j = 0; j < m; ++j) x[j] = (double)U[d * m + j];
    SampleProblem p{sL, sa, aux, sv, 1.0 / variance, k, m};
    newton_cg(p, x, m, maxiter);
    for (int j = 0; j < m; ++j) U[d * m + j] = (T)snap_eps(x[j]);
}

// ---------------------------------------------------------------------------------------------------------
// signature embeddings: one CTA per signature; evaluations are block reductions over the samples
// ---------------------------------------------------------------------------------------------------------
constexpr int SIG_THREADS = 256;

template <typename T>
struct SignatureProblem {
    const T *U, *b, *auxT;  // U [D][m], b [D], auxT [D][k]
    double* red;            // shared [SIG_THREADS / 32][1 + MAXM + MAXM * MAXM]
    double* cl;             // shared [1 + MAXM + MAXM * MAXM]: this CTA's totals, read by the other CTAs of the cluster
    double s, inv_var;
    int64_t D;
    int k, m, j;
    int rank, nrank;        // position in the thread-block cluster that shares signature j (samples are interleaved)

    // fixed-order block reduction of n values per thread; result broadcast to all threads through shared memory
    __device__ void reduce(double* v, int n) const {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, stride = 1 + MAXM + MAXM * MAXM;
        for (int i = 0; i < n; ++i) {
            double t = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) red[w * stride + i] = t;
        }
        __syncthreads();
        for (int i = 0; i < n; ++i) {
            double t = 0.0;
            for (int ww = 0; ww < SIG_THREADS / 32; ++ww) t += red[ww * stride + i];
            v[i] = t;
        }
        __syncthreads();
        if (nrank > 1) {  // sum the CTAs' totals in rank order through distributed shared memory
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0)
                for (int i = 0; i < n; ++i) cl[i] = v[i];
            cluster.sync();
            for (int i = 0; i < n; ++i) {
                double t = 0.0;
                for (int r = 0; r < nrank; ++r) t += cluster.map_shared_rank(cl, r)[i];
                v[i] = t;
            }
            cluster.sync();
        }
    }
    __device__ double f_grad_hess(const double* x, double* g, double* A) const {
        double v[1 + MAXM + MAXM * MAXM];  // [f | gradient | Hessian]: one collective reduction for all of it
        const int n = 1 + m + m * m;
        for (int q = 0; q < n; ++q) v[q] = 0.0;
        for (int64_t d = (int64_t)rank * SIG_THREADS + threadIdx.x; d < D; d += (int64_t)nrank * SIG_THREADS) {
            double u[MAXM], sp = 0.0;
            for (int q = 0; q < m; ++q) u[q] = (double)U[d * m + q], sp += u[q] * x[q];
            const double e = exp(s + (double)b[d] + sp), ax = (double)auxT[d * k + j];
            v[0] += sp * ax - e;
            const double w = e - ax;
            for (int q = 0; q < m; ++q) v[1 + q] += w * u[q];
            for (int p = 0; p < m; ++p)
                for (int q = 0; q < m; ++q) v[1 + m + p * m + q] += e * u[p] * u[q];
        }
        reduce(v, n);
        double nrm = 0.0;
        for (int q = 0; q < m; ++q) nrm += x[q] * x[q], g[q] = v[1 + q] + x[q] * inv_var;
        for (int q = 0; q < m * m; ++q) A[q] = v[1 + m + q];
        for (int q = 0; q < m; ++q) A[q * m + q] += inv_var;
        return -(v[0] - 0.5 * nrm * inv_var);
    }
};

template <typename T>
__global__ void __launch_bounds__(SIG_THREADS) signature_embeddings_kernel(const T* auxT, const T* a, const T* b, T* L, const T* U,
                                                                          int64_t D, int k, int m, double variance, int sig_begin) {
    __shared__ double red[(SIG_THREADS / 32) * (1 + MAXM + MAXM * MAXM)];
    __shared__ double cl[1 + MAXM + MAXM * MAXM];
    cg::cluster_group cluster = cg::this_cluster();
    const int nrank = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int j = sig_begin + blockIdx.x / nrank;  // one cluster per signature; every CTA of it runs the same Newton-CG on the same numbers
    double x[MAXM];
    for (int q = 0; q < m; ++q) x[q] = (double)L[j * m + q];
    SignatureProblem<T> p{U, b, auxT, red, cl, (double)a[j], 1.0 / variance, D, k, m, j, rank, nrank};
    if (nrank > 1) cluster.sync();  // nobody writes L[j] before everybody has read it
    newton_cg(p, x, m, 200 * m);
    if (rank == 0 && threadIdx.x == 0)
        for (int q = 0; q < m; ++q) L[j * m + q] = (T)snap_eps(x[q]);
}

// ---------------------------------------------------------------------------------------------------------
// element-wise / reduction kernels
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) exposures_kernel(const T* a, const T* b, const T* L, const T* U, int64_t D, int k, int m, T* H) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * k) return;
    const int64_t d = i / k;
    const int j = (int)(i - d * k);
    double sp = 0.0;
    for (int q = 0; q < m; ++q) sp += (double)L[j * m + q] * (double)U[d * m + q];
    H[i] = (T)exp((double)a[j] + (double)b[d] + sp);
}

template <typename T>
__global__ void __launch_bounds__(256) row_sums_kernel(const T* X, int64_t D, int V, T* out) {
    const int64_t d = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (d >= D) return;
    double s = 0.0;
    for (int v = threadIdx.x & 31; v < V; v += 32) s += (double)X[d * V + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) out[d] = (T)s;
}

template <typename T>
__global__ void __launch_bounds__(256) sample_scalings_kernel(const T* xsum, const T* a, const T* L, const T* U, int64_t D, int k, int m,
                                                             T* b) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double s = 0.0;
    for (int j = 0; j < k; ++j) {
        double sp = 0.0;
        for (int q = 0; q < m; ++q) sp += (double)L[j * m + q] * (double)U[d * m + q];
        s += exp((double)a[j] + sp);
    }
    b[d] = (T)(log((double)xsum[d]) - log(s));
}

__device__ double block_sum_256(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}

// sums[j] = sum_d aux_dj, sums[k + j] = sum_d exp(b_d + l_j.u_d)  (this rank's samples).  SCAL_BLOCKS CTAs per signature
// (blockIdx.y = signature) each leave a partial pair; the last one to finish adds them in block order (deterministic).
constexpr int SCAL_BLOCKS = 32;
template <typename T>
__global__ void __launch_bounds__(256) signature_scalings_kernel(const T* auxT, const T* b, const T* L, const T* U, int64_t D, int k, int m,
                                                                double* partial, unsigned int* counter, double* sums) {
    __shared__ double s_red[8];
    __shared__ bool last;
    const int j = blockIdx.y;
    double l[MAXM];
    for (int q = 0; q < m; ++q) l[q] = (double)L[j * m + q];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t d = (int64_t)blockIdx.x * 256 + threadIdx.x; d < D; d += (int64_t)SCAL_BLOCKS * 256) {
        double sp = 0.0;
        for (int q = 0; q < m; ++q) sp += l[q] * (double)U[d * m + q];
        s1 += (double)auxT[d * k + j];
        s2 += exp((double)b[d] + sp);
    }
    s1 = block_sum_256(s1, s_red);
    s2 = block_sum_256(s2, s_red);
    if (threadIdx.x == 0) {
        partial[(2 * j) * SCAL_BLOCKS + blockIdx.x] = s1;
        partial[(2 * j + 1) * SCAL_BLOCKS + blockIdx.x] = s2;
        __threadfence();
        last = atomicAdd(&counter[j], 1u) == SCAL_BLOCKS - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < SCAL_BLOCKS; ++i) {
            t1 += ((volatile double*)partial)[(2 * j) * SCAL_BLOCKS + i];
            t2 += ((volatile double*)partial)[(2 * j + 1) * SCAL_BLOCKS + i];
        }
        sums[j] = t1, sums[k + j] = t2;
        counter[j] = 0;
    }
}

template <typename T>
__global__ void signature_scalings_finish_kernel(const double* sums, int k, T* a) {
    const int j = threadIdx.x;
    if (j < k) a[j] = (T)(log(sums[j]) - log(sums[k + j]));
}

// out[0] = sum L^2, out[1] = sum U^2 (this rank's samples), out[2] = sum_x lnGamma(1 + x) when X != null.
// blockIdx.y selects the quantity, NORM_BLOCKS blocks each leave a partial; the last block to finish adds them in order.
constexpr int NORM_BLOCKS = 64;
template <typename T>
__global__ void __launch_bounds__(256) norms_kernel(const T* L, int64_t nL, const T* U, int64_t nU, const T* X, int64_t nX,
                                                   double* partial, unsigned int* counter, double* out) {
    __shared__ double s_red[8];
    __shared__ bool last;
    const int q = blockIdx.y;
    const T* src = q == 0 ? L : q == 1 ? U : X;
    const int64_t n = q == 0 ? nL : q == 1 ? nU : nX;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)NORM_BLOCKS * 256) {
        const double v = (double)src[i];
        s += q < 2 ? v * v : lgamma(1.0 + v);
    }
    s = block_sum_256(s, s_red);
    if (threadIdx.x == 0) {
        partial[q * NORM_BLOCKS + blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(&counter[q], 1u) == NORM_BLOCKS - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (int b = 0; b < NORM_BLOCKS; ++b) t += ((volatile double*)partial)[q * NORM_BLOCKS + b];
        out[q] = t;
        counter[q] = 0;
    }
}

}  // namespace

#define SAL_DISPATCH_T(c, call_f, call_d) \
    do {                                  \
        if ((c)->dtype == SAL_F32) {      \
            call_f;                       \
        } else {                          \
            call_d;                       \
        }                                 \
    } while (0)

int sal_launch_corrnmf_exposures(sal_ctx* c, const void* a, const void* b, const void* L, const void* U, int m, void* H, cudaStream_t st) {
    const int64_t n = c->D * c->k;
    if (n == 0) return 0;
    const int grid = (int)((n + 255) / 256);
    SAL_DISPATCH_T(c, (exposures_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)L, (const float*)U, c->D, c->k, m, (float*)H)),
                   (exposures_kernel<double><<<grid, 256, 0, st>>>((const double*)a, (const double*)b, (const double*)L, (const double*)U, c->D, c->k, m, (double*)H)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_row_sums(sal_ctx* c, const void* X, void* out, cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 7) / 8);
    SAL_DISPATCH_T(c, (row_sums_kernel<float><<<grid, 256, 0, st>>>((const float*)X, c->D, c->V, (float*)out)),
                   (row_sums_kernel<double><<<grid, 256, 0, st>>>((const double*)X, c->D, c->V, (double*)out)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_sample_scalings(sal_ctx* c, const void* xsum, const void* a, const void* L, const void* U, int m, void* b,
                                       cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 255) / 256);
    SAL_DISPATCH_T(c, (sample_scalings_kernel<float><<<grid, 256, 0, st>>>((const float*)xsum, (const float*)a, (const float*)L, (const float*)U, c->D, c->k, m, (float*)b)),
                   (sample_scalings_kernel<double><<<grid, 256, 0, st>>>((const double*)xsum, (const double*)a, (const double*)L, (const double*)U, c->D, c->k, m, (double*)b)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

// counters of the "last block adds the partials" reductions: [0..3] norms, [4..4 + SAL_KMAX) signature scalings
static int ensure_counters(sal_ctx* c, cudaStream_t st) {
    if (!c->norm_counter) {
        SAL_CUDA(cudaMalloc((void**)&c->norm_counter, (4 + SAL_KMAX) * sizeof(unsigned int)));
        SAL_CUDA(cudaMemsetAsync(c->norm_counter, 0, (4 + SAL_KMAX) * sizeof(unsigned int), st));
    }
    return 0;
}

int sal_launch_corrnmf_signature_scalings_sums(sal_ctx* c, const void* auxT, const void* b, const void* L, const void* U, int m,
                                               double* sums, cudaStream_t st) {
    if (int e = ensure_counters(c, st)) return e;
    // scratch: 2 * k * SCAL_BLOCKS doubles at the start of the W-numerator partials (no pass is in flight on this stream)
    static_assert(2 * SAL_KMAX * SCAL_BLOCKS * sizeof(double) <= (size_t)148 * SAL_KMAX * SAL_VMAX * 4, "scratch too small");
    double* partial = (double*)c->partial_wnum;
    unsigned int* counter = c->norm_counter + 4;
    const dim3 grid(SCAL_BLOCKS, c->k);
    SAL_DISPATCH_T(c, (signature_scalings_kernel<float><<<grid, 256, 0, st>>>((const float*)auxT, (const float*)b, (const float*)L, (const float*)U, c->D, c->k, m, partial, counter, sums)),
                   (signature_scalings_kernel<double><<<grid, 256, 0, st>>>((const double*)auxT, (const double*)b, (const double*)L, (const double*)U, c->D, c->k, m, partial, counter, sums)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_signature_scalings_finish(sal_ctx* c, const double* sums, void* a, cudaStream_t st) {
    SAL_DISPATCH_T(c, (signature_scalings_finish_kernel<float><<<1, 32, 0, st>>>(sums, c->k, (float*)a)),
                   (signature_scalings_finish_kernel<double><<<1, 32, 0, st>>>(sums, c->k, (double*)a)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_sample_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, int b_is_matrix, const void* L,
                                         void* U, int m, double variance, int maxiter, cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 127) / 128);
    SAL_DISPATCH_T(c, (sample_embeddings_kernel<float><<<grid, 128, 0, st>>>((const float*)auxT, (const float*)a, (const float*)b, b_is_matrix, (const float*)L, (float*)U, c->D, c->k, m, variance, maxiter)),
                   (sample_embeddings_kernel<double><<<grid, 128, 0, st>>>((const double*)auxT, (const double*)a, (const double*)b, b_is_matrix, (const double*)L, (double*)U, c->D, c->k, m, variance, maxiter)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_signature_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                            double variance, int sig_begin, int sig_count, cudaStream_t st) {
    if (sig_count <= 0) return 0;
    // one thread-block cluster per signature: 8 CTAs share the sums over samples once there is enough work for them, 16
    // (non-portable size: one cluster per GPC, 8 GPCs) when there are few signatures and a lot of samples
    int csize = c->D >= 8 * 4 * SIG_THREADS ? 8 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(SIG_THREADS), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    if (sig_count <= 8 && c->D >= 16 * 8 * SIG_THREADS) {
        // (the occupancy query can take tens of milliseconds: asked once per device, dtype and signature count)
        static signed char cached[64][2][9];
        static bool cached_init = false;
        if (!cached_init) memset(cached, -1, sizeof(cached)), cached_init = true;
        signed char& ok16 = cached[c->device & 63][c->dtype == SAL_F32 ? 0 : 1][sig_count];
        if (ok16 < 0) {
            const void* fn = c->dtype == SAL_F32 ? (const void*)signature_embeddings_kernel<float> : (const void*)signature_embeddings_kernel<double>;
            int n_active = 0;
            attr.val.clusterDim.x = 16;
            cfg.gridDim = dim3(sig_count * 16);
            ok16 = (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                    cudaOccupancyMaxActiveClusters(&n_active, fn, &cfg) == cudaSuccess && n_active >= sig_count)
                       ? 1
                       : 0;
            (void)cudaGetLastError();  // a refused query is not an error of this call: fall back to the portable size
        }
        if (ok16 == 1) csize = 16;
    }
    if (const char* e = getenv("SAL_B200_SIG_CLUSTER")) {  // diagnostics: force the cluster size (1, 8 or 16)
        const int forced = atoi(e);
        if (forced == 1 || forced == 8 || (forced == 16 && csize == 16)) csize = forced;
    }
    attr.val.clusterDim.x = csize;
    cfg.gridDim = dim3(sig_count * csize);
    const int64_t D = c->D;
    const int k = c->k;
    if (c->dtype == SAL_F32)
        SAL_CUDA(cudaLaunchKernelEx(&cfg, signature_embeddings_kernel<float>, (const float*)auxT, (const float*)a, (const float*)b, (float*)L,
                                    (const float*)U, D, k, m, variance, sig_begin));
    else
        SAL_CUDA(cudaLaunchKernelEx(&cfg, signature_embeddings_kernel<double>, (const double*)auxT, (const double*)a, (const double*)b,
                                    (double*)L, (const double*)U, D, k, m, variance, sig_begin));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_norms(sal_ctx* c, const void* L, const void* U, int m, const void* X_or_null, double* out, cudaStream_t st) {
    // scratch: the objective partials of the pass (>= 296 doubles) hold the 3 x 64 partials; the counters follow out[3]
    static_assert(3 * NORM_BLOCKS <= 148 * SAL_KMAX, "partial_hsum has n_sm * SAL_KMAX slots at least");
    if (int e = ensure_counters(c, st)) return e;
    const dim3 grid(NORM_BLOCKS, X_or_null ? 3 : 2);
    SAL_DISPATCH_T(c, (norms_kernel<float><<<grid, 256, 0, st>>>((const float*)L, (int64_t)c->k * m, (const float*)U, c->D * m, (const float*)X_or_null, c->D * c->V, c->partial_hsum, c->norm_counter, out)),
                   (norms_kernel<double><<<grid, 256, 0, st>>>((const double*)L, (int64_t)c->k * m, (const double*)U, c->D * m, (const double*)X_or_null, c->D * c->V, c->partial_hsum, c->norm_counter, out)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_corrnmf_max_dim(void) { return MAXM; }
