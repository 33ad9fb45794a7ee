// MvNMF k x k steps as single-CTA kernels (k <= 32, V <= 96), always in float64 internally.
//
//   volume_logdet            models/mvnmf.py:19-24   ln det(W^T W + delta I), LU with partial pivoting
//   update_W_unconstrained   models/mvnmf.py:37-66   closed-form root given N = (X/(WH)) H^T and rowsums(H)
//   line-search candidate    models/mvnmf.py:80-81,85-88 + utils.py:155-158 (blend, normalise, clip)
//
// W is stored [k][V] (asignatures.X); the reference's W (V,k) is its transpose.
#include "sal_common.cuh"

namespace {

constexpr int NT = 128;
constexpr int KM = SAL_KMAX;
constexpr int GP = KM + 1;        // pitch of k x k matrices
constexpr int WP = SAL_VMAX + 1;  // pitch of the [k][V] copy

struct Smem {
    double Wd[KM * WP];      // W (or the trial W) as double, [k][V]
    double G[KM * GP];       // Gram + delta I, destroyed by the factorisation
    double Y[KM * GP];       // inverse (w_unconstrained only)
    double colsum[KM];
    double col[KM];
    double scal[4];
    int piv;
};

// G = Wd Wd^T + delta I   (k x k)
__device__ void gram(Smem& s, int V, int k, double delta) {
    for (int i = threadIdx.x; i < k * k; i += NT) {
        const int a = i / k, b = i - a * k;
        double t = 0.0;
        for (int v = 0; v < V; ++v) t += s.Wd[a * WP + v] * s.Wd[b * WP + v];
        s.G[a * GP + b] = t + (a == b ? delta : 0.0);
    }
    __syncthreads();
}

// In-place LU with partial pivoting of s.G; returns det (all threads).
__device__ double lu_det(Smem& s, int k) {
    double sign = 1.0;
    for (int c = 0; c < k; ++c) {
        if (threadIdx.x == 0) {
            int p = c;
            double best = fabs(s.G[c * GP + c]);
            for (int r = c + 1; r < k; ++r) {
                const double a = fabs(s.G[r * GP + c]);
                if (a > best) best = a, p = r;
            }
            s.piv = p;
        }
        __syncthreads();
        const int p = s.piv;
        if (p != c) {
            sign = -sign;
            for (int j = threadIdx.x; j < k; j += NT) {
                const double t = s.G[c * GP + j];
                s.G[c * GP + j] = s.G[p * GP + j];
                s.G[p * GP + j] = t;
            }
            __syncthreads();
        }
        const double d = s.G[c * GP + c];
        const int nr = k - c - 1, nc = k - c - 1;
        // factors first (column c below the diagonal), then the trailing update
        for (int r = threadIdx.x; r < nr; r += NT) s.G[(c + 1 + r) * GP + c] /= d;
        __syncthreads();
        for (int i = threadIdx.x; i < nr * nc; i += NT) {
            const int r = c + 1 + i / nc, j = c + 1 + i % nc;
            s.G[r * GP + j] -= s.G[r * GP + c] * s.G[c * GP + j];
        }
        __syncthreads();
    }
    double det = sign;
    for (int c = 0; c < k; ++c) det *= s.G[c * GP + c];
    return det;
}

// Y = G^-1 by Gauss-Jordan with partial pivoting (G destroyed).
__device__ void invert(Smem& s, int k) {
    for (int i = threadIdx.x; i < k * k; i += NT) s.Y[(i / k) * GP + i % k] = (i / k == i % k) ? 1.0 : 0.0;
    __syncthreads();
    for (int c = 0; c < k; ++c) {
        if (threadIdx.x == 0) {
            int p = c;
            double best = fabs(s.G[c * GP + c]);
            for (int r = c + 1; r < k; ++r) {
                const double a = fabs(s.G[r * GP + c]);
                if (a > best) best = a, p = r;
            }
            s.piv = p;
        }
        __syncthreads();
        const int p = s.piv;
        if (p != c) {
            for (int j = threadIdx.x; j < 2 * k; j += NT) {
                double* M = j < k ? s.G : s.Y;
                const int jj = j < k ? j : j - k;
                const double t = M[c * GP + jj];
                M[c * GP + jj] = M[p * GP + jj];
                M[p * GP + jj] = t;
            }
            __syncthreads();
        }
        const double inv_d = 1.0 / s.G[c * GP + c];
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * k; j += NT) {
            double* M = j < k ? s.G : s.Y;
            M[c * GP + (j < k ? j : j - k)] *= inv_d;
        }
        __syncthreads();
        // eliminate column c from every other row (row c itself is not touched in this step)
        for (int r = threadIdx.x; r < k; r += NT) s.col[r] = s.G[r * GP + c];
        __syncthreads();
        for (int i = threadIdx.x; i < k * 2 * k; i += NT) {
            const int r = i / (2 * k), j = i - r * 2 * k;
            if (r == c) continue;
            double* M = j < k ? s.G : s.Y;
            const int jj = j < k ? j : j - k;
            M[r * GP + jj] -= s.col[r] * M[c * GP + jj];
        }
        __syncthreads();
    }
}

template <typename T>
__device__ void load_w(Smem& s, const T* W, int V, int k) {
    for (int i = threadIdx.x; i < k * V; i += NT) s.Wd[(i / V) * WP + i % V] = (double)W[i];
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(NT) mvnmf_logdet_kernel(const T* W, int V, int k, double delta, double* out) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    load_w(s, W, V, k);
    gram(s, V, k, delta);
    const double det = lu_det(s, k);
    if (threadIdx.x == 0) *out = log(det);
}

template <typename T>
__global__ void __launch_bounds__(NT) mvnmf_w_unc_kernel(const T* W, const T* N, const T* hsum, int V, int k,
                                                        double lam, double delta, int n_given, T* W_unc) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    load_w(s, W, V, k);
    gram(s, V, k, delta);
    invert(s, k);
    for (int i = threadIdx.x; i < k * V; i += NT) {
        const int j = i / V, v = i - j * V;
        const double w = s.Wd[j * WP + v];
        double out;
        if (j < n_given) {
            out = w;
        } else {
            double wym = 0.0, wya = 0.0;
            for (int a = 0; a < k; ++a) {
                const double y = s.Y[a * GP + j];
                const double wa = s.Wd[a * WP + v];
                wym += wa * fmax(0.0, -y);
                wya += wa * fabs(y);
            }
            const double r = (double)hsum[j];
            const double a1 = r - 4.0 * lam * wym;
            const double s2 = 8.0 * lam * wya * (double)N[i];
            const double num = sqrt(a1 * a1 + s2) + (-r + 4.0 * lam * wym);
            out = fmax(w * num / (4.0 * lam * wya), (double)SAL_EPS_F32);
        }
        W_unc[i] = (T)out;
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) mvnmf_trial_kernel(const T* W, const T* W_unc, int V, int k, double gamma,
                                                        double delta, T* W_trial, T* h_scale, double* logdet_out) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    for (int i = threadIdx.x; i < k * V; i += NT) {
        const double wu = (double)W_unc[i];
        s.Wd[(i / V) * WP + i % V] = gamma < 0.0 ? wu : (1.0 - gamma) * (double)W[i] + gamma * wu;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += NT) {
        double t = 0.0;
        for (int v = 0; v < V; ++v) t += s.Wd[j * WP + v];
        s.colsum[j] = t;
        h_scale[j] = (T)t;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k * V; i += NT) {
        const int j = i / V, v = i - j * V;
        const T wt = (T)fmax(s.Wd[j * WP + v] / s.colsum[j], (double)SAL_EPS_F32);
        W_trial[i] = wt;
        s.Wd[j * WP + v] = (double)wt;  // logdet of what the objective pass will actually read
    }
    __syncthreads();
    gram(s, V, k, delta);
    const double det = lu_det(s, k);
    if (threadIdx.x == 0) *logdet_out = log(det);
}

template <typename K>
int set_smem(K kernel) {
    SAL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    return 0;
}

}  // namespace

int sal_launch_mvnmf_logdet(sal_ctx* c, const void* W, double delta, double* out, cudaStream_t st) {
    if (c->dtype == SAL_F32) {
        if (int e = set_smem(mvnmf_logdet_kernel<float>)) return e;
        mvnmf_logdet_kernel<float><<<1, NT, sizeof(Smem), st>>>((const float*)W, c->V, c->k, delta, out);
    } else {
        if (int e = set_smem(mvnmf_logdet_kernel<double>)) return e;
        mvnmf_logdet_kernel<double><<<1, NT, sizeof(Smem), st>>>((const double*)W, c->V, c->k, delta, out);
    }
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_mvnmf_w_unc(sal_ctx* c, const void* W, const void* N, const void* hsum, double lam,
                           double delta, int n_given, void* W_unc, cudaStream_t st) {
    if (c->dtype == SAL_F32) {
        if (int e = set_smem(mvnmf_w_unc_kernel<float>)) return e;
        mvnmf_w_unc_kernel<float><<<1, NT, sizeof(Smem), st>>>((const float*)W, (const float*)N, (const float*)hsum,
                                                              c->V, c->k, lam, delta, n_given, (float*)W_unc);
    } else {
        if (int e = set_smem(mvnmf_w_unc_kernel<double>)) return e;
        mvnmf_w_unc_kernel<double><<<1, NT, sizeof(Smem), st>>>((const double*)W, (const double*)N,
                                                               (const double*)hsum, c->V, c->k, lam, delta,
                                                               n_given, (double*)W_unc);
    }
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_mvnmf_trial(sal_ctx* c, const void* W, const void* W_unc, double gamma, double delta,
                           void* W_trial, void* h_scale, double* logdet_out, cudaStream_t st) {
    if (c->dtype == SAL_F32) {
        if (int e = set_smem(mvnmf_trial_kernel<float>)) return e;
        mvnmf_trial_kernel<float><<<1, NT, sizeof(Smem), st>>>((const float*)W, (const float*)W_unc, c->V, c->k,
                                                              gamma, delta, (float*)W_trial, (float*)h_scale,
                                                              logdet_out);
    } else {
        if (int e = set_smem(mvnmf_trial_kernel<double>)) return e;
        mvnmf_trial_kernel<double><<<1, NT, sizeof(Smem), st>>>((const double*)W, (const double*)W_unc, c->V, c->k,
                                                               gamma, delta, (double*)W_trial, (double*)h_scale,
                                                               logdet_out);
    }
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}
