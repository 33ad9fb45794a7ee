// MvNMF k x k steps as single-CTA kernels (k <= 32, V <= 96), always in float64 internally.
//
//   volume_logdet            models/mvnmf.py:19-24   ln det(W^T W + delta I), LU with partial pivoting
//   update_W_unconstrained   models/mvnmf.py:37-66   closed-form root given N = (X/(WH)) H^T and rowsums(H)
//   line-search candidate    models/mvnmf.py:80-81,85-88 + utils.py:155-158 (blend, normalise, clip)
//
// W is stored [k][V] (asignatures.X); the reference's W (V,k) is its transpose.
#include "sal_common.cuh"
#include "mvnmf_kk.cuh"

namespace {

constexpr int NT = 256;
constexpr int KM = SAL_KMAX;
constexpr int GP = KM + 1;        // pitch of k x k matrices
constexpr int WP = SAL_VMAX + 1;  // pitch of the [k][V] copy

struct Smem {
    double Wd[KM * WP];      // W (or the trial W) as double, [k][V]
    double G[KM * GP];       // Gram + delta I, destroyed by the factorisation
    double Y[KM * GP];       // inverse (w_unconstrained only)
    double Wu[KM * WP];      // W_unconstrained as the first line-search trial reads it (fused w_unconstrained + trial)
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

// LU determinant / Gauss-Jordan inverse of s.G (mvnmf_kk.cuh: pivot search on warp 0, row operations over the whole block)
__device__ double lu_det(Smem& s, int k) { return lu_det_block<NT>(s.G, &s.piv, GP, k); }
__device__ void invert(Smem& s, int k) { invert_block<NT>(s.G, s.Y, s.col, &s.piv, GP, k); }

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

// W_trial != null: the first line-search candidate (the full step: normalise + clip of W_unconstrained itself, gamma ignored,
// mvnmf.py:80-88) follows in the same launch -- what mvnmf_trial_kernel(gamma < 0) would do right behind this kernel
template <typename T>
__global__ void __launch_bounds__(NT) mvnmf_w_unc_kernel(const T* W, const T* N, const T* hsum, int V, int k,
                                                        double lam, double delta, int n_given, T* W_unc, T* W_trial, T* h_scale,
                                                        double* logdet_out) {
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
        if (W_trial) s.Wu[j * WP + v] = (double)(T)out;
    }
    if (!W_trial) return;
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += NT) {
        double t = 0.0;
        for (int v = 0; v < V; ++v) t += s.Wu[j * WP + v];
        s.colsum[j] = t;
        h_scale[j] = (T)t;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k * V; i += NT) {
        const int j = i / V, v = i - j * V;
        const T wt = (T)fmax(s.Wu[j * WP + v] / s.colsum[j], (double)SAL_EPS_F32);
        W_trial[i] = wt;
        s.Wd[j * WP + v] = (double)wt;  // logdet of what the objective pass will actually read (W itself is not needed any more)
    }
    __syncthreads();
    gram(s, V, k, delta);
    const double det = lu_det(s, k);
    if (threadIdx.x == 0) *logdet_out = log(det);
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
                           double delta, int n_given, void* W_unc, void* W_trial, void* h_scale, double* logdet_out, cudaStream_t st) {
    if (c->dtype == SAL_F32) {
        if (int e = set_smem(mvnmf_w_unc_kernel<float>)) return e;
        mvnmf_w_unc_kernel<float><<<1, NT, sizeof(Smem), st>>>((const float*)W, (const float*)N, (const float*)hsum,
                                                              c->V, c->k, lam, delta, n_given, (float*)W_unc, (float*)W_trial,
                                                              (float*)h_scale, logdet_out);
    } else {
        if (int e = set_smem(mvnmf_w_unc_kernel<double>)) return e;
        mvnmf_w_unc_kernel<double><<<1, NT, sizeof(Smem), st>>>((const double*)W, (const double*)N,
                                                               (const double*)hsum, c->V, c->k, lam, delta,
                                                               n_given, (double*)W_unc, (double*)W_trial, (double*)h_scale, logdet_out);
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
