// k x k factorisations on ONE warp (k <= 32: a lane per row, __syncwarp between the steps), shared by the single-CTA MvNMF
// kernels (mvnmf.cu), the persistent small-problem kernel and its cluster version (mvnmf_small.cu).  Same pivoting rule
// everywhere (first row of maximal magnitude: numpy.linalg.det / inv go through LAPACK getrf).
#pragma once

namespace {

// (with many threads every block-wide barrier of a column step costs more than the step itself)

// pivot row of column c among rows c .. k - 1 (lowest index among equals); result on every lane
__device__ __forceinline__ int pivot_row(const double* G, int GP, int k, int c) {
    const int lane = threadIdx.x & 31;
    double a = (lane >= c && lane < k) ? fabs(G[lane * GP + c]) : -1.0;
    int p = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oa = __shfl_xor_sync(0xffffffffu, a, o);
        const int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (oa > a || (oa == a && op < p)) a = oa, p = op;
    }
    return p;
}

// In-place LU with partial pivoting; returns det on every lane of warp 0 (call from warp 0 only)
__device__ double lu_det_warp(double* G, int GP, int k) {
    const int lane = threadIdx.x & 31;
    double sign = 1.0;
    for (int c = 0; c < k; ++c) {
        const int p = pivot_row(G, GP, k, c);
        if (p != c) {
            sign = -sign;
            if (lane < k) {
                const double t = G[c * GP + lane];
                G[c * GP + lane] = G[p * GP + lane];
                G[p * GP + lane] = t;
            }
            __syncwarp();
        }
        const double d = G[c * GP + c];
        const int nr = k - c - 1;
        if (lane < nr) G[(c + 1 + lane) * GP + c] /= d;
        __syncwarp();
        for (int i = lane; i < nr * nr; i += 32) {
            const int r = c + 1 + i / nr, j = c + 1 + i % nr;
            G[r * GP + j] -= G[r * GP + c] * G[c * GP + j];
        }
        __syncwarp();
    }
    double det = sign;
    for (int c = 0; c < k; ++c) det *= G[c * GP + c];
    return det;
}

// Y = G^-1 by Gauss-Jordan with partial pivoting (G destroyed); warp 0 only
__device__ void invert_warp(double* G, double* Y, double* col, int GP, int k) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < k * k; i += 32) Y[(i / k) * GP + i % k] = (i / k == i % k) ? 1.0 : 0.0;
    __syncwarp();
    for (int c = 0; c < k; ++c) {
        const int p = pivot_row(G, GP, k, c);
        if (p != c) {
            for (int j = lane; j < 2 * k; j += 32) {
                double* M = j < k ? G : Y;
                const int jj = j < k ? j : j - k;
                const double t = M[c * GP + jj];
                M[c * GP + jj] = M[p * GP + jj];
                M[p * GP + jj] = t;
            }
            __syncwarp();
        }
        const double inv_d = 1.0 / G[c * GP + c];
        __syncwarp();
        for (int j = lane; j < 2 * k; j += 32) {
            double* M = j < k ? G : Y;
            M[c * GP + (j < k ? j : j - k)] *= inv_d;
        }
        __syncwarp();
        if (lane < k) col[lane] = G[lane * GP + c];  // column c before the elimination
        __syncwarp();
        for (int i = lane; i < k * 2 * k; i += 32) {
            const int r = i / (2 * k), j = i - r * 2 * k;
            if (r == c) continue;
            double* M = j < k ? G : Y;
            const int jj = j < k ? j : j - k;
            M[r * GP + jj] -= col[r] * M[c * GP + jj];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Block-wide versions for CTAs of a few warps (the cluster kernels: 256 threads): the pivot search stays on warp 0, the
// row operations of a column step are spread over all threads -- one element per thread instead of seven rounds of one warp --
// with block barriers in between.  Same operations on the same values: bit-identical to the warp versions.  (With 1024 threads
// the barriers cost more than they save; the single-CTA kernel keeps the warp versions.)
// ---------------------------------------------------------------------------------------------------------
template <int NTH>
__device__ double lu_det_block(double* G, int* piv, int GP, int k) {  // all threads; det on every thread
    const int tid = threadIdx.x;
    double sign = 1.0;
    for (int c = 0; c < k; ++c) {
        if (tid < 32) {
            const int p = pivot_row(G, GP, k, c);
            if (tid == 0) *piv = p;
        }
        __syncthreads();
        const int p = *piv;
        if (p != c) {
            sign = -sign;
            if (tid < k) {
                const double t = G[c * GP + tid];
                G[c * GP + tid] = G[p * GP + tid];
                G[p * GP + tid] = t;
            }
            __syncthreads();
        }
        const double d = G[c * GP + c];
        const int nr = k - c - 1;
        if (tid < nr) G[(c + 1 + tid) * GP + c] /= d;
        __syncthreads();
        for (int i = tid; i < nr * nr; i += NTH) {
            const int r = c + 1 + i / nr, j = c + 1 + i % nr;
            G[r * GP + j] -= G[r * GP + c] * G[c * GP + j];
        }
        __syncthreads();
    }
    double det = sign;
    for (int c = 0; c < k; ++c) det *= G[c * GP + c];
    __syncthreads();  // (G may be overwritten by the caller now)
    return det;
}

template <int NTH>
__device__ void invert_block(double* G, double* Y, double* col, int* piv, int GP, int k) {  // all threads; Y = G^-1
    const int tid = threadIdx.x;
    for (int i = tid; i < k * k; i += NTH) Y[(i / k) * GP + i % k] = (i / k == i % k) ? 1.0 : 0.0;
    __syncthreads();
    for (int c = 0; c < k; ++c) {
        if (tid < 32) {
            const int p = pivot_row(G, GP, k, c);
            if (tid == 0) *piv = p;
        }
        __syncthreads();
        const int p = *piv;
        if (p != c) {
            for (int j = tid; j < 2 * k; j += NTH) {
                double* M = j < k ? G : Y;
                const int jj = j < k ? j : j - k;
                const double t = M[c * GP + jj];
                M[c * GP + jj] = M[p * GP + jj];
                M[p * GP + jj] = t;
            }
            __syncthreads();
        }
        const double inv_d = 1.0 / G[c * GP + c];
        if (tid < k && tid != c) col[tid] = G[tid * GP + c];  // column c before the elimination (row c is scaled below)
        __syncthreads();                                     // everybody has read G[c][c]
        for (int j = tid; j < 2 * k; j += NTH) {
            double* M = j < k ? G : Y;
            M[c * GP + (j < k ? j : j - k)] *= inv_d;
        }
        __syncthreads();
        for (int i = tid; i < k * 2 * k; i += NTH) {
            const int r = i / (2 * k), j = i - r * 2 * k;
            if (r == c) continue;
            double* M = j < k ? G : Y;
            const int jj = j < k ? j : j - k;
            M[r * GP + jj] -= col[r] * M[c * GP + jj];
        }
        __syncthreads();
    }
}

}  // namespace
