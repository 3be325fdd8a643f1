// ffx_pq_build.cuh — the build side of product quantization on the GPU.
//
// Replaces what the reference delegates to nanopq 0.2.1 / scipy on the host
// (quantizer/nanopq.py:30,98 `fit` -> scipy.cluster.vq.kmeans2 per subspace;
// quantizer/nanopq.py:41,109 `encode` -> scipy.cluster.vq.vq per subspace): nearest-codeword
// assignment and Lloyd iterations.  Off the scoring path, but hours of CPU at corpus scale
// (20 M x 768: 3.9e12 multiply-adds per encode pass) against a fraction of a second here.
//
//   ffx_pq_assign_kernel<DS>  codes[n, M] = argmin_k |x[n, m*Ds:(m+1)*Ds] - codewords[m, k]|^2
//   ffx_pq_lloyd_kernel<DS>   the same assignment, accumulated into per-codeword sums / counts
//   ffx_pq_means_kernel       codewords = sums / counts (a codeword without members keeps its
//                             value, like kmeans2(missing="warn"))
//
// One CTA works on one subspace m and a tile of vectors: the subspace's codebook (Ks x Ds
// floats, 8 KB at Ks = 256, Ds = 8) sits in shared memory and is read as broadcast float4s —
// 4 FMAs per lane per shared-memory wavefront, so the FP32 pipe is the bound, not the LSU.
// A thread keeps its sub-vector in registers (DS = 4, 8, 16, 32) or streams it (DS = 0: any Ds).
// Distances are accumulated directly as sum (x - c)^2 in fp32; ties go to the lower index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ffx {

constexpr int kPqThreads = 256;

template <int DS>
__device__ __forceinline__ int pq_nearest(const float *x_global, int Ds, const float *s_cw, int Ks) {
    float best = INFINITY;
    int arg = 0;
    if (DS > 0) {
        float x[DS > 0 ? DS : 1];
#pragma unroll
        for (int d = 0; d < DS; d += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x_global + d));
            x[d] = v.x;
            x[d + 1] = v.y;
            x[d + 2] = v.z;
            x[d + 3] = v.w;
        }
        for (int k = 0; k < Ks; k++) {
            const float4 *c4 = reinterpret_cast<const float4 *>(s_cw + k * DS);
            float dist = 0.f;
#pragma unroll
            for (int d = 0; d < DS; d += 4) {
                const float4 c = c4[d >> 2];
                const float e0 = x[d] - c.x, e1 = x[d + 1] - c.y, e2 = x[d + 2] - c.z, e3 = x[d + 3] - c.w;
                dist = fmaf(e0, e0, dist);
                dist = fmaf(e1, e1, dist);
                dist = fmaf(e2, e2, dist);
                dist = fmaf(e3, e3, dist);
            }
            if (dist < best) {
                best = dist;
                arg = k;
            }
        }
    } else {
        for (int k = 0; k < Ks; k++) {
            const float *c = s_cw + k * Ds;
            float dist = 0.f;
            for (int d = 0; d < Ds; d++) {
                const float e = __ldg(x_global + d) - c[d];
                dist = fmaf(e, e, dist);
            }
            if (dist < best) {
                best = dist;
                arg = k;
            }
        }
    }
    return arg;
}

// grid (tiles, M); vecs [n, M*Ds] row-major
template <int DS>
__global__ void __launch_bounds__(kPqThreads) ffx_pq_assign_kernel(const float *vecs, int64_t n, int M, int Ks, int Ds,
                                                                   const float *codewords, uint8_t *codes) {
    extern __shared__ __align__(16) float s_cw[];  // [Ks, Ds]
    const int m = blockIdx.y;
    for (int i = threadIdx.x; i < Ks * Ds; i += kPqThreads) s_cw[i] = codewords[static_cast<size_t>(m) * Ks * Ds + i];
    __syncthreads();
    const int64_t D = static_cast<int64_t>(M) * Ds;
    for (int64_t r = blockIdx.x * static_cast<int64_t>(kPqThreads) + threadIdx.x; r < n;
         r += static_cast<int64_t>(gridDim.x) * kPqThreads)
        codes[r * M + m] = static_cast<uint8_t>(pq_nearest<DS>(vecs + r * D + static_cast<int64_t>(m) * Ds, Ds, s_cw, Ks));
}

// One Lloyd accumulation pass: sums [M, Ks, Ds] (double), counts [M, Ks] (unsigned long long) must
// be zero on entry.  Per-CTA partial sums live in shared memory (fp32 atomics) and are flushed once.
template <int DS>
__global__ void __launch_bounds__(kPqThreads) ffx_pq_lloyd_kernel(const float *vecs, int64_t n, int M, int Ks, int Ds,
                                                                  const float *codewords, double *sums,
                                                                  unsigned long long *counts, int rows_per_cta) {
    extern __shared__ __align__(16) float s_mem[];  // [Ks, Ds] codebook | [Ks, Ds] partial sums | [Ks] partial counts
    float *s_cw = s_mem;
    float *s_sum = s_mem + Ks * Ds;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_sum + Ks * Ds);
    const int m = blockIdx.y;
    for (int i = threadIdx.x; i < Ks * Ds; i += kPqThreads) {
        s_cw[i] = codewords[static_cast<size_t>(m) * Ks * Ds + i];
        s_sum[i] = 0.f;
    }
    for (int i = threadIdx.x; i < Ks; i += kPqThreads) s_cnt[i] = 0u;
    __syncthreads();
    const int64_t D = static_cast<int64_t>(M) * Ds;
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
    const int64_t hi = min(n, lo + rows_per_cta);
    for (int64_t r = lo + threadIdx.x; r < hi; r += kPqThreads) {
        const float *x = vecs + r * D + static_cast<int64_t>(m) * Ds;
        const int k = pq_nearest<DS>(x, Ds, s_cw, Ks);
        atomicAdd(&s_cnt[k], 1u);
        for (int d = 0; d < Ds; d++) atomicAdd(&s_sum[k * Ds + d], __ldg(x + d));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Ks * Ds; i += kPqThreads)
        if (s_sum[i] != 0.f) atomicAdd(&sums[static_cast<size_t>(m) * Ks * Ds + i], static_cast<double>(s_sum[i]));
    for (int i = threadIdx.x; i < Ks; i += kPqThreads)
        if (s_cnt[i]) atomicAdd(&counts[static_cast<size_t>(m) * Ks + i], static_cast<unsigned long long>(s_cnt[i]));
}

__global__ void ffx_pq_means_kernel(float *codewords, const double *sums, const unsigned long long *counts, int64_t total,
                                    int Ds) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const unsigned long long c = counts[i / Ds];
        if (c) codewords[i] = static_cast<float>(sums[i] / static_cast<double>(c));
    }
}

// ---- OPQ rotation update (quantizer/nanopq.py:94-98 -> nanopq OPQ.fit) ----------------------------
// Each round of OPQ training is two matrix products around the k-means: X = vecs @ R  ([N, D] x
// [D, D]) and the Procrustes matrix vecs^T @ X_hat ([D, N] x [N, D]); the D x D SVD stays on the host.
// One register-tiled fp32 kernel serves both: C[m, n] (+)= op(A)[m, k] . B[k, n], 64 x 64 tile per
// CTA, 4 x 4 per thread, 16 values of k staged through shared memory; gridDim.z slices of k write
// their own partial C (reduced in fixed order by ffx_sgemm_reduce_kernel: deterministic, and the
// long sums over N are accumulated in shorter pieces).  fp32 FMA throughout — TF32 / BF16 tensor
// cores would drop the products to 10 / 7 mantissa bits, and the rotation must stay orthogonal to
// ~1e-6 for the ADC identity q.(dec(c) R^T) = (q R).dec(c) that scoring relies on.
constexpr int kGemmTile = 64, kGemmK = 16;

template <bool TRANS_A>
__global__ void __launch_bounds__(256) ffx_sgemm_kernel(const float *A, const float *B, float *C, int64_t M, int64_t N,
                                                        int64_t K, int64_t k_per_slice) {
    __shared__ float s_a[kGemmK][kGemmTile + 4];
    __shared__ float s_b[kGemmK][kGemmTile + 4];
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * kGemmTile, n0 = static_cast<int64_t>(blockIdx.x) * kGemmTile;
    const int64_t k_lo = static_cast<int64_t>(blockIdx.z) * k_per_slice;
    const int64_t k_hi = k_lo + k_per_slice < K ? k_lo + k_per_slice : K;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.f;
    for (int64_t k0 = k_lo; k0 < k_hi; k0 += kGemmK) {
#pragma unroll
        for (int e = threadIdx.x; e < kGemmK * kGemmTile; e += 256) {
            // op(A) tile: [16 k][64 m]
            if constexpr (TRANS_A) {  // A is [K, M]: rows of A run along m
                const int kk = e / kGemmTile, mm = e % kGemmTile;
                const int64_t k = k0 + kk, m = m0 + mm;
                s_a[kk][mm] = (k < k_hi && m < M) ? __ldg(A + k * M + m) : 0.f;
            } else {  // A is [M, K]: rows of A run along k
                const int mm = e / kGemmK, kk = e % kGemmK;
                const int64_t k = k0 + kk, m = m0 + mm;
                s_a[kk][mm] = (k < k_hi && m < M) ? __ldg(A + m * K + k) : 0.f;
            }
            const int kk = e / kGemmTile, nn = e % kGemmTile;
            const int64_t k = k0 + kk, n = n0 + nn;
            s_b[kk][nn] = (k < k_hi && n < N) ? __ldg(B + k * N + n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; r++) a[r] = s_a[kk][4 * ty + r];
#pragma unroll
            for (int c = 0; c < 4; c++) b[c] = s_b[kk][4 * tx + c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
    float *out = C + static_cast<int64_t>(blockIdx.z) * M * N;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int64_t m = m0 + 4 * ty + r, n = n0 + 4 * tx + c;
            if (m < M && n < N) out[m * N + n] = acc[r][c];
        }
}

// C[i] = sum over slices, in slice order
__global__ void ffx_sgemm_reduce_kernel(const float *partial, float *C, int64_t elems, int slices) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < elems;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float acc = partial[i];
        for (int s = 1; s < slices; s++) acc += partial[static_cast<int64_t>(s) * elems + i];
        C[i] = acc;
    }
}

}  // namespace ffx
