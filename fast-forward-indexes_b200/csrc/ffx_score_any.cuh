// ffx_score_any.cuh — the dot product for every dimension WITHOUT a uniform numpy tree (D = 100,
// 130, 300, 1000, ... up to 4096; the uniform ones have the lane-major layout of ffx_layout.h).
// The summation tree is DATA (ffx_any_plan, ffx_layout.h): rows stay in original element order,
// every lane owns CPL chains (leaf slot, accumulator j) of the numpy tree and reads their elements
// from the staged row with computed addresses; the query vector sits beside the ring in shared
// memory.  Rows of one or two leaves are shared by 4 / 8 lanes (8 / 4 rows per warp step).  The
// kernel around it is ffx_score_packed_kernel (ffx_score_packed.cuh, TreeDot).
//
// Replaces the thread-per-pair ffx_score_generic_kernel (index/base.py:279-314 for any D,
// bit-exact: N1 of DESIGN.md).
#pragma once
#include "ffx_score_tma.cuh"

namespace ffx {

template <int N>
struct AnyVec {
    float v[N];
};
template <int N>
__device__ __forceinline__ AnyVec<N> lds_vec(uint32_t addr) {
    AnyVec<N> r;
    if constexpr (N == 1) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r.v[0]) : "r"(addr));
    } else if constexpr (N == 2) {
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "r"(addr));
    } else if constexpr (N == 4) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                     : "r"(addr));
    } else {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                     : "r"(addr));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                     : "r"(addr + 16));
    }
    return r;
}

// One staged row (shared memory, original element order) against the query vector: numpy's
// np.sum(q * d) for this D.  Lane `sub` of the LPR lanes sharing the row owns chains
// sub*CPL .. sub*CPL+CPL-1 = accumulators j0 .. j0+CPL-1 of leaf slot (sub*CPL)/8:
//   chain sums (products rounded, sequential adds), balanced combine inside the lane, xor
//   butterfly over the lanes of the leaf, the last leaf's tail elements one by one, xor
//   butterfly over the leaf slots, 0 + total.
// A chain has at most 16 terms (a leaf is at most 128 elements).  The lane's slice of the query
// vector lives in registers when it is short (CPL <= 2), else beside the ring in shared memory.
constexpr int kAnyMaxSteps = 16;

// CPL <= 2 (up to 32 floats per lane): the query slice lives in registers and a row's elements are
// all loaded before the arithmetic; wider slices (D > 1024) stream both from shared memory step
// by step (measured: registers win 4.5 -> 6.8 TB/s at D = 1000, and lose 5.8 -> 2.5 TB/s at D = 1280)
template <int CPL>
struct AnyQuery {
    float v[CPL <= 2 ? CPL * kAnyMaxSteps : 1];
};

template <int CPL>
__device__ __forceinline__ void any_load_query(AnyQuery<CPL> &q, uint32_t qv, uint32_t my_byte, int my_steps) {
    if constexpr (CPL <= 2) {
#pragma unroll
        for (int s = 0; s < kAnyMaxSteps; s++) {
            AnyVec<CPL> x;
#pragma unroll
            for (int c = 0; c < CPL; c++) x.v[c] = 0.f;
            if (s < my_steps) x = lds_vec<CPL>(qv + my_byte + s * 32);
#pragma unroll
            for (int c = 0; c < CPL; c++) q.v[s * CPL + c] = x.v[c];
        }
    }
}

template <int CPL, int LPR>
__device__ __forceinline__ float any_row_dot(uint32_t row, uint32_t qv, const AnyQuery<CPL> &qr, uint32_t my_byte,
                                             int my_steps, int max_steps, bool tail_mine, uint32_t tail_byte,
                                             int tail_len) {
    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) acc[c] = 0.f;
    if constexpr (CPL <= 2) {
        AnyVec<CPL> d[kAnyMaxSteps];
#pragma unroll
        for (int s = 0; s < kAnyMaxSteps; s++) {
#pragma unroll
            for (int c = 0; c < CPL; c++) d[s].v[c] = 0.f;
            if (s < my_steps) d[s] = lds_vec<CPL>(row + my_byte + s * 32);
        }
#pragma unroll
        for (int s = 0; s < kAnyMaxSteps; s++) {
            if (s < my_steps) {
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    const float prod = __fmul_rn(qr.v[s * CPL + c], d[s].v[c]);
                    acc[c] = s == 0 ? prod : __fadd_rn(acc[c], prod);
                }
            }
        }
    } else {
#pragma unroll 4
        for (int s = 0; s < max_steps; s++) {  // warp-uniform bound: the longest chain of this dimension
            if (s < my_steps) {
                const AnyVec<CPL> d = lds_vec<CPL>(row + my_byte + s * 32);
                const AnyVec<CPL> q = lds_vec<CPL>(qv + my_byte + s * 32);
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    const float prod = __fmul_rn(q.v[c], d.v[c]);
                    acc[c] = s == 0 ? prod : __fadd_rn(acc[c], prod);
                }
            }
        }
    }
#pragma unroll
    for (int w = 1; w < CPL; w <<= 1) {
#pragma unroll
        for (int c = 0; c < CPL; c += 2 * w) acc[c] = __fadd_rn(acc[c], acc[c + w]);
    }
    float v = acc[0];
    constexpr int kLeafLanes = 8 / CPL;  // lanes that share one leaf slot
#pragma unroll
    for (int o = 1; o < kLeafLanes; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, o));
    if (tail_len) {  // warp-uniform
        if (tail_mine) {
            for (int t = 0; t < tail_len; t++) {
                const AnyVec<1> d = lds_vec<1>(row + tail_byte + 4 * t);
                const AnyVec<1> q = lds_vec<1>(qv + tail_byte + 4 * t);
                v = __fadd_rn(v, __fmul_rn(q.v[0], d.v[0]));
            }
        }
    }
#pragma unroll
    for (int o = kLeafLanes; o < LPR; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, o));
    return __fadd_rn(0.f, v);
}

// staging for dimensions that are not a multiple of 4: rows <-> the 16-byte padded store
__global__ void ffx_pad_rows_kernel(float *dst, const float *src, int64_t nrows, int dim, int stride, int to_store,
                                    const int64_t *rows) {
    const int64_t total = nrows * stride;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / stride;
        const int k = static_cast<int>(t % stride);
        if (to_store) {
            dst[r * stride + k] = k < dim ? src[r * dim + k] : 0.f;
        } else if (k < dim) {
            const int64_t sr = rows ? rows[r] : r;
            dst[r * dim + k] = src[sr * stride + k];
        }
    }
}

}  // namespace ffx
