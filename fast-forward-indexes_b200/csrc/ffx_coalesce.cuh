// ffx_coalesce.cuh — sequential coalescing of every document's passage vectors on the device.
//
// Replaces the per-document host loop of `create_coalesced_index` (util/__init__.py:51-101 of the
// reference) for its default distance, `cos_dist` (:40-48): walk a document's vectors in row
// order, keep a running group; a vector whose cosine distance to the group's mean is >= delta
// closes the group (its mean is emitted) and starts a new one.  One warp owns one document.
// Arithmetic follows numpy's: the group sum is accumulated row by row in fp32 and divided by the
// count in fp32 (`np.mean(A, axis=0)`), the distance is `1 - dot / (|a| |b|)` evaluated in fp32 from
// the dot product and the two squared norms.  Those three sums are accumulated here in double and
// rounded to fp32 once (the reference takes them from BLAS `sdot`, whose summation order is not
// specified): a grouping decision can differ from the reference's only when the distance is
// within fp32 rounding (~1e-7 relative) of delta.  Given the same decisions the emitted means are
// bit-identical.  The store's element permutation (lane-major rows) is irrelevant to the sums
// and undone when a mean is written out.  HBM-bound (every row is read once); offline tooling.
#pragma once
#include "ffx_kernels.cuh"

namespace ffx {

constexpr int kCoalesceWarps = 4;

struct CoalesceArgs {
    const float *vectors;     // the row store
    int stride;               // floats per stored row
    int dim;                  // floats per output row
    int cpl, steps, lanes;    // lane-major plan of the store (cpl == 0: original element order)
    const uint2 *doc_span;
    const int32_t *doc_rows;
    int indirect;
    int64_t doc0, n_docs;     // documents [doc0, doc0 + n_docs)
    const int64_t *out_off;   // [n_docs + 1] first output row of every document (cumulative row counts)
    double delta;
    float *out;               // [out_off[n_docs], dim] original element order
    int32_t *out_groups;      // [n_docs] groups emitted per document
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__global__ void __launch_bounds__(kCoalesceWarps * 32) ffx_coalesce_kernel(const CoalesceArgs a) {
    extern __shared__ float co_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sum = co_smem + static_cast<size_t>(warp) * 2 * a.stride;
    float *mean = sum + a.stride;
    const int64_t d = static_cast<int64_t>(blockIdx.x) * kCoalesceWarps + warp;
    if (d >= a.n_docs) return;
    const uint2 span = a.doc_span[a.doc0 + d];
    const uint32_t cnt = span.y;
    float *out = a.out + a.out_off[d] * a.dim;
    int groups = 0, members = 0;

    auto emit = [&]() {
        float *dst = out + static_cast<int64_t>(groups) * a.dim;
        for (int e = lane; e < a.stride; e += 32) {
            const int orig = a.cpl ? ffx_orig_index(a.cpl, a.steps, e, a.lanes) : e;
            if (orig < a.dim) dst[orig] = mean[e];
        }
        groups++;
    };

    for (uint32_t r = 0; r < cnt; r++) {
        const uint32_t row = a.indirect ? static_cast<uint32_t>(a.doc_rows[span.x + r]) : span.x + r;
        const float *v = a.vectors + static_cast<size_t>(row) * a.stride;
        bool fresh = r == 0;
        if (!fresh) {
            double dot = 0.0, nv = 0.0, nm = 0.0;
            for (int e = lane; e < a.stride; e += 32) {
                const double x = v[e], m = mean[e];
                dot += x * m;
                nv += x * x;
                nm += m * m;
            }
            dot = warp_sum_f64(dot);
            nv = warp_sum_f64(nv);
            nm = warp_sum_f64(nm);
            // util/__init__.py:48 in float32: 1 - np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
            const float denom = __fmul_rn(sqrtf(static_cast<float>(nv)), sqrtf(static_cast<float>(nm)));
            const float dist = __fsub_rn(1.0f, __fdiv_rn(static_cast<float>(dot), denom));
            if (static_cast<double>(dist) >= a.delta) {
                __syncwarp();
                emit();
                fresh = true;
            }
        }
        __syncwarp();
        members = fresh ? 1 : members + 1;
        const float k = static_cast<float>(members);
        for (int e = lane; e < a.stride; e += 32) {
            const float s = fresh ? v[e] : __fadd_rn(sum[e], v[e]);
            sum[e] = s;
            mean[e] = __fdiv_rn(s, k);
        }
        __syncwarp();
    }
    if (cnt > 0) emit();
    if (lane == 0) a.out_groups[d] = groups;
}

}  // namespace ffx
