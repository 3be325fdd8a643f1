// ffx_adc.cuh — asymmetric-distance scoring over PQ / OPQ codes.
//
// Replaces `Quantizer.decode` + the dot products of `Index._compute_scores`
// (quantizer/base.py:123-132, quantizer/nanopq.py:43-44,111-112, index/base.py:292-303):
// instead of materialising decoded [rows, D] fp32 vectors,
//     q . (dec(c) R^T) = (q R) . dec(c) = sum_m LUT[m][c_m],
//     LUT[m][k] = (qR)[m*Ds:(m+1)*Ds] . codewords[m][k]
// The LUT (M*Ks fp32, 96 KB for M=96, Ks=256) lives in shared memory, one query per CTA.
//
// Work decomposition: ONE THREAD PER PASSAGE ROW, not per document.  A warp takes 32
// candidates, flattens their rows (warp scan of the row counts), and walks the flattened
// sequence 32 rows at a time, so ragged documents (1..64 passages) cause no divergence;
// the per-document max / Kahan-mean / first is then folded in row order with shuffles.
// Bound: shared-memory lookups (96 per row, random banks), not HBM (96 B per row).
#pragma once
#include "ffx_kernels.cuh"

namespace ffx {

struct AdcArgs {
    const uint8_t *codes;     // [rows, M]
    const float *codewords;   // [M, Ks, Ds]
    const float *qeff;        // [nq, D]: q (PQ) or q @ R (OPQ)
    int M, Ks, Ds;
    const uint2 *doc_span;
    const int32_t *doc_rows;
    int indirect;
    int mode;
    const int64_t *q_off;
    const int32_t *cand;
    const float *lex;
    float alpha, beta;
    float *out_ff, *out_int, *rank_scores;
    int tiles_per_query, tile;
    uint32_t limit, base, count;
    int *err;
};

// qeff[q, j] = sum_i qvecs[q, i] * R[i, j]   (row vector times R; nanopq OPQ.rotate)
__global__ void __launch_bounds__(256) ffx_rotate_queries_kernel(const float *qvecs, const float *R,
                                                                 int D, float *qeff) {
    extern __shared__ float s_q[];
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_q[i] = qvecs[q * D + i];
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < D; i++) acc = fmaf(s_q[i], R[static_cast<size_t>(i) * D + j], acc);
        qeff[q * D + j] = acc;
    }
}

// One row: sum of M table entries.  V16 = M/16 when M is a multiple of 16 (all code bytes of
// the row are requested with V16 128-bit loads before the first lookup), 0 = any M.
template <int V16>
__device__ __forceinline__ float adc_row(const uint8_t *code, const float *lut, int M, int Ks) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (V16 > 0) {
        uint4 w[V16 > 0 ? V16 : 1];
        const uint4 *c16 = reinterpret_cast<const uint4 *>(code);
#pragma unroll
        for (int v = 0; v < V16; v++) w[v] = __ldg(c16 + v);
#pragma unroll
        for (int v = 0; v < V16; v++) {
            const uint32_t ws[4] = {w[v].x, w[v].y, w[v].z, w[v].w};
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const uint32_t cc = (ws[t >> 2] >> (8 * (t & 3))) & 0xffu;
                acc[t & 3] += lut[(v * 16 + t) * Ks + cc];
            }
        }
    } else {
        for (int m = 0; m < M; m++) acc[m & 3] += lut[m * Ks + code[m]];
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

template <int V16>
__global__ void __launch_bounds__(kThreads, 2) ffx_adc_kernel(const AdcArgs a) {
    extern __shared__ float s_lut[];  // [M, Ks]
    __shared__ int s_next;
    const int lane = threadIdx.x & 31;
    const int64_t q_idx = blockIdx.x / a.tiles_per_query;
    const int t_idx = blockIdx.x % a.tiles_per_query;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);
    const int c0 = t_idx * a.tile;
    const int n_tile = min(a.tile, n_query - c0);
    if (n_tile <= 0) return;
    // the scratch scores of a separate top-k pass are indexed relative to the launch's first pair
    float *rank = a.rank_scores ? a.rank_scores - a.q_off[0] : nullptr;
    if (threadIdx.x == 0) s_next = 0;

    // ---- per-query lookup table
    const int D = a.M * a.Ds;
    const float *qe = a.qeff + q_idx * D;
    for (int e = threadIdx.x; e < a.M * a.Ks; e += kThreads) {
        const int m = e / a.Ks;
        const float *cw = a.codewords + static_cast<size_t>(e) * a.Ds;
        float acc = 0.f;
        for (int d = 0; d < a.Ds; d++) acc = fmaf(__ldg(qe + m * a.Ds + d), __ldg(cw + d), acc);
        s_lut[e] = acc;
    }
    __syncthreads();

    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_next, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n_tile) break;
        const int nb = min(32, n_tile - base);
        const int64_t p = q_begin + c0 + base + lane;

        // lane j resolves candidate j
        uint32_t start = 0, cnt = 0;
        bool mine = false;
        if (lane < nb) {
            const int32_t u = __ldg(a.cand + p);
            uint32_t loc = 0;
            if (!candidate_ok(u, a.limit, a.err, p)) {
                mine = true;
            } else if (candidate_mine(u, a.base, a.count, &loc)) {
                mine = true;
                if (a.mode == FFX_MODE_PASSAGE) {
                    start = loc;
                    cnt = 1;
                } else {
                    const uint2 sp = __ldg(a.doc_span + loc);
                    start = sp.x;
                    cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
                }
            }
        }
        // flatten: inclusive scan of the row counts
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t pre = incl - cnt;
        const uint32_t total = __shfl_sync(kFull, incl, 31);

        DocReduce red;
        red.init();
        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t r = r0 + lane;
            // owner of row r: first candidate whose inclusive count exceeds r
            int c = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t t = __shfl_sync(kFull, incl, c + step - 1);
                if (t <= r) c += step;
            }
            c = min(c, 31);
            const uint32_t c_pre = __shfl_sync(kFull, pre, c);
            const uint32_t c_start = __shfl_sync(kFull, start, c);
            float s = 0.f;
            if (r < total) {
                uint32_t row = c_start + (r - c_pre);
                if (indirect) row = static_cast<uint32_t>(__ldg(a.doc_rows + row));
                s = adc_row<V16>(a.codes + static_cast<size_t>(row) * a.M, s_lut, a.M, a.Ks);
            }
            // fold this chunk's rows into their documents, in row order (N2 needs the order)
            const int lo = static_cast<int>(pre) - static_cast<int>(r0);
            const int hi = lo + static_cast<int>(cnt);
            const int n_here = static_cast<int>(min(32u, total - r0));
            for (int i = 0; i < n_here; i++) {
                const float v = __shfl_sync(kFull, s, i);
                if (i >= lo && i < hi) red.add(v, i == lo, a.mode);
            }
        }
        if (lane < nb) {
            if (mine) {
                const float ff = red.finish(cnt, a.mode);
                float inter = ff;
                if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, __ldg(a.lex + p)), __fmul_rn(a.beta, ff));
                if (a.out_ff) a.out_ff[p] = ff;
                if (a.out_int) a.out_int[p] = inter;
                if (rank) rank[p] = inter;
            } else if (rank) {
                rank[p] = __int_as_float(0x7fc00000);
            }
        }
    }
}

}  // namespace ffx
