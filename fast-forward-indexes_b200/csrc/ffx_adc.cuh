// ffx_adc.cuh — asymmetric-distance scoring over PQ / OPQ codes.
//
// Replaces `Quantizer.decode` + the dot products of `Index._compute_scores`
// (quantizer/base.py:123-132, quantizer/nanopq.py:43-44,111-112, index/base.py:292-303):
// instead of materialising decoded [rows, D] fp32 vectors,
//     q . (dec(c) R^T) = (q R) . dec(c) = sum_m LUT[m][c_m],
//     LUT[m][k] = (qR)[m*Ds:(m+1)*Ds] . codewords[m][k]
// The LUT (M*Ks fp32, 96 KB for M=96, Ks=256) lives in shared memory, one query per CTA.
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

__global__ void __launch_bounds__(kThreads, 2) ffx_adc_kernel(const AdcArgs a) {
    extern __shared__ float s_lut[];  // [M, Ks]
    const int64_t q_idx = blockIdx.x / a.tiles_per_query;
    const int t_idx = blockIdx.x % a.tiles_per_query;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);
    const int c0 = t_idx * a.tile;
    const int n_tile = min(a.tile, n_query - c0);
    if (n_tile <= 0) return;

    const int D = a.M * a.Ds;
    const float *qe = a.qeff + q_idx * D;
    for (int e = threadIdx.x; e < a.M * a.Ks; e += kThreads) {
        const int m = e / a.Ks;
        const float *cw = a.codewords + static_cast<size_t>(e) * a.Ds;
        float acc = 0.f;
        for (int d = 0; d < a.Ds; d++) acc = fmaf(qe[m * a.Ds + d], cw[d], acc);
        s_lut[e] = acc;
    }
    __syncthreads();

    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    for (int c = threadIdx.x; c < n_tile; c += kThreads) {
        const int64_t p = q_begin + c0 + c;
        const int32_t u = a.cand[p];
        uint32_t start = 0, cnt = 0, loc = 0;
        if (!candidate_ok(u, a.limit, a.err, p)) {
            cnt = 0;
        } else if (!candidate_mine(u, a.base, a.count, &loc)) {
            if (a.rank_scores) a.rank_scores[p] = __int_as_float(0x7fc00000);
            continue;
        } else if (a.mode == FFX_MODE_PASSAGE) {
            start = loc;
            cnt = 1;
        } else {
            const uint2 sp = a.doc_span[loc];
            start = sp.x;
            cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
        }
        DocReduce red;
        red.init();
        for (uint32_t r = 0; r < cnt; r++) {
            const uint32_t row = indirect ? static_cast<uint32_t>(a.doc_rows[start + r]) : start + r;
            const uint8_t *code = a.codes + static_cast<size_t>(row) * a.M;
            float s = 0.f;
            if ((a.M & 15) == 0) {
                const uint4 *c16 = reinterpret_cast<const uint4 *>(code);
                for (int m0 = 0; m0 < a.M; m0 += 16) {
                    const uint4 w = __ldg(c16 + (m0 >> 4));
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        const uint32_t cc = (ws[t >> 2] >> (8 * (t & 3))) & 0xffu;
                        s += s_lut[(m0 + t) * a.Ks + cc];
                    }
                }
            } else {
                for (int m = 0; m < a.M; m++) s += s_lut[m * a.Ks + code[m]];
            }
            red.add(s, r == 0, a.mode);
        }
        const float ff = red.finish(cnt, a.mode);
        float inter = ff;
        if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, a.lex[p]), __fmul_rn(a.beta, ff));
        if (a.out_ff) a.out_ff[p] = ff;
        if (a.out_int) a.out_int[p] = inter;
        if (a.rank_scores) a.rank_scores[p] = inter;
    }
}

}  // namespace ffx
