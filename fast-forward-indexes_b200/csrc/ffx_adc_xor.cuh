// ffx_adc_xor.cuh — asymmetric-distance scoring, thread-per-row with XOR-swizzled look-ups.
//
// Same contract as ffx_adc_kernel (ffx_adc.cuh): q . dec(c) = sum_m LUT[m][c_m] over uint8 PQ /
// OPQ codes, per-document max / mean / first, interpolation and (FUSE) the per-query top-k.
// Replaces quantizer/base.py:123-132 -> quantizer/nanopq.py:43-44,111-112 + index/base.py:292-312
// + ranking.py:319,115-117,285-291 of the reference.
//
// One thread scores one passage row (no cross-lane reduction), and every shared-memory
// look-up of a warp is bank-conflict free although the code values are random:
//   * the table is stored as lut[m / 32][c][m % 32]: sub-quantizer m of code value c lives in
//     bank m % 32 of a 128-byte line;
//   * at look-up step p (0..31 inside a 32-code chunk) lane t handles sub-quantizer p ^ t of ITS
//     row, so the 32 lanes of a warp always hit 32 different banks;
//   * "the byte at static position p is the code of sub-quantizer p ^ t" is arranged without
//     dynamic register indexing: the two 16-byte loads of a chunk swap places by bit 4 of t
//     (addresses), the four words of each load are exchanged by bits 2 and 3 of t (8 selects),
//     and bits 0..1 pick the byte through the dp4a selector (4 per-lane constants);
//   * the address of a look-up is one LOP3 + one dp4a:
//         dp4a(word, 0x80 << 8*((k ^ t) & 3), (table | 4t) ^ 4p)  =  table + 128*c + 4*(p ^ t).
// Rows of the 32 candidates a warp grabs are flattened (warp scan), lane r takes stream row r,
// a segmented scan folds row scores into documents in row order and lane j pulls candidate j.
//
// Shapes: M % 32 == 0, M <= 128 (NC = M / 32 chunks); shared memory M * Ks * 4 bytes of table
// (96 KB at M = 96, Ks = 256) + 4 B per candidate (FUSE).  1024 threads per CTA, one CTA per SM.
// Bound: shared-memory look-ups (3 conflict-free LDS per row at M = 96) and instruction issue.
#pragma once
#include "ffx_adc_warp.cuh"

namespace ffx {

constexpr int kAdcXorThreads = 1024;

// codewords [M][Ks][Ds] -> cw_x [M/32][Ks][32][Ds]: the order in which the table is written
__global__ void ffx_adc_xor_codewords_kernel(const float *cw, int M, int Ks, int Ds, float *cw_x) {
    const int64_t total = static_cast<int64_t>(M) * Ks * Ds;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(t % Ds);
        const int64_t e = t / Ds;
        const int b = static_cast<int>(e & 31);
        const int c = static_cast<int>((e >> 5) % Ks);
        const int j3 = static_cast<int>(e / (32ll * Ks));
        cw_x[t] = cw[(static_cast<size_t>(32 * j3 + b) * Ks + c) * Ds + d];
    }
}

__host__ __device__ inline size_t adc_xor_smem_bytes(int M, int Ks, int cpad_scores) {
    const size_t lut = static_cast<size_t>(M) * Ks * 4;
    const size_t keys = static_cast<size_t>(cpad_scores) * 8;  // sort keys overlay the dead table
    return ((static_cast<size_t>(cpad_scores) * 4 + 127) & ~static_cast<size_t>(127)) + (lut > keys ? lut : keys);
}

__device__ __forceinline__ uint4 ldg_u4(const void *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

// AdcWarpArgs::cw_t holds cw_x here.
template <int NC, bool FUSE>
__global__ void __launch_bounds__(kAdcXorThreads, 1) ffx_adc_xor_kernel(const AdcWarpArgs w) {
    const AdcArgs &a = w.base;
    extern __shared__ __align__(128) unsigned char adc_smem[];
    float *s_scores = reinterpret_cast<float *>(adc_smem);  // [cpad] (FUSE), padded so the table stays 128-byte aligned
    float *s_lut = reinterpret_cast<float *>(
        adc_smem + (FUSE ? ((static_cast<size_t>(w.cpad) * 4 + 127) & ~static_cast<size_t>(127)) : 0));
    __shared__ int s_next;

    const int lane = threadIdx.x & 31;
    const int64_t q_idx = blockIdx.x / a.tiles_per_query;
    const int t_idx = blockIdx.x % a.tiles_per_query;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);
    const int c0 = t_idx * a.tile;
    const int n_tile = min(a.tile, n_query - c0);
    if (!FUSE && n_tile <= 0) return;
    if (threadIdx.x == 0) s_next = 0;
    if (FUSE) {
        for (int i = threadIdx.x; i < n_query; i += blockDim.x) s_scores[i] = __int_as_float(0x7fc00000);
    }

    constexpr int M = NC * 32;
    // ---- per-query table: s_lut[(j3*Ks + c)*32 + b] = qeff[m*Ds..] . codewords[m][c],  m = 32*j3 + b
    {
        const float *qe = a.qeff + q_idx * (static_cast<int64_t>(M) * a.Ds);
        const int total = M * a.Ks;
        const int per_table = a.Ks * 32;
        const bool vec4 = (a.Ds & 3) == 0 && (reinterpret_cast<uintptr_t>(qe) & 15) == 0;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int m = 32 * (e / per_table) + (e & 31);
            const float *cw = w.cw_t + static_cast<size_t>(e) * a.Ds;
            const float *qm = qe + m * a.Ds;
            float acc = 0.f;
            if (vec4) {
                for (int d = 0; d < a.Ds; d += 4) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4 *>(cw + d));
                    const float4 q4 = __ldg(reinterpret_cast<const float4 *>(qm + d));
                    acc = fmaf(q4.x, c4.x, acc);
                    acc = fmaf(q4.y, c4.y, acc);
                    acc = fmaf(q4.z, c4.z, acc);
                    acc = fmaf(q4.w, c4.w, acc);
                }
            } else {
                for (int d = 0; d < a.Ds; d++) acc = fmaf(__ldg(qm + d), __ldg(cw + d), acc);
            }
            s_lut[e] = acc;
        }
    }
    __syncthreads();

    // ---- per-lane constants of the XOR swizzle (t = lane)
    const uint32_t t = static_cast<uint32_t>(lane);
    uint32_t sel[4];  // dp4a selector of static byte position k: 0x80 in byte (k ^ t) & 3
#pragma unroll
    for (int k = 0; k < 4; k++) sel[k] = 0x80u << (8u * ((static_cast<uint32_t>(k) ^ t) & 3u));
    const uint32_t tab_bytes = static_cast<uint32_t>(a.Ks) * 128u;
    uint32_t tab[NC];  // table j3 of this lane, low 7 bits = 4t (the table is 128-byte aligned)
#pragma unroll
    for (int j = 0; j < NC; j++)
        tab[j] = (static_cast<uint32_t>(__cvta_generic_to_shared(s_lut)) + static_cast<uint32_t>(j) * tab_bytes) | (t << 2);
    const uint32_t off_a = (t & 16u) ? 16u : 0u;  // which 16-byte half of a chunk lands in positions 0..3
    const uint32_t off_b = 16u - off_a;
    const bool sw1 = (t & 4u) != 0, sw2 = (t & 8u) != 0;
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    const bool segmented = a.mode == FFX_MODE_MAXP || a.mode == FFX_MODE_AVEP;
    const bool is_max = a.mode == FFX_MODE_MAXP;

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_next, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n_tile) break;
        const int nb = min(32, n_tile - base);
        const int64_t p = q_begin + c0 + base + lane;

        // lane j resolves candidate j of the batch
        uint32_t start = 0, cnt = 0;
        bool mine = false;
        if (lane < nb) {
            const int32_t u = __ldg(a.cand + p);
            uint32_t loc = 0;
            if (!candidate_ok(u, a.limit, a.err, p)) {
                mine = true;  // reported; scores as an empty document
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
        // flattened row stream of the batch: inclusive scan of the row counts
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t pre = incl - cnt;
        const uint32_t total = __shfl_sync(kFull, incl, 31);

        float acc = 0.f;  // lane j: running max / sum / first of candidate j
        bool acc_set = false;

        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t g = r0 + t;  // this lane's row of the stream
            // owner of stream row g: first candidate whose inclusive count exceeds it
            int c = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(kFull, incl, c + step - 1);
                if (v <= g) c += step;
            }
            c = min(c, 31);
            const uint32_t my_k = g - __shfl_sync(kFull, pre, c);  // position inside the document
            const uint32_t c_start = __shfl_sync(kFull, start, c);

            float s = 0.f;
            if (g < total) {
                uint32_t row = c_start + my_k;
                if (indirect) row = static_cast<uint32_t>(__ldg(a.doc_rows + row));
                const uint8_t *code = a.codes + static_cast<uint64_t>(row) * M;
                uint4 qa[NC], qb[NC];
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    qa[j] = ldg_u4(code + 32 * j + off_a);
                    qb[j] = ldg_u4(code + 32 * j + off_b);
                }
                float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    uint32_t wd[8] = {qa[j].x, qa[j].y, qa[j].z, qa[j].w, qb[j].x, qb[j].y, qb[j].z, qb[j].w};
#pragma unroll
                    for (int h = 0; h < 8; h += 4) {  // position i <- word i ^ ((t >> 2) & 3)
                        const uint32_t x0 = sw1 ? wd[h + 1] : wd[h + 0], x1 = sw1 ? wd[h + 0] : wd[h + 1];
                        const uint32_t x2 = sw1 ? wd[h + 3] : wd[h + 2], x3 = sw1 ? wd[h + 2] : wd[h + 3];
                        wd[h + 0] = sw2 ? x2 : x0;
                        wd[h + 1] = sw2 ? x3 : x1;
                        wd[h + 2] = sw2 ? x0 : x2;
                        wd[h + 3] = sw2 ? x1 : x3;
                    }
#pragma unroll
                    for (int wi = 0; wi < 8; wi++) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t pos = static_cast<uint32_t>(4 * wi + k);
                            acc4[k] += lds_f32(__dp4a(wd[wi], sel[k], tab[j] ^ (pos << 2)));
                        }
                    }
                }
                s = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
            }

            // fold rows into documents (segmented inclusive scan in row order), then lane j pulls
            // candidate j's partial from the lane of its last row in this block
            if (segmented) {
                const uint32_t dist = min(my_k, t);  // same-document rows to the left
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float v = __shfl_up_sync(kFull, s, d);
                    if (dist >= static_cast<uint32_t>(d)) s = is_max ? fmaxf(s, v) : s + v;
                }
            }
            const bool overlap = cnt > 0 && pre < r0 + 32 && pre + cnt > r0;
            const uint32_t last = min(pre + cnt - 1, r0 + 31) - r0;
            const float got = __shfl_sync(kFull, s, overlap ? last : 0);
            if (overlap) {
                if (!acc_set) acc = got;
                else acc = is_max ? fmaxf(acc, got) : acc + got;
                acc_set = true;
            }
        }

        if (lane < nb) {
            if (mine) {
                float ff = acc;
                if (a.mode == FFX_MODE_AVEP) ff = __fdiv_rn(acc, static_cast<float>(cnt));
                float inter = ff;
                if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, __ldg(a.lex + p)), __fmul_rn(a.beta, ff));
                if (a.out_ff) a.out_ff[p] = ff;
                if (a.out_int) a.out_int[p] = inter;
                if (a.rank_scores) a.rank_scores[p] = inter;
                if (FUSE) s_scores[c0 + base + lane] = inter;
            } else if (a.rank_scores) {
                a.rank_scores[p] = __int_as_float(0x7fc00000);
            }
        }
    }

    if (FUSE) {
        // the table is dead: build the 64-bit sort keys over it
        __syncthreads();
        unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(s_lut);
        for (int i = threadIdx.x; i < w.cpad; i += blockDim.x)
            s_keys[i] = i < n_query ? topk_key(s_scores[i], static_cast<uint32_t>(i)) : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_keys, w.cpad);
        write_topk(s_keys, n_query, w.k, w.topk_score + q_idx * w.k, w.topk_pos + q_idx * w.k);
    }
}

}  // namespace ffx
