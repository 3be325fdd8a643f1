// ffx_adc_warp.cuh — asymmetric-distance scoring, warp-per-row with bank-conflict-free tables.
//
// Same contract as ffx_adc_kernel (ffx_adc.cuh): q . dec(c) = sum_m LUT[m][c_m] over uint8 PQ /
// OPQ codes, per-document max / mean / first, interpolation and (FUSE) the per-query top-k.
// Replaces quantizer/base.py:123-132 -> quantizer/nanopq.py:43-44,111-112 + index/base.py:292-312
// + ranking.py:319,115-117,285-291 of the reference.
//
// Why a second kernel: with one thread per row the 96 table look-ups of a row hit random banks
// (the bank is the code value), ~3.5-way conflicts, and shared-memory look-ups are the bound
// of this path.  Here a warp works on ONE row at a time: lane l < W = M/4 loads the row's l-th
// 32-bit word (4 codes; one coalesced 4*W-byte request per row) and does 4 look-ups, byte j in
// table j.  The four tables are stored as lut[j][c][32]: sub-quantizer m = 4*l + j of code value
// c sits at word l of a 128-byte line, so the bank of a look-up is the LANE, whatever the codes
// are — every LDS is conflict-free — and the address of a look-up is ONE instruction:
// dp4a(word, 0x80 << 8j, table_j + 4*lane) = table_j + 128*c_j + 4*lane.  32 rows are processed
// per block, then one transposed butterfly (31 shuffles) hands lane r the score of row r, a
// segmented scan folds rows into their documents and lane j pulls the result of candidate j.
//
// Shapes: M % 4 == 0 and 16 <= M/4 <= 32 (M = 64..128); other M keep ffx_adc_kernel.
// Shared memory: 4 * Ks * 128 B of tables (128 KB at Ks = 256) + 4 B per candidate (FUSE).
// Bound: shared-memory look-ups + shuffles (LSU) and instruction issue; HBM sees M bytes per row.
#pragma once
#include "ffx_adc.cuh"

namespace ffx {

constexpr int kAdcWarpThreads = 512;  // 16 warps, one CTA per SM (the table takes ~half of shared memory)

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}

// codewords [M][Ks][Ds] -> cw_t [4][Ks][32][Ds]: entry (j, c, l) = codewords[4l + j][c] (zero for
// l >= W), the order in which the table build writes the tables.
__global__ void ffx_adc_transpose_codewords_kernel(const float *cw, int M, int Ks, int Ds, float *cw_t) {
    const int W = M / 4;
    const int64_t total = 4ll * Ks * 32 * Ds;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(t % Ds);
        const int64_t e = t / Ds;
        const int l = static_cast<int>(e % 32);
        const int c = static_cast<int>((e / 32) % Ks);
        const int j = static_cast<int>(e / (32ll * Ks));
        cw_t[t] = l < W ? cw[(static_cast<size_t>(4 * l + j) * Ks + c) * Ds + d] : 0.f;
    }
}

// qeff[q, j] = sum_i qvecs[q, i] * R[i, j] for 8 queries per CTA: R is read once per 8 queries
__global__ void __launch_bounds__(256) ffx_rotate_queries8_kernel(const float *qvecs, const float *R, int D,
                                                                  int64_t nq, float *qeff) {
    extern __shared__ float s_q8[];  // [8][D]
    const int64_t q0 = static_cast<int64_t>(blockIdx.x) * 8;
    const int nqb = nq - q0 < 8 ? static_cast<int>(nq - q0) : 8;
    for (int i = threadIdx.x; i < 8 * D; i += blockDim.x)
        s_q8[i] = (i / D) < nqb ? qvecs[q0 * D + i] : 0.f;
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < D; i++) {
            const float r = __ldg(R + static_cast<size_t>(i) * D + j);
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] = fmaf(s_q8[u * D + i], r, acc[u]);
        }
        for (int u = 0; u < nqb; u++) qeff[(q0 + u) * D + j] = acc[u];
    }
}

// The same rotation as a register-tiled SGEMM for many queries: a CTA owns a 64-query x 64-output
// tile, a thread a 4 x 4 block, operands staged through shared memory 16 values of i at a time.
// Every output still accumulates fmaf(q[i], R[i][j], acc) over i ASCENDING from 0, i.e. the bits
// of ffx_rotate_queries8_kernel; R is read nq / 64 times instead of nq / 8 (C4: 0.50 -> ~0.1 ms
// per step; fp32 FMA, no tensor cores: TF32 / BF16 products would not keep the 1e-5 contract).
// Needs D % 16 == 0.
constexpr int kRotTile = 64, kRotK = 16;
__global__ void __launch_bounds__(256) ffx_rotate_queries_tiled_kernel(const float *qvecs, const float *R, int D,
                                                                       int64_t nq, float *qeff) {
    __shared__ float s_a[kRotK][kRotTile + 4];  // [i][query]: padded against bank conflicts on the transposing store
    __shared__ float s_b[kRotK][kRotTile];      // [i][output]
    const int64_t q0 = static_cast<int64_t>(blockIdx.y) * kRotTile;
    const int j0 = blockIdx.x * kRotTile;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // outputs 4*tx.., queries 4*ty..
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.f;
    for (int i0 = 0; i0 < D; i0 += kRotK) {
        // A tile: 64 queries x 16 i (1024 floats, 4 per thread), stored transposed
        {
            const int e = threadIdx.x * 4, qq = e / kRotK, ii = e % kRotK;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q0 + qq < nq) v = __ldg(reinterpret_cast<const float4 *>(qvecs + (q0 + qq) * D + i0 + ii));
            s_a[ii][qq] = v.x, s_a[ii + 1][qq] = v.y, s_a[ii + 2][qq] = v.z, s_a[ii + 3][qq] = v.w;
        }
        // B tile: 16 i x 64 outputs
        {
            const int e = threadIdx.x * 4, ii = e / kRotTile, jj = e % kRotTile;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + jj < D) v = __ldg(reinterpret_cast<const float4 *>(R + static_cast<size_t>(i0 + ii) * D + j0 + jj));
            *reinterpret_cast<float4 *>(&s_b[ii][jj]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kRotK; kk++) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&s_a[kk][4 * ty]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&s_b[kk][4 * tx]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t q = q0 + 4 * ty + r;
        if (q < nq && j0 + 4 * tx < D)
            *reinterpret_cast<float4 *>(qeff + q * D + j0 + 4 * tx) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
}

struct AdcWarpArgs {
    AdcArgs base;
    const float *cw_t;   // [4][Ks][32][Ds]
    int k, cpad;         // FUSE
    float *topk_score;
    int32_t *topk_pos;
    const float *lut;    // XOR kernel: [nq][M * Ks] tables built by ffx_adc_xor_lut_kernel (nullptr: built in the kernel)
};

__host__ __device__ inline size_t adc_warp_smem_bytes(int Ks, int cpad_scores) {
    const size_t lut = static_cast<size_t>(4) * Ks * 128;
    const size_t keys = static_cast<size_t>(cpad_scores) * 8;  // sort keys overlay the dead tables
    // the score region is padded so that the tables (and the 64-bit keys built over them) stay aligned
    return ((static_cast<size_t>(cpad_scores) * 4 + 127) & ~static_cast<size_t>(127)) + (lut > keys ? lut : keys);
}

template <bool FUSE, bool INDIRECT>
__global__ void __launch_bounds__(kAdcWarpThreads, 1) ffx_adc_warp_kernel(const AdcWarpArgs w) {
    const AdcArgs &a = w.base;
    extern __shared__ __align__(128) unsigned char adc_smem[];
    float *s_scores = reinterpret_cast<float *>(adc_smem);  // [cpad] (FUSE)
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
    // the scratch scores of a separate top-k pass are indexed relative to the launch's first pair
    float *rank = a.rank_scores ? a.rank_scores - a.q_off[0] : nullptr;
    if (threadIdx.x == 0) s_next = 0;
    if (FUSE) {
        for (int i = threadIdx.x; i < n_query; i += blockDim.x) s_scores[i] = __int_as_float(0x7fc00000);
    }

    const int M = a.M, W = a.M >> 2;

    // ---- per-query tables: s_lut[(j*Ks + c)*32 + l] = qeff[m*Ds..] . codewords[m][c],  m = 4l + j
    {
        const float *qe = a.qeff + q_idx * (static_cast<int64_t>(M) * a.Ds);
        const int total = 4 * a.Ks * 32;
        const int per_table = a.Ks * 32;
        const bool vec4 = (a.Ds & 3) == 0 && (reinterpret_cast<uintptr_t>(qe) & 15) == 0;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int l = e & 31;
            float acc = 0.f;
            if (l < W) {
                const int m = 4 * l + e / per_table;
                const float *cw = w.cw_t + static_cast<size_t>(e) * a.Ds;
                const float *qm = qe + m * a.Ds;
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
            }
            s_lut[e] = acc;
        }
    }
    __syncthreads();

    const bool act = lane < W;
    // table j of this lane: + 128 B per code value (dp4a with 0x80 in byte j of the selector)
    const uint32_t tab0 = static_cast<uint32_t>(__cvta_generic_to_shared(s_lut)) + static_cast<uint32_t>(lane) * 4u;
    const uint32_t tab_bytes = static_cast<uint32_t>(a.Ks) * 128u;
    const uint32_t tab1 = tab0 + tab_bytes, tab2 = tab1 + tab_bytes, tab3 = tab2 + tab_bytes;
    // lanes >= W read word 0 of the row (a valid address); their table column is all zeros
    const uint8_t *codes_lane = a.codes + (act ? lane * 4 : 0);
    const uint32_t Mu = static_cast<uint32_t>(M);

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
            const uint32_t t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t pre = incl - cnt;
        const uint32_t total = __shfl_sync(kFull, incl, 31);

        float acc = 0.f;       // lane j: running max / sum / first of candidate j
        bool acc_set = false;

        // warp-uniform cursor over the stream: next row and the end of its document
        int cj = -1;
        uint32_t crow = 0, cend = 0;

        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const int n_here = static_cast<int>(min(32u, total - r0));
            uint32_t word[32];
            // ---- phase 1: issue the code loads of the block (one coalesced request per row)
            auto next_row = [&]() -> uint32_t {
                while (crow == cend) {  // warp-uniform
                    cj++;
                    crow = __shfl_sync(kFull, start, cj);
                    cend = crow + __shfl_sync(kFull, cnt, cj);
                }
                uint32_t row = crow++;
                if (INDIRECT) row = static_cast<uint32_t>(__ldg(a.doc_rows + row));
                return __ldg(reinterpret_cast<const uint32_t *>(codes_lane + static_cast<uint64_t>(row) * Mu));
            };
            if (n_here == 32) {
#pragma unroll
                for (int i = 0; i < 32; i++) word[i] = next_row();
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    word[i] = 0u;
                    if (i < n_here) word[i] = next_row();
                }
            }
            // ---- phase 2: 4 conflict-free look-ups per lane per row (address = one dp4a each)
            float part[32];
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const uint32_t cw = word[i];
                const float v0 = lds_f32(__dp4a(cw, 0x00000080u, tab0));
                const float v1 = lds_f32(__dp4a(cw, 0x00008000u, tab1));
                const float v2 = lds_f32(__dp4a(cw, 0x00800000u, tab2));
                const float v3 = lds_f32(__dp4a(cw, 0x80000000u, tab3));
                part[i] = (v0 + v1) + (v2 + v3);
            }
            // ---- phase 3: transposed butterfly: lane r ends with the sum over lanes of part[r]
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; i++) {
                    const float send = upper ? part[i] : part[i + off];
                    const float keep = upper ? part[i + off] : part[i];
                    part[i] = keep + __shfl_xor_sync(kFull, send, off);
                }
            }
            float s = part[0];  // score of block row `lane`
            // ---- phase 4: fold rows into documents (segmented inclusive scan in row order), then
            // lane j pulls candidate j's partial from the lane of its last row in this block
            if (a.mode == FFX_MODE_MAXP || a.mode == FFX_MODE_AVEP) {
                // owner of stream row r0 + lane: first candidate whose inclusive count exceeds it
                const uint32_t g = r0 + lane;
                int c = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t t = __shfl_sync(kFull, incl, c + step - 1);
                    if (t <= g) c += step;
                }
                c = min(c, 31);
                const uint32_t my_k = g - __shfl_sync(kFull, pre, c);  // position inside the document
                const uint32_t dist = min(my_k, static_cast<uint32_t>(lane));  // same-document rows to the left
                const bool is_max = a.mode == FFX_MODE_MAXP;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float t = __shfl_up_sync(kFull, s, d);
                    if (dist >= static_cast<uint32_t>(d)) s = is_max ? fmaxf(s, t) : s + t;
                }
            }
            const bool overlap = cnt > 0 && pre < r0 + 32 && pre + cnt > r0;
            const uint32_t last = min(pre + cnt - 1, r0 + 31) - r0;
            const float got = __shfl_sync(kFull, s, overlap ? last : 0);
            if (overlap) {
                if (!acc_set) acc = got;
                else acc = a.mode == FFX_MODE_MAXP ? fmaxf(acc, got) : acc + got;
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
                if (rank) rank[p] = inter;
                if (FUSE) s_scores[c0 + base + lane] = inter;
            } else if (rank) {
                rank[p] = __int_as_float(0x7fc00000);
            }
        }
    }

    if (FUSE) {
        // the tables are dead: build the 64-bit sort keys over them
        __syncthreads();
        unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(s_lut);
        rank_scores_topk(s_scores, n_query, s_keys, w.k, w.topk_score + q_idx * w.k, w.topk_pos + q_idx * w.k);
    }
}

}  // namespace ffx
