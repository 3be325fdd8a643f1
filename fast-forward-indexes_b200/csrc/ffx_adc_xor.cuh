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
// Code rows reach the SM through the bulk-copy engine: a thread reading its own row straight
// from global memory touches 32 different lines per warp request (~24 L1 wavefronts each, the
// first version spent a third of the L1 data pipe on that).  Instead every warp owns a
// 32-row x M-byte slot buffer in shared memory; for each 32-row block of the stream the lanes
// that own a document issue ONE cp.async.bulk per document (its rows are consecutive in the
// store and in the stream), completion is counted on the warp's mbarrier, and lane t then reads
// slot t with 32-bit loads in exactly the swizzled word order above — which is also bank-
// conflict free for the slot stride (M = 96: bank = 8*((j3 - t) mod 4) + (w ^ (t >> 2))).
// The next block's copies are issued as soon as the words are in registers, so they overlap the
// look-ups.
//
// Shapes: M % 32 == 0, M <= 128 (NC = M / 32 chunks).  Shared memory: M * Ks * 4 bytes of table
// (96 KB at M = 96, Ks = 256) + 32 * M bytes per warp of slots.  1024 threads per CTA, one CTA
// per SM.  FUSE adds 4 B per candidate of scores and builds the sort keys over the dead table
// and slots.
// Bound: shared-memory look-ups (3 conflict-free LDS per row at M = 96) and instruction issue.
#pragma once
#include "ffx_adc_warp.cuh"
#include "ffx_score_tma.cuh"

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

// The per-query tables, all queries of a launch at once: lut[q][(j3*Ks + c)*32 + b] =
// qeff[q][m*Ds ..] . codewords[m][c], m = 32*j3 + b — the layout ffx_adc_xor_kernel keeps in shared
// memory, so that a CTA fetches its query's table with ONE bulk copy instead of building it
// (building in place cost every query ~20 us of latency-bound codebook reads with the
// look-up pipe idle: 12 % of the kernel, ncu profiles/r2_adc_xor_full_raw.csv).  A thread owns
// one table entry, keeps its codeword in registers and walks QT queries: the codebook is read
// once per QT queries.  Same arithmetic as the in-kernel build (fmaf over d, ascending).
constexpr int kAdcLutThreads = 256;
constexpr int kAdcLutQueries = 32;

template <int DS>
__global__ void __launch_bounds__(kAdcLutThreads) ffx_adc_xor_lut_kernel(const float *cw_x, const float *qeff, int M, int Ks,
                                                                        int Ds, int64_t nq, float *lut) {
    const int total = M * Ks;
    const int e = blockIdx.x * kAdcLutThreads + threadIdx.x;
    if (e >= total) return;
    const int per_table = Ks * 32;
    const int m = 32 * (e / per_table) + (e & 31);
    const int ds = DS ? DS : Ds;
    float c[DS ? DS : 1];
    const float *cw = cw_x + static_cast<size_t>(e) * ds;
    if constexpr (DS != 0) {
#pragma unroll
        for (int d = 0; d < DS; d += 4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4 *>(cw + d));
            c[d] = c4.x, c[d + 1] = c4.y, c[d + 2] = c4.z, c[d + 3] = c4.w;
        }
    }
    const int64_t q0 = static_cast<int64_t>(blockIdx.y) * kAdcLutQueries;
    const int64_t q1 = q0 + kAdcLutQueries < nq ? q0 + kAdcLutQueries : nq;
    const int64_t D = static_cast<int64_t>(M) * ds;
    // 32 consecutive threads = 32 consecutive sub-quantizers m: their query slices are one
    // contiguous 32 * Ds floats, read as float4s (DS is a multiple of 4: rows of D floats stay 16-byte
    // aligned when D is)
    const bool vec = DS != 0 && (D & 3) == 0 && (reinterpret_cast<uintptr_t>(qeff) & 15) == 0;
#pragma unroll 4
    for (int64_t q = q0; q < q1; q++) {
        const float *qm = qeff + q * D + m * ds;
        float acc = 0.f;
        if constexpr (DS != 0) {
            if (vec) {
#pragma unroll
                for (int d = 0; d < DS; d += 4) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4 *>(qm + d));
                    acc = fmaf(q4.x, c[d], acc);
                    acc = fmaf(q4.y, c[d + 1], acc);
                    acc = fmaf(q4.z, c[d + 2], acc);
                    acc = fmaf(q4.w, c[d + 3], acc);
                }
            } else {
#pragma unroll
                for (int d = 0; d < DS; d++) acc = fmaf(__ldg(qm + d), c[d], acc);
            }
        } else {
            for (int d = 0; d < ds; d++) acc = fmaf(__ldg(qm + d), __ldg(cw + d), acc);
        }
        __stcs(lut + q * total + e, acc);  // read once, by another SM: streaming store
    }
}

// The same tables with the query slices BROADCAST instead of gathered: a CTA owns a 32 x 32 tile of one
// table (32 codewords c, the 32 sub-quantizers b of table j3 — 4 KB of consecutive entries), lane =
// codeword, warp = sub-quantizer (4 of them per warp), so that the 32 lanes of a warp read the SAME
// 32-byte query slice (one L1 wavefront instead of eight: the thread-per-entry kernel above keeps the
// L1 data pipe 95 % busy, profiles/r2_adc_lut_raw.csv).  The tile goes through shared memory
// (stride 33: conflict-free both ways) and leaves as four 128-byte rows per warp.  Same arithmetic
// per entry: fmaf over d, ascending, from 0.  Needs Ks % 32 == 0 and DS in {4, 8, 16}.
constexpr int kAdcLutTileThreads = 256;

template <int DS>
__global__ void __launch_bounds__(kAdcLutTileThreads) ffx_adc_xor_lut_tile_kernel(const float *cw_x, const float *qeff, int M,
                                                                                int Ks, int64_t nq, float *lut) {
    __shared__ float tile[2][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_per_table = Ks / 32;
    const int j3 = blockIdx.x / tiles_per_table;
    const int c0 = (blockIdx.x % tiles_per_table) * 32;
    const int64_t e0 = (static_cast<int64_t>(j3) * Ks + c0) * 32;  // first entry of the tile
    const int64_t total = static_cast<int64_t>(M) * Ks;
    const int64_t D = static_cast<int64_t>(M) * DS;
    float c[4][DS];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int b = warp + 8 * p;
        const float *cw = cw_x + (e0 + static_cast<int64_t>(lane) * 32 + b) * DS;
#pragma unroll
        for (int d = 0; d < DS; d += 4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4 *>(cw + d));
            c[p][d] = c4.x, c[p][d + 1] = c4.y, c[p][d + 2] = c4.z, c[p][d + 3] = c4.w;
        }
    }
    const int64_t q0 = static_cast<int64_t>(blockIdx.y) * kAdcLutQueries;
    const int64_t q1 = q0 + kAdcLutQueries < nq ? q0 + kAdcLutQueries : nq;
    for (int64_t q = q0; q < q1; q++) {
        float *t = tile[(q - q0) & 1];
        const float *qrow = qeff + q * D + static_cast<int64_t>(32 * j3) * DS;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int b = warp + 8 * p;
            const float *qm = qrow + b * DS;  // warp-uniform: a broadcast load
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < DS; d += 4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(qm + d));
                acc = fmaf(q4.x, c[p][d], acc);
                acc = fmaf(q4.y, c[p][d + 1], acc);
                acc = fmaf(q4.z, c[p][d + 2], acc);
                acc = fmaf(q4.w, c[p][d + 3], acc);
            }
            t[lane * 33 + b] = acc;  // entry (c = c0 + lane, b)
        }
        __syncthreads();  // (the other tile buffer is free again: its readers passed the previous barrier)
        float *out = lut + q * total + e0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int row = warp + 8 * p;  // codeword c0 + row: 32 consecutive entries
            __stcs(out + row * 32 + lane, t[row * 33 + lane]);
        }
    }
}

// [table][slots: warps x 32 x M][mbarriers][FUSE: cpad interpolated scores]; the sort keys of the
// fused top-k overlay table + slots once they are dead
__host__ __device__ inline size_t adc_xor_smem_bytes(int M, int Ks, int cpad_scores) {
    const size_t lut = static_cast<size_t>(M) * Ks * 4;
    const size_t rest = static_cast<size_t>(kAdcXorThreads / 32) * (32 * static_cast<size_t>(M) + 8);
    return lut + rest + static_cast<size_t>(cpad_scores) * 4 + 128;
}

// acc += (a, b) as one packed fp32x2 add (sm_100)
__device__ __forceinline__ void add_f32x2(float2 &acc, float a, float b) {
    asm("{\n\t"
        ".reg .b64 va, vb;\n\t"
        "mov.b64 va, {%0, %1};\n\t"
        "mov.b64 vb, {%2, %3};\n\t"
        "add.rn.f32x2 va, va, vb;\n\t"
        "mov.b64 {%0, %1}, va;\n\t"
        "}"
        : "+f"(acc.x), "+f"(acc.y)
        : "f"(a), "f"(b));
}

// ld.shared.f32 [addr + OFF] with the offset as an immediate of the instruction
template <uint32_t OFF>
__device__ __forceinline__ float lds_f32_imm(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(addr), "n"(OFF));
    return r;
}

// The 32 look-ups of word position `wi` in all NC chunks when chunk j uses table j and the tables
// are TB bytes apart (TB known at compile time: Ks = 256): the swizzled line offset
// (table 0 | 4t) ^ 4p is formed once per static position p and shared by the NC tables.
template <int NC, uint32_t TB, int J = 0>
__device__ __forceinline__ void adc_xor_word(const uint32_t (&wd)[NC][8], int wi, const uint32_t (&sel)[4], uint32_t x0,
                                             uint32_t x1, uint32_t x2, uint32_t x3, float2 &acc01, float2 &acc23) {
    if constexpr (J < NC) {
        const float v0 = lds_f32_imm<J * TB>(__dp4a(wd[J][wi], sel[0], x0));
        const float v1 = lds_f32_imm<J * TB>(__dp4a(wd[J][wi], sel[1], x1));
        const float v2 = lds_f32_imm<J * TB>(__dp4a(wd[J][wi], sel[2], x2));
        const float v3 = lds_f32_imm<J * TB>(__dp4a(wd[J][wi], sel[3], x3));
        add_f32x2(acc01, v0, v1);
        add_f32x2(acc23, v2, v3);
        adc_xor_word<NC, TB, J + 1>(wd, wi, sel, x0, x1, x2, x3, acc01, acc23);
    }
}

// L2 residency (ncu of the first version: 19.5 GB of DRAM reads for 16.2 GB of codes — the
// 786 KB codebook every query's table is built from was being evicted by the streaming code rows
// and re-fetched from HBM): code rows pass through L2 as evict-first, the codebook is evict-last.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4 *p, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}

// AdcWarpArgs::cw_t holds cw_x here.
template <int NC, bool FUSE>
__global__ void __launch_bounds__(kAdcXorThreads, 1) ffx_adc_xor_kernel(const AdcWarpArgs w) {
    const AdcArgs &a = w.base;
    constexpr int M = NC * 32;
    extern __shared__ __align__(128) unsigned char adc_smem[];
    float *s_lut = reinterpret_cast<float *>(adc_smem);  // 128-byte aligned (the XOR addressing relies on it)
    __shared__ int s_next, s_next_small;
    const uint64_t keep_l2 = l2_policy_evict_last(), stream_l2 = l2_policy_evict_first();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lut_bytes = static_cast<uint32_t>(M) * static_cast<uint32_t>(w.base.Ks) * 4u;
    const uint32_t slots = smem_u32(adc_smem) + lut_bytes + static_cast<uint32_t>(warp) * (32u * M);
    const uint32_t bars_off = lut_bytes + (kAdcXorThreads / 32) * (32u * M);
    const uint32_t bar = smem_u32(adc_smem) + bars_off + static_cast<uint32_t>(warp) * 8u;
    float *s_scores = reinterpret_cast<float *>(adc_smem + bars_off + (kAdcXorThreads / 32) * 8);  // [cpad] (FUSE)
    const int64_t q_idx = blockIdx.x / a.tiles_per_query;
    const int t_idx = blockIdx.x % a.tiles_per_query;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);
    const int c0 = t_idx * a.tile;
    const int n_tile = min(a.tile, n_query - c0);
    if (!FUSE && n_tile <= 0) return;
    // the scratch scores of a separate top-k pass are indexed relative to the launch's first pair
    float *rank = a.rank_scores ? a.rank_scores - a.q_off[0] : nullptr;
    if (threadIdx.x == 0) {
        s_next = 0;
        s_next_small = 0;
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (FUSE) {
        for (int i = threadIdx.x; i < n_query; i += blockDim.x) s_scores[i] = __int_as_float(0x7fc00000);
    }

    // ---- per-query table: s_lut[(j3*Ks + c)*32 + b] = qeff[m*Ds..] . codewords[m][c],  m = 32*j3 + b
    if (w.lut) {
        // built for the whole launch by ffx_adc_xor_lut_kernel: one bulk copy
        __shared__ __align__(8) unsigned long long s_lut_bar;
        const uint32_t lbar = smem_u32(&s_lut_bar);
        if (threadIdx.x == 0) {
            mbar_init(lbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(lbar, lut_bytes);
            bulk_g2s_hint(smem_u32(s_lut), w.lut + q_idx * (static_cast<int64_t>(M) * a.Ks), lut_bytes, lbar, stream_l2);
        }
        __syncthreads();
        mbar_wait(lbar, 0);
    } else {
        const float *qe = a.qeff + q_idx * (static_cast<int64_t>(M) * a.Ds);
        const int total = M * a.Ks;
        const int per_table = a.Ks * 32;
        const bool vec4 = (a.Ds & 3) == 0 && (reinterpret_cast<uintptr_t>(qe) & 15) == 0;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            // e / per_table without an integer division: at most NC = 4 tables
            const int j3 = (e >= per_table) + (e >= 2 * per_table) + (e >= 3 * per_table);
            const int m = 32 * j3 + (e & 31);
            const float *cw = w.cw_t + static_cast<size_t>(e) * a.Ds;
            const float *qm = qe + m * a.Ds;
            float acc = 0.f;
            if (vec4) {
                for (int d = 0; d < a.Ds; d += 4) {
                    const float4 c4 = ldg_f4_hint(reinterpret_cast<const float4 *>(cw + d), keep_l2);
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
    uint32_t tab[1];  // table 0 of this lane, low 7 bits = 4t (the table is 128-byte aligned)
    tab[0] = smem_u32(s_lut) | (t << 2);
    // slot t of the warp's buffer; register position wi of chunk step j reads word
    // 8*chunk(j) + (wi ^ (t >> 2)), chunk(j) = (j + rot) % NC with a per-lane rotation that keeps
    // the 32-bit reads conflict free for every slot stride (NC = 2: rot = bit 1 of t, NC = 4: t & 3)
    const uint32_t slot_x = (slots + t * M) | ((t >> 2) << 2);
    const uint32_t rot = NC == 2 ? ((t >> 1) & 1u) : (NC == 4 ? (t & 3u) : 0u);
    uint32_t tab_l[NC], chunk_off[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const uint32_t cj = (static_cast<uint32_t>(j) + rot) % NC;
        chunk_off[j] = 32u * cj;
        tab_l[j] = tab[0] + cj * tab_bytes;
    }
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    uint32_t phase = 0;  // parity of the warp's mbarrier
    const bool segmented = a.mode == FFX_MODE_MAXP || a.mode == FFX_MODE_AVEP;
    const bool is_max = a.mode == FFX_MODE_MAXP;

    // Guided self-scheduling: batches of 32 candidates up to `t_small`, batches of 8 for the last
    // half round, so that the warps of the CTA finish within a quarter batch of one another (with
    // 5000 candidates and 32 warps a warp takes ~5 batches: a whole-batch tail idled ~10 % of the SM).
    const int t_small = max(0, (n_tile - 16 * static_cast<int>(blockDim.x >> 5)) & ~31);
    bool small = false;
    for (;;) {
        int base = 0;
        if (!small) {
            if (lane == 0) base = atomicAdd(&s_next, 32);
            base = __shfl_sync(kFull, base, 0);
            small = base >= t_small;
        }
        if (small) {
            if (lane == 0) base = t_small + atomicAdd(&s_next_small, 8);
            base = __shfl_sync(kFull, base, 0);
        }
        if (base >= n_tile) break;
        const int nb = min(small ? 8 : 32, n_tile - base);
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

        // copies of stream rows [r0, r0+32) into the slots: one bulk copy per document piece
        auto issue_block = [&](uint32_t r0) {
            if (lane == 0) mbar_expect_tx(bar, min(32u, total - r0) * M);
            const uint32_t lo = max(pre, r0), hi = min(pre + cnt, r0 + 32);
            if (cnt > 0 && lo < hi) {
                const uint32_t dst = slots + (lo - r0) * M;
                const uint32_t first = start + (lo - pre);
                if (!indirect) {
                    bulk_g2s_hint(dst, a.codes + static_cast<uint64_t>(first) * M, (hi - lo) * M, bar, stream_l2);
                } else {
                    for (uint32_t i = 0; i < hi - lo; i++)
                        bulk_g2s_hint(dst + i * M, a.codes + static_cast<uint64_t>(__ldg(a.doc_rows + first + i)) * M, M, bar,
                                      stream_l2);
                }
            }
        };
        if (total > 0) issue_block(0);

        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t g = r0 + t;  // this lane's row of the stream
            // rows of the same document to the left of this lane inside the block: distance to the
            // nearest document start at or below lane t (bit i of `heads` = a document starts at
            // block row i), or t when the document began in an earlier block
            const bool starts_here = cnt > 0 && pre >= r0 && pre < r0 + 32;
            const uint32_t heads = __reduce_or_sync(kFull, starts_here ? 1u << (pre - r0) : 0u);
            const uint32_t below = heads & (0xffffffffu >> (31u - t));
            const uint32_t dist = below ? t - (31u - static_cast<uint32_t>(__clz(below))) : t;

            mbar_wait(bar, phase);
            phase ^= 1u;
            uint32_t wd[NC][8];
            if (g < total) {
#pragma unroll
                for (int j = 0; j < NC; j++) {
#pragma unroll
                    for (int wi = 0; wi < 8; wi++) {
                        const uint32_t addr = (slot_x ^ (static_cast<uint32_t>(wi) << 2)) +
                                              ((NC == 2 || NC == 4) ? chunk_off[j] : 32u * j);
                        wd[j][wi] = lds_u32(addr);
                    }
                }
            }
            // The slots are refilled by the bulk-copy engine (async proxy) while the words above were
            // read through the generic proxy and have not been consumed yet: without a cross-proxy
            // fence a copy can land before a queued ld.shared has executed (seen as ~1 wrong row in
            // 10^6 when the LSU queue is deep).  Fence, then let the warp agree that all reads are done.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (r0 + 32 < total) issue_block(r0 + 32);

            float s = 0.f;
            if (g < total) {
                float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
                if constexpr (NC == 2 || NC == 4) {
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const uint32_t tj = tab_l[j];
#pragma unroll
                        for (int wi = 0; wi < 8; wi++) {
                            const uint32_t pos = static_cast<uint32_t>(4 * wi);
                            const float v0 = lds_f32(__dp4a(wd[j][wi], sel[0], tj ^ ((pos + 0u) << 2)));
                            const float v1 = lds_f32(__dp4a(wd[j][wi], sel[1], tj ^ ((pos + 1u) << 2)));
                            const float v2 = lds_f32(__dp4a(wd[j][wi], sel[2], tj ^ ((pos + 2u) << 2)));
                            const float v3 = lds_f32(__dp4a(wd[j][wi], sel[3], tj ^ ((pos + 3u) << 2)));
                            add_f32x2(acc01, v0, v1);
                            add_f32x2(acc23, v2, v3);
                        }
                    }
                } else if (a.Ks == 256) {
                    // chunk j uses table j, 32 KB apart: one XOR per static position instead of one per look-up
#pragma unroll
                    for (int wi = 0; wi < 8; wi++) {
                        const uint32_t pos = static_cast<uint32_t>(4 * wi);
                        adc_xor_word<NC, 256u * 128u>(wd, wi, sel, tab[0] ^ ((pos + 0u) << 2), tab[0] ^ ((pos + 1u) << 2),
                                                      tab[0] ^ ((pos + 2u) << 2), tab[0] ^ ((pos + 3u) << 2), acc01, acc23);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const uint32_t tj = tab[0] + static_cast<uint32_t>(j) * tab_bytes;
#pragma unroll
                        for (int wi = 0; wi < 8; wi++) {
                            const uint32_t pos = static_cast<uint32_t>(4 * wi);
                            const float v0 = lds_f32(__dp4a(wd[j][wi], sel[0], tj ^ ((pos + 0u) << 2)));
                            const float v1 = lds_f32(__dp4a(wd[j][wi], sel[1], tj ^ ((pos + 1u) << 2)));
                            const float v2 = lds_f32(__dp4a(wd[j][wi], sel[2], tj ^ ((pos + 2u) << 2)));
                            const float v3 = lds_f32(__dp4a(wd[j][wi], sel[3], tj ^ ((pos + 3u) << 2)));
                            add_f32x2(acc01, v0, v1);
                            add_f32x2(acc23, v2, v3);
                        }
                    }
                }
                s = (acc01.x + acc01.y) + (acc23.x + acc23.y);
            }

            // fold rows into documents (segmented inclusive scan in row order), then lane j pulls
            // candidate j's partial from the lane of its last row in this block
            if (segmented) {
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
                if (rank) rank[p] = inter;
                if (FUSE) s_scores[c0 + base + lane] = inter;
            } else if (rank) {
                rank[p] = __int_as_float(0x7fc00000);
            }
        }
    }

    if (FUSE) {
        // table and slots are dead (every issued copy has been waited for): build the 64-bit sort
        // keys over them (NaN score = pair of another shard = not ranked)
        __syncthreads();
        unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(adc_smem);
        rank_scores_topk<8>(s_scores, n_query, s_keys, w.k, w.topk_score + q_idx * w.k, w.topk_pos + q_idx * w.k,
                            bars_off);  // table + slots
    }
}

}  // namespace ffx
