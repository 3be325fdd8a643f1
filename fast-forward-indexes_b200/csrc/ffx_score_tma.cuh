// ffx_score_tma.cuh — the hot kernel, TMA-staged.
//
// Same arithmetic and outputs as ffx_score_kernel (ffx_kernels.cuh), but passage rows reach
// the SM through the bulk-copy engine instead of registers: every warp owns a ring of NS row
// slots in shared memory, one elected lane issues `cp.async.bulk` (one instruction per 3 KB
// row, completion counted on an mbarrier) and runs NS rows ahead of the arithmetic.  Loads
// therefore stay in flight while the warp multiplies and reduces, bytes in flight per SM are
// set by shared memory (8 warps x 6 slots x 3 KB = 144 KB) instead of the register file, and
// the row stream is flattened across candidates — single-row documents (PASSAGE / FIRSTP) and
// odd tails keep the ring as full as long MAXP documents do.
//
// Replaces index/base.py:279-314 + ranking.py:319,115-117,285-291 of the reference, like
// ffx_score_kernel.  HBM-bound; no tensor cores.
#pragma once
#include "ffx_kernels.cuh"

namespace ffx {

constexpr int kTmaMaxThreads = 512;  // 8..16 warps per CTA (one CTA per SM), chosen at launch

struct __align__(16) CandDesc {
    uint32_t start, cnt;  // first row (or offset into doc_rows) and row count; cnt 0 = nothing to read
    float lex;
    uint32_t mine;        // 0: pair belongs to another shard (no outputs)
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// global -> shared bulk copy; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "r"(addr));
    return r;
}

// Dynamic shared memory of one CTA: [FUSE: cpad interpolated scores][ring][mbarriers][descriptors].
// The 64-bit sort keys of the fused top-k are built inside the ring once it has drained.
// Rows of more than 64 elements per lane (D >= 2560) do not fit the register file next to the
// query vector: the query is kept in shared memory too (one more row-sized region).
__host__ __device__ inline bool tma_streams_query(int row_bytes) { return row_bytes > 32 * 64 * 4; }

__host__ __device__ inline size_t tma_smem_bytes(int cpad_scores, int warps, int ns, int row_bytes) {
    size_t keys = (static_cast<size_t>(cpad_scores) * 4 + 127) & ~static_cast<size_t>(127);
    return keys + (tma_streams_query(row_bytes) ? static_cast<size_t>(row_bytes) : 0) +
           static_cast<size_t>(warps) * ns * row_bytes + static_cast<size_t>(warps) * ns * 8 +
           static_cast<size_t>(warps) * 2 * 32 * sizeof(CandDesc) + 128;
}

// One row against the query, both in shared memory in lane-major order: lane l owns one whole
// numpy leaf (CPL = 8 accumulators of S terms); same products, same chains, same combine order
// as lane_chain_sum, with 8 live accumulators instead of 2 x 8*S registers.
template <int CPL, int S>
__device__ __forceinline__ float lane_chain_sum_streamed(uint32_t q_addr, uint32_t row_addr) {
    constexpr int NV4 = CPL * S / 4;
    float acc[CPL];
#pragma unroll
    for (int i = 0; i < NV4; i++) {
        const float4 q = lds_f4(q_addr + i * 512);
        const float4 v = lds_f4(row_addr + i * 512);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int m = 4 * i + c, ch = m % CPL;
            const float prod = __fmul_rn(f4c(q, c), f4c(v, c));
            acc[ch] = (m / CPL == 0) ? prod : __fadd_rn(acc[ch], prod);
        }
    }
#pragma unroll
    for (int w = 1; w < CPL; w <<= 1) {
#pragma unroll
        for (int ch = 0; ch < CPL; ch += 2 * w) acc[ch] = __fadd_rn(acc[ch], acc[ch + w]);
    }
    return acc[0];
}

// A warp per row (D >= 384 with a uniform tree); LPR stays in the signature (always 32) so that the
// kernel's name in launch lists and tests is stable.  Shorter rows: ffx_score_packed.cuh.
template <int CPL, int S, bool FUSE, int LPR = 32>
__global__ void __launch_bounds__(kTmaMaxThreads, 1) ffx_score_tma_kernel(const ScoreArgs a, const int ns,
                                                                        const int batch) {
    constexpr int EPL = CPL * S;
    constexpr int NV4 = EPL / 4;
    constexpr uint32_t ROWB = static_cast<uint32_t>(LPR) * EPL * 4u;
    constexpr uint32_t SLOTB = ROWB;                     // a ring slot holds one row
    constexpr bool kPairRows = EPL <= 32;                // two rows of registers per lane only while they fit
    constexpr bool kStream = EPL > 64;                   // query vector in shared memory, rows streamed against it
    static_assert(EPL % 4 == 0, "lane slice must be whole float4s");
    static_assert(LPR == 32, "a warp per row");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_next;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    const int64_t q_idx = blockIdx.x / a.tiles_per_query;
    const int t_idx = blockIdx.x % a.tiles_per_query;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);
    const int c0 = t_idx * a.tile;
    const int n_tile = min(a.tile, n_query - c0);
    if (!FUSE && n_tile <= 0) return;

    // ---- carve shared memory: [scores][ring][mbarriers][descriptors]
    float *s_scores = reinterpret_cast<float *>(smem_raw);
    size_t off = FUSE ? ((static_cast<size_t>(a.cpad) * 4 + 127) & ~static_cast<size_t>(127)) : 0;
    float *s_q = reinterpret_cast<float *>(smem_raw + off);  // kStream: the query vector, lane-major
    if (kStream) off += ROWB;
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem_raw + off);  // = ring, after the drain
    const uint32_t ring = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * SLOTB;
    off += static_cast<size_t>(n_warps) * ns * SLOTB;
    const uint32_t bars = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * 8;
    off += static_cast<size_t>(n_warps) * ns * 8;
    off = (off + 15) & ~static_cast<size_t>(15);
    CandDesc *desc = reinterpret_cast<CandDesc *>(smem_raw + off) + warp * 64;  // [2][32]

    // the scratch scores of a separate top-k pass are indexed relative to the launch's first pair
    float *rank = a.rank_scores ? a.rank_scores - a.q_off[0] : nullptr;
    if (threadIdx.x == 0) s_next = 0;
    if (FUSE) {
        for (int i = threadIdx.x; i < n_query; i += blockDim.x) s_scores[i] = __int_as_float(0x7fc00000);
    }
    if (lane == 0) {
        for (int s = 0; s < ns; s++) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    // this lane's slice of the query vector: registers (lane-major order), or — long rows — a
    // lane-major copy of the whole vector in shared memory
    float q[kStream ? 1 : EPL];
    const float *qv = a.qvecs + q_idx * (LPR * EPL);
    if (kStream) {
        for (int k = threadIdx.x; k < 32 * EPL; k += blockDim.x) s_q[k] = __ldg(qv + ffx_orig_index(CPL, S, k));
    } else {
#pragma unroll
        for (int m = 0; m < (kStream ? 1 : EPL); m++) {
            const int g = (lane % LPR) * CPL + (m % CPL);
            q[m] = __ldg(qv + (g >> 3) * (8 * S) + 8 * (m / CPL) + (g & 7));
        }
    }
    const uint32_t q_addr = smem_u32(s_q) + lane * 16;
    __syncthreads();

    const char *rows_base = reinterpret_cast<const char *>(a.vectors);
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    const int64_t pair0 = q_begin + c0;

    // ---- candidate-batch pipeline: batch t+3 has its candidate ids in flight (g_*), batch
    // t+2 its spans (h_*), batches t+1 (slot B) and t (slot A) are published in shared memory.
    int g_base = 0, g_nb = 0, g_cand = 0;
    float g_lex = 0.f;
    int h_base = 0, h_nb = 0;
    uint32_t h_start = 0, h_cnt = 0, h_mine = 0;
    float h_lex = 0.f;

    auto grab = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_next, batch);
        base = __shfl_sync(kFull, base, 0);
        g_base = base;
        g_nb = max(0, min(batch, n_tile - base));
        g_cand = 0;
        g_lex = 0.f;
        if (lane < g_nb) {
            g_cand = __ldg(a.cand + pair0 + base + lane);
            if (a.lex) g_lex = __ldg(a.lex + pair0 + base + lane);
        }
    };
    auto resolve = [&]() {  // g_* -> h_*: candidate -> (first row, count)
        h_base = g_base;
        h_nb = g_nb;
        h_lex = g_lex;
        h_start = 0;
        h_cnt = 0;
        h_mine = 0;
        if (lane < g_nb) {
            uint32_t loc = 0;
            if (!candidate_ok(g_cand, a.limit, a.err, pair0 + g_base + lane)) {
                h_mine = 1;  // reported; scores as an empty document
            } else if (candidate_mine(g_cand, a.base, a.count, &loc)) {
                h_mine = 1;
                if (a.mode == FFX_MODE_PASSAGE) {
                    h_start = loc;
                    h_cnt = 1;
                } else {
                    const uint2 sp = __ldg(a.doc_span + loc);
                    h_start = sp.x;
                    h_cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
                }
            }
        }
    };
    auto publish = [&](int slot) {
        CandDesc d;
        d.start = h_start;
        d.cnt = h_cnt;
        d.lex = h_lex;
        d.mine = h_mine;
        desc[slot * 32 + lane] = d;
        __syncwarp();
        // candidates of the batch that have rows to read (a doc-id-range shard owns a fraction of
        // them: producer and consumer walk the set bits instead of all 32 entries)
        return __ballot_sync(kFull, lane < h_nb && h_cnt > 0);
    };

    int cur = 0;  // slot of batch A (being consumed); batch B lives in slot cur ^ 1
    int nbA, nbB, baseA, baseB;
    uint32_t liveA, liveB;
    grab();
    resolve();
    liveA = publish(0);
    nbA = h_nb;
    baseA = h_base;
    grab();
    resolve();
    liveB = publish(1);
    nbB = h_nb;
    baseB = h_base;
    grab();
    resolve();
    grab();

    // ---- producer cursor (warp-uniform): walks the same row sequence as the consumer, up to
    // `ns` rows ahead
    bool p_inB = false;
    uint32_t p_left = liveA;  // candidates of the producer's batch it has not entered yet
    uint32_t pk = 0, pcnt = 0, pstart = 0, p_rows = 0;
    int p_stage = 0, c_stage = 0, inflight = 0;
    uint32_t c_phase = 0;

    auto top_up = [&]() {
        while (inflight < ns) {
            // next row of the flattened (candidate, row) sequence, or stop
            bool have = false;
            for (;;) {
                if (pk < pcnt) {
                    have = true;
                    break;
                }
                if (p_left) {
                    const int pj = __ffs(p_left) - 1;
                    p_left &= p_left - 1;
                    const CandDesc d = desc[((p_inB ? cur ^ 1 : cur) << 5) + pj];
                    pstart = d.start;
                    pcnt = d.cnt;
                    pk = 0;
                    continue;
                }
                if (!p_inB && nbB > 0) {
                    p_inB = true;
                    p_left = liveB;
                    pk = 0;
                    pcnt = 0;
                    continue;
                }
                break;
            }
            if (!have) break;
            uint32_t row = pstart + pk;
            if (indirect) {
                if ((pk & 31u) == 0)
                    p_rows = (pk + lane < pcnt) ? static_cast<uint32_t>(__ldg(a.doc_rows + pstart + pk + lane)) : 0u;
                row = __shfl_sync(kFull, p_rows, pk & 31u);
            }
            pk++;
            if (lane == 0) {
                const uint32_t bar = bars + p_stage * 8;
                mbar_expect_tx(bar, ROWB);
                bulk_g2s(ring + p_stage * ROWB, rows_base + static_cast<size_t>(row) * ROWB, ROWB, bar);
            }
            p_stage = p_stage + 1 == ns ? 0 : p_stage + 1;
            inflight++;
        }
    };

    while (nbA > 0) {
        float my_ff = 0.f;
        for (uint32_t todo = liveA; todo; todo &= todo - 1) {
            const int cj = __ffs(todo) - 1;
            const uint32_t cnt = desc[(cur << 5) + cj].cnt;
            DocReduce red;
            red.init();
            for (uint32_t ck = 0; ck < cnt;) {
                // two rows of the document per step: their multiply/reduce chains interleave
                top_up();
                const bool two = kPairRows && ck + 1 < cnt;  // warp-uniform
                const int s0 = c_stage;
                const int s1 = s0 + 1 == ns ? 0 : s0 + 1;
                float sa, sb = 0.f;
                if constexpr (kStream) {
                    mbar_wait(bars + s0 * 8, (c_phase >> s0) & 1u);
                    sa = lane_chain_sum_streamed<CPL, S>(q_addr, ring + s0 * ROWB + lane * 16);
                } else {
                    float4 v0[NV4], v1[NV4];
                    mbar_wait(bars + s0 * 8, (c_phase >> s0) & 1u);
                    const uint32_t src0 = ring + s0 * ROWB + lane * 16;
#pragma unroll
                    for (int i = 0; i < NV4; i++) v0[i] = lds_f4(src0 + i * 512);
                    if (two) {
                        mbar_wait(bars + s1 * 8, (c_phase >> s1) & 1u);
                        const uint32_t src1 = ring + s1 * ROWB + lane * 16;
#pragma unroll
                        for (int i = 0; i < NV4; i++) v1[i] = lds_f4(src1 + i * 512);
                    }
                    sa = lane_chain_sum<CPL, S>(q, v0);
                    sb = two ? lane_chain_sum<CPL, S>(q, v1) : 0.f;
                }
                sa = warp_tree_sum(sa);
                if (two) sb = warp_tree_sum(sb);
                // every lane holds the rows' values in registers now: the slots may be refilled
                __syncwarp();
                c_phase ^= (1u << s0) | (two ? 1u << s1 : 0u);
                c_stage = two ? (s1 + 1 == ns ? 0 : s1 + 1) : s1;
                inflight -= two ? 2 : 1;
                red.add(sa, ck == 0, a.mode);
                if (two) red.add(sb, false, a.mode);
                ck += two ? 2u : 1u;
            }
            const float ff = red.finish(cnt, a.mode);
            if (lane == cj) my_ff = ff;
        }

        // lane j holds candidate j's score: coalesced epilogue for batch A
        if (lane < nbA) {
            const CandDesc d = desc[(cur << 5) + lane];
            const int64_t my_pair = pair0 + baseA + lane;
            if (d.mine) {
                float inter = my_ff;
                if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, d.lex), __fmul_rn(a.beta, my_ff));
                if (a.out_ff) a.out_ff[my_pair] = my_ff;
                if (a.out_int) a.out_int[my_pair] = inter;
                if (rank) rank[my_pair] = inter;
                if (FUSE) s_scores[c0 + baseA + lane] = inter;
            } else if (rank) {
                rank[my_pair] = __int_as_float(0x7fc00000);
            }
        }
        __syncwarp();

        // shift: B becomes A, the resolved batch is published into the freed slot, the
        // look-ahead loads advance by one batch
        const uint32_t live_new = publish(cur);
        cur ^= 1;
        nbA = nbB;
        baseA = baseB;
        liveA = liveB;
        nbB = h_nb;
        baseB = h_base;
        liveB = live_new;
        if (p_inB) {
            p_inB = false;  // the producer's position in old B is a position in new A
        } else {
            p_left = liveA;  // it had exhausted old A without entering B
            pk = 0;
            pcnt = 0;
        }
        resolve();
        grab();
    }

    if (FUSE) {
        // every row this CTA requested has been consumed, the ring is idle: build the sort keys
        // in it (NaN = pair of another shard or NaN score = not ranked)
        __syncthreads();
        float *out_s;
        int32_t *out_p;
        topk_destination(a, q_idx, &out_s, &out_p);
        rank_scores_topk<16>(s_scores, n_query, s_keys, a.k, out_s, out_p,
                             static_cast<size_t>(n_warps) * ns * SLOTB);  // the drained ring
        if (a.sc_world) __threadfence_system();  // peer stores: visible to the owner once the kernel ends
    }
}

}  // namespace ffx
