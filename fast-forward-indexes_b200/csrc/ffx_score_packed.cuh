// ffx_score_packed.cuh — the scoring kernel for SHORT rows (D <= 256: 2 .. 8 lanes per row) and for
// every dimension without a uniform numpy tree.
//
// Same contract and arithmetic as ffx_score_tma_kernel / ffx_score_any_kernel (one warp-shared
// ring of bulk-copied row slots per warp, candidate batches resolved ahead, fused interpolation
// and per-query top-k), but a warp step is issue-bound at these row sizes (a 512-byte row every
// ~20 clocks per SM), so the per-step instruction count is what is minimised here:
//   * a ring slot holds the next RPS = 32 / LPR rows of the batch's FLATTENED (candidate, row)
//     sequence, whatever documents they belong to — every step is full, single-row modes
//     (PASSAGE, FIRSTP) included;
//   * there is no candidate walk: lane j keeps candidate j of the batch in registers (first row,
//     row count, position `first` of its rows in the flattened sequence = a warp prefix sum).
//     The producer side of a step is lane-parallel (every lane whose candidate has rows inside
//     the step's window issues one bulk copy for them), the reduce side too (after the lane
//     groups' dot products are broadcast, lane j folds the ones inside its window into ITS
//     DocReduce, in row order: the Kahan mean of AVEP sees the same sequence as before);
//   * no descriptor array in shared memory.
// The dot product is a policy: LaneMajorDot (D = 64 .. 256 with a uniform numpy tree, rows
// permuted lane-major, 128-bit loads) or TreeDot (any other D <= 256: the tree as data,
// ffx_any_plan).  Replaces index/base.py:279-314 + ranking.py:319,115-117,285-291, like the
// kernels it specialises.  HBM-bound by design, instruction-issue-bound in practice.
#pragma once
#include "ffx_score_any.cuh"

namespace ffx {

// dynamic shared memory: [FUSE: cpad scores][query vector (TreeDot)][ring][mbarriers]
__host__ __device__ inline size_t packed_smem_bytes(int cpad_scores, int warps, int ns, int query_bytes, int slot_bytes) {
    const size_t keys = (static_cast<size_t>(cpad_scores) * 4 + 127) & ~static_cast<size_t>(127);
    const size_t qv = (static_cast<size_t>(query_bytes) + 127) & ~static_cast<size_t>(127);
    return keys + qv + static_cast<size_t>(warps) * ns * slot_bytes + static_cast<size_t>(warps) * ns * 8 + 128;
}

// Lane-major short rows (ffx_layout.h): the row is STORED for L0 = 8 or 16 lanes, stored lane c owning
// accumulator chain c (S terms) in float4 pieces (i * L0 + c).  Here LPR = L0 / CPL lanes share a
// row, each taking CPL adjacent chains — 32 / LPR = 4 .. 16 rows per warp step, CPL independent add
// chains per lane.  Register k of a lane holds chain (k ^ x), x from the lane's place in its
// quarter warp: the lanes of one 128-bit load phase then touch all eight 16-byte bank groups
// (no shared-memory conflicts), and the in-lane combine (k, k + w) still pairs the chains numpy
// pairs — fp32 addition commutes, the tree is unchanged.
template <int S, int L0, int CPL>
struct LaneMajorDot {
    static constexpr int LPR = L0 / CPL;
    static constexpr int NV4 = S / 4;
    static constexpr bool kPair = CPL == 1;  // one chain per lane: two ring slots per consumer iteration instead
    static_assert(S % 4 == 0 && (L0 == 8 || L0 == 16 || L0 == 32) && (CPL == 1 || CPL == 2 || CPL == 4), "lane-major rows");
    struct Plan {};
    float q[CPL][S];
    uint32_t piece[CPL];  // byte offset of chain k's first float4 inside a row
    __device__ static uint32_t row_bytes(const Plan &) { return L0 * S * 4u; }
    __device__ static uint32_t query_bytes(const Plan &) { return 0; }
    __device__ void stage_query(const Plan &, const ScoreArgs &, int64_t, float *) const {}
    __device__ void init(const Plan &, const ScoreArgs &a, int64_t q_idx, int lane, uint32_t) {
        const float *qv = a.qvecs + q_idx * (L0 * S);
        const int x = ((lane & 7) * CPL) >> 3;
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            const int c = (lane % LPR) * CPL + (k ^ x);  // stored lane = chain
            piece[k] = static_cast<uint32_t>(c) * 16u;
#pragma unroll
            for (int m = 0; m < S; m++) q[k][m] = __ldg(qv + (c >> 3) * (8 * S) + 8 * m + (c & 7));
        }
    }
    __device__ void load(uint32_t row, float4 (&v)[CPL][NV4]) const {
#pragma unroll
        for (int i = 0; i < NV4; i++) {
#pragma unroll
            for (int k = 0; k < CPL; k++) v[k][i] = lds_f4(row + piece[k] + i * (L0 * 16));
        }
    }
    // `row` = shared address of this lane group's row; every lane of the group returns the row's dot product
    __device__ float operator()(uint32_t row, int) const {
        float4 v[CPL][NV4];
        load(row, v);
        float acc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; k++) acc[k] = __fmul_rn(q[k][0], f4c(v[k][0], 0));
#pragma unroll
        for (int m = 1; m < S; m++) {
#pragma unroll
            for (int k = 0; k < CPL; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(q[k][m], f4c(v[k][m >> 2], m & 3)));
        }
#pragma unroll
        for (int w = 1; w < CPL; w <<= 1) {
#pragma unroll
            for (int k = 0; k < CPL; k += 2 * w) acc[k] = __fadd_rn(acc[k], acc[k + w]);
        }
        float part = acc[0];
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) part = __fadd_rn(part, __shfl_xor_sync(kFull, part, o));
        return __fadd_rn(0.f, part);
    }
    // CPL == 1: two rows at once, their (order-bound) add chains and butterflies interleave
    __device__ void pair(uint32_t row0, uint32_t row1, int, float &out0, float &out1) const {
        float4 v0[CPL][NV4], v1[CPL][NV4];
        load(row0, v0);
        load(row1, v1);
        float a0 = __fmul_rn(q[0][0], f4c(v0[0][0], 0)), a1 = __fmul_rn(q[0][0], f4c(v1[0][0], 0));
#pragma unroll
        for (int m = 1; m < S; m++) {
            a0 = __fadd_rn(a0, __fmul_rn(q[0][m], f4c(v0[0][m >> 2], m & 3)));
            a1 = __fadd_rn(a1, __fmul_rn(q[0][m], f4c(v1[0][m >> 2], m & 3)));
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {
            const float b0 = __shfl_xor_sync(kFull, a0, o), b1 = __shfl_xor_sync(kFull, a1, o);
            a0 = __fadd_rn(a0, b0);
            a1 = __fadd_rn(a1, b1);
        }
        out0 = __fadd_rn(0.f, a0);
        out1 = __fadd_rn(0.f, a1);
    }
};

// any other D: the numpy tree as data (ffx_any_plan), rows in original order (stride padded to 16
// bytes), CPL chains per lane
template <int CPL, int LPR_>
struct TreeDot {
    static constexpr int LPR = LPR_;
    static constexpr bool kPair = false;
    using Plan = ffx_any_plan;
    AnyQuery<CPL> q;
    uint32_t q_addr, my_byte, tail_byte;
    int my_steps, max_steps, tail_len;
    bool tail_mine;
    __device__ static uint32_t row_bytes(const Plan &p) { return static_cast<uint32_t>(p.stride) * 4u; }
    __device__ static uint32_t query_bytes(const Plan &p) { return static_cast<uint32_t>(p.stride) * 4u; }
    __device__ void stage_query(const Plan &p, const ScoreArgs &a, int64_t q_idx, float *s_q) const {
        const float *qsrc = a.qvecs + q_idx * p.dim;
        for (int k = threadIdx.x; k < p.stride; k += blockDim.x) s_q[k] = k < p.dim ? __ldg(qsrc + k) : 0.f;
    }
    // after the staged query vector is visible (__syncthreads)
    __device__ void init(const Plan &p, const ScoreArgs &, int64_t, int lane, uint32_t s_q_addr) {
        const int sub = lane % LPR;
        const int slot = (sub * CPL) >> 3;
        q_addr = s_q_addr;
        my_byte = static_cast<uint32_t>(p.start[slot] + ((sub * CPL) & 7)) * 4u;
        my_steps = p.steps[slot];
        max_steps = p.max_steps;
        tail_mine = slot == p.tail_slot;
        tail_byte = static_cast<uint32_t>(p.tail_start) * 4u;
        tail_len = p.tail_len;
        any_load_query<CPL>(q, q_addr, my_byte, my_steps);
    }
    __device__ float operator()(uint32_t row, int) const {
        return any_row_dot<CPL, LPR>(row, q_addr, q, my_byte, my_steps, max_steps, tail_mine, tail_byte, tail_len);
    }
    __device__ void pair(uint32_t row0, uint32_t row1, int lane, float &out0, float &out1) const {
        out0 = (*this)(row0, lane);
        out1 = (*this)(row1, lane);
    }
};

// candidate j of a batch, held by lane j
struct PackedCand {
    uint32_t start, cnt;  // first row (or offset into doc_rows) and row count; cnt 0 = nothing to read
    uint32_t first;       // position of its first row in the batch's flattened row sequence
    float lex;
    uint32_t mine;        // 0: pair belongs to another shard (no outputs)
};

// The host keeps batch * (longest document) below 2^31: positions in a batch fit 32 bits.
template <class Dot, bool FUSE>
__global__ void __launch_bounds__(kTmaMaxThreads, 1) ffx_score_packed_kernel(const ScoreArgs a, const typename Dot::Plan plan,
                                                                           const int ns, const int batch) {
    constexpr int LPR = Dot::LPR;
    constexpr int RPS = 32 / LPR;  // rows per warp step = rows per ring slot
    constexpr bool kPair = Dot::kPair;
    static_assert(LPR == 2 || LPR == 4 || LPR == 8 || LPR == 16 || LPR == 32, "lanes per row");
    const uint32_t ROWB = Dot::row_bytes(plan);
    const uint32_t SLOTB = ROWB * RPS;

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

    float *s_scores = reinterpret_cast<float *>(smem_raw);
    size_t off = FUSE ? ((static_cast<size_t>(a.cpad) * 4 + 127) & ~static_cast<size_t>(127)) : 0;
    float *s_q = reinterpret_cast<float *>(smem_raw + off);
    off += (static_cast<size_t>(Dot::query_bytes(plan)) + 127) & ~static_cast<size_t>(127);
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem_raw + off);  // = ring, after the drain
    const uint32_t ring = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * SLOTB;
    off += static_cast<size_t>(n_warps) * ns * SLOTB;
    const uint32_t bars = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * 8;

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
    Dot dot;
    dot.stage_query(plan, a, q_idx, s_q);
    __syncthreads();
    dot.init(plan, a, q_idx, lane, smem_u32(s_q));

    const char *rows_base = reinterpret_cast<const char *>(a.vectors);
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    const int64_t pair0 = q_begin + c0;
    const int grp = lane / LPR;

    // ---- candidate-batch pipeline: batch t+3 has its candidate ids in flight (g_*), batch t+2 its
    // spans (h), batches t+1 (B) and t (A) are scanned: every lane knows where its candidate's rows
    // sit in the batch's row sequence
    int g_base = 0, g_nb = 0, g_cand = 0;
    float g_lex = 0.f;
    int h_base = 0, h_nb = 0;
    PackedCand h{};

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
    auto resolve = [&]() {  // g_* -> h: candidate -> (first row, count)
        h_base = g_base;
        h_nb = g_nb;
        h.lex = g_lex;
        h.start = 0;
        h.cnt = 0;
        h.mine = 0;
        if (lane < g_nb) {
            uint32_t loc = 0;
            if (!candidate_ok(g_cand, a.limit, a.err, pair0 + g_base + lane)) {
                h.mine = 1;  // reported; scores as an empty document
            } else if (candidate_mine(g_cand, a.base, a.count, &loc)) {
                h.mine = 1;
                if (a.mode == FFX_MODE_PASSAGE) {
                    h.start = loc;
                    h.cnt = 1;
                } else {
                    const uint2 sp = __ldg(a.doc_span + loc);
                    h.start = sp.x;
                    h.cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
                }
            }
        }
    };
    // h.first = rows of the lower lanes' candidates; returns the batch's row count
    auto scan = [&]() {
        uint32_t inc = h.cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += up;
        }
        h.first = inc - h.cnt;
        return __shfl_sync(kFull, inc, 31);
    };

    PackedCand A, B;
    int nbA, nbB, baseA, baseB;
    uint32_t rowsA, rowsB;
    grab();
    resolve();
    rowsA = scan();
    A = h;
    nbA = h_nb;
    baseA = h_base;
    grab();
    resolve();
    rowsB = scan();
    B = h;
    nbB = h_nb;
    baseB = h_base;
    grab();
    resolve();
    grab();

    // ---- producer cursor (warp-uniform): the next step of batch A (or B) to request, up to `ns`
    // steps ahead of the consumer
    bool p_inB = false;
    uint32_t p_row = 0;  // first row of the producer's next step inside its batch
    int p_stage = 0, c_stage = 0, inflight = 0;
    uint32_t c_phase = 0;

    auto top_up = [&]() {
        while (inflight < ns) {
            const uint32_t total = p_inB ? rowsB : rowsA;
            if (p_row >= total) {
                if (!p_inB && nbB > 0) {
                    p_inB = true;
                    p_row = 0;
                    continue;
                }
                break;
            }
            const uint32_t nr = min(static_cast<uint32_t>(RPS), total - p_row);
            const uint32_t bar = bars + p_stage * 8;
            if (lane == 0) mbar_expect_tx(bar, nr * ROWB);
            // this lane's candidate: its rows inside the step's window [p_row, p_row + nr)
            const uint32_t c_first = p_inB ? B.first : A.first, c_cnt = p_inB ? B.cnt : A.cnt;
            const uint32_t c_start = p_inB ? B.start : A.start;
            const uint32_t lo = max(c_first, p_row), hi = min(c_first + c_cnt, p_row + nr);
            if (lo < hi) {
                const uint32_t dst = ring + p_stage * SLOTB + (lo - p_row) * ROWB;
                if (!indirect) {
                    bulk_g2s(dst, rows_base + static_cast<size_t>(c_start + (lo - c_first)) * ROWB, (hi - lo) * ROWB, bar);
                } else {
                    for (uint32_t r = lo; r < hi; r++) {
                        const uint32_t row = static_cast<uint32_t>(__ldg(a.doc_rows + c_start + (r - c_first)));
                        bulk_g2s(dst + (r - lo) * ROWB, rows_base + static_cast<size_t>(row) * ROWB, ROWB, bar);
                    }
                }
            }
            __syncwarp();
            p_row += nr;
            p_stage = p_stage + 1 == ns ? 0 : p_stage + 1;
            inflight++;
        }
    };

    while (nbA > 0) {
        DocReduce red;
        red.init();
        for (uint32_t r0 = 0; r0 < rowsA;) {
            // two steps (ring slots) per iteration while the batch has them; groups beyond the batch's
            // last row — and the second slot of an odd last step — run on stale bytes: nobody's window
            // holds their value
            top_up();
            const bool two = kPair && r0 + RPS < rowsA;  // warp-uniform
            const int s0 = c_stage;
            const int s1 = s0 + 1 == ns ? 0 : s0 + 1;
            float part0, part1 = 0.f;
            mbar_wait(bars + s0 * 8, (c_phase >> s0) & 1u);
            if constexpr (kPair) {
                if (two) mbar_wait(bars + s1 * 8, (c_phase >> s1) & 1u);
                dot.pair(ring + s0 * SLOTB + grp * ROWB, ring + s1 * SLOTB + grp * ROWB, lane, part0, part1);
            } else {
                part0 = dot(ring + s0 * SLOTB + grp * ROWB, lane);
            }
            __syncwarp();  // every lane has consumed its rows: the slots may be refilled
            c_phase ^= (1u << s0) | (two ? 1u << s1 : 0u);
            c_stage = two ? (s1 + 1 == ns ? 0 : s1 + 1) : s1;
            inflight -= two ? 2 : 1;
            const int rel = static_cast<int>(r0 - A.first);  // row r0 + g is row rel + g of this lane's candidate
#pragma unroll
            for (int g = 0; g < (kPair ? 2 : 1) * RPS; g++) {
                const float part = g < RPS ? part0 : part1;
                const float v = RPS == 1 ? part : __shfl_sync(kFull, part, (g % RPS) * LPR);  // a whole-warp row: every lane has it
                if (static_cast<uint32_t>(rel + g) < A.cnt) red.add(v, rel + g == 0, a.mode);
            }
            r0 += two ? 2 * RPS : RPS;
        }
        const float my_ff = A.cnt ? red.finish(A.cnt, a.mode) : 0.f;

        // lane j holds candidate j's score: coalesced epilogue for batch A
        if (lane < nbA) {
            const int64_t my_pair = pair0 + baseA + lane;
            if (A.mine) {
                float inter = my_ff;
                if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, A.lex), __fmul_rn(a.beta, my_ff));
                if (a.out_ff) a.out_ff[my_pair] = my_ff;
                if (a.out_int) a.out_int[my_pair] = inter;
                if (rank) rank[my_pair] = inter;
                if (FUSE) s_scores[c0 + baseA + lane] = inter;
            } else if (rank) {
                rank[my_pair] = __int_as_float(0x7fc00000);
            }
        }

        // shift: B becomes A, the resolved batch is scanned into B, the look-ahead loads advance
        const uint32_t rows_new = scan();
        A = B;
        rowsA = rowsB;
        nbA = nbB;
        baseA = baseB;
        B = h;
        rowsB = rows_new;
        nbB = h_nb;
        baseB = h_base;
        if (p_inB) {
            p_inB = false;  // the producer's position in old B is a position in new A
        } else {
            p_row = 0;      // it had finished old A without entering B
        }
        resolve();
        grab();
    }

    if (FUSE) {
        __syncthreads();
        float *out_s;
        int32_t *out_p;
        topk_destination(a, q_idx, &out_s, &out_p);
        rank_scores_topk<16>(s_scores, n_query, s_keys, a.k, out_s, out_p, static_cast<size_t>(n_warps) * ns * SLOTB);
        if (a.sc_world) __threadfence_system();  // peer stores: visible to the owner once the kernel ends
    }
}

}  // namespace ffx
