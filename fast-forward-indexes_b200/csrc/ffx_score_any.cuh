// ffx_score_any.cuh — the scoring kernel for every dimension WITHOUT a uniform numpy tree
// (D = 100, 130, 300, 1000, ... up to 4096; the uniform ones have the lane-major kernels of
// ffx_score_tma.cuh).  Same contract, same pipeline as ffx_score_tma_kernel — per-warp rings of
// row slots filled by `cp.async.bulk`, candidate batches resolved three deep, dynamic balance
// over ragged documents, fused interpolation and per-query top-k — but the summation tree is
// DATA (ffx_any_plan, ffx_layout.h): rows stay in original element order, every lane owns CPL
// chains (leaf slot, accumulator j) of the numpy tree and reads their elements from the staged
// row with computed addresses; the query vector sits beside the ring in shared memory.  Short
// rows (one or two leaves) are shared by 8 / 16 lanes, 4 / 2 rows per warp step.
//
// Replaces the thread-per-pair ffx_score_generic_kernel (index/base.py:279-314 for any D,
// bit-exact: N1 of DESIGN.md).  HBM-bound; no tensor cores.
#pragma once
#include "ffx_score_tma.cuh"

namespace ffx {

// `slot_bytes`: what one warp step consumes — a row, or the 2 / 4 rows of a short-row step
__host__ __device__ inline size_t any_smem_bytes(int cpad_scores, int warps, int ns, int row_bytes, int slot_bytes) {
    const size_t keys = (static_cast<size_t>(cpad_scores) * 4 + 127) & ~static_cast<size_t>(127);
    const size_t qv = (static_cast<size_t>(row_bytes) + 127) & ~static_cast<size_t>(127);
    return keys + qv + static_cast<size_t>(warps) * ns * slot_bytes + static_cast<size_t>(warps) * ns * 8 +
           static_cast<size_t>(warps) * 2 * 32 * sizeof(CandDesc) + 128;
}

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

template <int CPL, int LPR, bool FUSE>
__global__ void __launch_bounds__(kTmaMaxThreads, 1) ffx_score_any_kernel(const ScoreArgs a, const ffx_any_plan plan,
                                                                        const int ns, const int batch) {
    constexpr int RPS = 32 / LPR;  // rows per warp step
    static_assert(LPR == 32 || CPL == 1, "short rows: one chain per lane");
    const uint32_t ROWB = static_cast<uint32_t>(plan.stride) * 4u;
    const uint32_t SLOTB = ROWB * RPS;  // a ring slot holds the up to RPS rows of one warp step

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

    // ---- carve shared memory: [scores][query vector][ring][mbarriers][descriptors]
    float *s_scores = reinterpret_cast<float *>(smem_raw);
    size_t off = FUSE ? ((static_cast<size_t>(a.cpad) * 4 + 127) & ~static_cast<size_t>(127)) : 0;
    float *s_q = reinterpret_cast<float *>(smem_raw + off);
    off += (static_cast<size_t>(ROWB) + 127) & ~static_cast<size_t>(127);
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem_raw + off);  // = ring, after the drain
    const uint32_t ring = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * SLOTB;
    off += static_cast<size_t>(n_warps) * ns * SLOTB;
    const uint32_t bars = smem_u32(smem_raw + off) + static_cast<uint32_t>(warp) * ns * 8;
    off += static_cast<size_t>(n_warps) * ns * 8;
    off = (off + 15) & ~static_cast<size_t>(15);
    CandDesc *desc = reinterpret_cast<CandDesc *>(smem_raw + off) + warp * 64;  // [2][32]

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
    const float *qsrc = a.qvecs + q_idx * plan.dim;
    for (int k = threadIdx.x; k < plan.stride; k += blockDim.x) s_q[k] = k < plan.dim ? __ldg(qsrc + k) : 0.f;

    // this lane's place in the tree
    const int sub = lane % LPR;
    const int slot = (sub * CPL) >> 3;
    const uint32_t my_byte = static_cast<uint32_t>(plan.start[slot] + ((sub * CPL) & 7)) * 4u;
    const int my_steps = plan.steps[slot];
    const bool tail_mine = slot == plan.tail_slot;
    const uint32_t tail_byte = static_cast<uint32_t>(plan.tail_start) * 4u;
    const uint32_t q_addr = smem_u32(s_q);
    __syncthreads();
    AnyQuery<CPL> q_regs;
    any_load_query<CPL>(q_regs, q_addr, my_byte, my_steps);

    const char *rows_base = reinterpret_cast<const char *>(a.vectors);
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    const int64_t pair0 = q_begin + c0;

    // ---- candidate-batch pipeline (as in ffx_score_tma_kernel)
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
    auto resolve = [&]() {
        h_base = g_base;
        h_nb = g_nb;
        h_lex = g_lex;
        h_start = 0;
        h_cnt = 0;
        h_mine = 0;
        if (lane < g_nb) {
            uint32_t loc = 0;
            if (!candidate_ok(g_cand, a.limit, a.err, pair0 + g_base + lane)) {
                h_mine = 1;
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
    auto publish = [&](int slot_) {
        CandDesc d;
        d.start = h_start;
        d.cnt = h_cnt;
        d.lex = h_lex;
        d.mine = h_mine;
        desc[slot_ * 32 + lane] = d;
        __syncwarp();
        // candidates of the batch that have rows to read (a doc-id-range shard owns a fraction of
        // them: producer and consumer walk the set bits instead of all 32 entries)
        return __ballot_sync(kFull, lane < h_nb && h_cnt > 0);
    };

    int cur = 0;
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

    bool p_inB = false;
    uint32_t p_left = liveA;  // candidates of the producer's batch it has not entered yet
    uint32_t pk = 0, pcnt = 0, pstart = 0, p_rows = 0;
    int p_stage = 0, c_stage = 0, inflight = 0;
    uint32_t c_phase = 0;

    auto top_up = [&]() {
        while (inflight < ns) {
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
            // the next min(RPS, rows left) rows of the document into one slot: one bulk copy when the
            // document's rows are consecutive, one per row otherwise; one barrier either way
            const uint32_t nr = min(static_cast<uint32_t>(RPS), pcnt - pk);
            const uint32_t bar = bars + p_stage * 8;
            if (lane == 0) mbar_expect_tx(bar, nr * ROWB);
            if (!indirect) {
                if (lane == 0)
                    bulk_g2s(ring + p_stage * SLOTB, rows_base + static_cast<size_t>(pstart + pk) * ROWB, nr * ROWB, bar);
                pk += nr;
            } else {
                for (uint32_t g = 0; g < nr; g++, pk++) {
                    if ((pk & 31u) == 0 || g == 0)
                        p_rows = ((pk & ~31u) + lane < pcnt) ? static_cast<uint32_t>(__ldg(a.doc_rows + pstart + (pk & ~31u) + lane)) : 0u;
                    const uint32_t row = __shfl_sync(kFull, p_rows, pk & 31u);
                    if (lane == 0)
                        bulk_g2s(ring + p_stage * SLOTB + g * ROWB, rows_base + static_cast<size_t>(row) * ROWB, ROWB, bar);
                }
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
                // lane group g takes row ck + g of the document, all from ring slot c_stage
                top_up();
                const int nr = static_cast<int>(min(static_cast<uint32_t>(RPS), cnt - ck));  // warp-uniform
                const int grp = lane / LPR;
                mbar_wait(bars + c_stage * 8, (c_phase >> c_stage) & 1u);
                // idle groups run the arithmetic on their (stale) part of the slot: the shuffles stay convergent
                const float part = any_row_dot<CPL, LPR>(ring + c_stage * SLOTB + grp * ROWB, q_addr, q_regs, my_byte,
                                                         my_steps, plan.max_steps, tail_mine, tail_byte, plan.tail_len);
                __syncwarp();  // every lane has consumed its row: the slot may be refilled
                c_phase ^= 1u << c_stage;
                for (int g = 0; g < nr; g++) red.add(__shfl_sync(kFull, part, g * LPR), ck + g == 0, a.mode);
                c_stage = c_stage + 1 == ns ? 0 : c_stage + 1;
                inflight--;
                ck += nr;
            }
            const float ff = red.finish(cnt, a.mode);
            if (lane == cj) my_ff = ff;
        }

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

        const uint32_t live_new = publish(cur);
        cur ^= 1;
        nbA = nbB;
        baseA = baseB;
        liveA = liveB;
        nbB = h_nb;
        baseB = h_base;
        liveB = live_new;
        if (p_inB) {
            p_inB = false;
        } else {
            p_left = liveA;
            pk = 0;
            pcnt = 0;
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
        if (a.sc_world) __threadfence_system();
    }
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
