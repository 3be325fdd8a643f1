// ffx_early_stop.cuh — `Index._early_stopping` (index/base.py:316-387) as ONE kernel launch.
//
// The reference scores a ranking in depth intervals [a, b) and, before every interval after
// the first, drops the queries whose `cutoff`-th best interpolated score so far can no longer
// be beaten:   kth_best(int_score) >= alpha * lexical(last scored row) + (1 - alpha) * max(ff_score)
// (base.py:351-356).  It does that with one pandas groupby/filter + one `_compute_scores` call
// per depth.  Queries are independent, so here one CTA owns one query and walks the depth
// list by itself: score the interval (same gather / exact pairwise dot / per-document reduce
// as ffx_score_kernel), keep the interpolated scores in shared memory, evaluate the criterion
// with an in-CTA sort of the scores so far, stop or go on.  A stopped query simply stops
// issuing loads — that is where early stopping saves HBM traffic.
//
// Shared memory: cpad fp32 interpolated scores (by position) + cpad 32-bit sort keys.
#pragma once
#include "ffx_kernels.cuh"

namespace ffx {

constexpr int kMaxEsDepths = 32;

struct EsPlan {
    int n_depths;
    int cutoff;
    int depths[kMaxEsDepths];  // ascending, every one >= cutoff, strictly increasing
    int32_t *out_scored;       // [nq] rows scored per query (a prefix of the query's block)
};

// fp32 -> uint32 whose unsigned order is the float order; -0.0 folded onto +0.0; NaN -> 0
// (pandas' nlargest / max skip NaN)
__device__ __forceinline__ uint32_t score_key32(float s) {
    if (s != s) return 0u;
    if (s == 0.f) s = 0.f;
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key32_score(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

__device__ inline void bitonic_sort_desc_u32(uint32_t *keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool desc = (lo & k) == 0;
                const uint32_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc && a != b) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

template <int CPL, int S>
__global__ void __launch_bounds__(kThreads, 2) ffx_score_es_kernel(const ScoreArgs a, const EsPlan es) {
    constexpr int EPL = CPL * S;
    constexpr int NV4 = EPL / 4;
    constexpr int ROW_F4 = 32 * NV4;

    extern __shared__ __align__(16) unsigned char es_smem[];
    float *s_int = reinterpret_cast<float *>(es_smem);                                   // [cpad]
    uint32_t *s_sort = reinterpret_cast<uint32_t *>(es_smem + static_cast<size_t>(a.cpad) * 4);  // [cpad]
    __shared__ int s_next;
    __shared__ float s_wmax[kWarps];
    __shared__ int s_go;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t q_idx = blockIdx.x;
    const int64_t q_begin = a.q_off[q_idx];
    const int n_query = static_cast<int>(a.q_off[q_idx + 1] - q_begin);

    float q[EPL];
    {
        const float *qv = a.qvecs + q_idx * (32 * EPL);
#pragma unroll
        for (int m = 0; m < EPL; m++) {
            const int g = lane * CPL + (m % CPL);
            q[m] = __ldg(qv + (g >> 3) * (8 * S) + 8 * (m / CPL) + (g & 7));
        }
    }

    const float4 *rows4 = reinterpret_cast<const float4 *>(a.vectors);
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;

    float run_max = -INFINITY;  // max ff_score over the candidates this lane finished (NaN skipped)
    int done = 0;

    for (int d = 0; d < es.n_depths; d++) {
        const int hi = min(es.depths[d], n_query);
        if (hi <= done) break;  // nothing left of this query (base.py:360-366: empty chunk)

        if (done > 0) {
            // ---- early-stopping criterion over rows [0, done)  (base.py:351-356)
            float wm = run_max;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) wm = fmaxf(wm, __shfl_xor_sync(kFull, wm, off));
            if (lane == 0) s_wmax[warp] = wm;
            int np2 = 1;
            while (np2 < done) np2 <<= 1;
            for (int i = threadIdx.x; i < np2; i += kThreads) s_sort[i] = i < done ? score_key32(s_int[i]) : 0u;
            __syncthreads();
            bitonic_sort_desc_u32(s_sort, np2);
            if (threadIdx.x == 0) {
                float max_ff = s_wmax[0];
                for (int w = 1; w < kWarps; w++) max_ff = fmaxf(max_ff, s_wmax[w]);
                // nlargest(cutoff).iat[-1]: the cutoff-th best, or the worst when fewer were scored
                const float kth = key32_score(s_sort[min(es.cutoff, done) - 1]);
                const float last_lex = __ldg(a.lex + q_begin + done - 1);
                const float bound = __fadd_rn(__fmul_rn(a.alpha, last_lex), __fmul_rn(a.beta, max_ff));
                s_go = kth < bound ? 1 : 0;
            }
            __syncthreads();
            if (!s_go) break;
        }

        if (threadIdx.x == 0) s_next = done;
        __syncthreads();

        // ---- score candidates [done, hi): warp per 32 candidates, as in ffx_score_kernel
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_next, 32);
            base = __shfl_sync(kFull, base, 0);
            if (base >= hi) break;
            const int nb = min(32, hi - base);
            const int64_t my_pair = q_begin + base + lane;

            uint32_t my_start = 0, my_cnt = 0;
            float my_lex = 0.f;
            if (lane < nb) {
                const int32_t u = __ldg(a.cand + my_pair);
                if (candidate_ok(u, a.limit, a.err, my_pair)) {
                    if (a.mode == FFX_MODE_PASSAGE) {
                        my_start = static_cast<uint32_t>(u);
                        my_cnt = 1;
                    } else {
                        const uint2 sp = __ldg(a.doc_span + u);
                        my_start = sp.x;
                        my_cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
                    }
                }
                my_lex = __ldg(a.lex + my_pair);
            }

            float my_ff = 0.f;
            for (int j = 0; j < nb; j++) {
                const uint32_t start = __shfl_sync(kFull, my_start, j);
                const uint32_t cnt = __shfl_sync(kFull, my_cnt, j);
                DocReduce red;
                red.init();
                for (uint32_t r0 = 0; r0 < cnt; r0 += 32) {
                    const uint32_t nr = min(32u, cnt - r0);
                    uint32_t my_row = start + r0 + lane;
                    if (indirect) my_row = lane < nr ? __ldg(a.doc_rows + start + r0 + lane) : 0u;
                    for (uint32_t r = 0; r < nr; r += 2) {
                        const bool two = r + 1 < nr;
                        const uint32_t row_a = __shfl_sync(kFull, my_row, r);
                        const uint32_t row_b = __shfl_sync(kFull, my_row, two ? r + 1 : r);
                        const float4 *pa = rows4 + static_cast<size_t>(row_a) * ROW_F4 + lane;
                        const float4 *pb = rows4 + static_cast<size_t>(row_b) * ROW_F4 + lane;
                        float4 va[NV4], vb[NV4];
#pragma unroll
                        for (int i = 0; i < NV4; i++) va[i] = ldg_stream(pa + i * 32);
                        if (two) {
#pragma unroll
                            for (int i = 0; i < NV4; i++) vb[i] = ldg_stream(pb + i * 32);
                        }
                        const float sa = warp_tree_sum(lane_chain_sum<CPL, S>(q, va));
                        red.add(sa, r0 + r == 0, a.mode);
                        if (two) {
                            const float sb = warp_tree_sum(lane_chain_sum<CPL, S>(q, vb));
                            red.add(sb, false, a.mode);
                        }
                    }
                }
                const float ff = red.finish(cnt, a.mode);
                if (lane == j) my_ff = ff;
            }

            if (lane < nb) {
                const float inter = __fadd_rn(__fmul_rn(a.alpha, my_lex), __fmul_rn(a.beta, my_ff));
                if (a.out_ff) a.out_ff[my_pair] = my_ff;
                if (a.out_int) a.out_int[my_pair] = inter;
                s_int[base + lane] = inter;
                run_max = fmaxf(run_max, my_ff);  // fmaxf returns the non-NaN operand, like max(skipna)
            }
        }
        __syncthreads();
        done = hi;
    }
    if (threadIdx.x == 0) es.out_scored[q_idx] = done;
}

// ---------------------------------------------------------------------------------------
// Early stopping for every other index kind (PQ / OPQ codes, dimensions that exist in the
// TMA-staged kernel only, dimensions without a lane-major plan): the same walk as a short
// sequence of launches per depth, all on the device and stream ordered — no host round trip
// decides anything.  Per depth: es_criterion (who goes on) -> es_plan (compact pair offsets)
// -> es_gather (compact candidate list) -> the index's ordinary scoring kernel over the compact
// list -> es_scatter (scores back to their pairs, interpolation, progress).
// ---------------------------------------------------------------------------------------
struct EsWalk {
    const int64_t *q_off;  // [nq+1] the full candidate blocks
    const int32_t *cand;   // [n]
    const float *lex;      // [n]
    float alpha, beta;
    int cutoff;
    int32_t *done;         // [nq] rows scored so far
    int32_t *active;       // [nq]
    int32_t *take;         // [nq] rows to score at this depth
    int64_t *part_off;     // [nq+1] offsets of the compact list
    int32_t *cand_sub;     // compact candidates
    float *ff_sub;         // their scores
    float *ff, *inter;     // [n] scores by pair (the caller's out_ff / out_int or scratch)
};

__global__ void es_init_kernel(EsWalk w, int64_t nq) {
    for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < nq;
         q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        w.done[q] = 0;
        w.active[q] = 1;
    }
}

// One CTA per query: the stopping criterion over the rows scored so far (index/base.py:351-356).
// Dynamic shared memory: next_pow2(max rows scored) 32-bit keys.
__global__ void __launch_bounds__(256) es_criterion_kernel(EsWalk w) {
    extern __shared__ uint32_t s_sort32[];
    __shared__ float s_wmax[8];
    const int64_t q = blockIdx.x;
    const int done = w.done[q];
    if (!w.active[q] || done == 0) {  // CTA-uniform
        if (threadIdx.x == 0 && done == 0) w.active[q] = 0;  // a query without rows never restarts
        return;
    }
    const int64_t b = w.q_off[q];
    int np2 = 1;
    while (np2 < done) np2 <<= 1;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        s_sort32[i] = i < done ? score_key32(w.inter[b + i]) : 0u;
        if (i < done) mx = fmaxf(mx, w.ff[b + i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
    if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = mx;
    __syncthreads();
    bitonic_sort_desc_u32(s_sort32, np2);
    if (threadIdx.x == 0) {
        float max_ff = s_wmax[0];
        for (int i = 1; i < 8; i++) max_ff = fmaxf(max_ff, s_wmax[i]);
        const float kth = key32_score(s_sort32[min(w.cutoff, done) - 1]);
        const float bound = __fadd_rn(__fmul_rn(w.alpha, w.lex[b + done - 1]), __fmul_rn(w.beta, max_ff));
        w.active[q] = kth < bound ? 1 : 0;
    }
}

// take[q] and the exclusive scan part_off[] — one CTA, chunked (nq is at most a few 10^5)
__global__ void __launch_bounds__(1024) es_plan_kernel(EsWalk w, int64_t nq, int depth) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t q0 = 0; q0 < nq; q0 += blockDim.x) {
        const int64_t q = q0 + threadIdx.x;
        long long t = 0;
        if (q < nq) {
            const int count = static_cast<int>(w.q_off[q + 1] - w.q_off[q]);
            t = w.active[q] ? max(0, min(depth, count) - w.done[q]) : 0;
            w.take[q] = static_cast<int32_t>(t);
        }
        long long incl = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        long long before = s_carry;
        for (int x = 0; x < warp; x++) before += s_warp[x];
        if (q < nq) w.part_off[q] = before + incl - t;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) w.part_off[nq] = s_carry;
}

// one CTA per query: compact candidate list of this depth
__global__ void es_gather_kernel(EsWalk w) {
    const int64_t q = blockIdx.x;
    const int t = w.take[q];
    const int64_t src = w.q_off[q] + w.done[q], dst = w.part_off[q];
    for (int i = threadIdx.x; i < t; i += blockDim.x) w.cand_sub[dst + i] = w.cand[src + i];
}

// one CTA per query: scores back to their pairs, interpolation, progress
__global__ void es_scatter_kernel(EsWalk w) {
    const int64_t q = blockIdx.x;
    const int t = w.take[q];
    const int64_t dst = w.q_off[q] + w.done[q], src = w.part_off[q];
    for (int i = threadIdx.x; i < t; i += blockDim.x) {
        const float f = w.ff_sub[src + i];
        w.ff[dst + i] = f;
        w.inter[dst + i] = __fadd_rn(__fmul_rn(w.alpha, w.lex[dst + i]), __fmul_rn(w.beta, f));
    }
    __syncthreads();
    if (threadIdx.x == 0) w.done[q] += t;
}

__global__ void es_finish_kernel(EsWalk w, int64_t nq, int32_t *out_scored) {
    for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < nq;
         q += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out_scored[q] = w.done[q];
}

}  // namespace ffx
