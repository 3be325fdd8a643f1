// ffx_kernels.cuh — sm_100a kernels of the re-ranking hot path.
//
//   ffx_score_kernel<CPL,S,FUSE>  gather passage rows by id, q.p dots with numpy's exact
//                                 pairwise tree, segmented max / Kahan-mean / first per
//                                 document, interpolation with the lexical score and (FUSE)
//                                 the per-query top-k, in ONE pass over HBM
//   ffx_score_generic_kernel      same semantics for dimensions without a lane-major plan
//   ffx_topk_kernel               per-query top-k when a query is split over several CTAs
//   ffx_permute_rows_kernel       staging: original row order -> lane-major store (and back)
//
// Replaces index/base.py:279-314, index/util.py:84-113, ranking.py:319 and
// ranking.py:115-117,285-291 of the reference.  HBM-bound (0.5 FLOP/B): no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ffx.h"
#include "ffx_layout.h"

namespace ffx {

constexpr int kThreads = 256;            // 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxFusedCand = 16384;     // fused top-k keeps <= 128 KB of keys in smem

struct ScoreArgs {
    const float *vectors;     // lane-major fp32 rows (identity layout for the generic kernel)
    const uint2 *doc_span;    // per document: (first row | offset into doc_rows, count)
    const int32_t *doc_rows;  // only when indirect
    int indirect;
    int mode;
    int64_t dim;
    const float *qvecs;       // [nq, D] original element order
    const int64_t *q_off;     // [nq+1]
    const int32_t *cand;      // [n]
    const float *lex;         // [n] or nullptr
    float alpha, beta;        // fl32(alpha), fl32(1 - alpha)
    int k;
    float *out_ff, *out_int;  // [n] or nullptr
    float *rank_scores;       // [n] or nullptr: input of a separate top-k pass (NaN = not ranked)
    float *topk_score;        // [nq, k]
    int32_t *topk_pos;
    int tiles_per_query;      // 1 when FUSE
    int tile;                 // candidates per tile
    int cpad;                 // FUSE: next pow2 >= max candidates per query
    uint32_t limit;           // valid candidates are [0, limit): GLOBAL documents (rows in PASSAGE mode)
    uint32_t base, count;     // this index holds candidates [base, base+count) of them (a shard)
    int *err;                 // set to 1 + (pair index & 0x3fffffff) when a candidate is out of range
    // scatter plan of a sharded corpus (ffx_index_set_topk_scatter): query q's top-k list goes to
    // its owner rank's receive buffer over peer memory instead of topk_score / topk_pos
    int sc_world, sc_rank;
    int64_t sc_stride;              // queries per (owner, shard) block of a receive buffer
    const int64_t *sc_bounds;       // [sc_world + 1] owner o merges queries [bounds[o], bounds[o+1])
    float *const *sc_score;         // [sc_world] receive buffers [world][stride][k], peer pointers
    int32_t *const *sc_pos;
};

// where query q's ranked list is written: locally, or into the owner's receive buffer
__device__ __forceinline__ void topk_destination(const ScoreArgs &a, int64_t q, float **out_s, int32_t **out_p) {
    if (a.sc_world == 0) {
        *out_s = a.topk_score + q * a.k;
        *out_p = a.topk_pos + q * a.k;
        return;
    }
    int o = 0;
    while (o + 1 < a.sc_world && q >= a.sc_bounds[o + 1]) o++;
    const int64_t slot = (static_cast<int64_t>(a.sc_rank) * a.sc_stride + (q - a.sc_bounds[o])) * a.k;
    *out_s = a.sc_score[o] + slot;
    *out_p = a.sc_pos[o] + slot;
}

// An out-of-range candidate is never dereferenced: it scores as an empty document and is
// reported through the index's error flag (checked by the host at the next sync).
__device__ __forceinline__ bool candidate_ok(int32_t u, uint32_t limit, int *err, int64_t pair) {
    if (static_cast<uint32_t>(u) < limit) return true;
    if (err) atomicCAS(err, 0, 1 + static_cast<int>(pair & 0x3fffffff));
    return false;
}

// Doc-id-range shards: candidates are global ordinals, a shard owns
// [base, base+count).  A pair owned by another shard is skipped: no loads, no per-pair
// outputs (the caller zero-fills them and sums across shards), no top-k key.
__device__ __forceinline__ bool candidate_mine(int32_t u, uint32_t base, uint32_t count, uint32_t *local) {
    *local = static_cast<uint32_t>(u) - base;
    return *local < count;
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;  // read-once data: keep it out of L1
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ float f4c(const float4 &v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

// Monotone key: larger score first, then smaller position first, when sorted DESCENDING.
// -0.0 is folded onto +0.0 (pandas ties them).  Key 0 is "empty" (also used for NaN).
__device__ __forceinline__ unsigned long long topk_key(float s, uint32_t pos) {
    if (s != s) return 0ull;  // NaN is not ranked (Ranking drops NaN rows, ranking.py:103)
    if (s == 0.f) s = 0.f;
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(~pos);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
    uint32_t u = static_cast<uint32_t>(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ int32_t key_pos(unsigned long long key) {
    return static_cast<int32_t>(~static_cast<uint32_t>(key));
}

// In-CTA bitonic sort, descending; `keys` may live in shared or global memory.
__device__ inline void bitonic_sort_desc(unsigned long long *keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool desc = (lo & k) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc && a != b) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

// Same result as bitonic_sort_desc, for n == E * blockDim.x keys (blockDim.x a multiple of 32):
// every thread keeps E consecutive keys in registers.  Compare-exchange distances below E stay
// inside a thread, distances up to 16*E go through warp shuffles, and only the distances that
// cross warps use shared memory (as a striped exchange buffer: element i of thread t at
// keys[i*T + t], conflict free).  For n = 8192, T = 1024 that is 15 shared-memory stages and 30
// barriers instead of 91 of each; the fused ADC kernels, whose scoring is cheap, spent ~40 %
// of their shared-memory traffic in the plain sort.
template <int E>
__device__ inline void bitonic_sort_desc_regs(unsigned long long *keys, int n) {
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int base = tid * E;
    unsigned long long r[E];
#pragma unroll
    for (int i = 0; i < E; i++) r[i] = keys[base + i];
    for (int k = 2; k <= n; k <<= 1) {
        const bool desc_hi = (base & k) == 0;  // direction of this thread's keys when k >= 2E
        for (int j = k >> 1; j >= 32 * E; j >>= 1) {  // partner in another warp, same lane
            __syncthreads();
#pragma unroll
            for (int i = 0; i < E; i++) keys[i * T + tid] = r[i];
            __syncthreads();
            const int ptid = tid ^ (j / E);
            const bool take_max = ((base & j) == 0) == desc_hi;
#pragma unroll
            for (int i = 0; i < E; i++) {
                const unsigned long long o = keys[i * T + ptid];
                r[i] = take_max ? (o > r[i] ? o : r[i]) : (o < r[i] ? o : r[i]);
            }
        }
        for (int j = min(k >> 1, 16 * E); j >= E; j >>= 1) {  // partner lane of the same warp
            const int lm = j / E;
            const bool take_max = ((base & j) == 0) == desc_hi;
#pragma unroll
            for (int i = 0; i < E; i++) {
                const unsigned long long o = __shfl_xor_sync(kFull, r[i], lm);
                r[i] = take_max ? (o > r[i] ? o : r[i]) : (o < r[i] ? o : r[i]);
            }
        }
#pragma unroll
        for (int jj = E / 2; jj >= 1; jj >>= 1) {  // inside the thread
            if (jj <= (k >> 1)) {
#pragma unroll
                for (int i = 0; i < E; i++) {
                    if ((i & jj) == 0) {
                        const bool desc = ((base + i) & k) == 0;
                        const unsigned long long a = r[i], b = r[i | jj];
                        if ((a < b) == desc && a != b) {
                            r[i] = b;
                            r[i | jj] = a;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < E; i++) keys[base + i] = r[i];
    __syncthreads();
}

// picks the register variant when the shape allows
__device__ inline void block_sort_desc(unsigned long long *keys, int n) {
    const int T = blockDim.x;
    if ((T & 31) == 0 && n == 8 * T) bitonic_sort_desc_regs<8>(keys, n);
    else if ((T & 31) == 0 && n == 4 * T) bitonic_sort_desc_regs<4>(keys, n);
    else if ((T & 31) == 0 && n == 16 * T) bitonic_sort_desc_regs<16>(keys, n);
    else bitonic_sort_desc(keys, n);
}

__device__ __forceinline__ void write_topk(const unsigned long long *keys, int n_valid, int k,
                                           float *out_s, int32_t *out_p) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < n_valid && keys[i] != 0ull;  // key 0 = empty slot (sorts last)
        out_s[i] = ok ? key_score(keys[i]) : -INFINITY;
        out_p[i] = ok ? key_pos(keys[i]) : -1;
    }
}

// Fused top-k epilogue: `scores[i]`, i < n, are a query's interpolated scores by position (NaN =
// not ranked: the pair belongs to another shard, or the score is NaN).  Builds the sort keys of
// the ranked pairs only, compacted to the front of `keys`, sorts the next power of two of
// their count and writes the k best.  A shard that owns 1/8 of the candidates therefore sorts
// 1024 keys per query instead of 8192 (the sort was ~20 % of such a shard's time).
// All threads of the CTA must call it; `keys` must hold next_pow2(n) entries.
// keys[0, m): distinct non-zero keys.  Moves the kk = min(k, m) largest to keys[0, kk) (in their
// original relative order) without sorting: MSB-first radix select of the kk-th largest key in
// 8-bit digits (histogram in shared memory, at most 8 passes, usually 3-4: it stops as soon as
// the bucket of the kk-th key is wholly selected), then a stable in-place compaction.  For
// k << m this replaces a bitonic sort of next_pow2(m) keys by one of next_pow2(k).
__device__ inline void select_largest_keys(unsigned long long *keys, int m, int kk) {
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_rem, s_done, s_base;
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        s_prefix = 0ull;
        s_rem = kk;
        s_done = 0;
    }
    int shift = 56;
    for (; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += T) s_hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i = tid; i < m; i += T) {
            const unsigned long long key = keys[i];
            if (shift == 56 || (key >> (shift + 8)) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid < 32) {  // lane l owns digits 8l .. 8l+7; find the digit holding the s_rem-th largest
            const int rem = s_rem;
            int c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c[j] = s_hist[8 * lane + j];
                tot += c[j];
            }
            int incl = tot;  // keys in the digits of lanes >= lane
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_down_sync(kFull, incl, d);
                if (lane + d < 32) incl += v;
            }
            int acc = incl - tot;  // keys in strictly higher lanes
            if (acc < rem && rem <= incl) {
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    if (acc + c[j] >= rem) {
                        s_prefix = (prefix << 8) | static_cast<unsigned long long>(8 * lane + j);
                        s_rem = rem - acc;
                        s_done = (c[j] == rem - acc) ? 1 : 0;  // the whole bucket is selected
                        break;
                    }
                    acc += c[j];
                }
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    if (shift < 0) shift = 0;
    const unsigned long long threshold = s_prefix << shift;  // smallest selected key (or a lower bound of it)
    // stable in-place compaction, one block of T keys at a time (writes never pass the reads)
    if (tid == 0) s_base = 0;
    __syncthreads();
    __shared__ int s_warp_cnt[32];
    for (int i0 = 0; i0 < m; i0 += T) {
        const int i = i0 + tid;
        const unsigned long long key = i < m ? keys[i] : 0ull;
        const bool keep = key >= threshold && key != 0ull;
        const unsigned mask = __ballot_sync(kFull, keep);
        if (lane == 0) s_warp_cnt[tid >> 5] = __popc(mask);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < (tid >> 5); w++) before += s_warp_cnt[w];
        const int base = s_base;
        if (keep) keys[base + before + __popc(mask & ((1u << lane) - 1u))] = key;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < (T + 31) / 32; w++) total += s_warp_cnt[w];
            s_base = base + total;
        }
        __syncthreads();
    }
}

// Stable LSD radix sort of keys[0, n) by their UPPER 32 bits (the score part of topk_key),
// descending, 8-bit digits, 4 passes; keys whose upper halves are equal keep their incoming
// order.  With the keys laid out by candidate position that is exactly the descending order of
// the full 64-bit keys, at a cost linear in n: ~1 k instructions per thread for 5000 keys on
// 1024 threads, where the bitonic network on the next power of two (8192) spends ~9 k.
// Warp w owns keys [w*E*32, (w+1)*E*32), E = ceil(n / T) <= EMAX, and takes them 32 at a time in
// index order: rank inside the warp = keys of the same digit in earlier rounds (the warp's
// shared-memory counter) + same-digit lanes below (match.any).  One block scan over the
// (digit, warp) counters turns them into destinations.  Keys travel through registers, so one
// buffer suffices.  `hist`: (T/32)*512 + 128 bytes of shared memory.  Requires n < 65536.
template <int EMAX>
__device__ inline void block_radix_sort_desc_hi32(unsigned long long *keys, int n, unsigned short *hist) {
    const int T = blockDim.x, W = T >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = (n + T - 1) / T;
    const int wbase = warp * E * 32;
    unsigned short *wh = hist + warp * 256;
    const unsigned lt = (1u << lane) - 1u;
    unsigned int *s_warp_total = reinterpret_cast<unsigned int *>(hist + W * 256);  // [32], behind the counters
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        unsigned long long kreg[EMAX];
        unsigned short rk[EMAX];
        for (int i = lane; i < 128; i += 32) reinterpret_cast<unsigned int *>(wh)[i] = 0u;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EMAX; r++) {
            if (r < E) {  // warp-uniform
                const int idx = wbase + r * 32 + lane;
                const bool valid = idx < n;
                kreg[r] = valid ? keys[idx] : 0ull;
                const unsigned dig = 255u - static_cast<unsigned>((kreg[r] >> shift) & 255ull);
                const unsigned peers = __match_any_sync(kFull, valid ? dig : 256u + lane);
                const unsigned before = valid ? wh[dig] : 0u;
                rk[r] = static_cast<unsigned short>(before + __popc(peers & lt));
                __syncwarp();
                if (valid && (peers & lt) == 0u) wh[dig] = static_cast<unsigned short>(before + __popc(peers));
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan of the W*256 counters in (digit, warp) order: 8 consecutive ones per thread
        // sequence number seq = digit * W + warp  <->  hist[warp * 256 + digit]; stepped, not divided
        unsigned cnt[8], sum = 0;
        const int dig0 = (tid * 8) / W, w0 = (tid * 8) % W;
        {
            int dig = dig0, w = w0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                cnt[j] = hist[w * 256 + dig];
                sum += cnt[j];
                if (++w == W) w = 0, dig++;
            }
        }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp_total[warp] = incl;
        __syncthreads();
        unsigned offset = incl - sum;
        for (int w = 0; w < warp; w++) offset += s_warp_total[w];
        {
            int dig = dig0, w = w0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                hist[w * 256 + dig] = static_cast<unsigned short>(offset);
                offset += cnt[j];
                if (++w == W) w = 0, dig++;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < EMAX; r++) {
            if (r < E) {
                const int idx = wbase + r * 32 + lane;
                if (idx < n) {
                    const unsigned dig = 255u - static_cast<unsigned>((kreg[r] >> shift) & 255ull);
                    keys[wh[dig] + rk[r]] = kreg[r];
                }
            }
        }
        __syncthreads();
    }
}

// RADIX_E > 0 enables the radix sort for lists whose every pair is ranked (no other-shard / NaN
// entries), from 2048 keys up, when k is not small enough for the select path, the CTA has at
// most RADIX_E keys per thread and `key_bytes` (the shared memory available at `keys`; 0 =
// unknown / global memory) also holds the counters.
template <int RADIX_E = 0, class ScoreAt>
__device__ inline void rank_topk(ScoreAt score_at, int n, unsigned long long *keys, int k, float *out_s,
                                 int32_t *out_p, size_t key_bytes = 0) {
    __shared__ int s_ranked;
#ifndef FFX_NO_RADIX
#define FFX_NO_RADIX 0
#endif
    if constexpr (RADIX_E > 0 && !FFX_NO_RADIX) {
        const int T = blockDim.x;
        const size_t hist_off = (static_cast<size_t>(n) * 8 + 15) & ~static_cast<size_t>(15);
        if (n >= 2048 && n < 65536 && k > n / 4 && (T & 31) == 0 && (n + T - 1) / T <= RADIX_E &&
            hist_off + static_cast<size_t>(T >> 5) * 512 + 128 <= key_bytes) {
            int all = 1;
            for (int i = threadIdx.x; i < n; i += T) {
                const unsigned long long key = topk_key(score_at(i), static_cast<uint32_t>(i));
                keys[i] = key;
                all &= key != 0ull;
            }
            if (__syncthreads_and(all)) {
                block_radix_sort_desc_hi32<RADIX_E>(
                    keys, n, reinterpret_cast<unsigned short *>(reinterpret_cast<unsigned char *>(keys) + hist_off));
                write_topk(keys, n, k, out_s, out_p);
                return;
            }
        }
    }
    if (threadIdx.x == 0) s_ranked = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int i0 = (threadIdx.x & ~31); i0 < n; i0 += blockDim.x) {  // whole warps stay in the loop
        const int i = i0 + lane;
        const unsigned long long key = i < n ? topk_key(score_at(i), static_cast<uint32_t>(i)) : 0ull;
        const unsigned ranked = __ballot_sync(kFull, key != 0ull);
        int base = 0;
        if (lane == 0 && ranked) base = atomicAdd(&s_ranked, __popc(ranked));
        base = __shfl_sync(kFull, base, 0);
        if (key != 0ull) keys[base + __popc(ranked & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    int m = s_ranked;
    if (m >= 1024 && k <= m / 4) {  // a short list out of many: select first, sort only the survivors
        select_largest_keys(keys, m, k);
        m = k;
    }
    int mpad = 1;
    while (mpad < m) mpad <<= 1;
    for (int i = m + threadIdx.x; i < mpad; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    block_sort_desc(keys, mpad);
    write_topk(keys, m, k, out_s, out_p);
}

template <int RADIX_E = 0>
__device__ inline void rank_scores_topk(const float *scores, int n, unsigned long long *keys, int k,
                                        float *out_s, int32_t *out_p, size_t key_bytes = 0) {
    rank_topk<RADIX_E>([scores](int i) { return scores[i]; }, n, keys, k, out_s, out_p, key_bytes);
}

// ---------------------------------------------------------------------------------------
// one (query, passage) dot product, bit-identical to np.sum(q * d) (N1)
// ---------------------------------------------------------------------------------------
template <int CPL, int S>
__device__ __forceinline__ float lane_chain_sum(const float (&q)[CPL * S],
                                                const float4 (&v)[CPL * S / 4]) {
    float acc[CPL];
#pragma unroll
    for (int ch = 0; ch < CPL; ch++) acc[ch] = __fmul_rn(q[ch], f4c(v[ch >> 2], ch & 3));
#pragma unroll
    for (int s = 1; s < S; s++) {
#pragma unroll
        for (int ch = 0; ch < CPL; ch++) {
            const int m = s * CPL + ch;
            acc[ch] = __fadd_rn(acc[ch], __fmul_rn(q[m], f4c(v[m >> 2], m & 3)));
        }
    }
#pragma unroll
    for (int w = 1; w < CPL; w <<= 1) {
#pragma unroll
        for (int ch = 0; ch < CPL; ch += 2 * w) acc[ch] = __fadd_rn(acc[ch], acc[ch + w]);
    }
    return acc[0];
}

__device__ __forceinline__ float warp_tree_sum(float t) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) t = __fadd_rn(t, __shfl_xor_sync(kFull, t, off));
    return __fadd_rn(0.f, t);  // the reduction starts from the identity 0
}

// Running per-document reduce (index/base.py:306-312).  Rows arrive in insertion order.
struct DocReduce {
    float acc, comp;
    __device__ __forceinline__ void init() { acc = 0.f; comp = 0.f; }
    __device__ __forceinline__ void add(float s, bool first, int mode) {
        if (mode == FFX_MODE_AVEP) {  // pandas group_mean: fp32 Kahan (N2)
            const float y = __fsub_rn(s, comp);
            const float t = __fadd_rn(acc, y);
            comp = __fsub_rn(__fsub_rn(t, acc), y);
            if (comp != comp) comp = 0.f;
            acc = t;
        } else if (first || (mode == FFX_MODE_MAXP && s > acc)) {
            acc = s;
        }
    }
    __device__ __forceinline__ float finish(uint32_t cnt, int mode) const {
        return mode == FFX_MODE_AVEP ? __fdiv_rn(acc, static_cast<float>(cnt)) : acc;
    }
};

// ---------------------------------------------------------------------------------------
// the fused hot kernel
// ---------------------------------------------------------------------------------------
template <int CPL, int S, bool FUSE>
__global__ void __launch_bounds__(kThreads, 2) ffx_score_kernel(const ScoreArgs a) {
    constexpr int EPL = CPL * S;       // elements per lane per row
    constexpr int NV4 = EPL / 4;       // 128-bit loads per lane per row
    constexpr int ROW_F4 = 32 * NV4;   // float4s per row
    static_assert(EPL % 4 == 0, "lane slice must be whole float4s");

    extern __shared__ unsigned long long s_keys[];  // FUSE only: cpad keys
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
        for (int i = threadIdx.x; i < a.cpad; i += kThreads) s_keys[i] = 0ull;
    }

    // this lane's slice of the query vector, gathered once from the original order
    float q[EPL];
    {
        const float *qv = a.qvecs + q_idx * (32 * EPL);
#pragma unroll
        for (int m = 0; m < EPL; m++) {
            const int g = lane * CPL + (m % CPL);
            q[m] = __ldg(qv + (g >> 3) * (8 * S) + 8 * (m / CPL) + (g & 7));
        }
    }
    __syncthreads();

    const float4 *rows4 = reinterpret_cast<const float4 *>(a.vectors);
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_next, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n_tile) break;
        const int nb = min(32, n_tile - base);
        const int64_t my_pair = q_begin + c0 + base + lane;

        // lane j resolves candidate j of the batch: id -> (first row, count)
        uint32_t my_start = 0, my_cnt = 0;
        float my_lex = 0.f;
        bool mine = false;
        if (lane < nb) {
            const int32_t u = __ldg(a.cand + my_pair);
            uint32_t loc = 0;
            if (!candidate_ok(u, a.limit, a.err, my_pair)) {
                mine = true;  // reported; scores as an empty document
            } else if (!candidate_mine(u, a.base, a.count, &loc)) {
                my_cnt = 0;
            } else if (a.mode == FFX_MODE_PASSAGE) {
                mine = true;
                my_start = loc;
                my_cnt = 1;
            } else {
                mine = true;
                const uint2 sp = __ldg(a.doc_span + loc);
                my_start = sp.x;
                my_cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
            }
            if (a.lex) my_lex = __ldg(a.lex + my_pair);
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
                    const bool two = r + 1 < nr;  // warp-uniform
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

        // lane j now holds candidate j's score: coalesced epilogue
        if (lane < nb && !mine && rank) rank[my_pair] = __int_as_float(0x7fc00000);
        if (lane < nb && mine) {
            float inter = my_ff;
            if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, my_lex), __fmul_rn(a.beta, my_ff));
            if (a.out_ff) a.out_ff[my_pair] = my_ff;
            if (a.out_int) a.out_int[my_pair] = inter;
            if (rank) rank[my_pair] = inter;
            if (FUSE) {
                const uint32_t pos = static_cast<uint32_t>(c0 + base + lane);
                s_keys[pos] = topk_key(inter, pos);
            }
        }
    }

    if (FUSE) {
        __syncthreads();
        bitonic_sort_desc(s_keys, a.cpad);
        write_topk(s_keys, n_query, a.k, a.topk_score + q_idx * a.k, a.topk_pos + q_idx * a.k);
    }
}

// ---------------------------------------------------------------------------------------
// generic exact kernel (any D, identity layout): one thread per pair
// ---------------------------------------------------------------------------------------
__device__ inline float pairwise_dot_generic(const float *q, const float *d, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; i++) res = __fadd_rn(res, __fmul_rn(q[i], d[i]));
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __fmul_rn(q[j], d[j]);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], __fmul_rn(q[i + j], d[i + j]));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __fadd_rn(res, __fmul_rn(q[i], d[i]));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(pairwise_dot_generic(q, d, n2), pairwise_dot_generic(q + n2, d + n2, n - n2));
}

__global__ void __launch_bounds__(128) ffx_score_generic_kernel(const ScoreArgs a, int64_t nq,
                                                                int64_t n_pairs) {
    const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (p >= n_pairs || p >= a.q_off[nq]) return;  // n_pairs is only an upper bound
    // query of this pair: binary search in q_off
    int64_t lo = 0, hi = nq;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.q_off[mid] <= p) lo = mid; else hi = mid;
    }
    const float *qv = a.qvecs + lo * a.dim;
    const int32_t u = a.cand[p];
    uint32_t start = 0, cnt = 0, loc = 0;
    if (!candidate_ok(u, a.limit, a.err, p)) {
        cnt = 0;
    } else if (!candidate_mine(u, a.base, a.count, &loc)) {
        if (a.rank_scores) a.rank_scores[p - a.q_off[0]] = __int_as_float(0x7fc00000);
        return;
    } else if (a.mode == FFX_MODE_PASSAGE) {
        start = loc;
        cnt = 1;
    } else {
        const uint2 sp = a.doc_span[loc];
        start = sp.x;
        cnt = a.mode == FFX_MODE_FIRSTP ? 1u : sp.y;
    }
    const bool indirect = a.indirect && a.mode != FFX_MODE_PASSAGE;
    DocReduce red;
    red.init();
    for (uint32_t r = 0; r < cnt; r++) {
        const uint32_t row = indirect ? static_cast<uint32_t>(a.doc_rows[start + r]) : start + r;
        const float s = __fadd_rn(
            0.f, pairwise_dot_generic(qv, a.vectors + static_cast<size_t>(row) * a.dim,
                                      static_cast<int>(a.dim)));
        red.add(s, r == 0, a.mode);
    }
    const float ff = red.finish(cnt, a.mode);
    float inter = ff;
    if (a.lex) inter = __fadd_rn(__fmul_rn(a.alpha, a.lex[p]), __fmul_rn(a.beta, ff));
    if (a.out_ff) a.out_ff[p] = ff;
    if (a.out_int) a.out_int[p] = inter;
    if (a.rank_scores) a.rank_scores[p - a.q_off[0]] = inter;
}

// ---------------------------------------------------------------------------------------
// per-query top-k over already interpolated scores (split path, generic path, merges)
// ---------------------------------------------------------------------------------------
// keys: shared memory when cpad <= kMaxFusedCand, else `gkeys + q*cpad` in global memory.
// With `lex` the kernel first interpolates (ranking.py:319): s = fl(alpha*lex) + fl(beta*scores),
// optionally storing s to out_int — `Ranking.interpolate` + `Ranking.cut` over existing scores.
// `scores_rel` != 0: `scores` is a launch-local scratch vector whose element 0 is pair q_off[0].
__global__ void __launch_bounds__(1024) ffx_topk_kernel(const float *scores, int scores_rel, const float *lex,
                                                        float alpha, float beta,
                                                        const int64_t *q_off, int k, int cpad,
                                                        unsigned long long *gkeys,
                                                        float *out_int, float *out_s,
                                                        int32_t *out_p) {
    extern __shared__ unsigned long long s_keys[];
    const int64_t q = blockIdx.x;
    const int64_t b = q_off[q];
    const int n = static_cast<int>(q_off[q + 1] - b);
    const float *sc = scores + (b - (scores_rel ? q_off[0] : 0));
    unsigned long long *keys = gkeys ? gkeys + q * cpad : s_keys;
    auto score_at = [=](int i) {
        float s = sc[i];
        if (lex) {
            s = __fadd_rn(__fmul_rn(alpha, lex[b + i]), __fmul_rn(beta, s));
            if (out_int) out_int[b + i] = s;
        }
        return s;
    };
    if (k > 0) {
        // only the ranked pairs are sorted (NaN = pair of another shard / NaN score)
        rank_topk<8>(score_at, n, keys, k, out_s + q * k, out_p + q * k, gkeys ? 0 : static_cast<size_t>(cpad) * 8);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) score_at(i);
    }
}

// Merge per-shard top-k lists [n_shards, nq, k] -> [nq, k]; positions are global positions
// inside the query's candidate block, so the ordering rule is the same key.
__global__ void __launch_bounds__(1024) ffx_merge_topk_kernel(const float *sh_s,
                                                                  const int32_t *sh_p,
                                                                  int n_shards, int64_t nq, int k,
                                                                  int cpad, float *out_s,
                                                                  int32_t *out_p) {
    extern __shared__ unsigned long long s_keys[];
    const int64_t q = blockIdx.x;
    const int total = n_shards * k;
    for (int i = threadIdx.x; i < cpad; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < total) {
            const int sh = i / k, j = i % k;
            const int64_t src = (static_cast<int64_t>(sh) * nq + q) * k + j;
            const int32_t pos = sh_p[src];
            if (pos >= 0) key = topk_key(sh_s[src], static_cast<uint32_t>(pos));
        }
        s_keys[i] = key;
    }
    __syncthreads();
    // cpad == 4, 8 or 16 keys per thread: the register / shuffle / shared-memory hybrid network
    // (15 shared-memory stages instead of 91 at 8192 keys), else the plain one
    block_sort_desc(s_keys, cpad);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = s_keys[i] != 0ull;
        out_s[q * k + i] = ok ? key_score(s_keys[i]) : -INFINITY;
        out_p[q * k + i] = ok ? key_pos(s_keys[i]) : -1;
    }
}

// ---------------------------------------------------------------------------------------
// staging: element permutation between original order and the lane-major store
// ---------------------------------------------------------------------------------------
// to_store != 0: dst(store row, staged k) = src(row, orig(k)); else the inverse.
// `rows` (optional) lists the store rows to read when gathering back.
__global__ void ffx_permute_rows_kernel(float *dst, const float *src, int64_t nrows, int dim,
                                        int cpl, int steps, int lanes, int to_store, const int64_t *rows) {
    const int64_t total = nrows * dim;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / dim;
        const int k = static_cast<int>(t % dim);
        const int e = cpl ? ffx_orig_index(cpl, steps, k, lanes) : k;
        if (to_store) {
            dst[r * dim + k] = src[r * dim + e];
        } else {
            const int64_t sr = rows ? rows[r] : r;
            dst[r * dim + e] = src[sr * dim + k];
        }
    }
}

__global__ void ffx_gather_bytes_kernel(uint8_t *dst, const uint8_t *src, int64_t nrows,
                                        int64_t row_bytes, const int64_t *rows) {
    const int64_t total = nrows * row_bytes;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / row_bytes;
        dst[t] = src[rows[r] * row_bytes + (t % row_bytes)];
    }
}

}  // namespace ffx
