// ffx.cu — host side of libffx.so: the C ABI declared in include/ffx.h.
//
// Owns the HBM-resident index (row store, doc -> rows spans, PQ codebooks), stages rows
// through a pinned double buffer, plans and launches the kernels of ffx_kernels.cuh /
// ffx_adc.cuh.  No torch, no CPU compute path: every scoring entry point needs a CUDA
// device and fails with FFX_ERR_CUDA without one.
#include <cuda_runtime.h>

#include <cxxabi.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ffx.h"
#include "ffx_adc.cuh"
#include "ffx_adc_warp.cuh"
#include "ffx_adc_xor.cuh"
#include "ffx_early_stop.cuh"
#include "ffx_pq_build.cuh"
#include "ffx_kernels.cuh"
#include "ffx_score_tma.cuh"
#include "ffx_score_any.cuh"
#include "ffx_score_packed.cuh"

#include <nvtx3/nvToolsExt.h>
#include "ffx_coalesce.cuh"
#include "ffx_layout.h"

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
// the scoring kernel (fp32 or ADC) launched last, as a device-function pointer; ffx_last_kernel
// asks the runtime for its symbol and demangles it (bench.py's roofline.kernel)
std::atomic<const void *> g_last_kernel{nullptr};
thread_local std::string g_kernel_name;

template <typename K>
void note_kernel(K kern) { g_last_kernel.store(reinterpret_cast<const void *>(kern)); }

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define FFX_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            cudaGetLastError();                                                             \
            return fail(e_ == cudaErrorMemoryAllocation ? FFX_ERR_OOM : FFX_ERR_CUDA,       \
                        "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                   \
    } while (0)

#define FFX_TRY(expr)          \
    do {                       \
        int rc_ = (expr);      \
        if (rc_ != FFX_OK) return rc_; \
    } while (0)

constexpr int64_t kStageBytes = 32ll << 20;  // per pinned / device staging buffer
constexpr size_t kAdcLutScratch = 768ull << 20;  // tables of one launch built ahead (C4: 5193 x 96 KB = 510 MB)
constexpr size_t kSmemBudget = 227 * 1024 - 2048;  // opt-in shared memory per CTA on sm_100, minus static (<= 1.5 KB)

// Tuning / diagnostic knobs (ffx_set_option).  0 = automatic.
struct Tuning {
    int adc_lut = 0;     // XOR-swizzled ADC tables: 1 = built inside the scoring kernel, 2 = thread-per-entry kernel
    int kernel = 0;      // 1: register-staged ffx_score_kernel, 2: TMA-staged ffx_score_tma_kernel
    int tma_stages = 0;  // ring slots per warp
    int tma_warps = 0;   // warps per CTA of the TMA-staged kernel (8..16)
    int batch = 0;       // candidates a warp takes per grab
    int adc = 0;         // 1: ffx_adc_kernel, 2: warp-per-row, 3: XOR-swizzled thread-per-row; 0 = best the shape allows
    int chunk_waves = 0; // ffx_rerank_host: queries per pipelined chunk, in units of 2 x #SMs (0 = 1)
};

// which ADC kernel scores this index: 3 = XOR-swizzled (M % 32 == 0), 2 = warp-per-row (M = 64..128), 1 = generic
int adc_kind(const ffx_index *idx);
Tuning g_tune;

// NVTX range over a C-ABI call (header-only NVTX 3: a no-op unless a tool is attached)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct Scratch {
    void *p = nullptr;
    size_t bytes = 0;
};

int scratch_reserve(Scratch &s, size_t bytes) {
    if (bytes <= s.bytes) return FFX_OK;
    if (s.p) {
        FFX_CUDA(cudaDeviceSynchronize());
        FFX_CUDA(cudaFree(s.p));
        s.p = nullptr;
        s.bytes = 0;
    }
    bytes = (bytes + (1u << 20)) & ~static_cast<size_t>((1u << 20) - 1);
    FFX_CUDA(cudaMalloc(&s.p, bytes));
    s.bytes = bytes;
    return FFX_OK;
}

inline int next_pow2(int64_t v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

struct ffx_index {
    int device = 0;
    int row_kind = FFX_ROWS_F32;
    int64_t dim = 0;
    int64_t capacity = 0;
    int64_t num_rows = 0;
    ffx_plan plan{0, 0, 32};
    ffx_any_plan any{};   // fp32 rows of a dimension without a lane-major plan: the numpy tree as data
    size_t row_bytes = 0; // bytes per STORED row (fp32 rows of the any-plan are padded to 16 bytes)
    size_t src_row_bytes = 0;  // bytes per row as callers pass / receive them
    void *store = nullptr;
    int sm_count = 148;

    // doc -> rows
    int64_t n_docs = 0;
    uint2 *doc_span = nullptr;
    int32_t *doc_rows = nullptr;
    int indirect = 0;
    int64_t max_doc_rows = 1;  // rows of the longest document (bounds a candidate batch of the short-row kernel)

    // doc-id-range shard of a larger corpus (ffx_index_set_shard); off = whole corpus
    bool sharded = false;
    int64_t doc_base = 0, global_docs = 0, row_base = 0, global_rows = 0;

    // PQ / OPQ
    int M = 0, Ks = 0, Ds = 0;
    float *codewords = nullptr;
    float *R = nullptr;
    float *cw_t = nullptr;  // [4][Ks][32][Ds] transposed codebooks of the warp-per-row ADC kernel (M = 64..128)
    float *cw_x = nullptr;  // [M/32][Ks][32][Ds] codebooks in table order of the XOR-swizzled kernel (M = 32..128)

    // staging: pinned double buffer + device landing buffers (fp32 rows are permuted
    // from the landing buffer into the lane-major store)
    cudaStream_t stream = nullptr;
    void *pinned[2] = {nullptr, nullptr};
    void *landing[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool stage_pending = false;  // host rows copied out, device side of the staging still in flight

    Scratch work;     // kernel scratch of ffx_rerank (scores / keys / rotated queries)
    Scratch es_work;  // state of the multi-launch early-stopping walk
    Scratch hostio;   // device mirrors of ffx_rerank_host's host buffers

    // scatter plan (ffx_index_set_topk_scatter): host copy + device arrays refreshed, stream
    // ordered, by the next launch when the plan changed
    int sc_world = 0, sc_rank = 0;
    int64_t sc_stride = 0;
    bool sc_dirty = false;
    std::vector<int64_t> sc_bounds_h;
    std::vector<void *> sc_score_h, sc_pos_h;
    int64_t *sc_bounds = nullptr;
    void **sc_score = nullptr, **sc_pos = nullptr;

    int *err_flag = nullptr;   // device: first out-of-range candidate seen by a kernel (0 = none)
    int *err_host = nullptr;   // pinned mirror

    // ffx_rerank_host pipeline: H2D / two alternating compute streams / D2H
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;
};

namespace {

int adc_kind(const ffx_index *idx) {
    if (g_tune.adc == 1) return 1;
    if (idx->cw_x && g_tune.adc != 2) return 3;
    if (idx->cw_t && g_tune.adc != 3) return 2;
    return 1;
}

int bind(const ffx_index *idx) {
    FFX_CUDA(cudaSetDevice(idx->device));
    return FFX_OK;
}

// staging is asynchronous on idx->stream: wait for it before anything else reads the store
int settle(ffx_index *idx) {
    if (idx->stage_pending) {
        idx->stage_pending = false;
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
    }
    return FFX_OK;
}

// rows change shape between the caller's buffers and the store: lane-major permutation, or
// padding of the row stride to 16 bytes (any-plan dimensions that are not a multiple of 4)
bool transforms_rows(const ffx_index *idx) {
    return idx->row_kind == FFX_ROWS_F32 && (idx->plan.cpl || idx->row_bytes != idx->src_row_bytes);
}

int ensure_staging(ffx_index *idx) {
    if (idx->pinned[0]) return FFX_OK;
    for (int i = 0; i < 2; i++) {
        FFX_CUDA(cudaMallocHost(&idx->pinned[i], kStageBytes));
        if (transforms_rows(idx))
            FFX_CUDA(cudaMalloc(&idx->landing[i], kStageBytes));
        FFX_CUDA(cudaEventCreateWithFlags(&idx->done[i], cudaEventDisableTiming));
    }
    return FFX_OK;
}

void release_staging(ffx_index *idx) {
    for (int i = 0; i < 2; i++) {
        if (idx->pinned[i]) cudaFreeHost(idx->pinned[i]);
        if (idx->landing[i]) cudaFree(idx->landing[i]);
        if (idx->done[i]) cudaEventDestroy(idx->done[i]);
        idx->pinned[i] = idx->landing[i] = nullptr;
        idx->done[i] = nullptr;
    }
}

// Host rows -> pinned staging buffer on several cores: one memcpy thread moves ~14 GB/s, less
// than a third of what the PCIe 5 link behind the buffer takes.
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    constexpr size_t kMinPerThread = 4u << 20;
    unsigned n = std::min<unsigned>(6, std::max(1u, std::thread::hardware_concurrency() / 2));
    n = static_cast<unsigned>(std::min<size_t>(n, std::max<size_t>(1, bytes / kMinPerThread)));
    if (n <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t step = ((bytes + n - 1) / n + 4095) & ~static_cast<size_t>(4095);
    for (unsigned t = 1; t < n; t++) {
        const size_t lo = std::min(bytes, t * step), hi = std::min(bytes, lo + step);
        if (lo < hi)
            pool.emplace_back([=] { memcpy(static_cast<char *>(dst) + lo, static_cast<const char *>(src) + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, step));
    for (auto &th : pool) th.join();
}

int permute_grid(int64_t elems, int sm_count) {
    const int64_t blocks = (elems + 255) / 256;
    return static_cast<int>(std::min<int64_t>(blocks, static_cast<int64_t>(sm_count) * 16));
}

// original order (device) -> store rows [row0, row0+nrows)
int launch_to_store(ffx_index *idx, int64_t row0, int64_t nrows, const void *src_dev) {
    char *dst = static_cast<char *>(idx->store) + static_cast<size_t>(row0) * idx->row_bytes;
    if (idx->row_kind == FFX_ROWS_F32 && idx->plan.cpl) {
        ffx::ffx_permute_rows_kernel<<<permute_grid(nrows * idx->dim, idx->sm_count), 256, 0,
                                       idx->stream>>>(
            reinterpret_cast<float *>(dst), static_cast<const float *>(src_dev), nrows,
            static_cast<int>(idx->dim), idx->plan.cpl, idx->plan.steps, idx->plan.lanes, 1, nullptr);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
    } else if (transforms_rows(idx)) {
        ffx::ffx_pad_rows_kernel<<<permute_grid(nrows * idx->any.stride, idx->sm_count), 256, 0, idx->stream>>>(
            reinterpret_cast<float *>(dst), static_cast<const float *>(src_dev), nrows, static_cast<int>(idx->dim),
            idx->any.stride, 1, nullptr);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
    } else {
        FFX_CUDA(cudaMemcpyAsync(dst, src_dev, static_cast<size_t>(nrows) * idx->row_bytes,
                                 cudaMemcpyDeviceToDevice, idx->stream));
    }
    return FFX_OK;
}

// ---- kernel dispatch -------------------------------------------------------------------
template <int CPL, int S>
int launch_score(const ffx::ScoreArgs &a, bool fuse, int grid, size_t smem, cudaStream_t st) {
    if (fuse) {
        auto kern = ffx::ffx_score_kernel<CPL, S, true>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        kern<<<grid, ffx::kThreads, smem, st>>>(a);
        note_kernel(kern);
    } else {
        ffx::ffx_score_kernel<CPL, S, false><<<grid, ffx::kThreads, 0, st>>>(a);
        note_kernel(ffx::ffx_score_kernel<CPL, S, false>);
    }
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

template <int CPL, int S>
int launch_score_tma(const ffx::ScoreArgs &a, bool fuse, int grid, int warps, int ns, int batch,
                     cudaStream_t st) {
    const size_t smem = ffx::tma_smem_bytes(fuse ? a.cpad : 0, warps, ns, 32 * CPL * S * 4);
    if (fuse) {
        auto kern = ffx::ffx_score_tma_kernel<CPL, S, true>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, warps * 32, smem, st>>>(a, ns, batch);
        note_kernel(kern);
    } else {
        auto kern = ffx::ffx_score_tma_kernel<CPL, S, false>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, warps * 32, smem, st>>>(a, ns, batch);
        note_kernel(kern);
    }
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

template <int V16>
int launch_adc_v(const ffx::AdcArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
    auto kern = ffx::ffx_adc_kernel<V16>;
    FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, ffx::kThreads, smem, st>>>(a);
    note_kernel(kern);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

int launch_adc(const ffx::AdcArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
    if (a.M % 16 == 0) {
        switch (a.M / 16) {
            case 1: return launch_adc_v<1>(a, grid, smem, st);
            case 2: return launch_adc_v<2>(a, grid, smem, st);
            case 3: return launch_adc_v<3>(a, grid, smem, st);
            case 4: return launch_adc_v<4>(a, grid, smem, st);
            case 6: return launch_adc_v<6>(a, grid, smem, st);
            case 8: return launch_adc_v<8>(a, grid, smem, st);
            default: break;
        }
    }
    return launch_adc_v<0>(a, grid, smem, st);
}

// ---- launch planning for the fp32 lane-major path ------------------------------------------
// TMA-staged kernel shapes (tools/sweep.py on B200): the register file
// allows 16 warps per SM at ~122 registers, shared memory (227 KB) holds the row rings.
//   fused, small key arrays : 2 CTAs/SM x 8 warps x >=3 ring slots (tails of one query overlap
//                             the other CTA's stream)
//   fused, 5000 candidates  : 1 CTA/SM x 14 warps x 4 slots (32 KB of scores leave 172 KB of ring)
//   tiled (few queries)     : 4 warps per CTA, 4 CTAs/SM
struct ScorePlan {
    bool tma = false;
    bool packed = false;  // ffx_score_packed_kernel with LaneMajorDot (rows of up to 512 elements)
    int warps = 8, ns = 0, batch = 16;
};

int ring_slots(int cpad_scores, int warps, int row_bytes, int ctas_per_sm) {
    const size_t budget = kSmemBudget / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
    const size_t fixed = ffx::tma_smem_bytes(cpad_scores, warps, 0, row_bytes);
    if (fixed >= budget) return 0;
    int ns = static_cast<int>((budget - fixed) / (static_cast<size_t>(warps) * (row_bytes + 8)));
    ns = std::min(ns, 16);
    // the fused top-k sorts cpad 64-bit keys inside the drained ring
    if (static_cast<size_t>(ns) * warps * row_bytes < static_cast<size_t>(cpad_scores) * 8) return 0;
    return ns;
}

int any_ring_slots(int keys, int warps, int row_bytes, int slot_bytes, int ctas_per_sm);

ScorePlan plan_score(int mode, bool fuse, int cpad, int row_bytes, bool sparse, bool few_pairs, bool short_rows) {
    ScorePlan p;
    // candidates a warp publishes per batch: few rows per candidate (single-row modes, or a shard
    // that owns only a fraction of the candidates) want the larger batch so that the row ring
    // always has something to prefetch
    const bool one_row = mode == FFX_MODE_PASSAGE || mode == FFX_MODE_FIRSTP;
    p.batch = g_tune.batch > 0 ? std::min(32, g_tune.batch) : (one_row || sparse ? 32 : 16);
    const int keys = fuse ? cpad : 0;
    if (ffx::tma_streams_query(row_bytes)) {
        // 10-16 KB rows (D >= 2560): only the TMA-staged kernel exists; few warps, each with at least
        // two row slots, fill shared memory next to the query vector and the scores
        for (p.warps = 8; p.warps >= 2; p.warps -= 2) {
            p.ns = std::min(4, ring_slots(keys, p.warps, row_bytes, 1));
            if (p.ns >= 2) break;
        }
        if (g_tune.tma_stages > 0) p.ns = std::min(p.ns, std::max(2, g_tune.tma_stages));
        p.tma = p.ns >= 2;
        return p;
    }
    if (short_rows) {
        // D <= 256 (ffx_score_packed_kernel): `row_bytes` is the ring slot of one warp step here — 4 .. 16
        // rows, 3-4 KB.  Fused: the shape with the most slots in flight per SM, at least two per warp.
        // Full batches of 32 candidates: a batch ends with a partly filled step.
        if (g_tune.batch <= 0) p.batch = 32;
        if (fuse) {
            // (ties go to two CTAs: one query's top-k epilogue overlaps the other's row stream)
            const int shapes[][2] = {{8, 2}, {16, 1}, {12, 1}, {8, 1}, {4, 1}};
            int best = 0;
            p.ns = 0;
            for (const auto &shape : shapes) {
                const int ns = std::min(6, any_ring_slots(keys, shape[0], 0, row_bytes, shape[1]));
                if (ns < 2 || ns * shape[0] * shape[1] <= best) continue;
                best = ns * shape[0] * shape[1];
                p.warps = shape[0];
                p.ns = ns;
            }
        } else {
            p.warps = few_pairs ? 2 : 4;
            if (few_pairs && g_tune.batch <= 0) p.batch = 8;
            p.ns = std::min(6, any_ring_slots(0, p.warps, 0, row_bytes, 4));
            if (p.ns < 2) p.ns = std::min(6, any_ring_slots(0, p.warps, 0, row_bytes, 1));
        }
        if (g_tune.tma_stages > 0) p.ns = std::min(p.ns, std::max(2, g_tune.tma_stages));
        p.tma = p.ns >= 2;
        p.packed = true;
        return p;
    }
    if (g_tune.kernel == 1) return p;
    if (g_tune.tma_warps > 0) {  // explicit shape (sweeps)
        p.warps = g_tune.tma_warps;
        p.ns = ring_slots(keys, p.warps, row_bytes, 1);
    } else if (!fuse && few_pairs) {
        // a handful of queries (serving latency): small CTAs and small batches give enough tiles to
        // put rows in flight on every SM
        p.warps = 2;
        p.batch = g_tune.batch > 0 ? p.batch : 8;
        p.ns = std::min(8, ring_slots(0, 2, row_bytes, 4));
    } else if (!fuse) {
        p.warps = 4;
        p.ns = std::min(4, ring_slots(0, 4, row_bytes, 4));
    } else {
        p.warps = 8;
        p.ns = std::min(6, ring_slots(keys, 8, row_bytes, 2));
        if (p.ns < 3) {
            p.warps = 14;
            p.ns = std::min(6, ring_slots(keys, 14, row_bytes, 1));
        }
        if (p.ns < 2) {
            p.warps = 8;
            p.ns = ring_slots(keys, 8, row_bytes, 1);
        }
    }
    if (g_tune.tma_stages > 0) p.ns = std::min(p.ns, g_tune.tma_stages);
    p.tma = p.ns >= 2;
    return p;
}

// ---- short rows and dimensions without a lane-major plan: ffx_score_packed_kernel ----------------
template <class Dot>
int launch_score_packed(const ffx::ScoreArgs &a, const typename Dot::Plan &plan, int query_bytes, int slot_bytes, bool fuse,
                        int grid, int warps, int ns, int batch, cudaStream_t st) {
    const size_t smem = ffx::packed_smem_bytes(fuse ? a.cpad : 0, warps, ns, query_bytes, slot_bytes);
    if (fuse) {
        auto kern = ffx::ffx_score_packed_kernel<Dot, true>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, warps * 32, smem, st>>>(a, plan, ns, batch);
        note_kernel(kern);
    } else {
        auto kern = ffx::ffx_score_packed_kernel<Dot, false>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, warps * 32, smem, st>>>(a, plan, ns, batch);
        note_kernel(kern);
    }
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

int dispatch_score_any(const ffx_any_plan &p, const ScorePlan &sp, const ffx::ScoreArgs &a, bool fuse, int grid,
                       cudaStream_t st) {
    if (p.cpl == 2 && p.lpr == 4)
        return launch_score_packed<ffx::TreeDot<2, 4>>(a, p, p.stride * 4, p.stride * 32, fuse, grid, sp.warps, sp.ns, sp.batch, st);
    if (p.cpl == 2 && p.lpr == 8)
        return launch_score_packed<ffx::TreeDot<2, 8>>(a, p, p.stride * 4, p.stride * 16, fuse, grid, sp.warps, sp.ns, sp.batch, st);
    if (p.cpl == 2 && p.lpr == 16)
        return launch_score_packed<ffx::TreeDot<2, 16>>(a, p, p.stride * 4, p.stride * 8, fuse, grid, sp.warps, sp.ns, sp.batch, st);
#define FFX_CASE(C) \
    if (p.cpl == C && p.lpr == 32) \
        return launch_score_packed<ffx::TreeDot<C, 32>>(a, p, p.stride * 4, p.stride * 4, fuse, grid, sp.warps, sp.ns, sp.batch, st)
    FFX_CASE(2);
    FFX_CASE(4);
    FFX_CASE(8);
#undef FFX_CASE
    return fail(FFX_ERR_UNSUPPORTED, "no kernel for tree plan (%d chains per lane, %d lanes per row)", p.cpl, p.lpr);
}

// ring slots per warp for `ctas_per_sm` resident CTAs; 0 if the shape does not fit (the fused
// top-k sorts its 64-bit keys inside the drained ring)
int any_ring_slots(int keys, int warps, int row_bytes, int slot_bytes, int ctas_per_sm) {
    const size_t budget = kSmemBudget / ctas_per_sm - 1024;  // static shared memory of the fused epilogue: 2.3 KB
    const size_t fixed = ffx::packed_smem_bytes(keys, warps, 0, row_bytes, slot_bytes);
    if (fixed >= budget) return 0;
    int ns = static_cast<int>((budget - fixed) / (static_cast<size_t>(warps) * (slot_bytes + 8)));
    ns = std::min(ns, 32);
    if (static_cast<size_t>(ns) * warps * slot_bytes < static_cast<size_t>(keys) * 8) return 0;
    return ns;
}

ScorePlan plan_score_any(const ffx_any_plan &p, int mode, bool fuse, int cpad, bool sparse, bool few_pairs) {
    ScorePlan sp;
    const int row_bytes = p.stride * 4, rps = 32 / p.lpr, slot_bytes = row_bytes * rps, keys = fuse ? cpad : 0;
    const bool one_row = mode == FFX_MODE_PASSAGE || mode == FFX_MODE_FIRSTP;
    sp.batch = g_tune.batch > 0 ? std::min(32, g_tune.batch) : (one_row || sparse ? 32 : 16);
    const int need = 2;  // a warp step consumes one slot while the next one loads
    const int want = slot_bytes >= 2048 ? 6 : (slot_bytes >= 1024 ? 8 : 16);
    if (fuse) {
        // the shape that keeps the most ring slots (bytes) in flight per SM: two 8-warp CTAs, or one
        // CTA of up to 16 warps when the keys and the query vector leave too little for two
        const int shapes[][2] = {{8, 2}, {16, 1}, {14, 1}, {12, 1}, {8, 1}, {4, 1}};
        int best = 0;
        sp.ns = 0;
        for (const auto &shape : shapes) {
            const int ns = std::min(want, any_ring_slots(keys, shape[0], row_bytes, slot_bytes, shape[1]));
            if (ns < need) continue;
            const int slots = ns * shape[0] * shape[1];
            if (slots > best) {
                best = slots;
                sp.warps = shape[0];
                sp.ns = ns;
            }
        }
    } else {
        sp.warps = few_pairs ? 2 : 4;
        if (few_pairs && g_tune.batch <= 0) sp.batch = 8;
        sp.ns = std::min(want, any_ring_slots(0, sp.warps, row_bytes, slot_bytes, 4));
        if (sp.ns < need) sp.ns = std::min(want, any_ring_slots(0, sp.warps, row_bytes, slot_bytes, 1));
    }
    if (g_tune.tma_stages > 0) sp.ns = std::max(need, std::min(sp.ns, g_tune.tma_stages));
    if (g_tune.tma_warps > 0 && any_ring_slots(keys, g_tune.tma_warps, row_bytes, slot_bytes, 1) >= sp.ns) sp.warps = g_tune.tma_warps;
    sp.tma = sp.ns >= need && sp.warps >= 1;
    return sp;
}

int dispatch_score(const ffx_plan &p, const ScorePlan &sp, const ffx::ScoreArgs &a, bool fuse, int grid,
                   cudaStream_t st) {
    if (sp.tma && sp.packed) {
        // rows of up to 512 elements: several rows per warp step (ffx_score_packed_kernel)
#define FFX_CASE(S_, L, C)                                                                                   \
    if (p.lanes == L && p.steps == S_ && ffx_packed_cpl(p) == C)                                                 \
        return launch_score_packed<ffx::LaneMajorDot<S_, L, C>>(a, {}, 0, 32 * C * S_ * 4, fuse, grid, sp.warps, sp.ns, \
                                                                sp.batch, st)
        FFX_CASE(8, 8, 4);    // D = 64: 2 lanes per row, 16 rows per warp step
        FFX_CASE(12, 8, 2);   // D = 96: 4 lanes, 8 rows
        FFX_CASE(16, 8, 2);   // D = 128
        FFX_CASE(12, 16, 2);  // D = 192: 8 lanes, 4 rows
        FFX_CASE(16, 16, 2);  // D = 256
        FFX_CASE(12, 32, 2);  // D = 384: 16 lanes, 2 rows
        FFX_CASE(16, 32, 2);  // D = 512
#undef FFX_CASE
    }
    if (sp.tma) {
#define FFX_CASE(C, S_) \
    if (p.lanes == 32 && p.cpl == C && p.steps == S_) \
        return launch_score_tma<C, S_>(a, fuse, grid, sp.warps, sp.ns, sp.batch, st)
        FFX_CASE(1, 12);
        FFX_CASE(1, 16);
        FFX_CASE(2, 10);
        FFX_CASE(2, 12);
        FFX_CASE(2, 14);
        FFX_CASE(2, 16);
        FFX_CASE(4, 12);
        FFX_CASE(4, 16);
        FFX_CASE(8, 10);
        FFX_CASE(8, 12);
        FFX_CASE(8, 14);
        FFX_CASE(8, 16);
#undef FFX_CASE
    }
    const size_t smem = fuse ? static_cast<size_t>(a.cpad) * 8 : 0;
#define FFX_CASE(C, S_) \
    if (p.lanes == 32 && p.cpl == C && p.steps == S_) return launch_score<C, S_>(a, fuse, grid, smem, st)
    FFX_CASE(1, 12);
    FFX_CASE(1, 16);
    FFX_CASE(2, 10);
    FFX_CASE(2, 12);
    FFX_CASE(2, 14);
    FFX_CASE(2, 16);
    FFX_CASE(4, 12);
    FFX_CASE(4, 16);
#undef FFX_CASE
    return fail(FFX_ERR_UNSUPPORTED, "no lane-major kernel for plan (%d,%d)", p.cpl, p.steps);
}

// One CTA per query with the top-k fused into the scoring kernel: needs the lane-major fast
// path, keys that fit shared memory, and enough queries to fill the machine.
bool will_fuse(const ffx_index *idx, int64_t nq, int k, int cpad) {
    if (k <= 0 || cpad > ffx::kMaxFusedCand) return false;
    if (idx->row_kind == FFX_ROWS_PQ_U8) {  // fused ADC kernels: one CTA per SM
        const int kind = adc_kind(idx);
        if (kind == 1 || nq < static_cast<int64_t>(idx->sm_count)) return false;
        if (kind == 3)  // sort keys overlay table + slots
            return ffx::adc_xor_smem_bytes(idx->M, idx->Ks, cpad) <= kSmemBudget &&
                   static_cast<size_t>(cpad) * 8 <= ffx::adc_xor_smem_bytes(idx->M, idx->Ks, 0) - 128;
        return ffx::adc_warp_smem_bytes(idx->Ks, cpad) <= kSmemBudget;
    }
    return (idx->plan.cpl != 0 || idx->any.valid) && nq >= static_cast<int64_t>(idx->sm_count) * 2;
}

template <int NC>
int launch_adc_xor(const ffx::AdcWarpArgs &w, bool fuse, unsigned grid, size_t smem, cudaStream_t st) {
    if (fuse) {
        auto kern = ffx::ffx_adc_xor_kernel<NC, true>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, ffx::kAdcXorThreads, smem, st>>>(w);
        note_kernel(kern);
    } else {
        auto kern = ffx::ffx_adc_xor_kernel<NC, false>;
        FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, ffx::kAdcXorThreads, smem, st>>>(w);
        note_kernel(kern);
    }
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

template <bool FUSE, bool INDIRECT>
int launch_adc_warp(const ffx::AdcWarpArgs &w, unsigned grid, size_t smem, cudaStream_t st) {
    auto kern = ffx::ffx_adc_warp_kernel<FUSE, INDIRECT>;
    FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, ffx::kAdcWarpThreads, smem, st>>>(w);
    note_kernel(kern);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

int take_error(ffx_index *idx, cudaStream_t st) {
    FFX_CUDA(cudaMemcpyAsync(idx->err_host, idx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    FFX_CUDA(cudaStreamSynchronize(st));
    const int seen = *idx->err_host;
    if (seen == 0) return FFX_OK;
    FFX_CUDA(cudaMemsetAsync(idx->err_flag, 0, sizeof(int), st));
    FFX_CUDA(cudaStreamSynchronize(st));
    return fail(FFX_ERR_INVALID, "candidate out of range near pair %d: not a document ordinal / row "
                "of this index", seen - 1);
}

cudaEvent_t event_at(ffx_index *idx, size_t i) {
    while (idx->events.size() <= i) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        idx->events.push_back(e);
    }
    return idx->events[i];
}

int launch_topk(const float *scores, bool scores_rel, const float *lex, float alpha, float beta, float *out_int,
                const int64_t *q_off, int64_t nq, int k, int cpad, Scratch &work, size_t work_off,
                float *out_s, int32_t *out_p, cudaStream_t st) {
    unsigned long long *gkeys = nullptr;
    size_t smem = static_cast<size_t>(cpad) * 8;
    if (cpad > ffx::kMaxFusedCand) {
        gkeys = reinterpret_cast<unsigned long long *>(static_cast<char *>(work.p) + work_off);
        smem = 0;
    }
    if (smem > 48 * 1024)
        FFX_CUDA(cudaFuncSetAttribute(ffx::ffx_topk_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    // 8 keys per thread where possible: the register / shuffle sort applies from 2048 keys up
    const int threads = std::max(256, std::min(1024, cpad / 8));
    ffx::ffx_topk_kernel<<<static_cast<unsigned>(nq), threads, smem, st>>>(
        scores, scores_rel ? 1 : 0, lex, alpha, beta, q_off, k, cpad, gkeys, out_int, out_s, out_p);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

template <int CPL, int S>
int launch_es(const ffx::ScoreArgs &a, const ffx::EsPlan &es, unsigned grid, size_t smem, cudaStream_t st) {
    auto kern = ffx::ffx_score_es_kernel<CPL, S>;
    FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, ffx::kThreads, smem, st>>>(a, es);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

// base.py:339-343,365-366: ascending, depths below the cutoff skipped, the walk ends at the first
// depth that cannot add rows (a repeat)
int plan_depths(const int32_t *depths, int n_depths, int cutoff, ffx::EsPlan *es) {
    std::vector<int32_t> d(depths, depths + n_depths);
    std::sort(d.begin(), d.end());
    es->n_depths = 0;
    es->cutoff = cutoff;
    int prev = 0;
    for (int32_t b : d) {
        if (b < cutoff) continue;
        if (b <= prev) break;
        if (es->n_depths == ffx::kMaxEsDepths)
            return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank_early_stop: more than %d depths", ffx::kMaxEsDepths);
        es->depths[es->n_depths++] = b;
        prev = b;
    }
    return FFX_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
extern "C" {

int ffx_abi_version(void) { return FFX_ABI_VERSION; }

// used by the other translation units of the library (ffx_ids.cpp) to report through ffx_last_error
int ffx_set_error_message(int code, const char *msg) { return fail(code, "%s", msg ? msg : ""); }

const char *ffx_last_error(void) { return g_err.c_str(); }

int ffx_set_option(const char *name, int value) {
    if (!name) return fail(FFX_ERR_INVALID, "ffx_set_option: NULL name");
    const std::string key(name);
    if (key == "kernel" && value >= 0 && value <= 2) g_tune.kernel = value;
    else if (key == "tma_stages" && value >= 0 && value <= 16) g_tune.tma_stages = value;
    else if (key == "batch" && value >= 0 && value <= 32) g_tune.batch = value;
    else if (key == "adc" && value >= 0 && value <= 3) g_tune.adc = value;
    else if (key == "adc_lut" && value >= 0 && value <= 2) g_tune.adc_lut = value;
    else if (key == "chunk_waves" && value >= 0 && value <= 64) g_tune.chunk_waves = value;
    else if (key == "tma_warps" && (value == 0 || (value >= 1 && value <= 16))) g_tune.tma_warps = value;
    else return fail(FFX_ERR_INVALID, "ffx_set_option: unknown option or bad value (%s=%d)", name, value);
    return FFX_OK;
}

int ffx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t ffx_launch_count(void) { return g_launches.load(); }

const char *ffx_last_kernel(void) {
    g_kernel_name.clear();
    const void *fn = g_last_kernel.load();
    const char *mangled = nullptr;
    if (fn && cudaFuncGetName(&mangled, fn) == cudaSuccess && mangled) {
        int status = 0;
        char *plain = abi::__cxa_demangle(mangled, nullptr, nullptr, &status);
        g_kernel_name = (status == 0 && plain) ? plain : mangled;
        free(plain);
    } else {
        cudaGetLastError();
    }
    return g_kernel_name.c_str();
}

int ffx_host_alloc(void **out, int64_t bytes) {
    if (!out || bytes < 0) return fail(FFX_ERR_INVALID, "ffx_host_alloc: bad arguments");
    *out = nullptr;
    if (bytes == 0) return FFX_OK;
    FFX_CUDA(cudaMallocHost(out, static_cast<size_t>(bytes)));
    return FFX_OK;
}

int ffx_host_free(void *p) {
    if (p) FFX_CUDA(cudaFreeHost(p));
    return FFX_OK;
}

// ---- index ---------------------------------------------------------------------------
int ffx_index_create(int device, int row_kind, int64_t dim, int64_t capacity_rows,
                     ffx_index **out) {
    if (!out) return fail(FFX_ERR_INVALID, "ffx_index_create: out is NULL");
    *out = nullptr;
    if (row_kind != FFX_ROWS_F32 && row_kind != FFX_ROWS_PQ_U8)
        return fail(FFX_ERR_INVALID, "ffx_index_create: unknown row kind %d", row_kind);
    if (dim <= 0 || dim > (1 << 20) || capacity_rows < 0 || capacity_rows > 0xffffffffll)
        return fail(FFX_ERR_INVALID, "ffx_index_create: bad dim/capacity (%lld, %lld)",
                    static_cast<long long>(dim), static_cast<long long>(capacity_rows));
    const int n_dev = ffx_device_count();
    if (n_dev <= 0) return fail(FFX_ERR_CUDA, "no CUDA device (ffx has no CPU path)");
    if (device < 0 || device >= n_dev)
        return fail(FFX_ERR_INVALID, "ffx_index_create: device %d of %d", device, n_dev);
    FFX_CUDA(cudaSetDevice(device));

    ffx_index *idx = new ffx_index();
    idx->device = device;
    idx->row_kind = row_kind;
    idx->dim = dim;
    if (row_kind == FFX_ROWS_F32) {
        idx->plan = ffx_plan_for_dim(dim);
        idx->row_bytes = static_cast<size_t>(dim) * 4;
        if (!idx->plan.cpl) {
            idx->any = ffx_any_plan_for_dim(dim);
            if (idx->any.valid) idx->row_bytes = static_cast<size_t>(idx->any.stride) * 4;
        }
    } else {
        idx->row_bytes = static_cast<size_t>(dim);
    }
    idx->src_row_bytes = static_cast<size_t>(dim) * (row_kind == FFX_ROWS_F32 ? 4 : 1);
    cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete idx;
        return fail(FFX_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = idx;
    bool ok = cudaMalloc(&idx->err_flag, sizeof(int)) == cudaSuccess &&
              cudaMemset(idx->err_flag, 0, sizeof(int)) == cudaSuccess &&
              cudaMallocHost(&idx->err_host, sizeof(int)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&idx->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&idx->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&idx->s_comp[0], cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&idx->s_comp[1], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        ffx_index_destroy(idx);
        *out = nullptr;
        return fail(FFX_ERR_CUDA, "ffx_index_create: could not create streams / error flag");
    }
    int rc = ffx_index_reserve(idx, capacity_rows);
    if (rc != FFX_OK) {
        ffx_index_destroy(idx);
        *out = nullptr;
    }
    return rc;
}

int ffx_index_destroy(ffx_index *idx) {
    if (!idx) return FFX_OK;
    cudaSetDevice(idx->device);
    cudaDeviceSynchronize();
    release_staging(idx);
    cudaFree(idx->store);
    cudaFree(idx->doc_span);
    cudaFree(idx->doc_rows);
    cudaFree(idx->codewords);
    cudaFree(idx->R);
    cudaFree(idx->cw_t);
    cudaFree(idx->cw_x);
    cudaFree(idx->sc_bounds);
    cudaFree(idx->sc_score);
    cudaFree(idx->sc_pos);
    cudaFree(idx->work.p);
    cudaFree(idx->es_work.p);
    cudaFree(idx->hostio.p);
    cudaFree(idx->err_flag);
    if (idx->err_host) cudaFreeHost(idx->err_host);
    for (cudaEvent_t e : idx->events) cudaEventDestroy(e);
    for (cudaStream_t st : {idx->s_h2d, idx->s_d2h, idx->s_comp[0], idx->s_comp[1]})
        if (st) cudaStreamDestroy(st);
    if (idx->stream) cudaStreamDestroy(idx->stream);
    cudaGetLastError();
    delete idx;
    return FFX_OK;
}

int ffx_index_reserve(ffx_index *idx, int64_t capacity_rows) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_index_reserve: NULL index");
    if (capacity_rows > 0xffffffffll)
        return fail(FFX_ERR_INVALID, "ffx_index_reserve: more than 2^32 rows");
    if (capacity_rows <= idx->capacity) return FFX_OK;
    FFX_TRY(bind(idx));
    FFX_TRY(settle(idx));
    void *fresh = nullptr;
    // 256 B slack: the kernels may issue (masked) vector loads just past the last row
    FFX_CUDA(cudaMalloc(&fresh, static_cast<size_t>(capacity_rows) * idx->row_bytes + 256));
    if (idx->store) {
        if (idx->num_rows > 0)
            FFX_CUDA(cudaMemcpyAsync(fresh, idx->store,
                                     static_cast<size_t>(idx->num_rows) * idx->row_bytes,
                                     cudaMemcpyDeviceToDevice, idx->stream));
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
        FFX_CUDA(cudaFree(idx->store));
    }
    idx->store = fresh;
    idx->capacity = capacity_rows;
    return FFX_OK;
}

int ffx_index_copy_rows(ffx_index *dst, ffx_index *src) {
    NvtxRange nvtx_range("ffx_index_copy_rows");
    if (!dst || !src || dst == src) return fail(FFX_ERR_INVALID, "ffx_index_copy_rows: bad arguments");
    if (dst->row_kind != src->row_kind || dst->dim != src->dim)
        return fail(FFX_ERR_INVALID, "ffx_index_copy_rows: the two indexes hold different kinds of rows");
    if (dst->num_rows != 0) return fail(FFX_ERR_STATE, "ffx_index_copy_rows: the destination is not empty");
    FFX_TRY(bind(src));
    FFX_TRY(settle(src));
    FFX_CUDA(cudaStreamSynchronize(src->stream));
    if (src->num_rows == 0) return FFX_OK;
    FFX_TRY(ffx_index_reserve(dst, src->num_rows));
    FFX_TRY(bind(dst));
    // the store's own layout (lane-major / padded) is the same on both sides: a plain byte copy
    const size_t bytes = static_cast<size_t>(src->num_rows) * src->row_bytes;
    if (dst->device == src->device)
        FFX_CUDA(cudaMemcpyAsync(dst->store, src->store, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    else
        FFX_CUDA(cudaMemcpyPeerAsync(dst->store, dst->device, src->store, src->device, bytes, dst->stream));
    FFX_CUDA(cudaStreamSynchronize(dst->stream));
    dst->num_rows = src->num_rows;
    return FFX_OK;
}

int ffx_index_stage_rows(ffx_index *idx, int64_t row0, int64_t nrows, const void *rows,
                         int src_on_device) {
    NvtxRange nvtx_range("ffx_index_stage_rows");
    if (!idx || (!rows && nrows > 0) || row0 < 0 || nrows < 0)
        return fail(FFX_ERR_INVALID, "ffx_index_stage_rows: bad arguments");
    if (row0 + nrows > idx->capacity)
        return fail(FFX_ERR_INVALID, "ffx_index_stage_rows: rows [%lld, %lld) exceed capacity %lld",
                    static_cast<long long>(row0), static_cast<long long>(row0 + nrows),
                    static_cast<long long>(idx->capacity));
    if (nrows == 0) return FFX_OK;
    FFX_TRY(bind(idx));
    if (src_on_device) {
        // the caller's producer ran on some other stream: order after everything queued so far
        FFX_CUDA(cudaDeviceSynchronize());
        FFX_TRY(launch_to_store(idx, row0, nrows, rows));
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
    } else {
        FFX_TRY(ensure_staging(idx));
        const int64_t rows_per_buf = std::max<int64_t>(1, kStageBytes / static_cast<int64_t>(idx->row_bytes));
        if (static_cast<int64_t>(idx->row_bytes) > kStageBytes)
            return fail(FFX_ERR_UNSUPPORTED, "row of %zu bytes exceeds the staging buffer", idx->row_bytes);
        const bool permute = transforms_rows(idx);
        int b = 0;
        for (int64_t r = 0; r < nrows; r += rows_per_buf, b ^= 1) {
            const int64_t nr = std::min(rows_per_buf, nrows - r);
            const size_t bytes = static_cast<size_t>(nr) * idx->src_row_bytes;
            FFX_CUDA(cudaEventSynchronize(idx->done[b]));  // buffer b free again
            parallel_memcpy(idx->pinned[b], static_cast<const char *>(rows) + static_cast<size_t>(r) * idx->src_row_bytes, bytes);
            if (permute) {
                FFX_CUDA(cudaMemcpyAsync(idx->landing[b], idx->pinned[b], bytes,
                                         cudaMemcpyHostToDevice, idx->stream));
                FFX_TRY(launch_to_store(idx, row0 + r, nr, idx->landing[b]));
            } else {
                FFX_CUDA(cudaMemcpyAsync(static_cast<char *>(idx->store) +
                                             static_cast<size_t>(row0 + r) * idx->row_bytes,
                                         idx->pinned[b], bytes, cudaMemcpyHostToDevice, idx->stream));
            }
            FFX_CUDA(cudaEventRecord(idx->done[b], idx->stream));
        }
        // the caller's rows have all been copied out: return now, the H2D + permute tail overlaps the
        // caller's next chunk; whoever touches the store next settles the stream first
        idx->stage_pending = true;
    }
    idx->num_rows = std::max(idx->num_rows, row0 + nrows);
    return FFX_OK;
}

int ffx_index_read_rows(ffx_index *idx, const int64_t *rows, int64_t n, void *out) {
    if (!idx || n < 0 || (n > 0 && (!rows || !out)))
        return fail(FFX_ERR_INVALID, "ffx_index_read_rows: bad arguments");
    for (int64_t i = 0; i < n; i++)
        if (rows[i] < 0 || rows[i] >= idx->num_rows)
            return fail(FFX_ERR_INVALID, "ffx_index_read_rows: row %lld out of range",
                        static_cast<long long>(rows[i]));
    if (n == 0) return FFX_OK;
    FFX_TRY(bind(idx));
    FFX_TRY(settle(idx));
    const int64_t per = std::max<int64_t>(1, kStageBytes / static_cast<int64_t>(idx->row_bytes));
    const size_t chunk_rows = static_cast<size_t>(std::min(per, n));
    const size_t ids_bytes = (chunk_rows * 8 + 255) & ~static_cast<size_t>(255);
    FFX_TRY(scratch_reserve(idx->work, ids_bytes + chunk_rows * idx->row_bytes));
    const bool padded = idx->row_kind == FFX_ROWS_F32 && !idx->plan.cpl && idx->row_bytes != idx->src_row_bytes;
    int64_t *d_rows = static_cast<int64_t *>(idx->work.p);
    char *d_out = static_cast<char *>(idx->work.p) + ids_bytes;
    for (int64_t r = 0; r < n; r += per) {
        const int64_t nr = std::min(per, n - r);
        FFX_CUDA(cudaMemcpyAsync(d_rows, rows + r, static_cast<size_t>(nr) * 8,
                                 cudaMemcpyHostToDevice, idx->stream));
        if (padded) {
            ffx::ffx_pad_rows_kernel<<<permute_grid(nr * idx->any.stride, idx->sm_count), 256, 0, idx->stream>>>(
                reinterpret_cast<float *>(d_out), static_cast<const float *>(idx->store), nr, static_cast<int>(idx->dim),
                idx->any.stride, 0, d_rows);
        } else if (idx->row_kind == FFX_ROWS_F32) {
            ffx::ffx_permute_rows_kernel<<<permute_grid(nr * idx->dim, idx->sm_count), 256, 0,
                                           idx->stream>>>(
                reinterpret_cast<float *>(d_out), static_cast<const float *>(idx->store), nr,
                static_cast<int>(idx->dim), idx->plan.cpl, idx->plan.steps, idx->plan.lanes, 0, d_rows);
        } else {
            ffx::ffx_gather_bytes_kernel<<<permute_grid(nr * idx->dim, idx->sm_count), 256, 0,
                                           idx->stream>>>(
                reinterpret_cast<uint8_t *>(d_out), static_cast<const uint8_t *>(idx->store), nr,
                static_cast<int64_t>(idx->row_bytes), d_rows);
        }
        g_launches++;
        FFX_CUDA(cudaGetLastError());
        FFX_CUDA(cudaMemcpyAsync(static_cast<char *>(out) + static_cast<size_t>(r) * idx->src_row_bytes,
                                 d_out, static_cast<size_t>(nr) * idx->src_row_bytes,
                                 cudaMemcpyDeviceToHost, idx->stream));
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
    }
    return FFX_OK;
}

int64_t ffx_index_num_rows(const ffx_index *idx) { return idx ? idx->num_rows : -1; }
int64_t ffx_index_capacity(const ffx_index *idx) { return idx ? idx->capacity : -1; }
int64_t ffx_index_dim(const ffx_index *idx) { return idx ? idx->dim : -1; }
int ffx_index_has_fast_path(const ffx_index *idx) {
    return idx && idx->row_kind == FFX_ROWS_F32 && (idx->plan.cpl != 0 || idx->any.valid);
}

int ffx_index_set_docs(ffx_index *idx, int64_t n_docs, const int64_t *doc_off,
                       const int64_t *doc_rows) {
    NvtxRange nvtx_range("ffx_index_set_docs");
    if (!idx || n_docs < 0 || (n_docs > 0 && !doc_off))
        return fail(FFX_ERR_INVALID, "ffx_index_set_docs: bad arguments");
    if (n_docs > 0x7fffffffll) return fail(FFX_ERR_INVALID, "ffx_index_set_docs: too many documents");
    FFX_TRY(bind(idx));
    FFX_CUDA(cudaDeviceSynchronize());
    cudaFree(idx->doc_span);
    cudaFree(idx->doc_rows);
    idx->doc_span = nullptr;
    idx->doc_rows = nullptr;
    idx->n_docs = 0;
    idx->indirect = 0;
    idx->max_doc_rows = 1;
    if (n_docs == 0) return FFX_OK;

    const int64_t total = doc_off[n_docs];
    if (doc_off[0] != 0 || total < 0 || total > 0xffffffffll)
        return fail(FFX_ERR_INVALID, "ffx_index_set_docs: bad offsets");
    bool contiguous = true;
    for (int64_t d = 0; d < n_docs; d++) {
        const int64_t b = doc_off[d], e = doc_off[d + 1];
        if (e <= b) return fail(FFX_ERR_INVALID, "ffx_index_set_docs: document %lld has no rows",
                                static_cast<long long>(d));
        if (doc_rows) {
            for (int64_t i = b; i < e; i++) {
                if (doc_rows[i] < 0 || doc_rows[i] >= idx->capacity)
                    return fail(FFX_ERR_INVALID, "ffx_index_set_docs: row %lld out of range",
                                static_cast<long long>(doc_rows[i]));
                if (i > b && doc_rows[i] != doc_rows[i - 1] + 1) contiguous = false;
            }
        } else if (e > idx->capacity) {
            return fail(FFX_ERR_INVALID, "ffx_index_set_docs: offsets exceed the row store");
        }
    }
    std::vector<uint2> span(static_cast<size_t>(n_docs));
    for (int64_t d = 0; d < n_docs; d++) {
        const int64_t b = doc_off[d];
        const uint32_t cnt = static_cast<uint32_t>(doc_off[d + 1] - b);
        idx->max_doc_rows = std::max<int64_t>(idx->max_doc_rows, cnt);
        const uint32_t first = contiguous ? static_cast<uint32_t>(doc_rows ? doc_rows[b] : b)
                                          : static_cast<uint32_t>(b);
        span[static_cast<size_t>(d)] = make_uint2(first, cnt);
    }
    FFX_CUDA(cudaMalloc(&idx->doc_span, span.size() * sizeof(uint2)));
    FFX_CUDA(cudaMemcpy(idx->doc_span, span.data(), span.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    if (!contiguous) {
        std::vector<int32_t> rows32(static_cast<size_t>(total));
        for (int64_t i = 0; i < total; i++)
            rows32[static_cast<size_t>(i)] = static_cast<int32_t>(static_cast<uint32_t>(doc_rows[i]));
        FFX_CUDA(cudaMalloc(&idx->doc_rows, rows32.size() * 4 + 256));
        FFX_CUDA(cudaMemcpy(idx->doc_rows, rows32.data(), rows32.size() * 4, cudaMemcpyHostToDevice));
        idx->indirect = 1;
    }
    idx->n_docs = n_docs;
    return FFX_OK;
}

int ffx_index_set_shard(ffx_index *idx, int64_t doc_base, int64_t global_docs, int64_t row_base,
                        int64_t global_rows) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_index_set_shard: NULL index");
    if (global_docs == 0 && global_rows == 0) {
        idx->sharded = false;
        return FFX_OK;
    }
    if (doc_base < 0 || row_base < 0 || global_docs > 0x7fffffffll || global_rows > 0xffffffffll ||
        doc_base + idx->n_docs > global_docs || row_base + idx->num_rows > global_rows)
        return fail(FFX_ERR_INVALID, "ffx_index_set_shard: shard [%lld,+%lld) docs / [%lld,+%lld) rows does "
                    "not fit the corpus (%lld docs, %lld rows)", static_cast<long long>(doc_base),
                    static_cast<long long>(idx->n_docs), static_cast<long long>(row_base),
                    static_cast<long long>(idx->num_rows), static_cast<long long>(global_docs),
                    static_cast<long long>(global_rows));
    idx->sharded = true;
    idx->doc_base = doc_base;
    idx->global_docs = global_docs;
    idx->row_base = row_base;
    idx->global_rows = global_rows;
    return FFX_OK;
}

int ffx_index_set_topk_scatter(ffx_index *idx, int world, int rank, int64_t stride, const int64_t *bounds,
                               void *const *peer_score, void *const *peer_pos) {
    constexpr int kMaxWorld = 64;
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_index_set_topk_scatter: NULL index");
    if (world == 0) {
        idx->sc_world = 0;
        return FFX_OK;
    }
    if (world < 0 || world > kMaxWorld || rank < 0 || rank >= world || stride <= 0 || !bounds || !peer_score ||
        !peer_pos)
        return fail(FFX_ERR_INVALID, "ffx_index_set_topk_scatter: bad arguments");
    for (int o = 0; o < world; o++)
        if (bounds[o + 1] < bounds[o] || bounds[o + 1] - bounds[o] > stride || !peer_score[o] || !peer_pos[o])
            return fail(FFX_ERR_INVALID, "ffx_index_set_topk_scatter: bad bounds / buffers for owner %d", o);
    FFX_TRY(bind(idx));
    if (!idx->sc_bounds) {
        FFX_CUDA(cudaMalloc(&idx->sc_bounds, (kMaxWorld + 1) * sizeof(int64_t)));
        FFX_CUDA(cudaMalloc(&idx->sc_score, kMaxWorld * sizeof(void *)));
        FFX_CUDA(cudaMalloc(&idx->sc_pos, kMaxWorld * sizeof(void *)));
    }
    idx->sc_bounds_h.assign(bounds, bounds + world + 1);
    idx->sc_score_h.assign(peer_score, peer_score + world);
    idx->sc_pos_h.assign(peer_pos, peer_pos + world);
    idx->sc_world = world;
    idx->sc_rank = rank;
    idx->sc_stride = stride;
    idx->sc_dirty = true;
    return FFX_OK;
}

int ffx_index_set_pq(ffx_index *idx, int M, int Ks, int Ds, const float *codewords, const float *R) {
    if (!idx || !codewords || M <= 0 || Ks <= 0 || Ds <= 0)
        return fail(FFX_ERR_INVALID, "ffx_index_set_pq: bad arguments");
    if (idx->row_kind != FFX_ROWS_PQ_U8)
        return fail(FFX_ERR_STATE, "ffx_index_set_pq: index does not hold PQ codes");
    if (M != idx->dim) return fail(FFX_ERR_INVALID, "ffx_index_set_pq: M=%d but rows have %lld codes",
                                   M, static_cast<long long>(idx->dim));
    if (Ks > 256) return fail(FFX_ERR_UNSUPPORTED, "ffx_index_set_pq: Ks=%d > 256 (uint8 codes)", Ks);
    if (static_cast<size_t>(M) * Ks * 4 > 200 * 1024)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_index_set_pq: LUT of %d x %d floats exceeds shared memory", M, Ks);
    FFX_TRY(bind(idx));
    FFX_CUDA(cudaDeviceSynchronize());
    cudaFree(idx->codewords);
    cudaFree(idx->R);
    cudaFree(idx->cw_t);
    cudaFree(idx->cw_x);
    idx->codewords = idx->R = idx->cw_t = idx->cw_x = nullptr;
    const size_t cw = static_cast<size_t>(M) * Ks * Ds * 4;
    FFX_CUDA(cudaMalloc(&idx->codewords, cw));
    FFX_CUDA(cudaMemcpy(idx->codewords, codewords, cw, cudaMemcpyHostToDevice));
    if (R) {
        const size_t D = static_cast<size_t>(M) * Ds;
        FFX_CUDA(cudaMalloc(&idx->R, D * D * 4));
        FFX_CUDA(cudaMemcpy(idx->R, R, D * D * 4, cudaMemcpyHostToDevice));
    }
    idx->M = M;
    idx->Ks = Ks;
    idx->Ds = Ds;
    // warp-per-row kernel with the transposed, conflict-free table: M = 64..128 in whole words
    if (M % 4 == 0 && M / 4 >= 16 && M / 4 <= 32 && ffx::adc_warp_smem_bytes(Ks, 0) <= kSmemBudget) {
        const size_t n = static_cast<size_t>(4) * Ks * 32 * Ds;
        FFX_CUDA(cudaMalloc(&idx->cw_t, n * 4));
        ffx::ffx_adc_transpose_codewords_kernel<<<permute_grid(static_cast<int64_t>(n), idx->sm_count), 256, 0,
                                                  idx->stream>>>(idx->codewords, M, Ks, Ds, idx->cw_t);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
    }
    if (M % 32 == 0 && M <= 128 && ffx::adc_xor_smem_bytes(M, Ks, 0) <= kSmemBudget) {
        const size_t n = static_cast<size_t>(M) * Ks * Ds;
        FFX_CUDA(cudaMalloc(&idx->cw_x, n * 4));
        ffx::ffx_adc_xor_codewords_kernel<<<permute_grid(static_cast<int64_t>(n), idx->sm_count), 256, 0,
                                            idx->stream>>>(idx->codewords, M, Ks, Ds, idx->cw_x);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
        FFX_CUDA(cudaStreamSynchronize(idx->stream));
    }
    return FFX_OK;
}

// ---- the hot path ----------------------------------------------------------------------
// `qeff_buf` (device, [nq, D], optional): where the OPQ-rotated queries of this launch go.  The
// default is the index's scratch, which launches on different streams would share — the
// pipelined host path gives every query chunk its own slice instead.
static int rerank_impl(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
                       const int32_t *cand, const float *lex, double alpha, int k, int64_t max_cand,
                       float *out_ff, float *out_int, float *out_topk_score, int32_t *out_topk_pos,
                       void *stream, float *qeff_buf, float *lut_buf = nullptr) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_rerank: NULL index");
    if (mode < FFX_MODE_PASSAGE || mode > FFX_MODE_AVEP)
        return fail(FFX_ERR_INVALID, "ffx_rerank: unknown mode %d", mode);
    if (nq < 0 || k < 0 || max_cand < 0 || nq > 0x7fffffffll || max_cand > (1ll << 30))
        return fail(FFX_ERR_INVALID, "ffx_rerank: bad sizes");
    if (nq == 0) return FFX_OK;
    if (!qvecs || !q_off || (!cand && max_cand > 0))
        return fail(FFX_ERR_INVALID, "ffx_rerank: NULL input");
    if (k > 0 && (!out_topk_score || !out_topk_pos) && idx->sc_world == 0)
        return fail(FFX_ERR_INVALID, "ffx_rerank: k > 0 needs top-k outputs");
    if (mode != FFX_MODE_PASSAGE && idx->n_docs == 0 && max_cand > 0)
        return fail(FFX_ERR_STATE, "ffx_rerank: document modes need ffx_index_set_docs first");
    if (idx->row_kind == FFX_ROWS_PQ_U8 && !idx->codewords)
        return fail(FFX_ERR_STATE, "ffx_rerank: PQ index without codebooks (ffx_index_set_pq)");
    FFX_TRY(bind(idx));
    FFX_TRY(settle(idx));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int cpad = next_pow2(std::max<int64_t>(max_cand, 1));
    const bool pq = idx->row_kind == FFX_ROWS_PQ_U8;
    const bool fast = !pq && idx->plan.cpl != 0;
    const bool any = !pq && !fast && idx->any.valid;  // the numpy tree as data (ffx_score_any_kernel)
    const int64_t D = pq ? static_cast<int64_t>(idx->M) * idx->Ds : idx->dim;
    const float alpha32 = static_cast<float>(alpha);
    const float beta32 = static_cast<float>(1.0 - alpha);

    const bool psg_mode = mode == FFX_MODE_PASSAGE;
    const uint32_t count = static_cast<uint32_t>(psg_mode ? idx->num_rows : idx->n_docs);
    const uint32_t base = idx->sharded ? static_cast<uint32_t>(psg_mode ? idx->row_base : idx->doc_base) : 0u;
    const uint32_t limit = idx->sharded ? static_cast<uint32_t>(psg_mode ? idx->global_rows : idx->global_docs)
                                        : count;

    // one CTA per query with the top-k fused needs enough queries to fill the machine
    const int64_t slots = static_cast<int64_t>(idx->sm_count) * 2;
    // nothing to score (every list empty): the separate top-k pass still writes the (-inf, -1) padding
    bool fuse = max_cand > 0 && will_fuse(idx, nq, k, cpad);
    const bool few_pairs = nq * max_cand < static_cast<int64_t>(idx->sm_count) * 8 * 64;
    ScorePlan sp_any{};
    if (any) {
        sp_any = plan_score_any(idx->any, mode, fuse, cpad, idx->sharded, few_pairs);
        if (fuse && !sp_any.tma) {  // the keys do not fit beside the ring: separate top-k pass
            fuse = false;
            sp_any = plan_score_any(idx->any, mode, false, cpad, idx->sharded, few_pairs);
        }
    }

    // scratch plan: [scores n_total?][keys nq*cpad?][qeff nq*D?]
    // a separate top-k pass reads out_int when the caller asked for it; a shard must not (pairs
    // of other shards stay untouched there), it ranks a scratch copy with NaN = "not mine"
    const bool need_scores = k > 0 && !fuse && (!out_int || idx->sharded);
    const bool need_gkeys = k > 0 && !fuse && cpad > ffx::kMaxFusedCand;
    size_t off_scores = 0, off_keys = 0, off_qeff = 0, total = 0;
    int64_t n_total = 0;
    if (need_scores) {
        // the pair count lives in device memory (q_off[nq]); bound it by nq * max_cand
        n_total = nq * max_cand;
        off_scores = total;
        total += (static_cast<size_t>(n_total) * 4 + 255) & ~static_cast<size_t>(255);
    }
    if (need_gkeys) {
        off_keys = total;
        total += static_cast<size_t>(nq) * cpad * 8;
    }
    if (pq && idx->R && !qeff_buf) {
        off_qeff = total;
        total += static_cast<size_t>(nq) * D * 4;
    }
    // XOR-swizzled ADC kernel: the tables of all queries are built by one launch beforehand
    // (ffx_adc_xor_lut_kernel) when they fit the scratch budget, else inside the scoring kernel
    size_t off_lut = 0;
    const size_t lut_floats = pq ? static_cast<size_t>(idx->M) * idx->Ks : 0;
    const bool lut_ahead = pq && max_cand > 0 && adc_kind(idx) == 3 && (lut_floats * 4) % 16 == 0 && g_tune.adc_lut != 1 &&
                           (lut_buf || static_cast<size_t>(nq) * lut_floats * 4 <= kAdcLutScratch);
    if (lut_ahead && !lut_buf) {
        total = (total + 255) & ~static_cast<size_t>(255);
        off_lut = total;
        total += static_cast<size_t>(nq) * lut_floats * 4;
    }
    if (total) FFX_TRY(scratch_reserve(idx->work, total));
    char *work = static_cast<char *>(idx->work.p);
    float *rank_scores = need_scores ? reinterpret_cast<float *>(work + off_scores) : nullptr;
    const float *topk_src = rank_scores ? rank_scores : out_int;  // input of a separate top-k pass

    // tiles: split a query over several CTAs when there are few queries
    // (short rows: the ring slot of one warp step, 32 / (lanes per row) rows)
    // (options kernel = 1 / 2 pick the register-staged / TMA-staged whole-warp kernels where they exist)
    const bool packed = ffx_packed_cpl(idx->plan) != 0 && !(g_tune.kernel != 0 && idx->plan.lanes == 32);
    ScorePlan sp = fast ? plan_score(mode, fuse, cpad,
                                     static_cast<int>(idx->dim) * 4 * (packed ? ffx_rows_per_step(idx->plan) : 1),
                                     idx->sharded, few_pairs, packed) : sp_any;
    // short-row kernel: positions inside a batch's flattened row sequence are 32-bit
    sp.batch = static_cast<int>(std::min<int64_t>(sp.batch, std::max<int64_t>(1, 0x7fffffffll / idx->max_doc_rows)));
    int tiles = 1, tile = static_cast<int>(std::max<int64_t>(max_cand, 1));
    if (!fuse) {
        // enough CTAs for ~2 waves at the plan's occupancy, but a tile keeps every warp of its
        // CTA busy with at least one candidate batch
        const int64_t want = slots * 4;
        const int64_t grain = ((fast || any) && sp.tma) ? static_cast<int64_t>(sp.warps) * sp.batch : 32;
        tiles = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((want + nq - 1) / nq,
                                                                        (max_cand + grain - 1) / grain)));
        tile = static_cast<int>((std::max<int64_t>(max_cand, 1) + tiles - 1) / tiles);
        tile = (tile + 31) & ~31;
        tiles = static_cast<int>((std::max<int64_t>(max_cand, 1) + tile - 1) / tile);
    }
    if (nq * tiles > 0x7fffffffll) return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank: grid too large");

    if (max_cand > 0) {
        if (pq) {
            const float *qeff = qvecs;
            if (idx->R) {
                float *qe = qeff_buf ? qeff_buf : reinterpret_cast<float *>(work + off_qeff);
                const size_t rot_smem = static_cast<size_t>(D) * 8 * 4;
                if (D % 16 == 0 && nq >= 64 && (reinterpret_cast<uintptr_t>(qvecs) & 15) == 0) {
                    const dim3 grid_r(static_cast<unsigned>((D + ffx::kRotTile - 1) / ffx::kRotTile),
                                      static_cast<unsigned>((nq + ffx::kRotTile - 1) / ffx::kRotTile));
                    ffx::ffx_rotate_queries_tiled_kernel<<<grid_r, 256, 0, st>>>(qvecs, idx->R, static_cast<int>(D), nq, qe);
                } else if (rot_smem <= 48 * 1024) {
                    ffx::ffx_rotate_queries8_kernel<<<static_cast<unsigned>((nq + 7) / 8), 256, rot_smem, st>>>(
                        qvecs, idx->R, static_cast<int>(D), nq, qe);
                } else {
                    ffx::ffx_rotate_queries_kernel<<<static_cast<unsigned>(nq), 256,
                                                     static_cast<size_t>(D) * 4, st>>>(
                        qvecs, idx->R, static_cast<int>(D), qe);
                }
                g_launches++;
                FFX_CUDA(cudaGetLastError());
                qeff = qe;
            }
            ffx::AdcArgs a{};
            a.codes = static_cast<const uint8_t *>(idx->store);
            a.codewords = idx->codewords;
            a.qeff = qeff;
            a.M = idx->M;
            a.Ks = idx->Ks;
            a.Ds = idx->Ds;
            a.doc_span = idx->doc_span;
            a.doc_rows = idx->doc_rows;
            a.indirect = idx->indirect;
            a.mode = mode;
            a.q_off = q_off;
            a.cand = cand;
            a.lex = lex;
            a.alpha = alpha32;
            a.beta = beta32;
            a.out_ff = out_ff;
            a.out_int = out_int;
            a.rank_scores = rank_scores;
            a.tiles_per_query = tiles;
            a.tile = tile;
            a.limit = limit;
            a.base = base;
            a.count = count;
            a.err = idx->err_flag;
            const int kind = adc_kind(idx);
            if (kind == 3) {
                ffx::AdcWarpArgs w{};
                w.base = a;
                w.cw_t = idx->cw_x;
                w.k = k;
                w.cpad = cpad;
                w.topk_score = out_topk_score;
                w.topk_pos = out_topk_pos;
                if (lut_ahead) {
                    float *lut = lut_buf ? lut_buf : reinterpret_cast<float *>(work + off_lut);
                    const unsigned q_blocks = static_cast<unsigned>((nq + ffx::kAdcLutQueries - 1) / ffx::kAdcLutQueries);
                    const bool tiled = g_tune.adc_lut != 2 && idx->Ks % 32 == 0 && (idx->Ds == 4 || idx->Ds == 8 || idx->Ds == 16) &&
                                       (reinterpret_cast<uintptr_t>(qeff) & 15) == 0 && q_blocks <= 65535u;
                    if (tiled) {
                        // a 32 x 32 tile of one table per CTA, query slices broadcast
                        const dim3 grid_t(static_cast<unsigned>(idx->M / 32 * (idx->Ks / 32)), q_blocks);
                        switch (idx->Ds) {
                            case 4: ffx::ffx_adc_xor_lut_tile_kernel<4><<<grid_t, ffx::kAdcLutTileThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, nq, lut); break;
                            case 8: ffx::ffx_adc_xor_lut_tile_kernel<8><<<grid_t, ffx::kAdcLutTileThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, nq, lut); break;
                            default: ffx::ffx_adc_xor_lut_tile_kernel<16><<<grid_t, ffx::kAdcLutTileThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, nq, lut); break;
                        }
                    } else {
                        const dim3 grid_l(static_cast<unsigned>((lut_floats + ffx::kAdcLutThreads - 1) / ffx::kAdcLutThreads),
                                          static_cast<unsigned>((nq + ffx::kAdcLutQueries - 1) / ffx::kAdcLutQueries));
                        switch (idx->Ds) {
                            case 4: ffx::ffx_adc_xor_lut_kernel<4><<<grid_l, ffx::kAdcLutThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, idx->Ds, nq, lut); break;
                            case 8: ffx::ffx_adc_xor_lut_kernel<8><<<grid_l, ffx::kAdcLutThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, idx->Ds, nq, lut); break;
                            case 16: ffx::ffx_adc_xor_lut_kernel<16><<<grid_l, ffx::kAdcLutThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, idx->Ds, nq, lut); break;
                            default: ffx::ffx_adc_xor_lut_kernel<0><<<grid_l, ffx::kAdcLutThreads, 0, st>>>(idx->cw_x, qeff, idx->M, idx->Ks, idx->Ds, nq, lut); break;
                        }
                    }
                    g_launches++;
                    FFX_CUDA(cudaGetLastError());
                    w.lut = lut;
                }
                const size_t smem = ffx::adc_xor_smem_bytes(idx->M, idx->Ks, fuse ? cpad : 0);
                const unsigned grid = static_cast<unsigned>(fuse ? nq : nq * tiles);
                switch (idx->M / 32) {
                    case 1: FFX_TRY(launch_adc_xor<1>(w, fuse, grid, smem, st)); break;
                    case 2: FFX_TRY(launch_adc_xor<2>(w, fuse, grid, smem, st)); break;
                    case 3: FFX_TRY(launch_adc_xor<3>(w, fuse, grid, smem, st)); break;
                    default: FFX_TRY(launch_adc_xor<4>(w, fuse, grid, smem, st)); break;
                }
            } else if (kind == 2) {
                ffx::AdcWarpArgs w{};
                w.base = a;
                w.cw_t = idx->cw_t;
                w.k = k;
                w.cpad = cpad;
                w.topk_score = out_topk_score;
                w.topk_pos = out_topk_pos;
                const size_t smem = ffx::adc_warp_smem_bytes(idx->Ks, fuse ? cpad : 0);
                const bool ind = idx->indirect && mode != FFX_MODE_PASSAGE;
                const unsigned grid = static_cast<unsigned>(fuse ? nq : nq * tiles);
                if (fuse && ind) FFX_TRY((launch_adc_warp<true, true>(w, grid, smem, st)));
                else if (fuse) FFX_TRY((launch_adc_warp<true, false>(w, grid, smem, st)));
                else if (ind) FFX_TRY((launch_adc_warp<false, true>(w, grid, smem, st)));
                else FFX_TRY((launch_adc_warp<false, false>(w, grid, smem, st)));
            } else {
                const size_t smem = static_cast<size_t>(idx->M) * idx->Ks * 4;
                FFX_TRY(launch_adc(a, static_cast<unsigned>(nq * tiles), smem, st));
            }
        } else {
            ffx::ScoreArgs a{};
            a.vectors = static_cast<const float *>(idx->store);
            a.doc_span = idx->doc_span;
            a.doc_rows = idx->doc_rows;
            a.indirect = idx->indirect;
            a.mode = mode;
            a.dim = idx->dim;
            a.qvecs = qvecs;
            a.q_off = q_off;
            a.cand = cand;
            a.lex = lex;
            a.alpha = alpha32;
            a.beta = beta32;
            a.k = k;
            a.out_ff = out_ff;
            a.out_int = out_int;
            a.rank_scores = rank_scores;
            a.topk_score = out_topk_score;
            a.topk_pos = out_topk_pos;
            a.tiles_per_query = tiles;
            a.tile = tile;
            a.cpad = cpad;
            a.limit = limit;
            a.base = base;
            a.count = count;
            a.err = idx->err_flag;
            if (idx->sc_world) {
                // the fused exchange lives in the epilogue of the fused TMA-staged kernel only
                if (!((fast || any) && fuse && sp.tma))
                    return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank: a scatter plan needs the fused fp32 kernel "
                                "(lane-major dimension, k > 0, >= 2 queries per SM, <= %d candidates per query)",
                                ffx::kMaxFusedCand);
                if (idx->sc_dirty) {  // pageable sources: the calls return once the data is staged
                    const size_t w = static_cast<size_t>(idx->sc_world);
                    FFX_CUDA(cudaMemcpyAsync(idx->sc_bounds, idx->sc_bounds_h.data(), (w + 1) * sizeof(int64_t),
                                             cudaMemcpyHostToDevice, st));
                    FFX_CUDA(cudaMemcpyAsync(idx->sc_score, idx->sc_score_h.data(), w * sizeof(void *),
                                             cudaMemcpyHostToDevice, st));
                    FFX_CUDA(cudaMemcpyAsync(idx->sc_pos, idx->sc_pos_h.data(), w * sizeof(void *),
                                             cudaMemcpyHostToDevice, st));
                    idx->sc_dirty = false;
                }
                a.sc_world = idx->sc_world;
                a.sc_rank = idx->sc_rank;
                a.sc_stride = idx->sc_stride;
                a.sc_bounds = idx->sc_bounds;
                a.sc_score = reinterpret_cast<float *const *>(idx->sc_score);
                a.sc_pos = reinterpret_cast<int32_t *const *>(idx->sc_pos);
            }
            if (fast) {
                FFX_TRY(dispatch_score(idx->plan, sp, a, fuse, static_cast<int>(nq * tiles), st));
            } else if (any) {
                if (!sp.tma) return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank: rows of %lld floats do not fit the row rings",
                                         static_cast<long long>(idx->dim));
                FFX_TRY(dispatch_score_any(idx->any, sp, a, fuse, static_cast<int>(nq * tiles), st));
            } else {
                // generic exact kernel: one thread per pair over the [0, nq*max_cand) bound;
                // the kernel reads the true pair count from q_off[nq]
                const int64_t bound = nq * max_cand;
                ffx::ffx_score_generic_kernel<<<static_cast<unsigned>((bound + 127) / 128), 128, 0, st>>>(
                    a, nq, bound);
                note_kernel(ffx::ffx_score_generic_kernel);
                g_launches++;
                FFX_CUDA(cudaGetLastError());
            }
        }
    }
    if (k > 0 && !fuse)
        FFX_TRY(launch_topk(topk_src, rank_scores != nullptr, nullptr, 0.f, 0.f, nullptr, q_off, nq, k, cpad,
                            idx->work, off_keys,
                            out_topk_score, out_topk_pos, st));
    return FFX_OK;
}

int ffx_rerank(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
               const int32_t *cand, const float *lex, double alpha, int k, int64_t max_cand,
               float *out_ff, float *out_int, float *out_topk_score, int32_t *out_topk_pos,
               void *stream) {
    NvtxRange nvtx_range("ffx_rerank");
    return rerank_impl(idx, mode, qvecs, nq, q_off, cand, lex, alpha, k, max_cand, out_ff, out_int,
                       out_topk_score, out_topk_pos, stream, nullptr);
}

int ffx_rerank_host(ffx_index *idx, int mode, const float *qvecs, int64_t nq,
                    const int64_t *q_off, const int32_t *cand, const float *lex, double alpha,
                    int k, float *out_ff, float *out_int, float *out_topk_score,
                    int32_t *out_topk_pos) {
    NvtxRange nvtx_range("ffx_rerank_host");
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_rerank_host: NULL index");
    if (nq < 0) return fail(FFX_ERR_INVALID, "ffx_rerank_host: nq < 0");
    if (nq == 0) return FFX_OK;
    if (!qvecs || !q_off) return fail(FFX_ERR_INVALID, "ffx_rerank_host: NULL input");
    if (q_off[0] != 0) return fail(FFX_ERR_INVALID, "ffx_rerank_host: q_off[0] must be 0");
    int64_t max_cand = 0;
    for (int64_t q = 0; q < nq; q++) {
        const int64_t c = q_off[q + 1] - q_off[q];
        if (c < 0) return fail(FFX_ERR_INVALID, "ffx_rerank_host: q_off not monotone");
        max_cand = std::max(max_cand, c);
    }
    const int64_t n = q_off[nq];
    if (n > 0 && !cand) return fail(FFX_ERR_INVALID, "ffx_rerank_host: NULL candidates");
    if (k > 0 && (!out_topk_score || !out_topk_pos))
        return fail(FFX_ERR_INVALID, "ffx_rerank_host: k > 0 needs top-k outputs");
    const bool pq = idx->row_kind == FFX_ROWS_PQ_U8;
    const int64_t D = pq ? static_cast<int64_t>(idx->M) * idx->Ds : idx->dim;
    FFX_TRY(bind(idx));

    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_q = pad(static_cast<size_t>(nq) * D * 4), b_off = pad(static_cast<size_t>(nq + 1) * 8);
    const size_t b_n = pad(static_cast<size_t>(n) * 4), b_k = pad(static_cast<size_t>(nq) * k * 4);
    const bool rotated = pq && idx->R;
    // tables of the XOR-swizzled ADC kernel, built ahead per chunk: every chunk needs its own slice
    const size_t lut_floats = pq && adc_kind(idx) == 3 ? static_cast<size_t>(idx->M) * idx->Ks : 0;
    const size_t b_lut = lut_floats && static_cast<size_t>(nq) * lut_floats * 4 <= kAdcLutScratch
                             ? pad(static_cast<size_t>(nq) * lut_floats * 4) : 0;
    size_t total = b_q + (rotated ? b_q : 0) + b_off + b_n /*cand*/ + (lex ? b_n : 0) + (out_ff ? b_n : 0) +
                   (out_int ? b_n : 0) + (k > 0 ? 2 * b_k : 0) + b_lut;
    FFX_TRY(scratch_reserve(idx->hostio, total));
    char *p = static_cast<char *>(idx->hostio.p);
    auto take = [&](size_t b) { char *r = p; p += b; return r; };
    float *d_q = reinterpret_cast<float *>(take(b_q));
    float *d_qe = rotated ? reinterpret_cast<float *>(take(b_q)) : nullptr;  // OPQ-rotated queries, per chunk
    int64_t *d_off = reinterpret_cast<int64_t *>(take(b_off));
    int32_t *d_cand = reinterpret_cast<int32_t *>(take(b_n));
    float *d_lex = lex ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_ff = out_ff ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_int = out_int ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_ts = k > 0 ? reinterpret_cast<float *>(take(b_k)) : nullptr;
    int32_t *d_tp = k > 0 ? reinterpret_cast<int32_t *>(take(b_k)) : nullptr;
    float *d_lut = b_lut ? reinterpret_cast<float *>(take(b_lut)) : nullptr;

    // Query chunks flow through H2D -> kernel -> D2H on separate streams, so the PCIe copies
    // of chunk i+1 / i-1 hide behind the kernel of chunk i.  Every chunk has its own slice of
    // the device mirrors (no buffer reuse, hence no hazards); two compute streams alternate
    // so one chunk's tail wave overlaps the next chunk's first wave.  Chunking only applies
    // when every chunk still takes the fused one-CTA-per-query kernel.
    const int cpad = next_pow2(std::max<int64_t>(max_cand, 1));
    const int64_t wave = static_cast<int64_t>(idx->sm_count) * 2;
    int64_t chunk_q = nq;
    // one wave of queries per chunk: only the first chunk's H2D and the last chunk's D2H (1/18 of the
    // C3 step each) are not hidden behind a kernel; measured e2e 70.7 / 69.6 / 68.8 ms at 4 / 2 / 1
    // waves against 68.1 ms with device-resident inputs
    const int64_t per_chunk = (g_tune.chunk_waves > 0 ? g_tune.chunk_waves : 1) * wave;
    if (n >= (1 << 20) && nq >= 2 * per_chunk && will_fuse(idx, per_chunk, k, cpad)) chunk_q = per_chunk;
    // whole chunks only; the remainder rides with the last one (a tail of a few queries would
    // fall below the fused kernel's query threshold)
    const int64_t n_chunks = std::max<int64_t>(1, nq / chunk_q);

    FFX_CUDA(cudaMemcpyAsync(d_off, q_off, static_cast<size_t>(nq + 1) * 8, cudaMemcpyHostToDevice, idx->s_h2d));
    size_t ev = 0;
    for (int64_t c = 0; c < n_chunks; c++) {
        const int64_t q0 = c * chunk_q, q1 = c + 1 == n_chunks ? nq : q0 + chunk_q, cq = q1 - q0;
        const int64_t p0 = q_off[q0], np_ = q_off[q1] - p0;
        cudaStream_t comp = idx->s_comp[c & 1];
        FFX_CUDA(cudaMemcpyAsync(d_q + q0 * D, qvecs + q0 * D, static_cast<size_t>(cq) * D * 4,
                                 cudaMemcpyHostToDevice, idx->s_h2d));
        if (np_ > 0) {
            FFX_CUDA(cudaMemcpyAsync(d_cand + p0, cand + p0, static_cast<size_t>(np_) * 4,
                                     cudaMemcpyHostToDevice, idx->s_h2d));
            if (lex)
                FFX_CUDA(cudaMemcpyAsync(d_lex + p0, lex + p0, static_cast<size_t>(np_) * 4,
                                         cudaMemcpyHostToDevice, idx->s_h2d));
        }
        cudaEvent_t in_ready = event_at(idx, ev++);
        FFX_CUDA(cudaEventRecord(in_ready, idx->s_h2d));
        FFX_CUDA(cudaStreamWaitEvent(comp, in_ready, 0));
        // q_off holds absolute pair offsets: pass the shifted offset pointer with the
        // unshifted per-pair arrays; per-query outputs are shifted to the chunk
        FFX_TRY(rerank_impl(idx, mode, d_q + q0 * D, cq, d_off + q0, d_cand, d_lex, alpha, k, max_cand,
                            d_ff, d_int, d_ts ? d_ts + q0 * k : nullptr, d_tp ? d_tp + q0 * k : nullptr,
                            comp, d_qe ? d_qe + q0 * D : nullptr, d_lut ? d_lut + q0 * lut_floats : nullptr));
        cudaEvent_t out_ready = event_at(idx, ev++);
        FFX_CUDA(cudaEventRecord(out_ready, comp));
        FFX_CUDA(cudaStreamWaitEvent(idx->s_d2h, out_ready, 0));
        if (np_ > 0) {
            if (out_ff)
                FFX_CUDA(cudaMemcpyAsync(out_ff + p0, d_ff + p0, static_cast<size_t>(np_) * 4,
                                         cudaMemcpyDeviceToHost, idx->s_d2h));
            if (out_int)
                FFX_CUDA(cudaMemcpyAsync(out_int + p0, d_int + p0, static_cast<size_t>(np_) * 4,
                                         cudaMemcpyDeviceToHost, idx->s_d2h));
        }
        if (k > 0) {
            FFX_CUDA(cudaMemcpyAsync(out_topk_score + q0 * k, d_ts + q0 * k, static_cast<size_t>(cq) * k * 4,
                                     cudaMemcpyDeviceToHost, idx->s_d2h));
            FFX_CUDA(cudaMemcpyAsync(out_topk_pos + q0 * k, d_tp + q0 * k, static_cast<size_t>(cq) * k * 4,
                                     cudaMemcpyDeviceToHost, idx->s_d2h));
        }
    }
    return take_error(idx, idx->s_d2h);
}

// The walk as a stream-ordered sequence of launches per depth (ffx_early_stop.cuh), for the index
// kinds the one-launch kernel does not cover.  Scores come from the index's ordinary kernels.
static int early_stop_walk(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
                           const int32_t *cand, const float *lex, double alpha, const ffx::EsPlan &es,
                           int64_t max_cand, float *out_ff, float *out_int, int32_t *out_scored, cudaStream_t st) {
    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t n_bound = static_cast<size_t>(nq) * static_cast<size_t>(max_cand);
    const size_t b_q = pad(static_cast<size_t>(nq) * 4), b_off = pad(static_cast<size_t>(nq + 1) * 8), b_n = pad(n_bound * 4);
    FFX_TRY(scratch_reserve(idx->es_work, 3 * b_q + b_off + 2 * b_n + (out_ff ? 0 : b_n) + (out_int ? 0 : b_n)));
    char *p = static_cast<char *>(idx->es_work.p);
    auto take = [&](size_t b) { char *r = p; p += b; return r; };
    ffx::EsWalk w{};
    w.q_off = q_off;
    w.cand = cand;
    w.lex = lex;
    w.alpha = static_cast<float>(alpha);
    w.beta = static_cast<float>(1.0 - alpha);
    w.cutoff = es.cutoff;
    w.done = reinterpret_cast<int32_t *>(take(b_q));
    w.active = reinterpret_cast<int32_t *>(take(b_q));
    w.take = reinterpret_cast<int32_t *>(take(b_q));
    w.part_off = reinterpret_cast<int64_t *>(take(b_off));
    w.cand_sub = reinterpret_cast<int32_t *>(take(b_n));
    w.ff_sub = reinterpret_cast<float *>(take(b_n));
    w.ff = out_ff ? out_ff : reinterpret_cast<float *>(take(b_n));
    w.inter = out_int ? out_int : reinterpret_cast<float *>(take(b_n));
    const unsigned per_query = static_cast<unsigned>(nq);
    const unsigned flat = static_cast<unsigned>(std::min<int64_t>((nq + 255) / 256, 1 << 16));
    ffx::es_init_kernel<<<flat, 256, 0, st>>>(w, nq);
    g_launches++;
    int prev = 0;
    for (int d = 0; d < es.n_depths; d++) {
        const int depth = es.depths[d];
        if (prev >= max_cand) break;  // every block is exhausted
        if (d > 0) {
            const size_t smem = static_cast<size_t>(next_pow2(std::min<int64_t>(prev, max_cand))) * 4;
            if (smem > 48 * 1024)
                FFX_CUDA(cudaFuncSetAttribute(ffx::es_criterion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(smem)));
            ffx::es_criterion_kernel<<<per_query, 256, smem, st>>>(w);
            g_launches++;
        }
        ffx::es_plan_kernel<<<1, 1024, 0, st>>>(w, nq, depth);
        ffx::es_gather_kernel<<<per_query, 128, 0, st>>>(w);
        g_launches += 2;
        FFX_CUDA(cudaGetLastError());
        const int64_t window = std::min<int64_t>(max_cand, depth) - std::min<int64_t>(max_cand, prev);
        FFX_TRY(rerank_impl(idx, mode, qvecs, nq, w.part_off, w.cand_sub, nullptr, 0.0, 0, std::max<int64_t>(window, 1),
                            w.ff_sub, nullptr, nullptr, nullptr, st, nullptr));
        ffx::es_scatter_kernel<<<per_query, 128, 0, st>>>(w);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
        prev = depth;
    }
    ffx::es_finish_kernel<<<flat, 256, 0, st>>>(w, nq, out_scored);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

int ffx_rerank_early_stop(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
                          const int32_t *cand, const float *lex, double alpha, int cutoff,
                          const int32_t *depths, int n_depths, int64_t max_cand, float *out_ff,
                          float *out_int, int32_t *out_scored, void *stream) {
    NvtxRange nvtx_range("ffx_rerank_early_stop");
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop: NULL index");
    if (mode < FFX_MODE_PASSAGE || mode > FFX_MODE_AVEP)
        return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop: unknown mode %d", mode);
    if (nq < 0 || nq > 0x7fffffffll || max_cand < 0 || cutoff < 1 || n_depths < 0 || (n_depths > 0 && !depths))
        return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop: bad sizes");
    if (nq == 0) return FFX_OK;
    if (!qvecs || !q_off || !out_scored || (max_cand > 0 && (!cand || !lex)))
        return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop: NULL input");
    if (idx->row_kind == FFX_ROWS_PQ_U8 && !idx->codewords)
        return fail(FFX_ERR_STATE, "ffx_rerank_early_stop: PQ index without codebooks (ffx_index_set_pq)");
    if (idx->sharded)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank_early_stop: not defined on a doc-id-range shard");
    if (mode != FFX_MODE_PASSAGE && idx->n_docs == 0 && max_cand > 0)
        return fail(FFX_ERR_STATE, "ffx_rerank_early_stop: document modes need ffx_index_set_docs first");
    const int cpad = next_pow2(std::max<int64_t>(max_cand, 1));
    if (cpad > ffx::kMaxFusedCand)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_rerank_early_stop: more than %d candidates per query",
                    ffx::kMaxFusedCand);
    ffx::EsPlan es{};
    FFX_TRY(plan_depths(depths, n_depths, cutoff, &es));
    es.out_scored = out_scored;
    FFX_TRY(bind(idx));
    FFX_TRY(settle(idx));
    // one launch for the register-staged lane-major dimensions; every other index kind walks the
    // depths as a stream-ordered sequence of launches
    const bool one_launch = idx->row_kind == FFX_ROWS_F32 && idx->plan.cpl != 0 && idx->plan.cpl <= 4 &&
                            idx->plan.lanes == 32;
    if (!one_launch)
        return early_stop_walk(idx, mode, qvecs, nq, q_off, cand, lex, alpha, es, max_cand, out_ff, out_int,
                               out_scored, static_cast<cudaStream_t>(stream));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    ffx::ScoreArgs a{};
    a.vectors = static_cast<const float *>(idx->store);
    a.doc_span = idx->doc_span;
    a.doc_rows = idx->doc_rows;
    a.indirect = idx->indirect;
    a.mode = mode;
    a.dim = idx->dim;
    a.qvecs = qvecs;
    a.q_off = q_off;
    a.cand = cand;
    a.lex = lex;
    a.alpha = static_cast<float>(alpha);
    a.beta = static_cast<float>(1.0 - alpha);
    a.out_ff = out_ff;
    a.out_int = out_int;
    a.tiles_per_query = 1;
    a.cpad = cpad;
    a.limit = a.count = static_cast<uint32_t>(mode == FFX_MODE_PASSAGE ? idx->num_rows : idx->n_docs);
    a.err = idx->err_flag;
    const size_t smem = static_cast<size_t>(cpad) * 8;
    const ffx_plan &p = idx->plan;
#define FFX_CASE(C, S_) \
    if (p.lanes == 32 && p.cpl == C && p.steps == S_) return launch_es<C, S_>(a, es, static_cast<unsigned>(nq), smem, st)
    FFX_CASE(1, 12);
    FFX_CASE(1, 16);
    FFX_CASE(2, 10);
    FFX_CASE(2, 12);
    FFX_CASE(2, 14);
    FFX_CASE(2, 16);
    FFX_CASE(4, 12);
    FFX_CASE(4, 16);
#undef FFX_CASE
    return fail(FFX_ERR_UNSUPPORTED, "no lane-major kernel for plan (%d,%d)", p.cpl, p.steps);
}

int ffx_rerank_early_stop_host(ffx_index *idx, int mode, const float *qvecs, int64_t nq,
                               const int64_t *q_off, const int32_t *cand, const float *lex,
                               double alpha, int cutoff, const int32_t *depths, int n_depths,
                               float *out_ff, float *out_int, int32_t *out_scored) {
    NvtxRange nvtx_range("ffx_rerank_early_stop_host");
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop_host: NULL index");
    if (nq < 0) return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop_host: nq < 0");
    if (nq == 0) return FFX_OK;
    if (!qvecs || !q_off || q_off[0] != 0 || !out_scored)
        return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop_host: bad input");
    int64_t max_cand = 0;
    for (int64_t q = 0; q < nq; q++) {
        const int64_t c = q_off[q + 1] - q_off[q];
        if (c < 0) return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop_host: q_off not monotone");
        max_cand = std::max(max_cand, c);
    }
    const int64_t n = q_off[nq];
    if (n > 0 && (!cand || !lex)) return fail(FFX_ERR_INVALID, "ffx_rerank_early_stop_host: NULL candidates / scores");
    FFX_TRY(bind(idx));
    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const int64_t D = idx->row_kind == FFX_ROWS_PQ_U8 ? static_cast<int64_t>(idx->M) * idx->Ds : idx->dim;
    const size_t b_q = pad(static_cast<size_t>(nq) * D * 4), b_off = pad(static_cast<size_t>(nq + 1) * 8);
    const size_t b_n = pad(static_cast<size_t>(n) * 4), b_s = pad(static_cast<size_t>(nq) * 4);
    FFX_TRY(scratch_reserve(idx->hostio, b_q + b_off + 4 * b_n + b_s));
    char *p = static_cast<char *>(idx->hostio.p);
    auto take = [&](size_t b) { char *r = p; p += b; return r; };
    float *d_q = reinterpret_cast<float *>(take(b_q));
    int64_t *d_off = reinterpret_cast<int64_t *>(take(b_off));
    int32_t *d_cand = reinterpret_cast<int32_t *>(take(b_n));
    float *d_lex = reinterpret_cast<float *>(take(b_n));
    float *d_ff = out_ff ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_int = out_int ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    int32_t *d_scored = reinterpret_cast<int32_t *>(take(b_s));
    cudaStream_t st = idx->stream;
    FFX_CUDA(cudaMemcpyAsync(d_q, qvecs, static_cast<size_t>(nq) * D * 4, cudaMemcpyHostToDevice, st));
    FFX_CUDA(cudaMemcpyAsync(d_off, q_off, static_cast<size_t>(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        FFX_CUDA(cudaMemcpyAsync(d_cand, cand, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, st));
        FFX_CUDA(cudaMemcpyAsync(d_lex, lex, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, st));
        // rows that are never scored read back as 0
        if (d_ff) FFX_CUDA(cudaMemsetAsync(d_ff, 0, static_cast<size_t>(n) * 4, st));
        if (d_int) FFX_CUDA(cudaMemsetAsync(d_int, 0, static_cast<size_t>(n) * 4, st));
    }
    FFX_TRY(ffx_rerank_early_stop(idx, mode, d_q, nq, d_off, d_cand, d_lex, alpha, cutoff, depths, n_depths,
                                  max_cand, d_ff, d_int, d_scored, st));
    if (n > 0) {
        if (out_ff) FFX_CUDA(cudaMemcpyAsync(out_ff, d_ff, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, st));
        if (out_int) FFX_CUDA(cudaMemcpyAsync(out_int, d_int, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, st));
    }
    FFX_CUDA(cudaMemcpyAsync(out_scored, d_scored, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, st));
    return take_error(idx, st);
}

// ---- product-quantizer build side --------------------------------------------------------
extern "C++" {
namespace {

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

template <int DS>
int launch_pq_assign(const float *vecs, int64_t n, int M, int Ks, int Ds, const float *cw, uint8_t *codes,
                     int sm_count) {
    const size_t smem = static_cast<size_t>(Ks) * Ds * 4;
    auto kern = ffx::ffx_pq_assign_kernel<DS>;
    FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int64_t tiles = std::min<int64_t>((n + ffx::kPqThreads - 1) / ffx::kPqThreads,
                                            std::max<int64_t>(1, static_cast<int64_t>(sm_count) * 16 / M));
    kern<<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(M)), ffx::kPqThreads, smem>>>(vecs, n, M, Ks, Ds,
                                                                                                  cw, codes);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

template <int DS>
int launch_pq_lloyd(const float *vecs, int64_t n, int M, int Ks, int Ds, const float *cw, double *sums,
                    unsigned long long *counts) {
    const size_t smem = static_cast<size_t>(Ks) * Ds * 8 + static_cast<size_t>(Ks) * 4;
    auto kern = ffx::ffx_pq_lloyd_kernel<DS>;
    FFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int rows_per_cta = 4096;
    kern<<<dim3(static_cast<unsigned>((n + rows_per_cta - 1) / rows_per_cta), static_cast<unsigned>(M)),
           ffx::kPqThreads, smem>>>(vecs, n, M, Ks, Ds, cw, sums, counts, rows_per_cta);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}

#define FFX_PQ_DISPATCH(fn, ...)                      \
    switch (Ds) {                                     \
        case 4: return fn<4>(__VA_ARGS__);            \
        case 8: return fn<8>(__VA_ARGS__);            \
        case 16: return fn<16>(__VA_ARGS__);          \
        case 32: return fn<32>(__VA_ARGS__);          \
        default: return fn<0>(__VA_ARGS__);           \
    }

int pq_assign(const float *vecs, int64_t n, int M, int Ks, int Ds, const float *cw, uint8_t *codes, int sm_count) {
    FFX_PQ_DISPATCH(launch_pq_assign, vecs, n, M, Ks, Ds, cw, codes, sm_count)
}
int pq_lloyd(const float *vecs, int64_t n, int M, int Ks, int Ds, const float *cw, double *sums,
             unsigned long long *counts) {
    FFX_PQ_DISPATCH(launch_pq_lloyd, vecs, n, M, Ks, Ds, cw, sums, counts)
}

int pq_check(const char *who, int device, const float *vecs, int64_t n, int M, int Ks, int Ds, const void *cw,
             size_t smem, int *sm_count) {
    if (!vecs || !cw || n < 0 || M <= 0 || Ks <= 0 || Ds <= 0) return fail(FFX_ERR_INVALID, "%s: bad arguments", who);
    if (Ks > 256) return fail(FFX_ERR_UNSUPPORTED, "%s: Ks=%d > 256 (uint8 codes)", who, Ks);
    if (smem > kSmemBudget) return fail(FFX_ERR_UNSUPPORTED, "%s: a %d x %d codebook exceeds shared memory", who, Ks, Ds);
    const int n_dev = ffx_device_count();
    if (n_dev <= 0) return fail(FFX_ERR_CUDA, "no CUDA device (ffx has no CPU path)");
    if (device < 0 || device >= n_dev) return fail(FFX_ERR_INVALID, "%s: device %d of %d", who, device, n_dev);
    FFX_CUDA(cudaSetDevice(device));
    cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device);
    return FFX_OK;
}

}  // namespace
}  // extern "C++"

int ffx_pq_encode(int device, const float *vecs, int64_t n, int M, int Ks, int Ds, const float *codewords,
                  uint8_t *codes) {
    int sm_count = 148;
    FFX_TRY(pq_check("ffx_pq_encode", device, vecs, n, M, Ks, Ds, codewords, static_cast<size_t>(Ks) * Ds * 4, &sm_count));
    if (n == 0) return FFX_OK;
    if (!codes) return fail(FFX_ERR_INVALID, "ffx_pq_encode: codes is NULL");
    const size_t D = static_cast<size_t>(M) * Ds;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (256ll << 20) / static_cast<int64_t>(D * 4)));
    DevBuf d_vec, d_codes, d_cw;
    FFX_CUDA(cudaMalloc(&d_vec.p, static_cast<size_t>(chunk) * D * 4));
    FFX_CUDA(cudaMalloc(&d_codes.p, static_cast<size_t>(chunk) * M));
    FFX_CUDA(cudaMalloc(&d_cw.p, static_cast<size_t>(M) * Ks * Ds * 4));
    FFX_CUDA(cudaMemcpy(d_cw.p, codewords, static_cast<size_t>(M) * Ks * Ds * 4, cudaMemcpyHostToDevice));
    for (int64_t r = 0; r < n; r += chunk) {
        const int64_t nr = std::min(chunk, n - r);
        FFX_CUDA(cudaMemcpy(d_vec.p, vecs + static_cast<size_t>(r) * D, static_cast<size_t>(nr) * D * 4, cudaMemcpyHostToDevice));
        FFX_TRY(pq_assign(static_cast<const float *>(d_vec.p), nr, M, Ks, Ds, static_cast<const float *>(d_cw.p),
                          static_cast<uint8_t *>(d_codes.p), sm_count));
        FFX_CUDA(cudaMemcpy(codes + static_cast<size_t>(r) * M, d_codes.p, static_cast<size_t>(nr) * M, cudaMemcpyDeviceToHost));
    }
    return FFX_OK;
}

int ffx_pq_kmeans(int device, const float *vecs, int64_t n, int M, int Ks, int Ds, float *codewords, int iters) {
    int sm_count = 148;
    FFX_TRY(pq_check("ffx_pq_kmeans", device, vecs, n, M, Ks, Ds, codewords,
                     static_cast<size_t>(Ks) * Ds * 8 + static_cast<size_t>(Ks) * 4, &sm_count));
    if (iters < 0) return fail(FFX_ERR_INVALID, "ffx_pq_kmeans: iters < 0");
    if (n == 0 || iters == 0) return FFX_OK;
    const size_t D = static_cast<size_t>(M) * Ds, n_cw = static_cast<size_t>(M) * Ks * Ds;
    DevBuf d_vec, d_cw, d_sum, d_cnt;
    FFX_CUDA(cudaMalloc(&d_vec.p, static_cast<size_t>(n) * D * 4));  // the training set stays resident
    FFX_CUDA(cudaMalloc(&d_cw.p, n_cw * 4));
    FFX_CUDA(cudaMalloc(&d_sum.p, n_cw * 8));
    FFX_CUDA(cudaMalloc(&d_cnt.p, static_cast<size_t>(M) * Ks * 8));
    FFX_CUDA(cudaMemcpy(d_vec.p, vecs, static_cast<size_t>(n) * D * 4, cudaMemcpyHostToDevice));
    FFX_CUDA(cudaMemcpy(d_cw.p, codewords, n_cw * 4, cudaMemcpyHostToDevice));
    for (int it = 0; it < iters; it++) {
        FFX_CUDA(cudaMemsetAsync(d_sum.p, 0, n_cw * 8));
        FFX_CUDA(cudaMemsetAsync(d_cnt.p, 0, static_cast<size_t>(M) * Ks * 8));
        FFX_TRY(pq_lloyd(static_cast<const float *>(d_vec.p), n, M, Ks, Ds, static_cast<const float *>(d_cw.p),
                         static_cast<double *>(d_sum.p), static_cast<unsigned long long *>(d_cnt.p)));
        ffx::ffx_pq_means_kernel<<<permute_grid(static_cast<int64_t>(n_cw), sm_count), 256>>>(
            static_cast<float *>(d_cw.p), static_cast<const double *>(d_sum.p),
            static_cast<const unsigned long long *>(d_cnt.p), static_cast<int64_t>(n_cw), Ds);
        g_launches++;
        FFX_CUDA(cudaGetLastError());
    }
    FFX_CUDA(cudaMemcpy(codewords, d_cw.p, n_cw * 4, cudaMemcpyDeviceToHost));
    return FFX_OK;
}

int ffx_sgemm(int device, int trans_a, int64_t m, int64_t n, int64_t k, const float *A, const float *B, float *C) {
    if (m <= 0 || n <= 0 || k <= 0 || !A || !B || !C) return fail(FFX_ERR_INVALID, "ffx_sgemm: bad arguments");
    const int n_dev = ffx_device_count();
    if (n_dev <= 0) return fail(FFX_ERR_CUDA, "no CUDA device (ffx has no CPU path)");
    if (device < 0 || device >= n_dev) return fail(FFX_ERR_INVALID, "ffx_sgemm: device %d of %d", device, n_dev);
    FFX_CUDA(cudaSetDevice(device));
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
    const int64_t tiles = ((m + ffx::kGemmTile - 1) / ffx::kGemmTile) * ((n + ffx::kGemmTile - 1) / ffx::kGemmTile);
    // slices of k: enough CTAs for ~4 waves, pieces of at least 1024 values of k, at most 64 partial results
    int64_t slices = std::max<int64_t>(1, std::min<int64_t>({64, (4 * sm_count + tiles - 1) / tiles, (k + 1023) / 1024}));
    const int64_t per_slice = ((k + slices - 1) / slices + ffx::kGemmK - 1) / ffx::kGemmK * ffx::kGemmK;
    slices = (k + per_slice - 1) / per_slice;
    DevBuf d_a, d_b, d_c, d_part;
    FFX_CUDA(cudaMalloc(&d_a.p, static_cast<size_t>(m) * k * 4));
    FFX_CUDA(cudaMalloc(&d_b.p, static_cast<size_t>(k) * n * 4));
    FFX_CUDA(cudaMalloc(&d_c.p, static_cast<size_t>(m) * n * 4));
    if (slices > 1) FFX_CUDA(cudaMalloc(&d_part.p, static_cast<size_t>(slices) * m * n * 4));
    FFX_CUDA(cudaMemcpy(d_a.p, A, static_cast<size_t>(m) * k * 4, cudaMemcpyHostToDevice));
    FFX_CUDA(cudaMemcpy(d_b.p, B, static_cast<size_t>(k) * n * 4, cudaMemcpyHostToDevice));
    const dim3 grid(static_cast<unsigned>((n + ffx::kGemmTile - 1) / ffx::kGemmTile),
                    static_cast<unsigned>((m + ffx::kGemmTile - 1) / ffx::kGemmTile), static_cast<unsigned>(slices));
    float *target = static_cast<float *>(slices > 1 ? d_part.p : d_c.p);
    if (trans_a)
        ffx::ffx_sgemm_kernel<true><<<grid, 256>>>(static_cast<const float *>(d_a.p), static_cast<const float *>(d_b.p), target, m, n, k, per_slice);
    else
        ffx::ffx_sgemm_kernel<false><<<grid, 256>>>(static_cast<const float *>(d_a.p), static_cast<const float *>(d_b.p), target, m, n, k, per_slice);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    if (slices > 1) {
        ffx::ffx_sgemm_reduce_kernel<<<permute_grid(m * n, sm_count), 256>>>(static_cast<const float *>(d_part.p),
                                                                             static_cast<float *>(d_c.p), m * n,
                                                                             static_cast<int>(slices));
        g_launches++;
        FFX_CUDA(cudaGetLastError());
    }
    FFX_CUDA(cudaMemcpy(C, d_c.p, static_cast<size_t>(m) * n * 4, cudaMemcpyDeviceToHost));
    return FFX_OK;
}

int ffx_index_coalesce(ffx_index *idx, int64_t doc0, int64_t n_docs, const int64_t *doc_off, double delta,
                       float *out_vectors, int32_t *out_groups) {
    NvtxRange nvtx_range("ffx_index_coalesce");
    if (!idx || doc0 < 0 || n_docs < 0 || !doc_off || !out_groups)
        return fail(FFX_ERR_INVALID, "ffx_index_coalesce: bad arguments");
    if (idx->row_kind != FFX_ROWS_F32) return fail(FFX_ERR_STATE, "ffx_index_coalesce: the index holds codes, not vectors");
    if (doc0 + n_docs > idx->n_docs) return fail(FFX_ERR_INVALID, "ffx_index_coalesce: documents out of range");
    if (n_docs == 0) return FFX_OK;
    const int64_t total = doc_off[n_docs];
    if (doc_off[0] != 0 || total <= 0 || !out_vectors) return fail(FFX_ERR_INVALID, "ffx_index_coalesce: bad offsets");
    FFX_TRY(bind(idx));
    FFX_TRY(settle(idx));
    const int stride = static_cast<int>(idx->row_bytes / 4);
    const size_t smem = static_cast<size_t>(ffx::kCoalesceWarps) * 2 * stride * 4;
    if (smem > kSmemBudget) return fail(FFX_ERR_UNSUPPORTED, "ffx_index_coalesce: rows of %d floats exceed shared memory", stride);
    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_off = pad(static_cast<size_t>(n_docs + 1) * 8), b_g = pad(static_cast<size_t>(n_docs) * 4);
    const size_t b_out = static_cast<size_t>(total) * idx->dim * 4;
    FFX_TRY(scratch_reserve(idx->work, b_off + b_g + b_out));
    char *p = static_cast<char *>(idx->work.p);
    int64_t *d_off = reinterpret_cast<int64_t *>(p);
    int32_t *d_groups = reinterpret_cast<int32_t *>(p + b_off);
    float *d_out = reinterpret_cast<float *>(p + b_off + b_g);
    cudaStream_t st = idx->stream;
    FFX_CUDA(cudaMemcpyAsync(d_off, doc_off, static_cast<size_t>(n_docs + 1) * 8, cudaMemcpyHostToDevice, st));
    ffx::CoalesceArgs a{};
    a.vectors = static_cast<const float *>(idx->store);
    a.stride = stride;
    a.dim = static_cast<int>(idx->dim);
    a.cpl = idx->plan.cpl;
    a.steps = idx->plan.steps;
    a.lanes = idx->plan.lanes;
    a.doc_span = idx->doc_span;
    a.doc_rows = idx->doc_rows;
    a.indirect = idx->indirect;
    a.doc0 = doc0;
    a.n_docs = n_docs;
    a.out_off = d_off;
    a.delta = delta;
    a.out = d_out;
    a.out_groups = d_groups;
    FFX_CUDA(cudaFuncSetAttribute(ffx::ffx_coalesce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ffx::ffx_coalesce_kernel<<<static_cast<unsigned>((n_docs + ffx::kCoalesceWarps - 1) / ffx::kCoalesceWarps),
                               ffx::kCoalesceWarps * 32, smem, st>>>(a);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    FFX_CUDA(cudaMemcpyAsync(out_groups, d_groups, static_cast<size_t>(n_docs) * 4, cudaMemcpyDeviceToHost, st));
    FFX_CUDA(cudaMemcpyAsync(out_vectors, d_out, b_out, cudaMemcpyDeviceToHost, st));
    FFX_CUDA(cudaStreamSynchronize(st));
    return FFX_OK;
}

int ffx_index_sync(ffx_index *idx, void *stream) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_index_sync: NULL index");
    FFX_TRY(bind(idx));
    return take_error(idx, static_cast<cudaStream_t>(stream));
}

int ffx_interpolate_topk(ffx_index *idx, const float *lex, const float *ff, int64_t nq,
                         const int64_t *q_off, double alpha, int k, int64_t max_cand, float *out_int,
                         float *out_topk_score, int32_t *out_topk_pos, void *stream) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_interpolate_topk: NULL index");
    if (nq < 0 || k < 0 || max_cand < 0 || nq > 0x7fffffffll || max_cand > (1ll << 30))
        return fail(FFX_ERR_INVALID, "ffx_interpolate_topk: bad sizes");
    if (nq == 0) return FFX_OK;
    if (!ff || !q_off) return fail(FFX_ERR_INVALID, "ffx_interpolate_topk: NULL input");
    if (k > 0 && (!out_topk_score || !out_topk_pos))
        return fail(FFX_ERR_INVALID, "ffx_interpolate_topk: k > 0 needs top-k outputs");
    FFX_TRY(bind(idx));
    const int cpad = next_pow2(std::max<int64_t>(max_cand, 1));
    if (cpad > ffx::kMaxFusedCand && k > 0)
        FFX_TRY(scratch_reserve(idx->work, static_cast<size_t>(nq) * cpad * 8));
    return launch_topk(ff, false, lex, static_cast<float>(alpha), static_cast<float>(1.0 - alpha), out_int,
                       q_off, nq, k, cpad, idx->work, 0, out_topk_score, out_topk_pos,
                       static_cast<cudaStream_t>(stream));
}

int ffx_interpolate_topk_host(ffx_index *idx, const float *lex, const float *ff, int64_t nq,
                              const int64_t *q_off, double alpha, int k, float *out_int,
                              float *out_topk_score, int32_t *out_topk_pos) {
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_interpolate_topk_host: NULL index");
    if (nq < 0) return fail(FFX_ERR_INVALID, "ffx_interpolate_topk_host: nq < 0");
    if (nq == 0) return FFX_OK;
    if (!ff || !q_off || q_off[0] != 0)
        return fail(FFX_ERR_INVALID, "ffx_interpolate_topk_host: bad input");
    int64_t max_cand = 0;
    for (int64_t q = 0; q < nq; q++) {
        const int64_t c = q_off[q + 1] - q_off[q];
        if (c < 0) return fail(FFX_ERR_INVALID, "ffx_interpolate_topk_host: q_off not monotone");
        max_cand = std::max(max_cand, c);
    }
    const int64_t n = q_off[nq];
    FFX_TRY(bind(idx));
    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_off = pad(static_cast<size_t>(nq + 1) * 8), b_n = pad(static_cast<size_t>(n) * 4);
    const size_t b_k = pad(static_cast<size_t>(nq) * k * 4);
    FFX_TRY(scratch_reserve(idx->hostio, b_off + 3 * b_n + 2 * b_k));
    char *p = static_cast<char *>(idx->hostio.p);
    auto take = [&](size_t b) { char *r = p; p += b; return r; };
    int64_t *d_off = reinterpret_cast<int64_t *>(take(b_off));
    float *d_ff = reinterpret_cast<float *>(take(b_n));
    float *d_lex = lex ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_int = (out_int && lex) ? reinterpret_cast<float *>(take(b_n)) : nullptr;
    float *d_ts = k > 0 ? reinterpret_cast<float *>(take(b_k)) : nullptr;
    int32_t *d_tp = k > 0 ? reinterpret_cast<int32_t *>(take(b_k)) : nullptr;
    cudaStream_t st = idx->stream;
    FFX_CUDA(cudaMemcpyAsync(d_off, q_off, static_cast<size_t>(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        FFX_CUDA(cudaMemcpyAsync(d_ff, ff, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, st));
        if (lex) FFX_CUDA(cudaMemcpyAsync(d_lex, lex, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, st));
    }
    FFX_TRY(ffx_interpolate_topk(idx, d_lex, d_ff, nq, d_off, alpha, k, max_cand, d_int, d_ts, d_tp, st));
    if (n > 0 && out_int)
        FFX_CUDA(cudaMemcpyAsync(out_int, lex ? d_int : d_ff, static_cast<size_t>(n) * 4,
                                 cudaMemcpyDeviceToHost, st));
    if (k > 0) {
        FFX_CUDA(cudaMemcpyAsync(out_topk_score, d_ts, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
        FFX_CUDA(cudaMemcpyAsync(out_topk_pos, d_tp, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
    }
    FFX_CUDA(cudaStreamSynchronize(st));
    return FFX_OK;
}

int ffx_merge_topk(int device, const float *shard_scores, const int32_t *shard_pos, int n_shards,
                   int64_t nq, int k, float *out_score, int32_t *out_pos, void *stream) {
    NvtxRange nvtx_range("ffx_merge_topk");
    if (n_shards <= 0 || nq < 0 || k <= 0 || !shard_scores || !shard_pos || !out_score || !out_pos)
        return fail(FFX_ERR_INVALID, "ffx_merge_topk: bad arguments");
    if (nq == 0) return FFX_OK;
    const int cpad = next_pow2(static_cast<int64_t>(n_shards) * k);
    const size_t smem = static_cast<size_t>(cpad) * 8;
    if (smem > 200 * 1024)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_merge_topk: %d shards x k=%d exceeds shared memory", n_shards, k);
    FFX_CUDA(cudaSetDevice(device));
    if (smem > 48 * 1024)
        FFX_CUDA(cudaFuncSetAttribute(ffx::ffx_merge_topk_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // 8 keys per thread where the list is long enough for the register sort (cpad >= 2048), never
    // fewer than 256 threads
    const int merge_threads = std::max(256, std::min(1024, cpad / 8));
    ffx::ffx_merge_topk_kernel<<<static_cast<unsigned>(nq), merge_threads, smem,
                                 static_cast<cudaStream_t>(stream)>>>(
        shard_scores, shard_pos, n_shards, nq, k, cpad, out_score, out_pos);
    g_launches++;
    FFX_CUDA(cudaGetLastError());
    return FFX_OK;
}


int ffx_merge_topk_host(ffx_index *idx, const float *shard_scores, const int32_t *shard_pos, int n_shards,
                        int64_t nq, int k, float *out_score, int32_t *out_pos) {
    NvtxRange nvtx_range("ffx_merge_topk_host");
    if (!idx) return fail(FFX_ERR_INVALID, "ffx_merge_topk_host: NULL index");
    if (n_shards <= 0 || nq < 0 || k <= 0 || !shard_scores || !shard_pos || !out_score || !out_pos)
        return fail(FFX_ERR_INVALID, "ffx_merge_topk_host: bad arguments");
    if (nq == 0) return FFX_OK;
    FFX_TRY(bind(idx));
    auto pad = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_in = pad(static_cast<size_t>(n_shards) * nq * k * 4), b_out = pad(static_cast<size_t>(nq) * k * 4);
    FFX_TRY(scratch_reserve(idx->hostio, 2 * b_in + 2 * b_out));
    char *p = static_cast<char *>(idx->hostio.p);
    float *d_s = reinterpret_cast<float *>(p);
    int32_t *d_p = reinterpret_cast<int32_t *>(p + b_in);
    float *d_os = reinterpret_cast<float *>(p + 2 * b_in);
    int32_t *d_op = reinterpret_cast<int32_t *>(p + 2 * b_in + b_out);
    cudaStream_t st = idx->stream;
    const size_t in_bytes = static_cast<size_t>(n_shards) * nq * k * 4, out_bytes = static_cast<size_t>(nq) * k * 4;
    FFX_CUDA(cudaMemcpyAsync(d_s, shard_scores, in_bytes, cudaMemcpyHostToDevice, st));
    FFX_CUDA(cudaMemcpyAsync(d_p, shard_pos, in_bytes, cudaMemcpyHostToDevice, st));
    FFX_TRY(ffx_merge_topk(idx->device, d_s, d_p, n_shards, nq, k, d_os, d_op, st));
    FFX_CUDA(cudaMemcpyAsync(out_score, d_os, out_bytes, cudaMemcpyDeviceToHost, st));
    FFX_CUDA(cudaMemcpyAsync(out_pos, d_op, out_bytes, cudaMemcpyDeviceToHost, st));
    FFX_CUDA(cudaStreamSynchronize(st));
    return FFX_OK;
}

}  // extern "C"
