// ffx_layout.h — the lane-major row layout shared by host staging code and the kernels.
//
// numpy reduces a D-element fp32 row with a fixed pairwise tree (DESIGN.md "N1"): the row is
// cut into NB = 2^t leaf blocks of 8*S elements (8*S <= 128); inside a leaf, accumulator j
// (0..7) sums elements j, 8+j, 16+j, ... sequentially ("chain" (b, j), S terms); the 8
// accumulators and then the NB leaves are combined by a balanced binary tree.
//
// A warp reproduces that tree when lane l owns CPL = NB/4 whole chains
//     g = l*CPL + ch,  block b = g / 8,  accumulator j = g % 8
// and the combine is the xor-butterfly 1,2,4,8,16.  So that a lane's chains arrive with
// perfectly coalesced 128-bit loads, the STORE keeps every row permuted: the i-th float4 of
// lane l sits at float offset (i*32 + l)*4 and holds lane-local elements m = 4i..4i+3, with
// lane-local element m = step s = m / CPL of chain ch = m % CPL.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FFX_HD __host__ __device__
#else
#define FFX_HD
#endif

struct ffx_plan {
    int cpl;    // chains per lane (1, 2, 4, 8); 0 = no fast plan (identity layout)
    int steps;  // S: terms per chain
    int lanes;  // lanes that share one row: 32, or 16 / 8 for short rows (several rows per warp step)
};

// D = lanes * cpl * steps.  Only shapes whose numpy tree is uniform qualify.
static inline ffx_plan ffx_plan_for_dim(int64_t dim) {
    static const struct { int dim, cpl, steps, lanes; } table[] = {
        {384, 1, 12, 32}, {512, 1, 16, 32}, {640, 2, 10, 32}, {768, 2, 12, 32},
        {896, 2, 14, 32}, {1024, 2, 16, 32}, {1536, 4, 12, 32}, {2048, 4, 16, 32},
        // 32 leaves of 8*S elements, one whole leaf per lane: the row is streamed against a
        // shared-memory copy of the query vector (ffx_score_tma.cuh, kStream)
        {2560, 8, 10, 32}, {3072, 8, 12, 32}, {3584, 8, 14, 32}, {4096, 8, 16, 32},
        // short rows: one leaf (8 chains) or two (16 chains) — a quarter / half warp per row
        {64, 1, 8, 8}, {96, 1, 12, 8}, {128, 1, 16, 8}, {192, 1, 12, 16}, {256, 1, 16, 16},
    };
    for (unsigned i = 0; i < sizeof(table) / sizeof(table[0]); i++)
        if (table[i].dim == dim) return ffx_plan{table[i].cpl, table[i].steps, table[i].lanes};
    return ffx_plan{0, 0, 32};
}

// Rows of up to 512 elements are CONSUMED by the packed kernel (ffx_score_packed.cuh): lanes / cpl
// lanes per row, each taking cpl adjacent chains (24 - 32 elements per lane), 32 * cpl / lanes rows
// per warp step.  0 = not a packed shape (a whole warp per row, ffx_score_tma.cuh).
static inline int ffx_packed_cpl(const ffx_plan &p) {
    if (p.cpl == 0) return 0;
    if (p.lanes == 32) return p.cpl == 1 ? 2 : 0;  // D = 384, 512: 16 lanes per row
    return p.steps <= 8 ? 4 : 2;                   // D = 64: 2 lanes; 96 .. 256: 4 or 8 lanes
}
// rows per warp step (= per ring slot)
static inline int ffx_rows_per_step(const ffx_plan &p) {
    const int cpl = ffx_packed_cpl(p);
    return cpl ? 32 * cpl / p.lanes : 1;
}

// staged float offset k inside a row  ->  original element index.  The i-th float4 of lane l
// (of the `lanes` lanes sharing the row) sits at float offset (i*lanes + l)*4.
FFX_HD static inline int ffx_orig_index(int cpl, int steps, int k, int lanes = 32) {
    const int i = k / (4 * lanes);        // which float4 of the lane
    const int l = (k % (4 * lanes)) >> 2; // lane
    const int c = k & 3;
    const int m = 4 * i + c;              // lane-local element
    const int s = m / cpl, ch = m % cpl;
    const int g = l * cpl + ch;
    return (g >> 3) * (8 * steps) + 8 * s + (g & 7);
}


// ---- any other dimension: the numpy tree as data ------------------------------------------------
// For a D without a uniform tree the recursion of numpy's pairwise sum (split n > 128 at
// n2 = n/2 - (n/2) % 8; a leaf of 8 <= n <= 128 elements keeps 8 accumulators over its whole groups
// of 8 and then adds its n % 8 last elements one by one; n < 8 is summed one by one from 0) gives
// leaves of different lengths at different depths.  The plan places every leaf in a slot of a
// BALANCED tree of 2^t slots (a leaf at depth d < t takes the first of its 2^(t-d) slots, the
// others stay empty: x + 0 == x), so that the combine is the same xor-butterfly as in the uniform
// case: 8 chains per slot, CPL = 8 * slots / lanes chains per lane.  Rows stay in ORIGINAL element
// order in the store (row stride = D rounded up to 4 floats, for the 16-byte bulk copies); a lane
// reads its chains' elements from the staged row with computed addresses.  Only the last leaf
// can have a tail (every left part of a split is a multiple of 8).
struct ffx_any_plan {
    int valid;       // 0: more than 32 leaf slots (D beyond ~4096): thread-per-pair kernel
    int dim, stride; // elements, stored floats per row
    int lpr;         // lanes sharing one row: 4, 8, 16, 32
    int cpl;         // chains per lane: 1, 2, 4, 8 (2 for the short rows of one or two leaves)
    int n_slots;     // leaf slots, a power of two <= 32
    int max_steps;   // longest chain (groups of 8 in the longest leaf)
    int tail_slot, tail_start, tail_len;
    int16_t start[32];  // first element of the slot's leaf
    int16_t steps[32];  // whole groups of 8 in the slot's leaf; 0 = empty slot (or a leaf of < 8 elements)
};

struct ffx_any_leaf_ { int start, len, depth, index; };

static inline int ffx_any_collect_(int start, int n, int depth, int index, ffx_any_leaf_ *out, int cap, int at) {
    if (n <= 128) {
        if (at < cap) out[at] = ffx_any_leaf_{start, n, depth, index};
        return at + 1;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    at = ffx_any_collect_(start, n2, depth + 1, 2 * index, out, cap, at);
    return ffx_any_collect_(start + n2, n - n2, depth + 1, 2 * index + 1, out, cap, at);
}

static inline ffx_any_plan ffx_any_plan_for_dim(int64_t dim) {
    ffx_any_plan p{};
    if (dim <= 0 || dim > 32 * 128) return p;
    ffx_any_leaf_ leaves[64];
    const int n = ffx_any_collect_(0, static_cast<int>(dim), 0, 0, leaves, 64, 0);
    if (n > 64) return p;
    int depth = 0;
    for (int i = 0; i < n; i++) depth = leaves[i].depth > depth ? leaves[i].depth : depth;
    if (depth > 5) return p;
    p.n_slots = 1 << depth;
    p.dim = static_cast<int>(dim);
    p.stride = (p.dim + 3) & ~3;
    p.tail_slot = -1;
    for (int i = 0; i < n; i++) {
        const int slot = leaves[i].index << (depth - leaves[i].depth);
        const int len = leaves[i].len;
        p.start[slot] = static_cast<int16_t>(leaves[i].start);
        p.steps[slot] = static_cast<int16_t>(len < 8 ? 0 : len / 8);
        if (p.steps[slot] > p.max_steps) p.max_steps = p.steps[slot];
        const int tail = len < 8 ? len : len % 8;
        if (tail) {
            p.tail_slot = slot;
            p.tail_start = leaves[i].start + len - tail;
            p.tail_len = tail;
        }
    }
    // up to four leaf slots (D <= 512): two chains per lane, 4 / 8 / 16 lanes per row (8 / 4 / 2 rows
    // per warp step) — at least ~16 elements per lane and step
    const int chains = 8 * p.n_slots;
    p.lpr = chains >= 64 ? 32 : chains / 2;
    p.cpl = chains / p.lpr;
    p.valid = 1;
    return p;
}
