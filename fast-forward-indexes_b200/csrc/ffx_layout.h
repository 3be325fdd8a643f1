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
