/* ffx.h — C ABI of the B200-native Fast-Forward re-ranking engine (libffx.so).
 *
 * The reference (mrjleo/fast-forward-indexes v0.8.0) is pure Python and has no FFI seam;
 * its seam is the `Index` class contract.  Each entry point below replaces one piece of
 * that contract on the re-ranking hot path; the citation is the reference code it stands
 * in for (paths relative to /root/reference/src/fast_forward/).  The Python drop-in shell
 * (fast-forward-indexes_b200/fast_forward/) binds these with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no torch / numpy types cross this boundary
 *   - every call returns 0 on success or a negative ffx_status; the message for the last
 *     failure on the calling thread is ffx_last_error(); nothing throws across the ABI
 *   - the caller owns every host buffer; the library owns all device memory it allocates
 *   - id strings never cross the ABI: the host shell maps ids to integers with the exact
 *     semantics of index/util.py:12-42 and raises IndexError itself
 *   - calls on one ffx_index must be serialised by the caller (the reference is
 *     single-threaded as well)
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 */
#ifndef FFX_H_
#define FFX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFX_ABI_VERSION 1

typedef struct ffx_index ffx_index;

typedef enum {
    FFX_OK = 0,
    FFX_ERR_INVALID = -1,    /* bad argument */
    FFX_ERR_CUDA = -2,       /* CUDA runtime error (message has the cudaError string) */
    FFX_ERR_OOM = -3,        /* device or pinned-host allocation failed */
    FFX_ERR_STATE = -4,      /* call not valid in the index's current state */
    FFX_ERR_UNSUPPORTED = -5 /* shape outside what the kernels implement */
} ffx_status;

/* index/base.py:18-24 (`Mode`); same numeric values */
typedef enum {
    FFX_MODE_PASSAGE = 1,
    FFX_MODE_MAXP = 2,
    FFX_MODE_FIRSTP = 3,
    FFX_MODE_AVEP = 4
} ffx_mode;

typedef enum {
    FFX_ROWS_F32 = 0,  /* fp32 vectors, `dim` floats per row */
    FFX_ROWS_PQ_U8 = 1 /* PQ/OPQ codes, `dim` = M code bytes per row (Ks <= 256) */
} ffx_row_kind;

/* ---- library ---------------------------------------------------------------------- */
int ffx_abi_version(void);
const char *ffx_last_error(void);
/* number of CUDA devices visible; 0 when there is none (then nothing else will work) */
int ffx_device_count(void);

/* Tuning / diagnostics (process-wide; 0 restores the automatic choice).  Not part of the
 * reference contract — it has no kernels to tune.
 *   "kernel"      1 = register-staged scoring kernel, 2 = TMA-staged whole-warp kernel, where they exist
 *                 (D >= 384 with a uniform summation tree); default: the packed kernel for rows of up to
 *                 512 elements and for every dimension without a uniform tree, the TMA-staged one above
 *   "tma_stages"  ring slots per warp of the TMA-staged kernel (capped by shared memory)
 *   "batch"       candidates a warp takes per grab (1..32)
 *   "tma_warps"   warps per CTA of the TMA-staged kernel
 *   "chunk_waves" ffx_rerank_host: queries per pipelined chunk, in units of 2 x #SMs (default 1)
 *   "adc"         1 = generic thread-per-row ADC kernel, 2 = warp-per-row with conflict-free tables
 *                 (M = 64..128), 3 = XOR-swizzled thread-per-row (M % 32 == 0, M <= 128); a kernel the
 *                 shape does not allow falls back to the next one
 *   "adc_lut"     tables of the XOR-swizzled ADC kernel: 1 = built inside the scoring kernel, 2 = by the
 *                 thread-per-entry kernel; default: ahead of the launch by the tiled kernel */
int ffx_set_option(const char *name, int value);

/* Pinned host memory for buffers that cross PCIe every call (candidate lists, query
 * vectors, outputs).  Plain malloc'ed buffers work too, just slower. */
int ffx_host_alloc(void **out, int64_t bytes);
int ffx_host_free(void *p);

/* ---- index storage ---------------------------------------------------------------- */
/* Replaces `InMemoryIndex.__init__` chunk allocation (index/memory.py:23-56): one
 * device-resident row store with room for `capacity_rows` rows on CUDA device `device`. */
int ffx_index_create(int device, int row_kind, int64_t dim, int64_t capacity_rows, ffx_index **out);
int ffx_index_destroy(ffx_index *idx);
/* Grow the store (index/memory.py:103-108, the alloc_size chunk growth); keeps contents. */
int ffx_index_reserve(ffx_index *idx, int64_t capacity_rows);

/* Copies all stored rows of `src` into the EMPTY index `dst` (same row kind and dimension; same or
 * another device: device-to-device / peer copy of the store as it is laid out, no host round
 * trip) — the rows of `OnDiskIndex.to_memory()` (index/disk.py:177-205), whose on-disk index is
 * already resident in HBM here.  Document tables and PQ tables are not copied (ffx_index_set_docs /
 * ffx_index_set_pq on the copy). */
int ffx_index_copy_rows(ffx_index *dst, ffx_index *src);
/* Replaces the chunk copy of `InMemoryIndex._add` (index/memory.py:97-119) and the slice
 * loop of `OnDiskIndex.to_memory` (index/disk.py:190-204): rows [row0, row0+nrows) take
 * `rows` (row-major, fp32 or uint8 per row_kind).  Host sources go through a pinned double
 * buffer; `src_on_device != 0` means `rows` is a device pointer on the index's device.
 * The store keeps fp32 rows in a lane-major permuted layout; callers only ever
 * see the original element order. */
int ffx_index_stage_rows(ffx_index *idx, int64_t row0, int64_t nrows, const void *rows,
                         int src_on_device);
/* Replaces `ChunkIndexer.__call__`'s fancy-index gather (index/util.py:97-113): copies the
 * listed rows, in order and in original element order, to host memory `out`. */
int ffx_index_read_rows(ffx_index *idx, const int64_t *rows, int64_t n, void *out);
int64_t ffx_index_num_rows(const ffx_index *idx);     /* highest staged row + 1 */
int64_t ffx_index_capacity(const ffx_index *idx);
int64_t ffx_index_dim(const ffx_index *idx);
/* 1 when `dim` has a warp-per-row kernel (lane-major plan, or the tree-as-data kernel for every
 * other dimension up to 4096), 0 when the thread-per-pair exact kernel is used */
int ffx_index_has_fast_path(const ffx_index *idx);

/* Replaces `doc_id_to_idx` (index/memory.py:86-88, index/disk.py:408-417): document
 * ordinal d owns rows doc_rows[doc_off[d] .. doc_off[d+1]) in insertion order.  When every
 * document's rows are consecutive the library keeps only (first,count) spans. */
int ffx_index_set_docs(ffx_index *idx, int64_t n_docs, const int64_t *doc_off,
                       const int64_t *doc_rows);

/* Declares this index to be one doc-id-range shard of a larger corpus (the
 * reference has no sharding — a corpus must fit one process, index/memory.py).  Candidates
 * passed to ffx_rerank are then GLOBAL ordinals: documents [doc_base, doc_base + #docs here) of
 * `global_docs`, rows [row_base, row_base + #rows here) of `global_rows` in PASSAGE mode.  Pairs
 * owned by another shard are skipped: their out_ff/out_int entries are left untouched (zero
 * them and sum across shards) and they take no top-k slot; positions stay positions in the
 * full candidate block, so per-shard lists merge with ffx_merge_topk.  global_docs ==
 * global_rows == 0 switches sharding off.  Call after ffx_index_set_docs / staging. */
int ffx_index_set_shard(ffx_index *idx, int64_t doc_base, int64_t global_docs, int64_t row_base,
                        int64_t global_rows);

/* Fused exchange of a doc-id-range sharded corpus (the reference has no counterpart).
 * With a scatter plan the fused scoring kernel writes every query's top-k list in its epilogue
 * straight into the receive buffer of the query's OWNER rank, over NVLink peer memory (plain
 * stores to a peer-mapped pointer), instead of into out_topk_* — compute and exchange are one
 * kernel, no all-to-all.  Owner o merges queries [bounds[o], bounds[o+1]) (host array
 * [world + 1]); peer_score[o] / peer_pos[o] are owner o's receive buffers as device pointers
 * valid on THIS device, laid out [world][stride][k] (stride >= every owner's query count); this
 * index writes block `rank` of each.  The caller synchronises the ranks (a device-side barrier
 * on the launching stream) before merging block-wise with ffx_merge_topk, and alternates two
 * buffer sets between calls.  Only the fused fp32 kernel carries the plan: ffx_rerank returns
 * FFX_ERR_UNSUPPORTED otherwise.  world = 0 removes the plan. */
int ffx_index_set_topk_scatter(ffx_index *idx, int world, int rank, int64_t stride, const int64_t *bounds,
                               void *const *peer_score, void *const *peer_pos);

/* Replaces `Quantizer.decode` on the scoring path (quantizer/base.py:123-132 ->
 * quantizer/nanopq.py:43-44,111-112): attaches nanopq-compatible codebooks
 * `codewords[M][Ks][Ds]` and, for OPQ, the rotation `R[D][D]` (NULL for plain PQ) to an
 * FFX_ROWS_PQ_U8 index.  Scoring then uses asymmetric distance computation. */
int ffx_index_set_pq(ffx_index *idx, int M, int Ks, int Ds, const float *codewords,
                     const float *R);

/* ---- host-side id coding ------------------------------------------------------------- */
/* Replaces the id dictionaries of the reference (`_doc_id_to_idx: dict[str, list[int]]`,
 * `_psg_id_to_idx: dict[str, int]`, index/memory.py:46-47,84-95; rebuilt by an O(N) Python loop
 * in index/disk.py:408-417) and the hashing of every candidate id of every call inside pandas
 * merges on string keys (index/base.py:291-298,314; index/util.py:29-41).  Pure host code (no
 * CUDA device needed).  Strings cross the ABI in Arrow layout: string i is
 * data[offsets[i] .. offsets[i+1]) (UTF-8), `validity` is an optional bitmap (bit
 * bit_offset + i, LSB first, NULL = all valid) — a pandas / pyarrow string column is coded in
 * place, on all host cores, without creating Python objects. */
typedef struct ffx_dict ffx_dict;
int ffx_dict_create(ffx_dict **out);
int ffx_dict_destroy(ffx_dict *d);
int64_t ffx_dict_size(const ffx_dict *d);       /* number of keys */
int64_t ffx_dict_key_bytes(const ffx_dict *d);  /* total length of all keys */
/* Document ids (index/memory.py:86-88): a new key gets the next ordinal (= number of keys so
 * far), a known key keeps its ordinal; out[i] = ordinal of string i, -1 for a null. */
int ffx_dict_insert_ordinal(ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                            int64_t bit_offset, int64_t n, int64_t *out);
/* Passage ids (index/memory.py:90-95): string i gets value first_value + i (its row number);
 * nulls are skipped.  A key that already exists — in the dictionary or earlier in the batch —
 * is an error: FFX_ERR_STATE, *first_dup = its index, and the dictionary is left unchanged
 * ("Passage ID ... already exists").  dry_run != 0 only checks. */
int ffx_dict_insert_unique(ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                           int64_t bit_offset, int64_t n, int64_t first_value, int dry_run, int64_t *first_dup);
/* Replaces index/util.py:29-41 for a whole column: out[i] = value of string i (as the int32
 * candidate ffx_rerank takes), -1 when absent or null; *first_missing = lowest such i or -1
 * (the id IndexError names).  n_threads 0 = all host cores. */
int ffx_dict_lookup(const ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                    int64_t bit_offset, int64_t n, int32_t *out, int64_t *first_missing, int n_threads);
/* Keys in insertion order (Arrow layout; offsets[size+1], data[key_bytes]) and their values. */
int ffx_dict_export(const ffx_dict *d, int64_t *offsets, char *data, int64_t *values);
/* doc -> rows CSR from the per-row document ordinals (-1 = row without document id): counting
 * sort, rows of a document in increasing (= insertion) order.  doc_off[n_docs+1]; doc_rows may
 * be NULL to get the offsets only. */
int ffx_csr_build(const int64_t *row_doc, int64_t n_rows, int64_t n_docs, int64_t *doc_off, int64_t *doc_rows);

/* Factorises a string column (Arrow layout, no nulls) on all host cores: codes[i] == codes[j] iff the
 * strings are equal, codes in [0, *n_keys); unlike ffx_dict_insert_ordinal the numbering is NOT the
 * order of first appearance (rows are partitioned by hash and deduplicated per partition) — what
 * `Ranking.__init__` needs for its id column (ranking.py:95-117 on integer codes).  The distinct
 * strings, in code order, come from ffx_factor_export (key_offsets[n_keys + 1], key_data[key_bytes]);
 * the input buffers must stay alive until then.  ffx_factor_free releases the handle. */
typedef struct ffx_factor ffx_factor;
int ffx_factorize(const int64_t *offsets, const char *data, int64_t n, int32_t *codes, ffx_factor **out, int64_t *n_keys,
                  int64_t *key_bytes, int n_threads);
int ffx_factor_export(const ffx_factor *f, int64_t *key_offsets, char *key_data);
void ffx_factor_free(ffx_factor *f);

/* A deep copy of a dictionary (the id tables of `OnDiskIndex.to_memory()`). */
int ffx_dict_clone(const ffx_dict *d, ffx_dict **out);

/* A column of fixed-width, NUL-padded byte ids (HDF5 `S{max_id_length}` datasets, "" = no id:
 * index/disk.py:152-165,414-417) as Arrow string buffers for the dictionaries above:
 * offsets[n+1], the concatenated bytes in `out` (room for n * width), a validity bitmap of
 * (n + 7) / 8 bytes (bit i clear = empty id), *n_valid = ids present. */
int ffx_fixed_width_to_arrow(const char *data, int64_t n, int width, int64_t *offsets, char *out,
                             uint8_t *validity, int64_t *n_valid);

/* ---- Ranking.__init__ on integer codes (host; ranking.py:67-121) ---------------------- */
/* The duplicate-pair check (ranking.py:95-98) over keys = q_code * n_ids + id_code:
 * *first = index of the first key that equals an earlier one, -1 if all are distinct. */
int ffx_first_repeat(const int64_t *keys, int64_t n, int64_t *first);
/* The frame order of ranking.py:115-117 — q_id descending (q_rank[i] = position of row i's
 * q_id among the distinct q_ids sorted descending), then score descending, ties in incoming
 * order (pandas' lexsort is stable; -0.0 ties with +0.0; NaN sorts after every number):
 * order[j] = the row that comes j-th.  A radix sort on all host cores (n_threads 0). */
int ffx_ranking_order(const int32_t *q_rank, const float *score, int64_t n, int64_t *order, int n_threads);
/* The (q_id, id) key order of a pandas outer merge (`Ranking.interpolate` / `__add__`,
 * ranking.py:188-217,310-318) on integer ranks: order[j] = the element that comes j-th by
 * ascending key, equal keys in index order. */
int ffx_order_u64(const uint64_t *keys, int64_t n, int64_t *order, int n_threads);
/* The join itself: pos[i] = index of want[i] in have[] (its first occurrence), -1 if absent. */
int ffx_match_keys(const int64_t *have, int64_t n_have, const int64_t *want, int64_t n_want, int64_t *pos);

/* Rankings on integer codes (host; the result side of index/base.py:461-469 and ranking.py:279-326,
 * which the reference builds with pandas merges on two string columns).
 * ffx_lut_gather: out[i] = lut[codes[i]] on all host cores — the candidate of every pair from the
 *   candidate of every DISTINCT id of a ranking (index/util.py:29-41 resolved once per id, not once
 *   per pair); *first_negative = lowest i with lut[codes[i]] < 0 (the id IndexError names), -1 if none.
 * ffx_topk_gather: the [nq, k] (position, score) lists of ffx_rerank -> the rows of the result
 *   ranking.  Query q keeps its first min(keep, valid) entries (lists are padded with pos = -1):
 *   out_off[nq + 1] = prefix sums of the kept counts, out_code[j] = src_code[src_off[q] + pos]
 *   (the id code of the source row), out_score[j] = score; either output may be NULL, both need
 *   room for nq * keep entries.  *n_ties = adjacent equal scores inside kept lists (the rows the
 *   host then orders by id, ranking.py:312-326); straddle[q] (optional, needs keep < k) = the
 *   keep-th and (keep+1)-th scores of q are equal, i.e. the cut falls inside a tie. */
/* ffx_tie_runs: the runs of equal scores inside the blocks [off[q], off[q+1]) of a ranked score
 *   column (blocks sorted descending): run_start[i] / run_len[i] (>= 2) in row order, *n_runs =
 *   their number; when that exceeds `cap` nothing is written (call again with room). */
int ffx_tie_runs(const float *score, const int64_t *off, int64_t nq, int64_t cap, int64_t *run_start,
                 int64_t *run_len, int64_t *n_runs, int n_threads);
int ffx_lut_gather(const int32_t *lut, int64_t n_lut, const int32_t *codes, int64_t n, int32_t *out,
                   int64_t *first_negative, int n_threads);
int ffx_topk_gather(const int32_t *pos, const float *score, int64_t nq, int64_t k, int64_t keep,
                    const int64_t *src_off, const int32_t *src_code, int64_t *out_off, int32_t *out_code,
                    float *out_score, int64_t *n_ties, uint8_t *straddle, int n_threads);

/* ---- TREC run files (host; ranking.py:348-366,388-409) --------------------------------------
 * ffx_run_open maps a whitespace-separated run file (q_id Q0 id rank score name), tokenises it
 * on n_threads cores (0 = all) and reports in info[16]:
 *   [0] rows, [1] bytes of all q_id tokens, [2] bytes of all id tokens, [3] lines without exactly six
 *   fields, [4] scores that are not plain decimals (or out of double range), [5] quote characters,
 *   [6..9] q_id column: tokens that look numeric / are canonical integers / are pandas NA strings /
 *   are boolean words, [10..13] the same for the id column, [14] bit 0/1/2: the first row's name looks
 *   numeric / NA / boolean, [15] length of the first row's name.
 * The caller decides from those whether pandas would have produced the same strings (it renumbers
 * all-numeric columns, drops NA tokens, ...) and then fetches the columns with ffx_run_read: Arrow
 * layout offsets[rows + 1] + data for the two id columns, scores as the doubles pandas' default
 * float parser produces (not correctly rounded — the reference's float32 scores are reproduced
 * bit for bit), the first row's name.  ffx_run_close unmaps.
 * ffx_run_write writes a ranking held as integer-coded columns (q_keys per block, block offsets,
 * the distinct ids + one code per row, float32 scores) as `q_id \t Q0 \t id \t rank \t score \t name`
 * lines, scores formatted as numpy / pandas print float32 (shortest round trip; positional for
 * 1e-4 <= |x| < 1e6, else scientific). */
typedef struct ffx_run ffx_run;
int ffx_run_open(const char *file_name, int n_threads, ffx_run **out, int64_t *info);
int ffx_run_read(ffx_run *run, int64_t *q_offsets, char *q_data, int64_t *id_offsets, char *id_data, double *score,
                 char *first_name);
void ffx_run_close(ffx_run *run);
int ffx_run_write(const char *file_name, int64_t nq, const int64_t *q_off, const int64_t *qk_offsets, const char *qk_data,
                  const int64_t *idk_offsets, const char *idk_data, const int32_t *id_code, const float *score,
                  const char *name, int n_threads);

/* ---- the hot path ----------------------------------------------------------------- */
/* Replaces, in one pass, `Index._compute_scores` (index/base.py:279-314) including
 * `_get_vectors` (index/memory.py:139-140), `Ranking.interpolate`'s arithmetic
 * (ranking.py:319) and the per-query `Ranking.cut` (ranking.py:115-117, :285-291).
 *
 *   qvecs   [nq, D] fp32 query vectors (D = vector dimension, also for PQ indexes)
 *   q_off   [nq+1]  pair offsets: query q owns pairs [q_off[q], q_off[q+1])
 *   cand    [n]     candidate per pair: a document ordinal (MAXP/FIRSTP/AVEP, as given to
 *                   ffx_index_set_docs) or a row number (PASSAGE)
 *   lex     [n]     first-stage scores, or NULL for "no interpolation"
 *   alpha           interpolated = fl32(fl32(alpha)*lex) + fl32(fl32(1-alpha)*ff)
 *   k               per-query cut-off; 0 = no top-k
 *   out_ff  [n]     semantic score per pair (may be NULL)
 *   out_int [n]     interpolated score per pair (may be NULL; equals ff when lex==NULL)
 *   out_topk_score / out_topk_pos [nq, k]: the k best interpolated scores per query in
 *                   descending order, ties by ascending position inside the query's block;
 *                   pos is that position; unused slots are (-inf, -1)
 *
 * fp32 scores are bit-identical to numpy's `np.sum(q * d, axis=1)` + pandas' groupby
 * max / mean / first; PQ scores are ADC sums (tolerance in DESIGN.md).
 *
 * ffx_rerank: all pointers are DEVICE pointers on the index's device, the work is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call returns
 * without synchronising.  `max_cand` must be >= the largest q_off[q+1]-q_off[q].
 * ffx_rerank_host: all pointers are HOST pointers; copies in, runs, copies out and
 * synchronises before returning. */
int ffx_rerank(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
               const int32_t *cand, const float *lex, double alpha, int k, int64_t max_cand,
               float *out_ff, float *out_int, float *out_topk_score, int32_t *out_topk_pos,
               void *stream);
int ffx_rerank_host(ffx_index *idx, int mode, const float *qvecs, int64_t nq,
                    const int64_t *q_off, const int32_t *cand, const float *lex, double alpha,
                    int k, float *out_ff, float *out_int, float *out_topk_score,
                    int32_t *out_topk_pos);

/* Replaces `Index._early_stopping` (index/base.py:316-387): scores every query's candidates in
 * the depth intervals [0,d0), [d0,d1), ... and, before each interval after the first, stops a
 * query once its `cutoff`-th best interpolated score so far is not below
 * fl32(fl32(alpha)*lex[last scored row]) + fl32(fl32(1-alpha)*max ff so far)  (base.py:351-356).
 * One launch, one CTA per query; a stopped query issues no further loads.
 *
 *   cand / lex   as ffx_rerank, each query's block in rank order (depth = position); lex required
 *   depths       HOST array of n_depths depths in any order.  As in the reference they are taken
 *                ascending, depths < cutoff are skipped (base.py:341-343) and the walk ends at the
 *                first depth that adds no rows (a repeated depth; base.py:365-366)
 *   out_ff / out_int [n]   written for the scored rows only (others untouched); may be NULL
 *   out_scored [nq]        rows scored per query: always a prefix of the query's block
 *
 * One launch, one CTA per query, on fp32 indexes of the register-staged lane-major dimensions
 * (384 ... 2048); every other index kind (PQ / OPQ codes, the other dimensions) walks the depths
 * as a stream-ordered sequence of launches per depth — criterion, compaction of the pairs still
 * to score, the index's ordinary scoring kernel, scatter — with no host round trip in between.
 * At most 16384 candidates per query and 32 distinct depths, not on a shard: anything else returns
 * FFX_ERR_UNSUPPORTED (the host shell then walks the depths itself with ffx_rerank).  ffx_rerank_early_stop takes DEVICE pointers (except `depths`) and is
 * asynchronous on `stream`; the _host variant takes host pointers and synchronises. */
int ffx_rerank_early_stop(ffx_index *idx, int mode, const float *qvecs, int64_t nq, const int64_t *q_off,
                          const int32_t *cand, const float *lex, double alpha, int cutoff,
                          const int32_t *depths, int n_depths, int64_t max_cand, float *out_ff,
                          float *out_int, int32_t *out_scored, void *stream);
int ffx_rerank_early_stop_host(ffx_index *idx, int mode, const float *qvecs, int64_t nq,
                               const int64_t *q_off, const int32_t *cand, const float *lex,
                               double alpha, int cutoff, const int32_t *depths, int n_depths,
                               float *out_ff, float *out_int, int32_t *out_scored);

/* ---- product-quantizer build side ------------------------------------------------------ */
/* Replace what the reference delegates to nanopq 0.2.1 / scipy on the host.  Host pointers;
 * synchronous; need a CUDA device.  Ks <= 256 (uint8 codes).
 * ffx_pq_encode: nanopq `PQ.encode` (quantizer/nanopq.py:41,109 -> scipy.cluster.vq.vq per
 *   subspace): codes[i, m] = argmin_k |vecs[i, m*Ds:(m+1)*Ds] - codewords[m, k]|^2, ties to the
 *   lower k.  `vecs` are already rotated for OPQ.
 * ffx_pq_kmeans: the Lloyd iterations of nanopq `PQ.fit` (quantizer/nanopq.py:30,98 ->
 *   scipy.cluster.vq.kmeans2 per subspace): `iters` rounds of assign + mean from the initial
 *   `codewords` (in/out, [M, Ks, Ds]); a codeword without members keeps its value
 *   (kmeans2 missing="warn").  The training set must fit the device. */
int ffx_pq_encode(int device, const float *vecs, int64_t n, int M, int Ks, int Ds, const float *codewords,
                  uint8_t *codes);
int ffx_pq_kmeans(int device, const float *vecs, int64_t n, int M, int Ks, int Ds, float *codewords, int iters);
/* The two matrix products of an OPQ rotation round (quantizer/nanopq.py:94-98 -> nanopq `OPQ.fit`:
 * X = vecs @ R and the Procrustes matrix vecs^T @ X_hat; the D x D SVD stays with the caller):
 * C[m, n] = op(A) . B in fp32 (FMA, no tensor cores), row-major host pointers, B is [k, n],
 * A is [m, k] (trans_a = 0) or [k, m] (trans_a != 0: C = A^T . B).  Deterministic. */
int ffx_sgemm(int device, int trans_a, int64_t m, int64_t n, int64_t k, const float *A, const float *B, float *C);

/* Replaces the per-document loop of `create_coalesced_index` (util/__init__.py:51-101) for its
 * default distance `cos_dist` (:40-48): sequential coalescing of the passage vectors of documents
 * [doc0, doc0 + n_docs) with threshold `delta`, one warp per document.  doc_off[n_docs + 1] (host)
 * = cumulative row counts of those documents; the means of document d's groups are written to rows
 * doc_off[d] .. doc_off[d] + out_groups[d] - 1 of out_vectors (host, [doc_off[n_docs], dim], original
 * element order; the remaining rows of a document's range are not written).  fp32 vector index only. */
int ffx_index_coalesce(ffx_index *idx, int64_t doc0, int64_t n_docs, const int64_t *doc_off, double delta,
                       float *out_vectors, int32_t *out_groups);

/* Synchronises `stream` and reports what the asynchronous launches on this index saw: the
 * kernels never dereference a candidate outside [0, #documents) (or [0, #rows) in PASSAGE
 * mode) — such a pair scores as an empty document and this call (like ffx_rerank_host)
 * returns FFX_ERR_INVALID naming the pair.  The host shell maps ids itself and raises
 * IndexError (index/util.py:38-39) long before; this is the last line of defence. */
int ffx_index_sync(ffx_index *idx, void *stream);

/* Replaces `Ranking.interpolate` + `Ranking.cut` (ranking.py:293-326, :279-291) over scores
 * that already exist: interpolated = fl32(fl32(alpha)*lex) + fl32(fl32(1-alpha)*ff) per pair
 * (lex == NULL: the scores in `ff` are ranked as they are), then the per-query top-k with the
 * ordering rule of ffx_rerank.  k = 0 only interpolates.  ffx_interpolate_topk takes device
 * pointers and is asynchronous on `stream`; the _host variant takes host pointers and
 * synchronises.  `idx` only provides the device and scratch memory. */
int ffx_interpolate_topk(ffx_index *idx, const float *lex, const float *ff, int64_t nq,
                         const int64_t *q_off, double alpha, int k, int64_t max_cand, float *out_int,
                         float *out_topk_score, int32_t *out_topk_pos, void *stream);
int ffx_interpolate_topk_host(ffx_index *idx, const float *lex, const float *ff, int64_t nq,
                              const int64_t *q_off, double alpha, int k, float *out_int,
                              float *out_topk_score, int32_t *out_topk_pos);

/* Exchange step of a doc-id-range sharded corpus: merges `n_shards` per-shard
 * top-k lists [n_shards, nq, k] (scores + GLOBAL positions, device pointers) into
 * [nq, k] with the same ordering rule.  Runs on `stream`. */
int ffx_merge_topk(int device, const float *shard_scores, const int32_t *shard_pos, int n_shards,
                   int64_t nq, int k, float *out_score, int32_t *out_pos, void *stream);

/* The same merge from HOST buffers (one process driving several GPUs: the per-device lists come
 * back through ffx_rerank_host); `idx` provides the device, stream and scratch.  Synchronises. */
int ffx_merge_topk_host(ffx_index *idx, const float *shard_scores, const int32_t *shard_pos, int n_shards,
                        int64_t nq, int k, float *out_score, int32_t *out_pos);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t ffx_launch_count(void);
/* Demangled symbol of the scoring kernel (fp32 gather-dot or ADC) launched last by this process,
 * "" before the first launch — what bench.py reports as roofline.kernel and what the ncu launch
 * list must show. */
const char *ffx_last_kernel(void);

/* ---- index files: the HDF5 layout OnDiskIndex writes, read without libhdf5 / h5py --------
 * Replaces the h5py reads of OnDiskIndex.load (index/disk.py:355-418), to_memory
 * (:177-205) and _get_vectors / _get_mmap_indexer (:94-121,309-336) as a staging source: the
 * file is mapped read-only and its structures (superblock, object headers, symbol-table
 * groups, v1 chunk B-trees, attributes, global heap) are walked in place.  Supported is what
 * such files contain (layout written by _create_ds, :138-165: uncompressed, chunks spanning
 * whole rows); compression, libver='latest' chunk indexes and dense attribute / link storage
 * return FFX_ERR_UNSUPPORTED.  Host only: no device is needed.  A handle is not thread-safe.
 * `path` arguments are '/'-separated object names ("vectors", "quantizer/data/codewords"). */
typedef struct ffx_h5 ffx_h5;
int ffx_h5_open(const char *file_name, ffx_h5 **out);
void ffx_h5_close(ffx_h5 *f);
/* *kind = 0 (no such object), 1 (group) or 2 (dataset)  — `"vectors" in fp`, disk.py:392 */
int ffx_h5_kind(ffx_h5 *f, const char *path, int *kind);
/* Member names of a group / attribute names of an object, each followed by '\n'.  Writes at
 * most `cap` bytes; *needed = the full length (call again with a larger buffer). */
int ffx_h5_list(ffx_h5 *f, const char *path, char *buf, int64_t cap, int64_t *needed);
int ffx_h5_attr_names(ffx_h5 *f, const char *path, char *buf, int64_t cap, int64_t *needed);
/* info[16] = { type class (0 integer, 1 float, 3 fixed-width string, 8 enum), element bytes,
 * signed, rank, dims[8], layout (0 compact, 1 contiguous, 2 chunked), chunk rows (axis 0),
 * bytes per row (= per index of axis 0), 0 }.  fp["vectors"].shape/.dtype/.chunks. */
int ffx_h5_dataset_info(ffx_h5 *f, const char *path, int64_t *info);
/* Copies rows [row0, row0 + nrows) of axis 0, row-major, little-endian as stored, into `dst`
 * (nrows * bytes-per-row bytes).  Chunks that were never written read as zeros. */
int ffx_h5_read_rows(ffx_h5 *f, const char *path, int64_t row0, int64_t nrows, void *dst);
/* The longest run of rows starting at row0 that is contiguous in the file: *rows points INTO
 * the mapping (valid until ffx_h5_close; NULL if that chunk was never written), *nrows is
 * its length.  One HDF5 chunk = one contiguous byte range = one ffx_index_stage_rows call
 * (the reference's memory-mapped read path, disk.py:94-121, as a staging source). */
int ffx_h5_row_span(ffx_h5 *f, const char *path, int64_t row0, const void **rows, int64_t *nrows);
/* One attribute: info[13] = { type class (0 integer, 1 float, 3 string, 8 enum), element
 * bytes, signed, rank, element count, dims[8] }; value bytes into buf (numbers raw
 * little-endian; strings — fixed or variable length — as their characters, several strings
 * separated by NUL), at most `cap` bytes, *needed = full length.  fp.attrs["num_vectors"],
 * dict(fp["quantizer/meta"].attrs) (disk.py:380-391). */
int ffx_h5_attr_read(ffx_h5 *f, const char *path, const char *name, int64_t *info, void *buf,
                     int64_t cap, int64_t *needed);

#ifdef __cplusplus
}
#endif
#endif /* FFX_H_ */
