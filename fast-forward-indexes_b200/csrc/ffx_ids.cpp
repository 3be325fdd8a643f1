// ffx_ids.cpp — host-side id coding of libffx (plain C++17, no CUDA).
//
// The reference keeps two Python dicts per index (`_doc_id_to_idx: dict[str, list[int]]`,
// `_psg_id_to_idx: dict[str, int]`, index/memory.py:46-47,84-95, rebuilt by an O(N) Python
// loop in index/disk.py:408-417) and hashes every candidate id of every call again inside
// pandas merges on string keys (index/base.py:291-298,314; index/util.py:29-41).  Here an id
// dictionary is one open-addressing hash table over a byte arena; ids cross the ABI in Arrow
// layout (int64 offsets into a UTF-8 byte buffer + optional validity bitmap), so a pandas /
// pyarrow string column is looked up in place, without creating Python objects, on all host
// cores.  What goes to the GPU afterwards is integers only.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <memory>
#include <vector>

#include "../../include/ffx.h"

extern "C" int ffx_set_error_message(int code, const char *msg);  // ffx.cu: fills ffx_last_error

struct ffx_dict {
    // slot: high 32 bits = hash tag, low 32 bits = entry index + 1 (0 = empty)
    std::vector<uint64_t> slots;
    uint64_t mask = 0;
    std::vector<int64_t> key_off{0};  // entry e occupies arena[key_off[e], key_off[e+1])
    std::vector<char> arena;
    std::vector<int64_t> values;
};

namespace {

int fail(int code, const std::string &msg) { return ffx_set_error_message(code, msg.c_str()); }

inline uint64_t mix(uint64_t h) {
    h ^= h >> 32;
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    return h;
}

inline uint64_t hash_bytes(const char *p, int64_t n) {
    uint64_t h = 0x243F6A8885A308D3ull ^ (static_cast<uint64_t>(n) * 0xFF51AFD7ED558CCDull);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = mix(h ^ w);
        p += 8;
        n -= 8;
    }
    if (n > 0) {
        uint64_t w = 0;
        memcpy(&w, p, static_cast<size_t>(n));
        h = mix(h ^ w);
    }
    return mix(h);
}

inline bool is_valid(const uint8_t *validity, int64_t bit_offset, int64_t i) {
    if (!validity) return true;
    const int64_t b = bit_offset + i;
    return (validity[b >> 3] >> (b & 7)) & 1;
}

// entry index of the key, or -1
inline int64_t find(const ffx_dict *d, const char *p, int64_t n, uint64_t h) {
    if (d->slots.empty()) return -1;
    const uint64_t tag = h >> 32;
    for (uint64_t s = h & d->mask;; s = (s + 1) & d->mask) {
        const uint64_t slot = d->slots[s];
        if (slot == 0) return -1;
        if ((slot >> 32) == tag) {
            const int64_t e = static_cast<int64_t>(slot & 0xffffffffu) - 1;
            const int64_t b = d->key_off[static_cast<size_t>(e)];
            if (d->key_off[static_cast<size_t>(e) + 1] - b == n && memcmp(d->arena.data() + b, p, static_cast<size_t>(n)) == 0)
                return e;
        }
    }
}

void rehash(ffx_dict *d, uint64_t capacity) {
    d->slots.assign(capacity, 0);
    d->mask = capacity - 1;
    const int64_t n = static_cast<int64_t>(d->values.size());
    for (int64_t e = 0; e < n; e++) {
        const int64_t b = d->key_off[static_cast<size_t>(e)];
        const uint64_t h = hash_bytes(d->arena.data() + b, d->key_off[static_cast<size_t>(e) + 1] - b);
        uint64_t s = h & d->mask;
        while (d->slots[s] != 0) s = (s + 1) & d->mask;
        d->slots[s] = ((h >> 32) << 32) | static_cast<uint64_t>(e + 1);
    }
}

void reserve_for(ffx_dict *d, int64_t extra) {
    const uint64_t need = static_cast<uint64_t>(d->values.size() + extra) * 2 + 16;
    uint64_t cap = d->slots.empty() ? 1024 : d->slots.size();
    while (cap < need) cap <<= 1;
    if (cap != d->slots.size()) rehash(d, cap);
}

// appends the key; caller guarantees it is absent and that capacity is reserved
inline int64_t put(ffx_dict *d, const char *p, int64_t n, uint64_t h, int64_t value) {
    const int64_t e = static_cast<int64_t>(d->values.size());
    d->arena.insert(d->arena.end(), p, p + n);
    d->key_off.push_back(static_cast<int64_t>(d->arena.size()));
    d->values.push_back(value);
    uint64_t s = h & d->mask;
    while (d->slots[s] != 0) s = (s + 1) & d->mask;
    d->slots[s] = ((h >> 32) << 32) | static_cast<uint64_t>(e + 1);
    return e;
}

// Building a dictionary is one dependent cache miss per key (the home slot of its hash).  The
// hashes are therefore computed kAhead keys early and their slots prefetched, which keeps that
// many misses in flight; valid while the table is not resized (callers reserve first).
struct HashAhead {
    static const int kAhead = 16;
    const ffx_dict *d;
    const int64_t *offsets;
    const char *data;
    const uint8_t *validity;
    int64_t bit_offset, n;
    uint64_t ring[kAhead];
    HashAhead(const ffx_dict *d_, const int64_t *o, const char *p, const uint8_t *v, int64_t b, int64_t n_)
        : d(d_), offsets(o), data(p), validity(v), bit_offset(b), n(n_) {
        for (int64_t i = 0; i < std::min<int64_t>(kAhead, n); i++) prime(i);
    }
    void prime(int64_t i) {
        if (!is_valid(validity, bit_offset, i)) return;
        const uint64_t h = hash_bytes(data + offsets[i], offsets[i + 1] - offsets[i]);
        ring[i % kAhead] = h;
        __builtin_prefetch(&d->slots[h & d->mask], 1);
    }
    // hash of key i (valid keys only); primes key i + kAhead
    uint64_t take(int64_t i) {
        const uint64_t h = ring[i % kAhead];
        if (i + kAhead < n) prime(i + kAhead);
        return h;
    }
    void skip(int64_t i) {
        if (i + kAhead < n) prime(i + kAhead);
    }
};

int worker_count(int requested, int64_t n) {
    int t = requested > 0 ? requested : static_cast<int>(std::thread::hardware_concurrency());
    t = std::max(1, std::min(t, 64));
    return static_cast<int>(std::min<int64_t>(t, std::max<int64_t>(1, n / 4096)));
}

template <class F>
void parallel_ranges(int64_t n, int threads, F &&body) {
    if (threads <= 1) {
        body(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t step = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const int64_t lo = t * step, hi = std::min(n, lo + step);
        if (lo >= hi) break;
        pool.emplace_back([&body, lo, hi] { body(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

// hash partitions for the parallel de-duplication passes: about 32k rows each (a table that stays in
// a core's L2), 2^6 .. 2^most of them
inline int partition_bits(int64_t n, int most) {
    int bits = 6;
    while (bits < most && (n >> bits) > 32768) bits++;
    return bits;
}

// n elements WITHOUT value-initialisation: the pages are first touched by the worker threads that
// fill them, not by one serial memset
template <typename T>
struct raw_array {
    std::unique_ptr<T[]> p;
    explicit raw_array(int64_t n) : p(new T[static_cast<size_t>(std::max<int64_t>(n, 1))]) {}
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    T *data() { return p.get(); }
    T *begin() { return p.get(); }
};

// order[j] = the element that comes j-th when n elements are sorted by key(i) ascending, equal
// keys in index order: LSD radix sort, 8-bit digits, digits that never vary are skipped, every
// pass (histogram, prefix, scatter) spread over the host cores.
template <class KeyOf>
int radix_order(int64_t n, int n_threads, int64_t *order, KeyOf &&key_of) {
    if (n == 0) return FFX_OK;
    struct Item {
        uint64_t key;
        uint32_t row;
    };
    raw_array<Item> a(n), b(n);
    const int threads = worker_count(n_threads, n / 16);
    std::vector<uint64_t> seen_or(static_cast<size_t>(threads), 0), seen_and(static_cast<size_t>(threads), ~uint64_t(0));
    const int64_t step = (n + threads - 1) / threads;
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        const size_t t = static_cast<size_t>(lo / step);
        uint64_t o = 0, an = ~uint64_t(0);
        for (int64_t i = lo; i < hi; i++) {
            const uint64_t k = key_of(i);
            a[static_cast<size_t>(i)] = Item{k, static_cast<uint32_t>(i)};
            o |= k;
            an &= k;
        }
        seen_or[t] = o;
        seen_and[t] = an;
    });
    uint64_t any = 0, all = ~uint64_t(0);
    for (int t = 0; t < threads; t++) {
        any |= seen_or[static_cast<size_t>(t)];
        all &= seen_and[static_cast<size_t>(t)];
    }
    const uint64_t varying = any & ~all;  // bits that differ between at least two keys
    std::vector<uint64_t> counts(static_cast<size_t>(threads) * 256);
    Item *src = a.data(), *dst = b.data();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        if (((varying >> shift) & 0xff) == 0) continue;  // every key has the same digit here
        parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
            uint64_t *c = counts.data() + static_cast<size_t>(lo / step) * 256;
            std::fill(c, c + 256, 0);
            for (int64_t i = lo; i < hi; i++) c[(src[i].key >> shift) & 0xff]++;
        });
        uint64_t at = 0;
        for (int d = 0; d < 256; d++)
            for (int t = 0; t < threads; t++) {
                uint64_t &c = counts[static_cast<size_t>(t) * 256 + d];
                const uint64_t mine = (static_cast<int64_t>(t) * step < n) ? c : 0;
                c = at;
                at += mine;
            }
        parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
            uint64_t *c = counts.data() + static_cast<size_t>(lo / step) * 256;
            for (int64_t i = lo; i < hi; i++) dst[c[(src[i].key >> shift) & 0xff]++] = src[i];
        });
        std::swap(src, dst);
    }
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) order[i] = src[i].row;
    });
    return FFX_OK;
}

bool bad_strings(const int64_t *offsets, const char *data, int64_t n) {
    return n < 0 || (n > 0 && (!offsets || (!data && offsets[n] > offsets[0])));
}

}  // namespace

extern "C" {

int ffx_dict_create(ffx_dict **out) {
    if (!out) return fail(FFX_ERR_INVALID, "ffx_dict_create: out is NULL");
    *out = new ffx_dict();
    return FFX_OK;
}

int ffx_dict_destroy(ffx_dict *d) {
    delete d;
    return FFX_OK;
}

int64_t ffx_dict_size(const ffx_dict *d) { return d ? static_cast<int64_t>(d->values.size()) : -1; }
int64_t ffx_dict_key_bytes(const ffx_dict *d) { return d ? static_cast<int64_t>(d->arena.size()) : -1; }

int ffx_dict_insert_ordinal(ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                            int64_t bit_offset, int64_t n, int64_t *out) {
    if (!d || bad_strings(offsets, data, n)) return fail(FFX_ERR_INVALID, "ffx_dict_insert_ordinal: bad arguments");
    if (static_cast<uint64_t>(d->values.size()) + static_cast<uint64_t>(n) > 0xfffffff0ull)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_dict_insert_ordinal: more than 2^32 keys");
    reserve_for(d, n);
    if (n > 0) {  // at most n new keys of offsets[n] - offsets[0] bytes: no regrowth inside the loop
        d->values.reserve(d->values.size() + static_cast<size_t>(n));
        d->key_off.reserve(d->key_off.size() + static_cast<size_t>(n));
        d->arena.reserve(d->arena.size() + static_cast<size_t>(offsets[n] - offsets[0]));
    }
    HashAhead hashes(d, offsets, data, validity, bit_offset, n);
    for (int64_t i = 0; i < n; i++) {
        if (!is_valid(validity, bit_offset, i)) {
            if (out) out[i] = -1;
            hashes.skip(i);
            continue;
        }
        const char *p = data + offsets[i];
        const int64_t len = offsets[i + 1] - offsets[i];
        const uint64_t h = hashes.take(i);
        int64_t e = find(d, p, len, h);
        if (e < 0) e = put(d, p, len, h, static_cast<int64_t>(d->values.size()));
        if (out) out[i] = d->values[static_cast<size_t>(e)];
    }
    return FFX_OK;
}

int ffx_dict_insert_unique(ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                           int64_t bit_offset, int64_t n, int64_t first_value, int dry_run, int64_t *first_dup) {
    if (!d || bad_strings(offsets, data, n) || !first_dup)
        return fail(FFX_ERR_INVALID, "ffx_dict_insert_unique: bad arguments");
    *first_dup = -1;
    if (static_cast<uint64_t>(d->values.size()) + static_cast<uint64_t>(n) > 0xfffffff0ull)
        return fail(FFX_ERR_UNSUPPORTED, "ffx_dict_insert_unique: more than 2^32 keys");
    // insert, remembering how to undo: a duplicate (against the dictionary or inside the batch)
    // leaves the dictionary exactly as it was
    const size_t keep_n = d->values.size(), keep_bytes = d->arena.size();
    reserve_for(d, n);
    if (n > 0) {  // at most n new keys of offsets[n] - offsets[0] bytes: no regrowth inside the loop
        d->values.reserve(d->values.size() + static_cast<size_t>(n));
        d->key_off.reserve(d->key_off.size() + static_cast<size_t>(n));
        d->arena.reserve(d->arena.size() + static_cast<size_t>(offsets[n] - offsets[0]));
    }
    HashAhead hashes(d, offsets, data, validity, bit_offset, n);
    for (int64_t i = 0; i < n; i++) {
        if (!is_valid(validity, bit_offset, i)) {
            hashes.skip(i);
            continue;
        }
        const char *p = data + offsets[i];
        const int64_t len = offsets[i + 1] - offsets[i];
        const uint64_t h = hashes.take(i);
        if (find(d, p, len, h) >= 0) {
            *first_dup = i;
            break;
        }
        put(d, p, len, h, first_value + i);
    }
    if (*first_dup >= 0 || dry_run) {
        if (d->values.size() != keep_n) {
            d->values.resize(keep_n);
            d->key_off.resize(keep_n + 1);
            d->arena.resize(keep_bytes);
            rehash(d, d->slots.size());
        }
        if (*first_dup >= 0) return fail(FFX_ERR_STATE, "ffx_dict_insert_unique: key already present");
    }
    return FFX_OK;
}

int ffx_dict_lookup(const ffx_dict *d, const int64_t *offsets, const char *data, const uint8_t *validity,
                    int64_t bit_offset, int64_t n, int32_t *out, int64_t *first_missing, int n_threads) {
    if (!d || bad_strings(offsets, data, n) || (n > 0 && !out))
        return fail(FFX_ERR_INVALID, "ffx_dict_lookup: bad arguments");
    std::atomic<int64_t> missing{INT64_MAX};
    parallel_ranges(n, worker_count(n_threads, n), [&](int64_t lo, int64_t hi) {
        constexpr int kBatch = 16;  // hash a batch, prefetch its slots, then probe
        uint64_t hs[kBatch];
        int64_t local_missing = INT64_MAX;
        for (int64_t b = lo; b < hi; b += kBatch) {
            const int m = static_cast<int>(std::min<int64_t>(kBatch, hi - b));
            for (int j = 0; j < m; j++) {
                const int64_t i = b + j;
                hs[j] = hash_bytes(data + offsets[i], offsets[i + 1] - offsets[i]);
                if (!d->slots.empty()) __builtin_prefetch(&d->slots[hs[j] & d->mask]);
            }
            for (int j = 0; j < m; j++) {
                const int64_t i = b + j;
                int64_t e = -1;
                if (is_valid(validity, bit_offset, i))
                    e = find(d, data + offsets[i], offsets[i + 1] - offsets[i], hs[j]);
                if (e < 0) {
                    out[i] = -1;
                    local_missing = std::min(local_missing, i);
                } else {
                    out[i] = static_cast<int32_t>(static_cast<uint32_t>(d->values[static_cast<size_t>(e)]));
                }
            }
        }
        int64_t cur = missing.load();
        while (local_missing < cur && !missing.compare_exchange_weak(cur, local_missing)) {
        }
    });
    if (first_missing) *first_missing = missing.load() == INT64_MAX ? -1 : missing.load();
    return FFX_OK;
}

int ffx_dict_clone(const ffx_dict *d, ffx_dict **out) {
    if (!d || !out) return fail(FFX_ERR_INVALID, "ffx_dict_clone: bad arguments");
    *out = new ffx_dict(*d);
    return FFX_OK;
}

int ffx_dict_export(const ffx_dict *d, int64_t *offsets, char *data, int64_t *values) {
    if (!d || !offsets) return fail(FFX_ERR_INVALID, "ffx_dict_export: bad arguments");
    memcpy(offsets, d->key_off.data(), d->key_off.size() * sizeof(int64_t));
    if (data && !d->arena.empty()) memcpy(data, d->arena.data(), d->arena.size());
    if (values && !d->values.empty()) memcpy(values, d->values.data(), d->values.size() * sizeof(int64_t));
    return FFX_OK;
}

int ffx_csr_build(const int64_t *row_doc, int64_t n_rows, int64_t n_docs, int64_t *doc_off, int64_t *doc_rows) {
    if (n_rows < 0 || n_docs < 0 || !doc_off || (n_rows > 0 && !row_doc))
        return fail(FFX_ERR_INVALID, "ffx_csr_build: bad arguments");
    std::fill(doc_off, doc_off + n_docs + 1, 0);
    for (int64_t r = 0; r < n_rows; r++) {
        const int64_t d = row_doc[r];
        if (d >= n_docs) return fail(FFX_ERR_INVALID, "ffx_csr_build: document ordinal out of range");
        if (d >= 0) doc_off[d + 1]++;
    }
    for (int64_t d = 0; d < n_docs; d++) doc_off[d + 1] += doc_off[d];
    if (doc_rows) {
        std::vector<int64_t> cursor(doc_off, doc_off + n_docs);
        for (int64_t r = 0; r < n_rows; r++) {  // increasing r: rows of a document stay in insertion order
            const int64_t d = row_doc[r];
            if (d >= 0) doc_rows[cursor[static_cast<size_t>(d)]++] = r;
        }
    }
    return FFX_OK;
}

// ---- fixed-width byte ids (the HDF5 id columns, index/disk.py:152-165,408-417) -> Arrow layout ----
int ffx_fixed_width_to_arrow(const char *data, int64_t n, int width, int64_t *offsets, char *out,
                             uint8_t *validity, int64_t *n_valid) {
    if (n < 0 || width < 1 || (n > 0 && (!data || !offsets || !out || !validity)) || !n_valid)
        return fail(FFX_ERR_INVALID, "ffx_fixed_width_to_arrow: bad arguments");
    int64_t at = 0, valid = 0;
    if (n > 0) std::fill(validity, validity + (n + 7) / 8, 0);
    for (int64_t i = 0; i < n; i++) {
        const char *p = data + i * static_cast<int64_t>(width);
        int len = width;
        while (len > 0 && p[len - 1] == 0) len--;  // NUL padding on the right (numpy 'S' semantics)
        offsets[i] = at;
        if (len > 0) {
            memcpy(out + at, p, static_cast<size_t>(len));
            at += len;
            validity[i >> 3] |= static_cast<uint8_t>(1u << (i & 7));
            valid++;
        }
    }
    if (offsets) offsets[n] = at;
    *n_valid = valid;
    return FFX_OK;
}

// ---- Ranking.__init__ on integer codes (ranking.py:95-98,115-117) ------------------------------
int ffx_first_repeat(const int64_t *keys, int64_t n, int64_t *first) {
    if (n < 0 || !first || (n > 0 && !keys)) return fail(FFX_ERR_INVALID, "ffx_first_repeat: bad arguments");
    *first = -1;
    if (n < 2) return FFX_OK;
    if (n >= (int64_t(1) << 32)) return fail(FFX_ERR_UNSUPPORTED, "ffx_first_repeat: more than 2^32 keys");
    // equal keys share a hash: rows are partitioned by its top bits, every partition is checked in its
    // own cache-sized table.  Reports the lowest row that repeats an earlier one INSIDE its partition,
    // minimised over partitions — the row pandas' duplicated() would flag first.
    const int kBits = partition_bits(n, 10), kParts = 1 << kBits;
    const int threads = worker_count(0, n / 4096);
    const int64_t step = (n + threads - 1) / threads;
    auto hash_of = [&](int64_t i) { return mix(static_cast<uint64_t>(keys[i]) * 0x9E3779B97F4A7C15ull); };
    std::vector<int64_t> counts(static_cast<size_t>(threads) * kParts, 0);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *mine = counts.data() + static_cast<size_t>(lo / step) * kParts;
        for (int64_t i = lo; i < hi; i++) mine[hash_of(i) >> (64 - kBits)]++;
    });
    std::vector<int64_t> begin(kParts + 1, 0);
    int64_t at = 0;
    for (int p = 0; p < kParts; p++) {
        begin[static_cast<size_t>(p)] = at;
        for (int t = 0; t < threads; t++) {
            int64_t &c = counts[static_cast<size_t>(t) * kParts + p];
            const int64_t mine = c;
            c = at;
            at += mine;
        }
    }
    begin[kParts] = at;
    struct Item {
        int64_t key;
        uint32_t row;
    };
    raw_array<Item> items(n);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *cursor = counts.data() + static_cast<size_t>(lo / step) * kParts;
        for (int64_t i = lo; i < hi; i++) items[static_cast<size_t>(cursor[hash_of(i) >> (64 - kBits)]++)] = Item{keys[i], static_cast<uint32_t>(i)};
    });
    std::atomic<int> next{0};
    std::atomic<int64_t> lowest{INT64_MAX};
    auto hash_key = [](int64_t k) { return mix(static_cast<uint64_t>(k) * 0x9E3779B97F4A7C15ull); };
    auto check = [&]() {
        std::vector<uint32_t> slots;  // position inside the partition + 1, 0 = free
        for (int p = next.fetch_add(1); p < kParts; p = next.fetch_add(1)) {
            const int64_t b = begin[static_cast<size_t>(p)], e = begin[static_cast<size_t>(p) + 1];
            uint64_t cap = 64;
            while (cap < static_cast<uint64_t>(e - b) * 2) cap <<= 1;
            slots.assign(cap, 0);
            const uint64_t mask = cap - 1;
            const Item *it = items.data() + b;
            const int64_t m = e - b;
            const int kAhead = 8;
            for (int64_t j = 0; j < m; j++) {
                if (j + kAhead < m) __builtin_prefetch(&slots[hash_key(it[j + kAhead].key) & mask], 1);
                bool repeat = false;
                for (uint64_t s = hash_key(it[j].key) & mask;; s = (s + 1) & mask) {
                    if (slots[s] == 0) {
                        slots[s] = static_cast<uint32_t>(j + 1);
                        break;
                    }
                    if (it[slots[s] - 1].key == it[j].key) {
                        repeat = true;
                        break;
                    }
                }
                if (repeat) {  // rows of a partition come in increasing order: this is its first repeat
                    const int64_t r = it[j].row;
                    int64_t cur = lowest.load();
                    while (r < cur && !lowest.compare_exchange_weak(cur, r)) {
                    }
                    break;
                }
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(check);
        check();
        for (auto &th : pool) th.join();
    }
    if (lowest.load() != INT64_MAX) *first = lowest.load();
    return FFX_OK;
}

int ffx_ranking_order(const int32_t *q_rank, const float *score, int64_t n, int64_t *order, int n_threads) {
    if (n < 0 || (n > 0 && (!q_rank || !score || !order)))
        return fail(FFX_ERR_INVALID, "ffx_ranking_order: bad arguments");
    if (n >= (int64_t(1) << 32)) return fail(FFX_ERR_UNSUPPORTED, "ffx_ranking_order: more than 2^32 rows");
    if (n == 0) return FFX_OK;
    // key: query rank ascending, then score DEscending (-0.0 == +0.0, NaN after every number),
    // the incoming row order breaking ties
    auto desc_key = [&](int64_t i) {
        float f = score[i];
        uint32_t u;
        if (f == 0.0f) f = 0.0f;  // drops the sign of -0.0
        memcpy(&u, &f, 4);
        const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending in f
        return (f != f) ? 0xffffffffu : ~asc;
    };
    const int threads = worker_count(n_threads, n / 16);
    const int64_t step = (n + threads - 1) / threads;
    std::vector<int32_t> top(static_cast<size_t>(threads), 0), low(static_cast<size_t>(threads), 0);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int32_t mx = 0, mn = 0;
        for (int64_t i = lo; i < hi; i++) {
            mx = std::max(mx, q_rank[i]);
            mn = std::min(mn, q_rank[i]);
        }
        top[static_cast<size_t>(lo / step)] = mx;
        low[static_cast<size_t>(lo / step)] = mn;
    });
    if (*std::min_element(low.begin(), low.end()) < 0) return fail(FFX_ERR_INVALID, "ffx_ranking_order: negative query rank");
    const int64_t n_ranks = static_cast<int64_t>(*std::max_element(top.begin(), top.end())) + 1;
    if (n_ranks * threads > (int64_t(1) << 24))  // too many queries for per-thread histograms
        return radix_order(n, n_threads, order, [&](int64_t i) {
            return (static_cast<uint64_t>(static_cast<uint32_t>(q_rank[i])) << 32) | desc_key(i);
        });
    // one stable counting pass groups the rows by query; every query's block (its rows, in incoming
    // order) is then sorted on its own, (score key, row) packed in one word
    std::vector<int64_t> counts(static_cast<size_t>(threads * n_ranks), 0);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *mine = counts.data() + (lo / step) * n_ranks;
        for (int64_t i = lo; i < hi; i++) mine[q_rank[i]]++;
    });
    std::vector<int64_t> begin(static_cast<size_t>(n_ranks) + 1, 0);
    int64_t at = 0;
    for (int64_t r = 0; r < n_ranks; r++) {
        begin[static_cast<size_t>(r)] = at;
        for (int t = 0; t < threads; t++) {
            int64_t &c = counts[static_cast<size_t>(t * n_ranks + r)];
            const int64_t mine = c;
            c = at;
            at += mine;
        }
    }
    begin[static_cast<size_t>(n_ranks)] = at;
    raw_array<uint64_t> items(n);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *cursor = counts.data() + (lo / step) * n_ranks;
        for (int64_t i = lo; i < hi; i++)
            items[static_cast<size_t>(cursor[q_rank[i]]++)] = (static_cast<uint64_t>(desc_key(i)) << 32) | static_cast<uint32_t>(i);
    });
    std::atomic<int64_t> next{0};
    const int64_t kChunk = 16;
    auto sort_blocks = [&]() {
        for (int64_t r0 = next.fetch_add(kChunk); r0 < n_ranks; r0 = next.fetch_add(kChunk))
            for (int64_t r = r0; r < std::min(n_ranks, r0 + kChunk); r++) {
                const int64_t b = begin[static_cast<size_t>(r)], e = begin[static_cast<size_t>(r) + 1];
                std::sort(items.begin() + b, items.begin() + e);
                for (int64_t i = b; i < e; i++) order[i] = static_cast<int64_t>(items[static_cast<size_t>(i)] & 0xffffffffu);
            }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(sort_blocks);
        sort_blocks();
        for (auto &th : pool) th.join();
    }
    return FFX_OK;
}

int ffx_order_u64(const uint64_t *keys, int64_t n, int64_t *order, int n_threads) {
    if (n < 0 || (n > 0 && (!keys || !order))) return fail(FFX_ERR_INVALID, "ffx_order_u64: bad arguments");
    if (n >= (int64_t(1) << 32)) return fail(FFX_ERR_UNSUPPORTED, "ffx_order_u64: more than 2^32 rows");
    return radix_order(n, n_threads, order, [&](int64_t i) { return keys[i]; });
}

int ffx_match_keys(const int64_t *have, int64_t n_have, const int64_t *want, int64_t n_want, int64_t *pos) {
    if (n_have < 0 || n_want < 0 || (n_have > 0 && !have) || (n_want > 0 && (!want || !pos)))
        return fail(FFX_ERR_INVALID, "ffx_match_keys: bad arguments");
    uint64_t cap = 1024;
    while (cap < static_cast<uint64_t>(n_have) * 2) cap <<= 1;
    const uint64_t mask = cap - 1;
    std::vector<int64_t> slots(cap, -1);
    const int kAhead = 16;
    auto home = [&](int64_t key) { return mix(static_cast<uint64_t>(key) * 0x9E3779B97F4A7C15ull) & mask; };
    for (int64_t i = 0; i < n_have; i++) {
        if (i + kAhead < n_have) __builtin_prefetch(&slots[home(have[i + kAhead])], 1);
        uint64_t s = home(have[i]);
        while (slots[s] >= 0 && have[slots[s]] != have[i]) s = (s + 1) & mask;
        if (slots[s] < 0) slots[s] = i;  // the first of equal keys stays
    }
    for (int64_t i = 0; i < n_want; i++) {
        if (i + kAhead < n_want) __builtin_prefetch(&slots[home(want[i + kAhead])], 0);
        uint64_t s = home(want[i]);
        while (slots[s] >= 0 && have[slots[s]] != want[i]) s = (s + 1) & mask;
        pos[i] = slots[s];
    }
    return FFX_OK;
}

// ---- rankings on integer codes: the result side of Index.__call__ / rerank -------------------
int ffx_lut_gather(const int32_t *lut, int64_t n_lut, const int32_t *codes, int64_t n, int32_t *out,
                   int64_t *first_negative, int n_threads) {
    if (n < 0 || n_lut < 0 || !first_negative || (n > 0 && (!lut || !codes || !out)))
        return fail(FFX_ERR_INVALID, "ffx_lut_gather: bad arguments");
    const int threads = worker_count(n_threads, n / 16);
    std::vector<int64_t> first(static_cast<size_t>(std::max(threads, 1)), -1);
    std::atomic<int> bad{0};
    const int64_t step = threads > 0 ? (n + threads - 1) / threads : n;
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t miss = -1;
        for (int64_t i = lo; i < hi; i++) {
            const int32_t c = codes[i];
            if (c < 0 || c >= n_lut) {
                bad.store(1, std::memory_order_relaxed);
                out[i] = -1;
                continue;
            }
            const int32_t v = lut[c];
            out[i] = v;
            if (v < 0 && miss < 0) miss = i;
        }
        first[static_cast<size_t>(step > 0 ? lo / step : 0)] = miss;
    });
    if (bad.load()) return fail(FFX_ERR_INVALID, "ffx_lut_gather: code outside the table");
    *first_negative = -1;
    for (int64_t m : first)
        if (m >= 0 && (*first_negative < 0 || m < *first_negative)) *first_negative = m;
    return FFX_OK;
}

int ffx_topk_gather(const int32_t *pos, const float *score, int64_t nq, int64_t k, int64_t keep,
                    const int64_t *src_off, const int32_t *src_code, int64_t *out_off, int32_t *out_code,
                    float *out_score, int64_t *n_ties, uint8_t *straddle, int n_threads) {
    if (nq < 0 || k < 0 || keep < 0 || keep > k || !out_off || (nq > 0 && k > 0 && (!pos || !score || !src_off)) ||
        (out_code && !src_code))
        return fail(FFX_ERR_INVALID, "ffx_topk_gather: bad arguments");
    const int threads = worker_count(n_threads, nq * std::max<int64_t>(k, 1) / 16);
    // pass 1: entries kept per query (lists are padded with pos = -1 at the end)
    out_off[0] = 0;
    parallel_ranges(nq, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t q = lo; q < hi; q++) {
            const int32_t *p = pos + q * k;
            int64_t c = 0;
            while (c < keep && p[c] >= 0) c++;
            out_off[q + 1] = c;
            if (straddle)
                straddle[q] = keep > 0 && keep < k && c == keep && p[keep] >= 0 &&
                              score[q * k + keep - 1] == score[q * k + keep];
        }
    });
    for (int64_t q = 0; q < nq; q++) out_off[q + 1] += out_off[q];
    // pass 2: compact rows
    std::atomic<int64_t> ties{0};
    std::atomic<int> bad{0};
    parallel_ranges(nq, threads, [&](int64_t lo, int64_t hi) {
        int64_t my_ties = 0;
        for (int64_t q = lo; q < hi; q++) {
            const int32_t *p = pos + q * k;
            const float *s = score + q * k;
            const int64_t base = src_off[q], width = src_off[q + 1] - base, at = out_off[q];
            const int64_t c = out_off[q + 1] - at;
            for (int64_t j = 0; j < c; j++) {
                if (p[j] >= width) {
                    bad.store(1, std::memory_order_relaxed);
                    continue;
                }
                if (out_code) out_code[at + j] = src_code[base + p[j]];
                if (out_score) out_score[at + j] = s[j];
                if (j > 0 && s[j] == s[j - 1]) my_ties++;
            }
        }
        ties.fetch_add(my_ties, std::memory_order_relaxed);
    });
    if (bad.load()) return fail(FFX_ERR_INVALID, "ffx_topk_gather: a position lies outside its query's block");
    if (n_ties) *n_ties = ties.load();
    return FFX_OK;
}

int ffx_tie_runs(const float *score, const int64_t *off, int64_t nq, int64_t cap, int64_t *run_start,
                 int64_t *run_len, int64_t *n_runs, int n_threads) {
    if (nq < 0 || cap < 0 || !n_runs || (nq > 0 && (!score || !off)) || (cap > 0 && (!run_start || !run_len)))
        return fail(FFX_ERR_INVALID, "ffx_tie_runs: bad arguments");
    const int64_t n = nq > 0 ? off[nq] : 0;
    const int threads = worker_count(n_threads, n / 16);
    std::vector<std::vector<int64_t>> found(static_cast<size_t>(std::max(threads, 1)));
    const int64_t step = threads > 0 ? (nq + threads - 1) / threads : nq;
    parallel_ranges(nq, threads, [&](int64_t lo, int64_t hi) {
        std::vector<int64_t> &mine = found[static_cast<size_t>(step > 0 ? lo / step : 0)];
        for (int64_t q = lo; q < hi; q++) {
            int64_t i = off[q];
            const int64_t end = off[q + 1];
            while (i + 1 < end) {
                if (score[i + 1] != score[i]) {
                    i++;
                    continue;
                }
                int64_t j = i + 1;
                while (j + 1 < end && score[j + 1] == score[i]) j++;
                mine.push_back(i);
                mine.push_back(j - i + 1);
                i = j + 1;
            }
        }
    });
    int64_t total = 0;
    for (const auto &v : found) total += static_cast<int64_t>(v.size() / 2);
    *n_runs = total;
    if (total > cap) return FFX_OK;  // the caller retries with room for *n_runs
    int64_t at = 0;
    for (const auto &v : found)
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            run_start[at] = v[i];
            run_len[at++] = v[i + 1];
        }
    return FFX_OK;
}

// ---- factorising a string column on all cores --------------------------------------------------
// `Ranking.__init__` needs ONE integer per id string, equal strings <-> equal integers; which
// integer is irrelevant (the distinct strings are kept beside the codes).  The dictionary above
// assigns ordinals in order of first appearance and is therefore serial; here rows are
// partitioned by the top bits of their hash (64 partitions), every partition is deduplicated in
// its own small open-addressing table by one thread at a time, and a string's code is
// (distinct strings of the lower partitions) + (its rank of first appearance inside its partition).
}  // extern "C"

struct ffx_factor {
    const int64_t *offsets = nullptr;
    const char *data = nullptr;
    std::vector<std::vector<uint32_t>> first_row;  // per partition: the row a distinct string was first seen in
    std::vector<int64_t> base;                     // per partition: code of its first distinct string
    int64_t n_keys = 0, key_bytes = 0;
};


extern "C" {

int ffx_factorize(const int64_t *offsets, const char *data, int64_t n, int32_t *codes, ffx_factor **out, int64_t *n_keys,
                  int64_t *key_bytes, int n_threads) {
    if (!out || !n_keys || !key_bytes || bad_strings(offsets, data, n) || (n > 0 && !codes))
        return fail(FFX_ERR_INVALID, "ffx_factorize: bad arguments");
    if (n >= (int64_t(1) << 31)) return fail(FFX_ERR_UNSUPPORTED, "ffx_factorize: more than 2^31 rows");
    const int kFactorBits = partition_bits(n, 8), kFactorParts = 1 << kFactorBits;
    ffx_factor *f = new ffx_factor();
    f->offsets = offsets;
    f->data = data;
    f->first_row.resize(kFactorParts);
    f->base.assign(kFactorParts + 1, 0);
    *out = f;
    const int threads = worker_count(n_threads, n / 8);
    const int64_t step = (n + threads - 1) / std::max(threads, 1);
    // pass 1: hash every row, count rows per (thread, partition)
    raw_array<uint32_t> hash_lo(n);
    raw_array<uint8_t> part(n);
    std::vector<int64_t> counts(static_cast<size_t>(threads) * kFactorParts, 0);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *mine = counts.data() + static_cast<size_t>(step > 0 ? lo / step : 0) * kFactorParts;
        for (int64_t i = lo; i < hi; i++) {
            const uint64_t h = hash_bytes(data + offsets[i], offsets[i + 1] - offsets[i]);
            hash_lo[static_cast<size_t>(i)] = static_cast<uint32_t>(h);
            const uint8_t p = static_cast<uint8_t>(h >> (64 - kFactorBits));
            part[static_cast<size_t>(i)] = p;
            mine[p]++;
        }
    });
    // pass 2: rows grouped by partition, in increasing row order inside a partition
    std::vector<int64_t> part_begin(kFactorParts + 1, 0);
    {
        int64_t at = 0;
        for (int p = 0; p < kFactorParts; p++) {
            part_begin[static_cast<size_t>(p)] = at;
            for (int t = 0; t < threads; t++) {
                int64_t &c = counts[static_cast<size_t>(t) * kFactorParts + p];
                const int64_t mine = c;
                c = at;
                at += mine;
            }
        }
        part_begin[kFactorParts] = at;
    }
    raw_array<uint32_t> rows(n);
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        int64_t *cursor = counts.data() + static_cast<size_t>(step > 0 ? lo / step : 0) * kFactorParts;
        for (int64_t i = lo; i < hi; i++) rows[static_cast<size_t>(cursor[part[static_cast<size_t>(i)]]++)] = static_cast<uint32_t>(i);
    });
    // pass 3: one table per partition; partitions are handed out dynamically
    std::atomic<int> next{0};
    auto dedupe = [&]() {
        std::vector<uint32_t> slots;
        for (int p = next.fetch_add(1); p < kFactorParts; p = next.fetch_add(1)) {
            const int64_t b = part_begin[static_cast<size_t>(p)], e = part_begin[static_cast<size_t>(p) + 1];
            uint64_t cap = 64;
            while (cap < static_cast<uint64_t>(e - b) * 2) cap <<= 1;
            slots.assign(cap, 0);  // local code + 1
            const uint64_t mask = cap - 1;
            std::vector<uint32_t> &first = f->first_row[static_cast<size_t>(p)];
            const int kAhead = 8;
            for (int64_t j = b; j < e; j++) {
                if (j + kAhead < e) __builtin_prefetch(&slots[mix(hash_lo[rows[static_cast<size_t>(j + kAhead)]]) & mask], 1);
                const uint32_t r = rows[static_cast<size_t>(j)];
                const char *s = data + offsets[r];
                const int64_t len = offsets[r + 1] - offsets[r];
                uint64_t at = mix(hash_lo[r]) & mask;
                for (;; at = (at + 1) & mask) {
                    const uint32_t v = slots[at];
                    if (v == 0) {
                        first.push_back(r);
                        slots[at] = static_cast<uint32_t>(first.size());
                        codes[r] = static_cast<int32_t>(first.size() - 1);
                        break;
                    }
                    const uint32_t o = first[v - 1];
                    if (offsets[o + 1] - offsets[o] == len && memcmp(data + offsets[o], s, static_cast<size_t>(len)) == 0) {
                        codes[r] = static_cast<int32_t>(v - 1);
                        break;
                    }
                }
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(dedupe);
        dedupe();
        for (auto &th : pool) th.join();
    }
    // pass 4: partition-local codes -> global codes
    for (int p = 0; p < kFactorParts; p++) {
        f->base[static_cast<size_t>(p) + 1] = f->base[static_cast<size_t>(p)] + static_cast<int64_t>(f->first_row[static_cast<size_t>(p)].size());
        for (uint32_t r : f->first_row[static_cast<size_t>(p)]) f->key_bytes += offsets[r + 1] - offsets[r];
    }
    f->n_keys = f->base[kFactorParts];
    if (f->n_keys >= (int64_t(1) << 31)) return fail(FFX_ERR_UNSUPPORTED, "ffx_factorize: more than 2^31 distinct strings");
    parallel_ranges(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) codes[i] += static_cast<int32_t>(f->base[part[static_cast<size_t>(i)]]);
    });
    *n_keys = f->n_keys;
    *key_bytes = f->key_bytes;
    return FFX_OK;
}

int ffx_factor_export(const ffx_factor *f, int64_t *key_offsets, char *key_data) {
    if (!f || !key_offsets || (f->key_bytes > 0 && !key_data)) return fail(FFX_ERR_INVALID, "ffx_factor_export: bad arguments");
    int64_t code = 0, at = 0;
    for (size_t p = 0; p < f->first_row.size(); p++)
        for (uint32_t r : f->first_row[p]) {
            const int64_t len = f->offsets[r + 1] - f->offsets[r];
            key_offsets[code++] = at;
            memcpy(key_data + at, f->data + f->offsets[r], static_cast<size_t>(len));
            at += len;
        }
    key_offsets[code] = at;
    return FFX_OK;
}

void ffx_factor_free(ffx_factor *f) { delete f; }

}  // extern "C"
