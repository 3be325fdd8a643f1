// ffx_h5.cpp — read-only access to the HDF5 files that OnDiskIndex writes, without libhdf5.
//
// The reference keeps an index as one HDF5 file (src/fast_forward/index/disk.py:83-85,138-165):
// root attributes `num_vectors` / `ff_version`; `vectors` (capacity, dim), chunked
// (chunk_size, dim), uncompressed; `doc_ids` / `psg_ids` fixed-width byte strings, chunked;
// a `quantizer/{meta,attributes,data}` group tree whose state sits in attributes and small
// contiguous datasets (disk.py:123-136).  Because nothing is compressed, every chunk of
// `vectors` is one contiguous row-major byte range of the file — the natural unit for the
// pinned staging buffers.  This unit maps the file and walks exactly the structures such a
// file is made of, as the HDF5 File Format Specification (version 3.0) lays them out:
//
//   superblock v0/v1 (root symbol-table entry) and v2/v3 (root object header address)
//   object headers v1 and v2 ("OHDR"/"OCHK"), with continuation blocks
//   groups: symbol table message -> v1 B-tree ("TREE", type 0) -> "SNOD" nodes + local "HEAP";
//           compact link messages (new-style groups with few links)
//   datasets: dataspace v1/v2, datatype (fixed-point, float, string, enum, vlen string),
//             data layout v3: compact, contiguous, chunked via v1 B-tree ("TREE", type 1)
//   attributes v1/v2/v3, variable-length strings through global heap collections ("GCOL")
//
// Anything else (filters/compression, layout v4 chunk indexes of libver='latest', dense link or
// attribute storage, shared messages) is reported as FFX_ERR_UNSUPPORTED rather than guessed at.
// Every read is bounds-checked against the mapping: a truncated or corrupt file yields an error
// message, never a fault.
//
// Host only; no CUDA.  Rows leave through ffx_h5_read_rows (copy) or ffx_h5_row_span (pointer
// into the mapping, handed to ffx_index_stage_rows so that the only host copy is the one into
// the pinned buffer).

#include "../../include/ffx.h"

#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

extern "C" int ffx_set_error_message(int code, const char *msg);  // ffx.cu: fills ffx_last_error

namespace {

const uint64_t UNDEF = ~uint64_t(0);

struct Unsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Msg {
    int type;
    int flags;
    uint64_t off;  // file offset of the message body
    uint64_t len;
};

struct Dtype {
    int cls = -1;       // HDF5 datatype class: 0 fixed-point, 1 float, 3 string, 8 enum, 9 vlen
    int sign = 0;       // fixed-point / enum: two's complement signed
    uint32_t size = 0;  // bytes per element
    bool vlen_string = false;
    bool big_endian = false;
};

struct Space {
    int rank = 0;
    bool null_space = false;
    uint64_t dims[8] = {0};
    uint64_t count() const {
        if (null_space) return 0;
        uint64_t n = 1;
        for (int i = 0; i < rank; i++) {
            if (dims[i] != 0 && n > ~uint64_t(0) / dims[i]) throw std::runtime_error("dataspace extent overflows 64 bits");
            n *= dims[i];
        }
        return n;
    }
};

struct Dataset {
    Dtype type;
    Space space;
    int layout = -1;  // 0 compact, 1 contiguous, 2 chunked
    uint64_t addr = UNDEF;      // contiguous: data; compact: body offset inside the header
    uint64_t bytes = 0;
    uint64_t chunk[8] = {0};
    uint64_t row_bytes = 0;     // bytes of one index along axis 0
    std::vector<uint64_t> chunk_addr;  // chunked: file address of chunk i along axis 0 (UNDEF = never written)
};

struct Attr {
    std::string name;
    Dtype type;
    Space space;
    uint64_t off = 0;  // file offset of the raw value
    uint64_t len = 0;
};

struct Object {
    uint64_t addr = 0;
    std::vector<Msg> msgs;
    bool is_dataset = false;
    bool parsed_dataset = false;
    Dataset ds;
    bool parsed_attrs = false;
    std::vector<Attr> attrs;
};

}  // namespace

struct ffx_h5 {
    int fd = -1;
    const uint8_t *map = nullptr;
    uint64_t size = 0;
    uint64_t base = 0;  // superblock base address: every file address is relative to it
    int O = 8, L = 8;   // size of offsets / size of lengths
    uint64_t root = UNDEF;
    std::map<std::string, Object> cache;

    const uint8_t *at(uint64_t off, uint64_t len) const {
        if (off > size || len > size - off) throw std::runtime_error("HDF5 structure points past the end of the file");
        return map + off;
    }
    const uint8_t *addr(uint64_t a, uint64_t len) const {
        if (a == UNDEF || a > UNDEF - base) throw std::runtime_error("undefined address in an HDF5 structure");
        return at(a + base, len);
    }
};

namespace {

uint64_t le(const uint8_t *p, int n) {
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}

uint64_t mul_checked(uint64_t a, uint64_t b) {  // sizes come from the file: never let them wrap
    if (a != 0 && b > UNDEF / a) throw std::runtime_error("size in an HDF5 structure overflows 64 bits");
    return a * b;
}

uint64_t le_addr(const uint8_t *p, int n) {  // all-ones of any width = the undefined address
    uint64_t v = le(p, n);
    if (n < 8 && v == ((uint64_t(1) << (8 * n)) - 1)) return UNDEF;
    return v;
}

// ---- object headers -----------------------------------------------------------------------
void read_header(const ffx_h5 &f, uint64_t a, std::vector<Msg> &out) {
    const uint8_t *p = f.addr(a, 16);
    std::vector<std::pair<uint64_t, uint64_t>> blocks;  // absolute offsets: [begin, end)
    if (p[0] == 1) {
        // v1: version, reserved, #messages (2), reference count (4), header size (4), pad to 8
        const unsigned n_msgs = (unsigned)le(p + 2, 2);
        const uint64_t first = le(p + 8, 4);
        blocks.push_back({a + f.base + 16, a + f.base + 16 + first});
        for (size_t b = 0; b < blocks.size(); b++) {
            uint64_t cur = blocks[b].first;
            const uint64_t end = blocks[b].second;
            f.at(cur, end - cur);
            while (cur + 8 <= end && out.size() < n_msgs) {
                const uint8_t *m = f.map + cur;
                Msg msg{(int)le(m, 2), m[4], cur + 8, le(m + 2, 2)};
                if (msg.off + msg.len > end) throw std::runtime_error("object header message overruns its block");
                if (msg.type == 0x10) {
                    const uint8_t *c = f.at(msg.off, f.O + f.L);
                    const uint64_t ca = le_addr(c, f.O), cl = le(c + f.O, f.L);
                    f.addr(ca, cl);
                    blocks.push_back({ca + f.base, ca + f.base + cl});
                }
                out.push_back(msg);
                cur = msg.off + msg.len;
            }
            if (blocks.size() > 4096) throw std::runtime_error("object header continuation loop");
        }
        return;
    }
    if (memcmp(p, "OHDR", 4) == 0 && p[4] == 2) {
        const int flags = p[5];
        uint64_t cur = 6;
        if (flags & 0x20) cur += 16;  // access, modification, change, birth times
        if (flags & 0x10) cur += 4;   // max compact / min dense attribute counts
        const int w = 1 << (flags & 3);
        const uint8_t *q = f.addr(a + cur, w);
        const uint64_t first = le(q, w);
        cur += w;
        const int mh = 4 + ((flags & 4) ? 2 : 0);  // type(1) size(2) flags(1) [creation order(2)]
        blocks.push_back({a + f.base + cur, a + f.base + cur + first});
        for (size_t b = 0; b < blocks.size(); b++) {
            uint64_t c = blocks[b].first;
            const uint64_t end = blocks[b].second;
            f.at(c, end - c);
            while (c + mh <= end) {
                const uint8_t *m = f.map + c;
                Msg msg{m[0], m[3], c + mh, le(m + 1, 2)};
                if (msg.off + msg.len > end) throw std::runtime_error("object header message overruns its block");
                if (msg.type == 0x10) {
                    const uint8_t *cc = f.at(msg.off, f.O + f.L);
                    const uint64_t ca = le_addr(cc, f.O), cl = le(cc + f.O, f.L);
                    const uint8_t *blk = f.addr(ca, cl);
                    if (cl < 8 || memcmp(blk, "OCHK", 4) != 0) throw std::runtime_error("bad object header continuation block");
                    blocks.push_back({ca + f.base + 4, ca + f.base + cl - 4});  // signature .. checksum
                }
                out.push_back(msg);
                c = msg.off + msg.len;
            }
            if (blocks.size() > 4096) throw std::runtime_error("object header continuation loop");
        }
        return;
    }
    throw std::runtime_error("not an object header (version 1 or 2) where one was expected");
}

// ---- groups ---------------------------------------------------------------------------------
typedef std::vector<std::pair<std::string, uint64_t>> Links;

std::string heap_string(const ffx_h5 &f, uint64_t heap_data, uint64_t heap_size, uint64_t off) {
    if (off >= heap_size) throw std::runtime_error("link name lies outside the local heap");
    const uint8_t *s = f.addr(heap_data + off, 1);
    const uint64_t room = heap_size - off;
    f.addr(heap_data + off, room);
    const void *z = memchr(s, 0, room);
    if (!z) throw std::runtime_error("unterminated link name in the local heap");
    return std::string((const char *)s, (const uint8_t *)z - s);
}

void walk_group_tree(const ffx_h5 &f, uint64_t node, uint64_t heap_data, uint64_t heap_size, Links &out, int depth) {
    if (depth > 32) throw std::runtime_error("group B-tree too deep");
    const uint8_t *p = f.addr(node, 8 + 2 * f.O);
    if (memcmp(p, "TREE", 4) != 0 || p[4] != 0) throw std::runtime_error("bad group B-tree node");
    const int level = p[5];
    const unsigned used = (unsigned)le(p + 6, 2);
    const uint64_t body = node + 8 + 2 * f.O;
    f.addr(body, (uint64_t)used * (f.L + f.O) + f.L);
    for (unsigned i = 0; i < used; i++) {
        const uint64_t child = le_addr(f.addr(body + (uint64_t)i * (f.L + f.O) + f.L, f.O), f.O);
        if (level > 0) {
            walk_group_tree(f, child, heap_data, heap_size, out, depth + 1);
            continue;
        }
        const uint8_t *s = f.addr(child, 8);
        if (memcmp(s, "SNOD", 4) != 0) throw std::runtime_error("bad symbol table node");
        const unsigned n = (unsigned)le(s + 6, 2);
        const uint64_t entry = 2 * f.O + 24;  // name offset, header address, cache type, reserved, scratch
        f.addr(child + 8, n * entry);
        for (unsigned k = 0; k < n; k++) {
            const uint8_t *e = f.addr(child + 8 + k * entry, entry);
            out.push_back({heap_string(f, heap_data, heap_size, le(e, f.O)), le_addr(e + f.O, f.O)});
        }
    }
}

Links group_links(const ffx_h5 &f, const Object &g) {
    Links out;
    for (const Msg &m : g.msgs) {
        if (m.type == 0x11) {  // symbol table: B-tree address, local heap address
            const uint8_t *p = f.at(m.off, 2 * f.O);
            const uint64_t tree = le_addr(p, f.O), heap = le_addr(p + f.O, f.O);
            const uint8_t *h = f.addr(heap, 8 + 2 * f.L + f.O);
            if (memcmp(h, "HEAP", 4) != 0) throw std::runtime_error("bad local heap");
            const uint64_t hsize = le(h + 8, f.L), hdata = le_addr(h + 8 + 2 * f.L, f.O);
            walk_group_tree(f, tree, hdata, hsize, out, 0);
        } else if (m.type == 0x06) {  // link message
            const uint8_t *p = f.at(m.off, m.len);
            const uint8_t *end = p + m.len;
            if (m.len < 3 || p[0] != 1) throw std::runtime_error("bad link message");
            const int fl = p[1];
            const uint8_t *q = p + 2;
            int kind = 0;
            if (fl & 8) kind = *q++;
            if (fl & 4) q += 8;  // creation order
            if (fl & 16) q += 1; // character set
            const int w = 1 << (fl & 3);
            if (q + w > end) throw std::runtime_error("bad link message");
            const uint64_t nlen = le(q, w);
            q += w;
            if (nlen > (uint64_t)(end - q)) throw std::runtime_error("bad link message");
            std::string name((const char *)q, nlen);
            q += nlen;
            if (kind != 0) continue;  // soft / external links: not something an index file holds
            if (q + f.O > end) throw std::runtime_error("bad link message");
            out.push_back({name, le_addr(q, f.O)});
        } else if (m.type == 0x02) {  // link info: dense storage lives in a fractal heap
            const uint8_t *p = f.at(m.off, m.len);
            uint64_t c = 2 + ((p[1] & 1) ? 8 : 0);
            if (c + f.O <= m.len && le_addr(p + c, f.O) != UNDEF)
                throw Unsupported("group with dense link storage (fractal heap)");
        }
    }
    return out;
}

bool is_group(const Object &o) {
    for (const Msg &m : o.msgs)
        if (m.type == 0x11 || m.type == 0x02 || m.type == 0x06 || m.type == 0x0A) return true;
    return false;
}

Object &open_object(ffx_h5 &f, const std::string &path) {
    std::string norm, word;
    std::vector<std::string> parts;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i < path.size() && path[i] != '/') {
            word.push_back(path[i]);
            continue;
        }
        if (!word.empty()) parts.push_back(word);
        word.clear();
    }
    for (const std::string &p : parts) norm += "/" + p;
    if (norm.empty()) norm = "/";
    auto hit = f.cache.find(norm);
    if (hit != f.cache.end()) return hit->second;

    uint64_t a = f.root;
    Object cur;
    cur.addr = a;
    read_header(f, a, cur.msgs);
    for (const std::string &p : parts) {
        if (!is_group(cur)) throw std::out_of_range("no object '" + norm + "' in the file");
        uint64_t next = UNDEF;
        for (auto &l : group_links(f, cur))
            if (l.first == p) next = l.second;
        if (next == UNDEF) throw std::out_of_range("no object '" + norm + "' in the file");
        cur = Object();
        cur.addr = next;
        read_header(f, next, cur.msgs);
    }
    for (const Msg &m : cur.msgs)
        if (m.type == 0x08) cur.is_dataset = true;
    return f.cache[norm] = cur;
}

// ---- datatype / dataspace ------------------------------------------------------------------
Dtype parse_dtype(const ffx_h5 &f, uint64_t off, uint64_t len) {
    if (len < 8) throw std::runtime_error("short datatype message");
    const uint8_t *p = f.at(off, len);
    Dtype t;
    t.cls = p[0] & 15;
    const unsigned bits = (unsigned)le(p + 1, 3);
    t.size = (uint32_t)le(p + 4, 4);
    switch (t.cls) {
    case 0:
        t.big_endian = bits & 1;
        t.sign = (bits >> 3) & 1;
        break;
    case 1:
        t.big_endian = bits & 1;
        break;
    case 3:
        break;
    case 8: {  // enum over an integer base type (h5py stores numpy bool this way)
        if (len < 16) throw std::runtime_error("short enum datatype");
        if ((p[8] & 15) != 0) throw Unsupported("enum over a non-integer base type");
        const unsigned base_bits = (unsigned)le(p + 9, 3);
        t.big_endian = base_bits & 1;
        t.sign = (base_bits >> 3) & 1;
        break;
    }
    case 9:
        if ((bits & 15) != 1) throw Unsupported("variable-length sequence datatype");
        t.vlen_string = true;
        break;
    default:
        throw Unsupported("datatype class " + std::to_string(t.cls));
    }
    if (t.big_endian && t.size > 1) throw Unsupported("big-endian data");
    return t;
}

Space parse_space(const ffx_h5 &f, uint64_t off, uint64_t len) {
    if (len < 4) throw std::runtime_error("short dataspace message");
    const uint8_t *p = f.at(off, len);
    Space s;
    const int version = p[0];
    s.rank = p[1];
    uint64_t c;
    if (version == 1)
        c = 8;
    else if (version == 2) {
        c = 4;
        s.null_space = p[3] == 2;
    } else
        throw Unsupported("dataspace message version " + std::to_string(version));
    if (s.rank > 8) throw Unsupported("more than 8 dimensions");
    if (c + (uint64_t)s.rank * f.L > len) throw std::runtime_error("short dataspace message");
    for (int i = 0; i < s.rank; i++) s.dims[i] = le(p + c + (uint64_t)i * f.L, f.L);
    return s;
}

// ---- datasets -------------------------------------------------------------------------------
void walk_chunk_tree(const ffx_h5 &f, uint64_t node, Dataset &d, int depth) {
    if (depth > 32) throw std::runtime_error("chunk B-tree too deep");
    const uint8_t *p = f.addr(node, 8 + 2 * f.O);
    if (memcmp(p, "TREE", 4) != 0 || p[4] != 1) throw std::runtime_error("bad chunk B-tree node");
    const int level = p[5];
    const unsigned used = (unsigned)le(p + 6, 2);
    const int rank = d.space.rank;
    const uint64_t key = 8 + 8 * (uint64_t)(rank + 1);  // chunk bytes, filter mask, offsets (+ element offset)
    const uint64_t body = node + 8 + 2 * f.O;
    f.addr(body, used * (key + f.O) + key);
    for (unsigned i = 0; i < used; i++) {
        const uint8_t *k = f.addr(body + i * (key + f.O), key + f.O);
        const uint64_t child = le_addr(k + key, f.O);
        if (level > 0) {
            walk_chunk_tree(f, child, d, depth + 1);
            continue;
        }
        const uint64_t nbytes = le(k, 4);
        const uint64_t row = le(k + 8, 8);
        for (int a = 1; a < rank; a++)
            if (le(k + 8 + 8 * a, 8) != 0) throw Unsupported("chunks that do not span whole rows");
        if (row % d.chunk[0]) throw std::runtime_error("chunk offset is not a multiple of the chunk shape");
        const uint64_t ci = row / d.chunk[0];
        if (ci >= d.chunk_addr.size()) continue;  // beyond the current extent (dataset was shrunk)
        if (nbytes != mul_checked(d.chunk[0], d.row_bytes)) throw Unsupported("chunk with a filtered (compressed) size");
        f.addr(child, nbytes);
        d.chunk_addr[ci] = child;
    }
}

Dataset &dataset_of(ffx_h5 &f, Object &o, const std::string &path) {
    if (!o.is_dataset) throw std::out_of_range("'" + path + "' is not a dataset");
    if (o.parsed_dataset) return o.ds;
    Dataset d;
    const Msg *layout = nullptr;
    bool have_type = false, have_space = false;
    for (const Msg &m : o.msgs) {
        if ((m.flags & 2) && (m.type == 0x01 || m.type == 0x03)) throw Unsupported("shared (committed) datatype or dataspace");
        if (m.type == 0x03) d.type = parse_dtype(f, m.off, m.len), have_type = true;
        if (m.type == 0x01) d.space = parse_space(f, m.off, m.len), have_space = true;
        if (m.type == 0x08) layout = &m;
        if (m.type == 0x0B) {
            const uint8_t *p = f.at(m.off, 2);
            if (p[1] != 0) throw Unsupported("dataset with a filter pipeline (compression)");
        }
    }
    if (!layout || !have_type || !have_space) throw std::runtime_error("dataset header lacks datatype, dataspace or layout");
    if (d.type.vlen_string) throw Unsupported("dataset of variable-length strings");
    d.row_bytes = d.type.size;
    for (int i = 1; i < d.space.rank; i++) d.row_bytes = mul_checked(d.row_bytes, d.space.dims[i]);
    const uint8_t *p = f.at(layout->off, layout->len);
    if (layout->len < 2) throw std::runtime_error("short layout message");
    if (p[0] != 3 && !(p[0] == 4 && p[1] != 2))
        throw Unsupported(p[0] == 4 ? "layout version 4 chunk index (file written with libver='latest')"
                                    : "data layout message version " + std::to_string(p[0]));
    d.layout = p[1];
    const uint64_t total = mul_checked(d.space.count(), d.type.size);
    if (d.layout == 0) {
        const uint64_t n = le(f.at(layout->off + 2, 2), 2);
        if (n < total || 4 + n > layout->len) throw std::runtime_error("compact dataset smaller than its dataspace");
        d.addr = layout->off + 4;  // absolute offset in the mapping (already includes base)
        d.bytes = n;
    } else if (d.layout == 1) {
        const uint8_t *q = f.at(layout->off + 2, f.O + f.L);
        d.addr = le_addr(q, f.O);
        d.bytes = le(q + f.O, f.L);
        if (d.addr != UNDEF) {
            if (d.bytes < total) throw std::runtime_error("contiguous dataset smaller than its dataspace");
            f.addr(d.addr, total);
        }
    } else if (d.layout == 2) {
        const int dimensionality = p[2];
        if (dimensionality != d.space.rank + 1 || d.space.rank < 1) throw std::runtime_error("chunk rank does not match the dataspace");
        const uint8_t *q = f.at(layout->off + 3, f.O + 4 * (uint64_t)dimensionality);
        const uint64_t tree = le_addr(q, f.O);
        for (int i = 0; i < dimensionality; i++) d.chunk[i] = le(q + f.O + 4 * i, 4);
        if (d.chunk[d.space.rank] != d.type.size) throw std::runtime_error("chunk element size differs from the datatype");
        for (int i = 1; i < d.space.rank; i++)
            if (d.chunk[i] != d.space.dims[i]) throw Unsupported("chunks that do not span whole rows");
        if (d.chunk[0] == 0) throw std::runtime_error("zero chunk shape");
        const uint64_t n_chunks = (d.space.dims[0] + d.chunk[0] - 1) / d.chunk[0];
        if (n_chunks > (uint64_t(1) << 28)) throw Unsupported("more than 2^28 chunks in one dataset");
        d.chunk_addr.assign(n_chunks, UNDEF);
        if (tree != UNDEF) walk_chunk_tree(f, tree, d, 0);
    } else
        throw Unsupported("data layout class " + std::to_string(d.layout));
    o.ds = d;
    o.parsed_dataset = true;
    return o.ds;
}

void copy_rows(const ffx_h5 &f, const Dataset &d, uint64_t row0, uint64_t n, uint8_t *dst) {
    if (d.layout != 2) {
        if (d.addr == UNDEF) {  // storage never allocated: fill value
            memset(dst, 0, n * d.row_bytes);
            return;
        }
        const uint8_t *src = d.layout == 0 ? f.at(d.addr, d.bytes) : f.addr(d.addr, d.bytes);
        if (mul_checked(row0 + n, d.row_bytes) > d.bytes) throw std::runtime_error("rows outside the stored data");
        memcpy(dst, src + row0 * d.row_bytes, n * d.row_bytes);
        return;
    }
    uint64_t row = row0;
    const uint64_t end = row0 + n;
    while (row < end) {
        const uint64_t ci = row / d.chunk[0], in = row % d.chunk[0];
        const uint64_t take = std::min(end - row, d.chunk[0] - in);
        uint8_t *to = dst + (row - row0) * d.row_bytes;
        if (d.chunk_addr[ci] == UNDEF)
            memset(to, 0, take * d.row_bytes);
        else
            memcpy(to, f.addr(d.chunk_addr[ci], d.chunk[0] * d.row_bytes) + in * d.row_bytes, take * d.row_bytes);
        row += take;
    }
}

// ---- attributes -----------------------------------------------------------------------------
uint64_t pad8(uint64_t n) { return (n + 7) & ~uint64_t(7); }

std::vector<Attr> &attrs_of(ffx_h5 &f, Object &o) {
    if (o.parsed_attrs) return o.attrs;
    std::vector<Attr> out;
    for (const Msg &m : o.msgs) {
        if (m.type == 0x15) {  // attribute info: dense storage lives in a fractal heap
            const uint8_t *p = f.at(m.off, m.len);
            uint64_t c = 2 + ((p[1] & 1) ? 2 : 0);
            if (c + f.O <= m.len && le_addr(p + c, f.O) != UNDEF) throw Unsupported("object with dense attribute storage");
        }
        if (m.type != 0x0C) continue;
        if (m.flags & 2) throw Unsupported("shared attribute message");
        const uint8_t *p = f.at(m.off, m.len);
        if (m.len < 8) throw std::runtime_error("short attribute message");
        const int version = p[0];
        if (version < 1 || version > 3) throw Unsupported("attribute message version " + std::to_string(version));
        if (version >= 2 && (p[1] & 3)) throw Unsupported("attribute with a shared datatype or dataspace");
        const uint64_t nlen = le(p + 2, 2), tlen = le(p + 4, 2), slen = le(p + 6, 2);
        uint64_t c = version == 3 ? 9 : 8;
        const uint64_t nstep = version == 1 ? pad8(nlen) : nlen;
        const uint64_t tstep = version == 1 ? pad8(tlen) : tlen;
        const uint64_t sstep = version == 1 ? pad8(slen) : slen;
        if (c + nstep + tstep + sstep > m.len || nlen == 0) throw std::runtime_error("attribute message overruns");
        Attr a;
        a.name.assign((const char *)p + c, strnlen((const char *)p + c, nlen));
        c += nstep;
        a.type = parse_dtype(f, m.off + c, tlen);
        c += tstep;
        a.space = parse_space(f, m.off + c, slen);
        c += sstep;
        a.off = m.off + c;
        a.len = a.space.count() * a.type.size;
        if (c + a.len > m.len) throw std::runtime_error("attribute value overruns its message");
        out.push_back(a);
    }
    o.attrs = out;
    o.parsed_attrs = true;
    return o.attrs;
}

// one variable-length string: length (4), global heap collection address (O), object index (4)
std::string vlen_string(const ffx_h5 &f, const uint8_t *ref) {
    const uint64_t n = le(ref, 4);
    const uint64_t col = le_addr(ref + 4, f.O);
    const unsigned want = (unsigned)le(ref + 4 + f.O, 4);
    if (n == 0 || col == UNDEF || col == 0) return std::string();
    const uint8_t *h = f.addr(col, 8 + f.L);
    if (memcmp(h, "GCOL", 4) != 0 || h[4] != 1) throw std::runtime_error("bad global heap collection");
    const uint64_t csize = le(h + 8, f.L);
    f.addr(col, csize);
    uint64_t c = 8 + f.L;
    const uint64_t oh = 8 + f.L;  // index (2), reference count (2), reserved (4), size (L)
    while (c + oh <= csize) {
        const uint8_t *o = f.addr(col + c, oh);
        const unsigned idx = (unsigned)le(o, 2);
        const uint64_t osz = le(o + 8, f.L);
        if (idx == 0) break;  // free space: the rest of the collection
        if (c + oh + osz > csize) throw std::runtime_error("global heap object overruns its collection");
        if (idx == want) {
            if (n > osz) throw std::runtime_error("variable-length string longer than its heap object");
            return std::string((const char *)f.addr(col + c + oh, osz), n);
        }
        c += oh + pad8(osz);
    }
    throw std::runtime_error("global heap object not found");
}

// the value of an attribute as bytes: numbers raw little-endian; strings as their characters,
// several strings separated by NUL
std::string attr_bytes(const ffx_h5 &f, const Attr &a) {
    const uint8_t *v = f.at(a.off, a.len);
    const uint64_t n = a.space.count();
    if (a.type.vlen_string || a.type.cls == 3) {
        std::string out;
        for (uint64_t i = 0; i < n; i++) {
            if (i) out.push_back('\0');
            if (a.type.vlen_string)
                out += vlen_string(f, v + i * a.type.size);
            else
                out.append((const char *)v + i * a.type.size, strnlen((const char *)v + i * a.type.size, a.type.size));
        }
        return out;
    }
    return std::string((const char *)v, a.len);
}

template <typename F>
int guarded(const char *what, F body) {
    try {
        return body();
    } catch (const Unsupported &e) {
        return ffx_set_error_message(FFX_ERR_UNSUPPORTED, (std::string(what) + ": " + e.what()).c_str());
    } catch (const std::out_of_range &e) {
        return ffx_set_error_message(FFX_ERR_STATE, (std::string(what) + ": " + e.what()).c_str());
    } catch (const std::exception &e) {
        return ffx_set_error_message(FFX_ERR_INVALID, (std::string(what) + ": " + e.what()).c_str());
    }
}

int put_text(const std::string &s, char *buf, int64_t cap, int64_t *needed) {
    if (needed) *needed = (int64_t)s.size();
    if (buf && cap > 0) memcpy(buf, s.data(), std::min<size_t>(s.size(), (size_t)cap));
    return FFX_OK;
}

}  // namespace

extern "C" {

int ffx_h5_open(const char *path, ffx_h5 **out) {
    if (!path || !out) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_open: bad arguments");
    *out = nullptr;
    ffx_h5 *f = new ffx_h5();
    int rc = guarded("ffx_h5_open", [&]() {
        f->fd = open(path, O_RDONLY);
        if (f->fd < 0) throw std::runtime_error(std::string("cannot open ") + path);
        struct stat st;
        if (fstat(f->fd, &st) != 0) throw std::runtime_error("fstat failed");
        f->size = (uint64_t)st.st_size;
        if (f->size < 48) throw std::runtime_error("not an HDF5 file (too short)");
        void *m = mmap(nullptr, f->size, PROT_READ, MAP_SHARED, f->fd, 0);
        if (m == MAP_FAILED) throw std::runtime_error("mmap failed");
        f->map = (const uint8_t *)m;
        static const uint8_t magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t sb = UNDEF;
        for (uint64_t o = 0; o + 8 <= f->size; o = o ? o * 2 : 512)  // 0, 512, 1024, 2048, ...
            if (memcmp(f->map + o, magic, 8) == 0) {
                sb = o;
                break;
            }
        if (sb == UNDEF) throw std::runtime_error("not an HDF5 file (no superblock signature)");
        const uint8_t *p = f->at(sb, 16);
        const int version = p[8];
        if (version <= 1) {
            f->O = p[13];
            f->L = p[14];
            if ((f->O != 4 && f->O != 8) || (f->L != 4 && f->L != 8)) throw Unsupported("offset/length sizes other than 4 or 8");
            uint64_t c = sb + 24 + (version == 1 ? 4 : 0);
            const uint8_t *q = f->at(c, 4 * (uint64_t)f->O + 2 * (uint64_t)f->O + 8 + 16);
            f->base = le_addr(q, f->O);
            // free-space address, end of file, driver block; then the root symbol table entry
            const uint8_t *entry = q + 4 * f->O;
            f->root = le_addr(entry + f->O, f->O);
        } else if (version <= 3) {
            f->O = p[9];
            f->L = p[10];
            if ((f->O != 4 && f->O != 8) || (f->L != 4 && f->L != 8)) throw Unsupported("offset/length sizes other than 4 or 8");
            const uint8_t *q = f->at(sb + 12, 4 * (uint64_t)f->O);
            f->base = le_addr(q, f->O);
            f->root = le_addr(q + 3 * f->O, f->O);
        } else
            throw Unsupported("superblock version " + std::to_string(version));
        if (f->base == UNDEF) f->base = 0;
        if (sb != 0 && f->base == 0) f->base = sb;  // user block in front of the file
        std::vector<Msg> probe;
        read_header(*f, f->root, probe);
        return FFX_OK;
    });
    if (rc != FFX_OK) {
        ffx_h5_close(f);
        return rc;
    }
    *out = f;
    return FFX_OK;
}

void ffx_h5_close(ffx_h5 *f) {
    if (!f) return;
    if (f->map) munmap((void *)f->map, f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
}

int ffx_h5_kind(ffx_h5 *f, const char *path, int *kind) {
    if (!f || !path || !kind) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_kind: bad arguments");
    *kind = 0;
    try {
        Object &o = open_object(*f, path);
        *kind = o.is_dataset ? 2 : 1;
    } catch (const std::out_of_range &) {
        return FFX_OK;  // absent: kind stays 0
    } catch (const Unsupported &e) {
        return ffx_set_error_message(FFX_ERR_UNSUPPORTED, (std::string("ffx_h5_kind: ") + e.what()).c_str());
    } catch (const std::exception &e) {
        return ffx_set_error_message(FFX_ERR_INVALID, (std::string("ffx_h5_kind: ") + e.what()).c_str());
    }
    return FFX_OK;
}

int ffx_h5_list(ffx_h5 *f, const char *path, char *buf, int64_t cap, int64_t *needed) {
    if (!f || !path) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_list: bad arguments");
    return guarded("ffx_h5_list", [&]() {
        Object &o = open_object(*f, path);
        if (o.is_dataset) throw std::out_of_range(std::string("'") + path + "' is not a group");
        std::string s;
        for (auto &l : group_links(*f, o)) s += l.first + '\n';
        return put_text(s, buf, cap, needed);
    });
}

int ffx_h5_dataset_info(ffx_h5 *f, const char *path, int64_t *info) {
    if (!f || !path || !info) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_dataset_info: bad arguments");
    return guarded("ffx_h5_dataset_info", [&]() {
        Object &o = open_object(*f, path);
        Dataset &d = dataset_of(*f, o, path);
        info[0] = d.type.cls;
        info[1] = d.type.size;
        info[2] = d.type.sign;
        info[3] = d.space.rank;
        for (int i = 0; i < 8; i++) info[4 + i] = i < d.space.rank ? (int64_t)d.space.dims[i] : 0;
        info[12] = d.layout;
        info[13] = d.layout == 2 ? (int64_t)d.chunk[0] : 0;
        info[14] = (int64_t)d.row_bytes;
        info[15] = 0;
        return FFX_OK;
    });
}

int ffx_h5_read_rows(ffx_h5 *f, const char *path, int64_t row0, int64_t nrows, void *dst) {
    if (!f || !path || row0 < 0 || nrows < 0 || (nrows && !dst))
        return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_read_rows: bad arguments");
    return guarded("ffx_h5_read_rows", [&]() {
        Object &o = open_object(*f, path);
        Dataset &d = dataset_of(*f, o, path);
        const uint64_t rows = d.space.rank ? d.space.dims[0] : 1;
        if ((uint64_t)row0 + (uint64_t)nrows > rows) throw std::runtime_error("rows outside the dataset");
        if (d.space.null_space || nrows == 0) return (int)FFX_OK;
        const uint64_t bytes = mul_checked((uint64_t)nrows, d.row_bytes);
        unsigned workers = bytes >= (32u << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
        if (workers <= 1) {
            copy_rows(*f, d, row0, nrows, (uint8_t *)dst);
            return (int)FFX_OK;
        }
        // first touch of a mapped file is a page fault per 4 KB: spread it over a few threads
        std::vector<std::thread> pool;
        std::vector<std::string> errors(workers);
        const uint64_t per = ((uint64_t)nrows + workers - 1) / workers;
        for (unsigned w = 0; w < workers; w++) {
            const uint64_t lo = std::min<uint64_t>(w * per, nrows), hi = std::min<uint64_t>(lo + per, nrows);
            if (lo == hi) continue;
            pool.emplace_back([&, w, lo, hi]() {
                try {
                    copy_rows(*f, d, row0 + lo, hi - lo, (uint8_t *)dst + lo * d.row_bytes);
                } catch (const std::exception &e) {
                    errors[w] = e.what();
                }
            });
        }
        for (auto &t : pool) t.join();
        for (auto &e : errors)
            if (!e.empty()) throw std::runtime_error(e);
        return (int)FFX_OK;
    });
}

int ffx_h5_row_span(ffx_h5 *f, const char *path, int64_t row0, const void **rows, int64_t *nrows) {
    if (!f || !path || row0 < 0 || !rows || !nrows)
        return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_row_span: bad arguments");
    *rows = nullptr;
    *nrows = 0;
    return guarded("ffx_h5_row_span", [&]() {
        Object &o = open_object(*f, path);
        Dataset &d = dataset_of(*f, o, path);
        const uint64_t total = d.space.rank ? d.space.dims[0] : 1;
        if ((uint64_t)row0 >= total) throw std::runtime_error("row outside the dataset");
        if (d.layout == 2) {
            const uint64_t ci = row0 / d.chunk[0], in = row0 % d.chunk[0];
            *nrows = (int64_t)std::min(d.chunk[0] - in, total - row0);
            if (d.chunk_addr[ci] != UNDEF)
                *rows = f->addr(d.chunk_addr[ci], d.chunk[0] * d.row_bytes) + in * d.row_bytes;
        } else {
            *nrows = (int64_t)(total - row0);
            if (d.addr != UNDEF) {
                if (mul_checked(total, d.row_bytes) > d.bytes) throw std::runtime_error("rows outside the stored data");
                *rows = (d.layout == 0 ? f->at(d.addr, d.bytes) : f->addr(d.addr, d.bytes)) + row0 * d.row_bytes;
            }
        }
        return FFX_OK;
    });
}

int ffx_h5_attr_names(ffx_h5 *f, const char *path, char *buf, int64_t cap, int64_t *needed) {
    if (!f || !path) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_attr_names: bad arguments");
    return guarded("ffx_h5_attr_names", [&]() {
        Object &o = open_object(*f, path);
        std::string s;
        for (auto &a : attrs_of(*f, o)) s += a.name + '\n';
        return put_text(s, buf, cap, needed);
    });
}

int ffx_h5_attr_read(ffx_h5 *f, const char *path, const char *name, int64_t *info, void *buf, int64_t cap,
                     int64_t *needed) {
    if (!f || !path || !name || !info) return ffx_set_error_message(FFX_ERR_INVALID, "ffx_h5_attr_read: bad arguments");
    return guarded("ffx_h5_attr_read", [&]() {
        Object &o = open_object(*f, path);
        for (auto &a : attrs_of(*f, o)) {
            if (a.name != name) continue;
            info[0] = a.type.vlen_string ? 3 : a.type.cls;  // both string kinds read as class 3
            info[1] = a.type.size;
            info[2] = a.type.sign;
            info[3] = a.space.rank;
            info[4] = (int64_t)a.space.count();
            for (int i = 0; i < 8; i++) info[5 + i] = i < a.space.rank ? (int64_t)a.space.dims[i] : 0;
            return put_text(attr_bytes(*f, a), (char *)buf, cap, needed);
        }
        throw std::out_of_range(std::string("no attribute '") + name + "' on '" + path + "'");
    });
}

}  // extern "C"
