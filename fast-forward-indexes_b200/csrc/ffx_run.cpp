// ffx_run.cpp — TREC run files on all host cores (plain C++17, no CUDA).
//
// The reference reads a run with `pd.read_csv(f, sep=r"\s+", header=None, names=[q_id, q0, id,
// rank, score, name])` and writes one with `DataFrame.to_csv(sep="\t", header=False)`
// (ranking.py:348-366,388-409): at 2.6 x 10^7 rows that is a minute of single-threaded parsing and
// formatting around a 70 ms GPU pass.  Here the file is mapped, cut at line boundaries into one
// piece per core, and tokenised in place; the two id columns come out in Arrow layout (what the
// id dictionaries take), the scores as the doubles pandas' default parser (`precise_xstrtod`)
// produces — digit accumulation capped at 17 digits, one scaling by a power of ten — so that the
// scores of the resulting ranking are bit-identical to the reference's.
// Anything whose pandas semantics are not reproduced here (quotes, NA tokens, numeric id columns
// that pandas would renumber, wrong field counts, non-decimal scores) is reported through
// counters; the Python shell then lets pandas read that file.
//
// Writing formats float32 scores exactly like numpy / pandas do (shortest round-trip digits;
// positional for 1e-4 <= |x| < 1e6, else scientific with a two-digit exponent).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ffx.h"

extern "C" int ffx_set_error_message(int code, const char *msg);

namespace {

int fail(int code, const std::string &msg) { return ffx_set_error_message(code, msg.c_str()); }

struct Token {
    const char *p;
    int32_t len;
};

struct Piece {  // what one thread found in its part of the file
    std::vector<Token> q, id;
    std::vector<double> score;
    int64_t q_bytes = 0, id_bytes = 0;
    int64_t bad_fields = 0, bad_score = 0, quotes = 0;
    // per id column (0 = q_id, 1 = id, 2 = name of the first row only)
    int64_t numeric_like[2] = {0, 0}, canonical_int[2] = {0, 0}, na_like[2] = {0, 0}, bool_like[2] = {0, 0};
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

bool equals_any(const char *p, int n, const char *const *words, int count) {
    for (int i = 0; i < count; i++)
        if (static_cast<int>(strlen(words[i])) == n && memcmp(p, words[i], static_cast<size_t>(n)) == 0) return true;
    return false;
}

// pandas' default NA strings (io/parsers: STR_NA_VALUES)
bool is_na_token(const char *p, int n) {
    static const char *const na[] = {"#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                                     "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    return n == 0 || equals_any(p, n, na, static_cast<int>(sizeof na / sizeof na[0]));
}

bool is_bool_token(const char *p, int n) {
    static const char *const b[] = {"True", "TRUE", "true", "False", "FALSE", "false"};
    return equals_any(p, n, b, 6);
}

// [+-]?digits*[.digits*]([eE][+-]?digits+)? with at least one digit: what xstrtod accepts whole
bool plain_decimal(const char *p, int n, bool *is_int) {
    int i = 0, digits = 0;
    *is_int = true;
    if (i < n && (p[i] == '+' || p[i] == '-')) i++;
    while (i < n && is_digit(p[i])) i++, digits++;
    if (i < n && p[i] == '.') {
        *is_int = false;
        i++;
        while (i < n && is_digit(p[i])) i++, digits++;
    }
    if (digits == 0) return false;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        *is_int = false;
        i++;
        if (i < n && (p[i] == '+' || p[i] == '-')) i++;
        int ed = 0;
        while (i < n && is_digit(p[i])) i++, ed++;
        if (ed == 0) return false;
    }
    return i == n;
}

bool special_float(const char *p, int n) {
    static const char *const w[] = {"inf", "-inf", "+inf", "Inf", "-Inf", "+Inf", "INF", "-INF", "infinity", "-infinity",
                                    "Infinity", "-Infinity", "+Infinity", "+infinity"};
    return equals_any(p, n, w, static_cast<int>(sizeof w / sizeof w[0]));
}

// -?(0|[1-9][0-9]{0,17}): an integer whose decimal string IS the token (pandas would parse the
// column to int64 and print it back unchanged)
bool canonical_int(const char *p, int n) {
    int i = 0;
    if (i < n && p[i] == '-') i++;
    const int digits = n - i;
    if (digits < 1 || digits > 18) return false;
    if (p[i] == '0') return digits == 1 && i == 0;
    for (; i < n; i++)
        if (!is_digit(p[i])) return false;
    return true;
}

// pandas/_libs/src/parser/tokenizer.c `precise_xstrtod` — the C engine's default float parser since
// pandas 1.2 (float_precision=None == "high") — for a plain decimal token: at most 17 digits
// (leading zeros count) are accumulated into a double, further integer digits only raise the
// exponent, then ONE multiplication or division by a correctly rounded power of ten.
const double *powers_of_ten() {
    static double table[309];
    static bool ready = [] {
        char text[16];
        for (int i = 0; i <= 308; i++) {
            std::snprintf(text, sizeof text, "1e%d", i);
            table[i] = std::strtod(text, nullptr);
        }
        return true;
    }();
    (void)ready;
    return table;
}

double xstrtod_like(const char *p, int n, bool *range_error) {
    const int max_digits = 17;
    const double *e10 = powers_of_ten();
    int i = 0;
    bool negative = false;
    if (i < n && (p[i] == '-' || p[i] == '+')) negative = p[i++] == '-';
    double number = 0.;
    int exponent = 0, num_digits = 0, num_decimals = 0;
    while (i < n && is_digit(p[i])) {
        if (num_digits < max_digits) {
            number = number * 10. + (p[i] - '0');
            num_digits++;
        } else {
            exponent++;
        }
        i++;
    }
    if (i < n && p[i] == '.') {
        i++;
        while (i < n && is_digit(p[i])) {
            if (num_digits < max_digits) {
                number = number * 10. + (p[i] - '0');
                num_digits++;
                num_decimals++;
            }
            i++;
        }
        exponent -= num_decimals;
    }
    if (negative) number = -number;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        i++;
        bool neg_e = false;
        if (i < n && (p[i] == '-' || p[i] == '+')) neg_e = p[i++] == '-';
        int e = 0;
        while (i < n && is_digit(p[i])) {
            if (e < 100000) e = e * 10 + (p[i] - '0');
            i++;
        }
        exponent += neg_e ? -e : e;
    }
    if (exponent < -300 || exponent > 300) {  // the reference's overflow / underflow paths: let pandas decide
        *range_error = true;
        return 0.;
    }
    if (exponent > 0) number *= e10[exponent];
    else if (exponent < 0) number /= e10[-exponent];
    if (std::isinf(number)) *range_error = true;
    return number;
}

void classify(Piece &pc, int col, const Token &t) {
    bool is_int = false;
    if (is_na_token(t.p, t.len)) pc.na_like[col]++;
    if (is_bool_token(t.p, t.len)) pc.bool_like[col]++;
    if (plain_decimal(t.p, t.len, &is_int) || special_float(t.p, t.len)) pc.numeric_like[col]++;
    if (canonical_int(t.p, t.len)) pc.canonical_int[col]++;
}

void scan_piece(const char *begin, const char *end, Piece &pc) {
    const char *p = begin;
    while (p < end) {
        const char *eol = static_cast<const char *>(memchr(p, '\n', static_cast<size_t>(end - p)));
        if (!eol) eol = end;
        Token f[7];
        int nf = 0;
        const char *c = p;
        while (c < eol) {
            while (c < eol && is_space(*c)) c++;
            if (c >= eol) break;
            const char *s = c;
            while (c < eol && !is_space(*c)) {
                if (*c == '"') pc.quotes++;
                c++;
            }
            if (nf < 7) f[nf] = Token{s, static_cast<int32_t>(c - s)};
            nf++;
        }
        if (nf != 0) {  // blank lines are skipped (skip_blank_lines=True)
            if (nf != 6) {
                pc.bad_fields++;
            } else {
                bool is_int = false, range = false;
                double v = 0.;
                if (plain_decimal(f[4].p, f[4].len, &is_int)) {
                    // an all-integer score column is parsed to int64 by pandas: exact below 2^53
                    if (is_int && f[4].len > 15) pc.bad_score++;
                    v = xstrtod_like(f[4].p, f[4].len, &range);
                    if (range) pc.bad_score++;
                } else {
                    pc.bad_score++;
                }
                pc.q.push_back(f[0]);
                pc.id.push_back(f[2]);
                pc.score.push_back(v);
                pc.q_bytes += f[0].len;
                pc.id_bytes += f[2].len;
                classify(pc, 0, f[0]);
                classify(pc, 1, f[2]);
            }
        }
        p = eol + 1;
    }
}

int worker_count(int requested, int64_t bytes) {
    int t = requested > 0 ? requested : static_cast<int>(std::thread::hardware_concurrency());
    t = std::max(1, std::min(t, 64));
    return static_cast<int>(std::min<int64_t>(t, std::max<int64_t>(1, bytes / (1 << 20))));
}

// ---- float32 -> text as numpy / pandas print it ---------------------------------------------------
int format_f32(float x, char *out) {
    if (std::isnan(x)) return static_cast<int>(std::snprintf(out, 8, "nan"));
    if (std::isinf(x)) return static_cast<int>(std::snprintf(out, 8, x < 0 ? "-inf" : "inf"));
    char sci[48];
    auto r = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);  // shortest round-trip digits
    *r.ptr = '\0';
    // sci = [-]d[.ddd]e[+-]XX
    const char *s = sci;
    char *o = out;
    if (*s == '-') *o++ = *s++;
    char digits[24];
    int nd = 0;
    digits[nd++] = *s++;
    if (*s == '.') {
        s++;
        while (*s != 'e') digits[nd++] = *s++;
    }
    s++;  // 'e'
    const int exp10 = std::atoi(s);
    const double a = std::fabs(static_cast<double>(x));
    if (a == 0.0 || (a < 1e6 && a >= 1e-4)) {  // positional, at least one digit after the point
        if (exp10 >= 0) {
            for (int i = 0; i <= exp10; i++) *o++ = i < nd ? digits[i] : '0';
            *o++ = '.';
            if (nd > exp10 + 1) {
                for (int i = exp10 + 1; i < nd; i++) *o++ = digits[i];
            } else {
                *o++ = '0';
            }
        } else {
            *o++ = '0';
            *o++ = '.';
            for (int i = 0; i < -exp10 - 1; i++) *o++ = '0';
            for (int i = 0; i < nd; i++) *o++ = digits[i];
        }
    } else {  // scientific: d[.ddd]e[+-]XX, no trailing ".0"
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            for (int i = 1; i < nd; i++) *o++ = digits[i];
        }
        *o++ = 'e';
        *o++ = exp10 < 0 ? '-' : '+';
        const int e = exp10 < 0 ? -exp10 : exp10;
        if (e < 10) *o++ = '0';
        o += std::snprintf(o, 8, "%d", e);
    }
    return static_cast<int>(o - out);
}

}  // namespace

struct ffx_run {
    int fd = -1;
    const char *map = nullptr;
    size_t size = 0;
    std::vector<Piece> pieces;
    int64_t rows = 0, q_bytes = 0, id_bytes = 0;
    std::string first_name;
};

extern "C" {

int ffx_run_open(const char *file_name, int n_threads, ffx_run **out, int64_t *info) {
    if (!file_name || !out || !info) return fail(FFX_ERR_INVALID, "ffx_run_open: bad arguments");
    *out = nullptr;
    const int fd = open(file_name, O_RDONLY);
    if (fd < 0) return fail(FFX_ERR_INVALID, std::string("ffx_run_open: cannot open ") + file_name);
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(FFX_ERR_INVALID, "ffx_run_open: fstat failed");
    }
    ffx_run *run = new ffx_run();
    run->fd = fd;
    run->size = static_cast<size_t>(st.st_size);
    if (run->size > 0) {
        void *m = mmap(nullptr, run->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            delete run;
            return fail(FFX_ERR_INVALID, "ffx_run_open: mmap failed");
        }
        run->map = static_cast<const char *>(m);
        madvise(m, run->size, MADV_SEQUENTIAL);
    }
    const int threads = worker_count(n_threads, static_cast<int64_t>(run->size));
    run->pieces.resize(static_cast<size_t>(threads));
    // cut at line boundaries
    std::vector<size_t> cut(static_cast<size_t>(threads) + 1, run->size);
    cut[0] = 0;
    for (int t = 1; t < threads; t++) {
        size_t at = run->size / static_cast<size_t>(threads) * static_cast<size_t>(t);
        at = std::max(at, cut[static_cast<size_t>(t) - 1]);
        const char *nl = at < run->size ? static_cast<const char *>(memchr(run->map + at, '\n', run->size - at)) : nullptr;
        cut[static_cast<size_t>(t)] = nl ? static_cast<size_t>(nl - run->map) + 1 : run->size;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([run, &cut, t] {
            scan_piece(run->map + cut[static_cast<size_t>(t)], run->map + cut[static_cast<size_t>(t) + 1],
                       run->pieces[static_cast<size_t>(t)]);
        });
    for (auto &th : pool) th.join();
    // info: rows, q bytes, id bytes, bad field counts, bad scores, quotes,
    //       then per key column: numeric-like, canonical ints, NA-like, bool-like; then the first name's flags
    std::fill(info, info + 16, 0);
    for (const Piece &pc : run->pieces) {
        run->rows += static_cast<int64_t>(pc.q.size());
        run->q_bytes += pc.q_bytes;
        run->id_bytes += pc.id_bytes;
        info[3] += pc.bad_fields;
        info[4] += pc.bad_score;
        info[5] += pc.quotes;
        for (int c = 0; c < 2; c++) {
            info[6 + 4 * c] += pc.numeric_like[c];
            info[7 + 4 * c] += pc.canonical_int[c];
            info[8 + 4 * c] += pc.na_like[c];
            info[9 + 4 * c] += pc.bool_like[c];
        }
    }
    info[0] = run->rows;
    info[1] = run->q_bytes;
    info[2] = run->id_bytes;
    // the name of the first row (ranking.py:403 `df["name"][0]`): the 6th token of the first non-blank line
    if (run->rows > 0 || info[3] > 0) {
        const char *p = run->map, *end = run->map + run->size;
        while (p < end) {
            const char *eol = static_cast<const char *>(memchr(p, '\n', static_cast<size_t>(end - p)));
            if (!eol) eol = end;
            std::vector<Token> f;
            const char *c = p;
            while (c < eol) {
                while (c < eol && is_space(*c)) c++;
                if (c >= eol) break;
                const char *s = c;
                while (c < eol && !is_space(*c)) c++;
                f.push_back(Token{s, static_cast<int32_t>(c - s)});
            }
            if (!f.empty()) {
                if (f.size() == 6) {
                    run->first_name.assign(f[5].p, static_cast<size_t>(f[5].len));
                    bool is_int = false;
                    const bool numeric = plain_decimal(f[5].p, f[5].len, &is_int) || special_float(f[5].p, f[5].len);
                    info[14] = (numeric ? 1 : 0) | (is_na_token(f[5].p, f[5].len) ? 2 : 0) | (is_bool_token(f[5].p, f[5].len) ? 4 : 0);
                }
                break;
            }
            p = eol + 1;
        }
    }
    info[15] = static_cast<int64_t>(run->first_name.size());
    *out = run;
    return FFX_OK;
}

int ffx_run_read(ffx_run *run, int64_t *q_offsets, char *q_data, int64_t *id_offsets, char *id_data, double *score,
                 char *first_name) {
    if (!run || !q_offsets || !id_offsets || !score) return fail(FFX_ERR_INVALID, "ffx_run_read: bad arguments");
    // where every piece starts in the outputs
    std::vector<int64_t> row0(run->pieces.size() + 1, 0), qb0(run->pieces.size() + 1, 0), ib0(run->pieces.size() + 1, 0);
    for (size_t t = 0; t < run->pieces.size(); t++) {
        row0[t + 1] = row0[t] + static_cast<int64_t>(run->pieces[t].q.size());
        qb0[t + 1] = qb0[t] + run->pieces[t].q_bytes;
        ib0[t + 1] = ib0[t] + run->pieces[t].id_bytes;
    }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < run->pieces.size(); t++)
        pool.emplace_back([&, t] {
            const Piece &pc = run->pieces[t];
            int64_t r = row0[t], qb = qb0[t], ib = ib0[t];
            for (size_t i = 0; i < pc.q.size(); i++, r++) {
                q_offsets[r] = qb;
                memcpy(q_data + qb, pc.q[i].p, static_cast<size_t>(pc.q[i].len));
                qb += pc.q[i].len;
                id_offsets[r] = ib;
                memcpy(id_data + ib, pc.id[i].p, static_cast<size_t>(pc.id[i].len));
                ib += pc.id[i].len;
                score[r] = pc.score[i];
            }
        });
    for (auto &th : pool) th.join();
    q_offsets[run->rows] = run->q_bytes;
    id_offsets[run->rows] = run->id_bytes;
    if (first_name) memcpy(first_name, run->first_name.data(), run->first_name.size());
    return FFX_OK;
}

void ffx_run_close(ffx_run *run) {
    if (!run) return;
    if (run->map) munmap(const_cast<char *>(run->map), run->size);
    if (run->fd >= 0) close(run->fd);
    delete run;
}

int ffx_run_write(const char *file_name, int64_t nq, const int64_t *q_off, const int64_t *qk_offsets, const char *qk_data,
                  const int64_t *idk_offsets, const char *idk_data, const int32_t *id_code, const float *score,
                  const char *name, int n_threads) {
    if (!file_name || nq < 0 || !name || (nq > 0 && (!q_off || !qk_offsets || !qk_data || !idk_offsets || !idk_data)))
        return fail(FFX_ERR_INVALID, "ffx_run_write: bad arguments");
    const int64_t n = nq > 0 ? q_off[nq] : 0;
    if (n > 0 && (!id_code || !score)) return fail(FFX_ERR_INVALID, "ffx_run_write: bad arguments");
    FILE *fp = std::fopen(file_name, "wb");
    if (!fp) return fail(FFX_ERR_INVALID, std::string("ffx_run_write: cannot create ") + file_name);
    const size_t name_len = strlen(name);
    int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
    threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(threads, 64), n / 65536 + 1)));
    // blocks of queries per thread, balanced by rows; written in order
    std::vector<int64_t> cut(static_cast<size_t>(threads) + 1, nq);
    cut[0] = 0;
    for (int t = 1; t < threads; t++) {
        const int64_t target = n / threads * t;
        cut[static_cast<size_t>(t)] = std::lower_bound(q_off, q_off + nq + 1, target) - q_off;
        cut[static_cast<size_t>(t)] = std::max(cut[static_cast<size_t>(t)], cut[static_cast<size_t>(t) - 1]);
        cut[static_cast<size_t>(t)] = std::min(cut[static_cast<size_t>(t)], nq);
    }
    std::vector<std::string> text(static_cast<size_t>(threads));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            std::string &buf = text[static_cast<size_t>(t)];
            const int64_t b0 = cut[static_cast<size_t>(t)], b1 = cut[static_cast<size_t>(t) + 1];
            if (b1 > b0) buf.reserve(static_cast<size_t>((q_off[b1] - q_off[b0]) * 48));
            char num[64];
            for (int64_t b = b0; b < b1; b++) {
                const char *qs = qk_data + qk_offsets[b];
                const size_t ql = static_cast<size_t>(qk_offsets[b + 1] - qk_offsets[b]);
                for (int64_t r = q_off[b]; r < q_off[b + 1]; r++) {
                    buf.append(qs, ql);
                    buf.append("\tQ0\t", 4);
                    const int32_t c = id_code[r];
                    buf.append(idk_data + idk_offsets[c], static_cast<size_t>(idk_offsets[c + 1] - idk_offsets[c]));
                    buf.push_back('\t');
                    auto rk = std::to_chars(num, num + sizeof num, r - q_off[b] + 1);
                    buf.append(num, static_cast<size_t>(rk.ptr - num));
                    buf.push_back('\t');
                    buf.append(num, static_cast<size_t>(format_f32(score[r], num)));
                    buf.push_back('\t');
                    buf.append(name, name_len);
                    buf.push_back('\n');
                }
            }
        });
    for (auto &th : pool) th.join();
    bool ok = true;
    for (const std::string &buf : text)
        if (!buf.empty() && std::fwrite(buf.data(), 1, buf.size(), fp) != buf.size()) ok = false;
    if (std::fclose(fp) != 0) ok = false;
    return ok ? FFX_OK : fail(FFX_ERR_INVALID, std::string("ffx_run_write: short write to ") + file_name);
}

}  // extern "C"
