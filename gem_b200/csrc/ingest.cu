// gem_b200/csrc/ingest.cu -- wire formats on either side of the hot path, native and multi-threaded (HOST code only;
// it lives in libgemb200.so so that one library serves the whole path, and needs no GPU).
//
// SURVEY 8(f) rank 2.  The reference reads and writes these files with per-line Python loops
// (gem/utils/graph_util.py:129-169): at the 20 M edges of BASELINE configs[1] that is minutes before the first kernel
// can start.  Same bytes, same values:
//   gemb_edge_list_scan / _parse   loadGraphFromEdgeListTxt (:143-158): every non-blank line is "src dst [weight]",
//                                  any run of blanks / tabs separates tokens, exactly 3 tokens -> float(weight),
//                                  otherwise weight 1.0 (the reference's rule, :151-154)
//   gemb_edge_list_write           saveGraphToEdgeListTxt (:129-134, two header lines) and
//                                  saveGraphToEdgeListTxtn2v (:137-140): one "%d %d %f\n" per edge
//   gemb_emb_read / gemb_emb_write loadEmbedding (:161-169) and the ".emb" text SNAP's WriteOutput produces
//                                  ("<rows> <d>" then "<id> v1 ... vd", 6 significant digits)
// The file is mmap'ed and cut at line boundaries into one piece per thread; numbers of the plain [+-]digits[.digits]
// form are parsed inline (exactly: integer mantissa / power of ten, both exact in fp64 up to 15 digits, so the
// result equals strtod's), everything else (exponents, inf, nan, > 15 digits) goes through strtod.
#include "common.cuh"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <thread>

namespace {

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap((void *)p, n);
        if (fd >= 0) close(fd);
    }
};

int map_file(const char *path, Mapped &m) {
    m.fd = open(path, O_RDONLY);
    if (m.fd < 0) { gemb::set_error("cannot open %s: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    struct stat st;
    if (fstat(m.fd, &st) != 0) { gemb::set_error("fstat %s: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    m.n = (size_t)st.st_size;
    if (m.n == 0) return GEMB_OK;
    void *q = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
    if (q == MAP_FAILED) { m.n = 0; gemb::set_error("mmap %s: %s", path, strerror(errno)); return GEMB_ERR_NOMEM; }
    m.p = (const char *)q;
    return GEMB_OK;
}

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// [begin, end) of the data after `skip` lines
size_t skip_lines(const Mapped &m, int64_t skip) {
    size_t pos = 0;
    for (int64_t i = 0; i < skip && pos < m.n; i++) {
        const void *nl = memchr(m.p + pos, '\n', m.n - pos);
        pos = nl ? (size_t)((const char *)nl - m.p) + 1 : m.n;
    }
    return pos;
}

// piece boundaries at line starts
std::vector<size_t> cut(const Mapped &m, size_t begin, int pieces) {
    std::vector<size_t> b(pieces + 1, m.n);
    b[0] = begin;
    for (int i = 1; i < pieces; i++) {
        size_t pos = begin + (m.n - begin) / pieces * i;
        if (pos < b[i - 1]) pos = b[i - 1];
        const void *nl = pos < m.n ? memchr(m.p + pos, '\n', m.n - pos) : nullptr;
        b[i] = nl ? (size_t)((const char *)nl - m.p) + 1 : m.n;
    }
    return b;
}

int n_threads(size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    int t = (int)std::min<size_t>(hw ? hw : 4, 64);
    const char *e = getenv("GEMB_IO_THREADS");
    if (e && atoi(e) > 0) t = atoi(e);
    const size_t by_size = bytes / ((size_t)4 << 20) + 1;     // at least 4 MB per thread
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)t, by_size));
}

static const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18};

// token [s, e) -> double; fast path for [+-]ddd[.ddd] with <= 15 significant digits, strtod otherwise
bool parse_double(const char *s, const char *e, double *out) {
    const char *p = s;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    bool seen_dot = false, any = false, simple = true;
    for (; p < e; p++) {
        const char c = *p;
        if (c >= '0' && c <= '9') {
            any = true;
            if (mant || c != '0') digits++;
            if (digits > 15) { simple = false; break; }
            mant = mant * 10 + (uint64_t)(c - '0');
            if (seen_dot) frac++;
        } else if (c == '.' && !seen_dot) {
            seen_dot = true;
        } else { simple = false; break; }
    }
    if (simple && any && frac <= 18) {
        const double v = (double)mant / kPow10[frac];    // both operands exact -> correctly rounded quotient
        *out = neg ? -v : v;
        return true;
    }
    char buf[96];
    const size_t len = (size_t)(e - s);
    if (len == 0 || len >= sizeof buf) return false;
    memcpy(buf, s, len);
    buf[len] = 0;
    char *endp = nullptr;
    const double v = strtod(buf, &endp);
    if (endp != buf + len) return false;
    *out = v;
    return true;
}

bool parse_int(const char *s, const char *e, int64_t *out) {
    const char *p = s;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    if (p >= e || e - p > 18) return false;
    int64_t v = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = neg ? -v : v;
    return true;
}

struct PieceResult {
    int64_t lines = 0;       // non-blank lines
    int64_t bad_line = -1;   // byte offset of the first malformed line, -1 if none
    int all_unit = 1;
};

// one pass over [b, e): count non-blank lines (fill == false) or parse them into the arrays starting at `at`
void edge_piece(const Mapped &m, size_t b, size_t e, bool fill, int64_t at, int64_t *src, int64_t *dst, double *w,
                PieceResult *res) {
    const char *p = m.p + b, *end = m.p + e;
    PieceResult r;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        // tokens of the line
        const char *tok_s[4], *tok_e[4];
        int nt = 0;
        const char *q = p;
        while (q < le) {
            while (q < le && is_blank(*q)) q++;
            if (q >= le) break;
            const char *t0 = q;
            while (q < le && !is_blank(*q)) q++;
            if (nt < 4) { tok_s[nt] = t0; tok_e[nt] = q; }
            nt++;
        }
        if (nt > 0) {
            if (fill) {
                int64_t a = 0, c = 0;
                double ww = 1.0;
                bool ok = nt >= 2 && parse_int(tok_s[0], tok_e[0], &a) && parse_int(tok_s[1], tok_e[1], &c);
                if (ok && nt == 3) ok = parse_double(tok_s[2], tok_e[2], &ww);
                if (!ok) { if (r.bad_line < 0) r.bad_line = (int64_t)(p - m.p); }
                else {
                    src[at + r.lines] = a;
                    dst[at + r.lines] = c;
                    if (w) w[at + r.lines] = ww;
                    if (ww != 1.0) r.all_unit = 0;
                }
            }
            r.lines++;
        }
        p = nl ? nl + 1 : end;
    }
    *res = r;
}

int run_edge_pass(const Mapped &m, size_t begin, bool fill, const std::vector<int64_t> *starts, int64_t *src,
                  int64_t *dst, double *w, std::vector<PieceResult> &res, const std::vector<size_t> &b) {
    const int T = (int)b.size() - 1;
    res.assign(T, PieceResult());
    std::vector<std::thread> th;
    for (int i = 0; i < T; i++)
        th.emplace_back(edge_piece, std::cref(m), b[i], b[i + 1], fill, starts ? (*starts)[i] : 0, src, dst, w, &res[i]);
    for (auto &t : th) t.join();
    (void)begin;
    return GEMB_OK;
}

}  // namespace

extern "C" {

int gemb_edge_list_scan(const char *path, int64_t skip, int64_t *n_edges) {
    GEMB_ARG(path && n_edges && skip >= 0, "path/n_edges/skip");
    Mapped m;
    GEMB_TRY(map_file(path, m));
    const size_t begin = skip_lines(m, skip);
    const std::vector<size_t> b = cut(m, begin, n_threads(m.n - begin));
    std::vector<PieceResult> res;
    run_edge_pass(m, begin, false, nullptr, nullptr, nullptr, nullptr, res, b);
    int64_t total = 0;
    for (auto &r : res) total += r.lines;
    *n_edges = total;
    return GEMB_OK;
}

int gemb_edge_list_parse(const char *path, int64_t skip, int64_t n_edges, int64_t *src, int64_t *dst, double *w,
                         int32_t *all_unit) {
    GEMB_ARG(path && skip >= 0 && n_edges >= 0 && (n_edges == 0 || (src && dst)), "path/arrays");
    Mapped m;
    GEMB_TRY(map_file(path, m));
    const size_t begin = skip_lines(m, skip);
    const std::vector<size_t> b = cut(m, begin, n_threads(m.n - begin));
    std::vector<PieceResult> res;
    run_edge_pass(m, begin, false, nullptr, nullptr, nullptr, nullptr, res, b);
    std::vector<int64_t> starts(res.size(), 0);
    int64_t total = 0;
    for (size_t i = 0; i < res.size(); i++) { starts[i] = total; total += res[i].lines; }
    if (total != n_edges) {
        gemb::set_error("gemb_edge_list_parse: %s has %lld edge lines, caller expected %lld", path, (long long)total, (long long)n_edges);
        return GEMB_ERR_ARG;
    }
    run_edge_pass(m, begin, true, &starts, src, dst, w, res, b);
    int unit = 1;
    for (auto &r : res) {
        if (r.bad_line >= 0) {
            const char *ls = m.p + r.bad_line;
            const char *le = (const char *)memchr(ls, '\n', m.n - (size_t)r.bad_line);
            const int len = (int)std::min<size_t>(le ? (size_t)(le - ls) : m.n - (size_t)r.bad_line, 60);
            gemb::set_error("gemb_edge_list_parse: malformed line at byte %lld of %s: '%.*s'", (long long)r.bad_line, path, len, ls);
            return GEMB_ERR_ARG;
        }
        unit &= r.all_unit;
    }
    if (all_unit) *all_unit = unit;
    return GEMB_OK;
}

// "%d %d %f\n" per edge; header_nodes >= 0 writes the two header lines of saveGraphToEdgeListTxt first
int gemb_edge_list_write(const char *path, int64_t n_edges, const int64_t *src, const int64_t *dst, const double *w,
                         int64_t header_nodes) {
    GEMB_ARG(path && n_edges >= 0 && (n_edges == 0 || (src && dst)), "path/arrays");
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads((size_t)n_edges * 24), n_edges / 65536 + 1));
    std::vector<std::string> out(T);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            const int64_t b0 = n_edges * t / T, b1 = n_edges * (t + 1) / T;
            std::string &s = out[t];
            s.reserve((size_t)(b1 - b0) * 28);
            char buf[400];
            for (int64_t i = b0; i < b1; i++) {
                const int len = snprintf(buf, sizeof buf, "%lld %lld %f\n", (long long)src[i], (long long)dst[i], w ? w[i] : 1.0);
                s.append(buf, (size_t)len);
            }
        });
    for (auto &t : th) t.join();
    FILE *f = fopen(path, "wb");
    if (!f) { gemb::set_error("cannot create %s: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    if (header_nodes >= 0) fprintf(f, "%lld\n%lld\n", (long long)header_nodes, (long long)n_edges);
    bool ok = true;
    for (auto &s : out) ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { gemb::set_error("write to %s failed: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    return GEMB_OK;
}

// ".emb": first line "<rows> <d>", then "<id> v1 ... vd".  X == NULL: only the header is read.
// X is rows x d fp64, zero-initialised by the caller; row = id (loadEmbedding, graph_util.py:168).
int gemb_emb_read(const char *path, int64_t *rows, int32_t *d, double *X) {
    GEMB_ARG(path && rows && d, "path/rows/d");
    Mapped m;
    GEMB_TRY(map_file(path, m));
    const char *nl = m.n ? (const char *)memchr(m.p, '\n', m.n) : nullptr;
    const char *he = nl ? nl : m.p + m.n;
    int64_t r = 0, dd = 0;
    {
        const char *q = m.p;
        while (q < he && is_blank(*q)) q++;
        const char *a0 = q;
        while (q < he && !is_blank(*q)) q++;
        const char *a1 = q;
        while (q < he && is_blank(*q)) q++;
        const char *b0 = q;
        while (q < he && !is_blank(*q)) q++;
        if (!parse_int(a0, a1, &r) || !parse_int(b0, q, &dd) || r < 0 || dd <= 0 || dd > (1 << 24)) {
            gemb::set_error("gemb_emb_read: bad header in %s", path);
            return GEMB_ERR_ARG;
        }
    }
    *rows = r;
    *d = (int32_t)dd;
    if (!X) return GEMB_OK;
    const size_t begin = nl ? (size_t)(nl - m.p) + 1 : m.n;
    const std::vector<size_t> b = cut(m, begin, n_threads(m.n - begin));
    const int T = (int)b.size() - 1;
    std::vector<int64_t> bad(T, -1);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            const char *p = m.p + b[t], *end = m.p + b[t + 1];
            while (p < end) {
                const char *l_nl = (const char *)memchr(p, '\n', (size_t)(end - p));
                const char *le = l_nl ? l_nl : end;
                const char *q = p;
                while (q < le && is_blank(*q)) q++;
                if (q < le) {
                    const char *t0 = q;
                    while (q < le && !is_blank(*q)) q++;
                    int64_t id = -1;
                    bool ok = parse_int(t0, q, &id) && id >= 0 && id < r;
                    double *row = ok ? X + (size_t)id * dd : nullptr;
                    int64_t c = 0;
                    while (ok) {
                        while (q < le && is_blank(*q)) q++;
                        if (q >= le) break;
                        const char *v0 = q;
                        while (q < le && !is_blank(*q)) q++;
                        double v = 0.0;
                        if (c >= dd || !parse_double(v0, q, &v)) { ok = false; break; }
                        row[c++] = v;
                    }
                    if (!ok || c != dd) { if (bad[t] < 0) bad[t] = (int64_t)(p - m.p); }
                }
                p = l_nl ? l_nl + 1 : end;
            }
        });
    for (auto &t : th) t.join();
    for (int t = 0; t < T; t++)
        if (bad[t] >= 0) {
            gemb::set_error("gemb_emb_read: malformed line at byte %lld of %s (id outside [0, rows) or not %lld values)",
                            (long long)bad[t], path, (long long)dd);
            return GEMB_ERR_ARG;
        }
    return GEMB_OK;
}

// ids == NULL: rows 0..n_ids-1 in order.  Values with 6 significant digits ("%g": what C++ ostream << double prints).
int gemb_emb_write(const char *path, int64_t n_ids, const int64_t *ids, int32_t d, const double *X, int64_t header_rows) {
    GEMB_ARG(path && n_ids >= 0 && d > 0 && (n_ids == 0 || X), "path/arrays");
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads((size_t)n_ids * d * 10), n_ids / 4096 + 1));
    std::vector<std::string> out(T);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            const int64_t b0 = n_ids * t / T, b1 = n_ids * (t + 1) / T;
            std::string &s = out[t];
            s.reserve((size_t)(b1 - b0) * ((size_t)d * 12 + 12));
            char buf[64];
            for (int64_t i = b0; i < b1; i++) {
                const int64_t id = ids ? ids[i] : i;
                int len = snprintf(buf, sizeof buf, "%lld", (long long)id);
                s.append(buf, (size_t)len);
                const double *row = X + (size_t)id * d;
                for (int c = 0; c < d; c++) {
                    len = snprintf(buf, sizeof buf, " %g", row[c]);
                    s.append(buf, (size_t)len);
                }
                s.push_back('\n');
            }
        });
    for (auto &t : th) t.join();
    FILE *f = fopen(path, "wb");
    if (!f) { gemb::set_error("cannot create %s: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    fprintf(f, "%lld %d\n", (long long)header_rows, (int)d);
    bool ok = true;
    for (auto &s : out) ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { gemb::set_error("write to %s failed: %s", path, strerror(errno)); return GEMB_ERR_ARG; }
    return GEMB_OK;
}

}  // extern "C"
