// sdpa_reader.cpp -- fast SDPA sparse-format (.dat-s) ingest with the output convention of the reference reader
// LReadSDPA (src_semi/io/lorads_file_io.c:21-418): per PSD block a CSC matrix with m+1 columns over the packed
// lower-triangular index PACK_IDX(n, max(i,j), min(i,j)) (lorads_utils.h:45), column 0 = objective block NEGATED
// (:279-281), entries with |v| < 1e-12 dropped (:250-256), a trailing negative block dimension = LP block (:149-151),
// duplicates not summed, entries of a column kept in file order (dcs_compress is a stable counting sort).
// The reference parses with fgets + sscanf per line and CSparse triplet growth; here the file is read once into
// memory, numbers are parsed in place and the CSC arrays are filled by a two-pass counting sort.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lorads_b200.h"

struct lb2_sdpa {
    lb2_int m = 0, nLpCols = 0;
    std::vector<lb2_int> dims;
    std::vector<double> rhs;
    struct Block { std::vector<lb2_int> beg, idx; std::vector<double> elem; };
    std::vector<Block> blocks;
    Block lp;      // LP block as CSC over columns 0..m (row = LP column index), kept for completeness
    lb2_int nElems = 0;
};

namespace {

thread_local std::string g_reader_err;

inline void skip_ws(const char *&p, const char *end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
}
inline void skip_line(const char *&p, const char *end) {
    while (p < end && *p != '\n') ++p;
    if (p < end) ++p;
}
// Both number parsers stop at the end of the current line: strtoll / strtod would skip the newline as white space and
// silently take a token of the next entry (or of the next thread's chunk), where the reference's per-line sscanf fails.
inline bool parse_int(const char *&p, const char *end, long long &v) {
    skip_ws(p, end);
    if (p >= end || *p == '\n') return false;
    char *q = nullptr;
    errno = 0;
    v = std::strtoll(p, &q, 10);
    if (q == p) return false;
    p = q;
    return true;
}
inline bool parse_double(const char *&p, const char *end, double &v) {
    skip_ws(p, end);
    if (p >= end || *p == '\n') return false;
    char *q = nullptr;
    v = std::strtod(p, &q);
    if (q == p) return false;
    p = q;
    return true;
}
inline bool is_punct(char c) { return c == '{' || c == '}' || c == '(' || c == ')' || c == ',' || c == '\''; }

// Binary instance cache (lb2_sdpa_save / auto-detected by lb2_read_sdpa): the reader's output arrays as they are, native
// little-endian: magic, m, #PSD blocks, nLpCols, nElems, dims[], rhs[], then per block (PSD blocks, then the LP block
// when nLpCols > 0): nnz, beg[m+2], idx[nnz], elem[nnz].
const char kBinMagic[9] = "LB2SDPA1";

int load_binary(const char *p, const char *end, lb2_sdpa **out) {
    try {
        auto take = [&](void *dst, size_t bytes) {
            if ((size_t)(end - p) < bytes) throw std::runtime_error("binary instance file is truncated");
            std::memcpy(dst, p, bytes);
            p += bytes;
        };
        p += 8;
        std::unique_ptr<lb2_sdpa> S(new lb2_sdpa());
        lb2_int hdr[4];
        take(hdr, sizeof(hdr));
        S->m = hdr[0]; S->nLpCols = hdr[2]; S->nElems = hdr[3];
        const lb2_int nblk = hdr[1];
        if (S->m <= 0 || nblk < 0 || nblk > (1 << 24) || S->nLpCols < 0) throw std::runtime_error("bad binary instance header");
        S->dims.resize((size_t)nblk);
        take(S->dims.data(), sizeof(lb2_int) * (size_t)nblk);
        S->rhs.resize((size_t)S->m);
        take(S->rhs.data(), sizeof(double) * (size_t)S->m);
        S->blocks.resize((size_t)nblk);
        auto block = [&](lb2_sdpa::Block &B, lb2_int index_limit) {
            lb2_int nnz = 0;
            take(&nnz, sizeof(nnz));
            if (nnz < 0 || (size_t)nnz > (size_t)(end - p) / sizeof(double)) throw std::runtime_error("bad binary instance block");
            B.beg.resize((size_t)S->m + 2); B.idx.resize((size_t)nnz); B.elem.resize((size_t)nnz);
            take(B.beg.data(), sizeof(lb2_int) * B.beg.size());
            take(B.idx.data(), sizeof(lb2_int) * (size_t)nnz);
            take(B.elem.data(), sizeof(double) * (size_t)nnz);
            // a cache file is data from outside: column pointers must ascend from 0 to nnz, indices must be in range
            if (B.beg.front() != 0 || B.beg.back() != nnz) throw std::runtime_error("bad binary instance block");
            for (size_t c = 0; c + 1 < B.beg.size(); ++c)
                if (B.beg[c] > B.beg[c + 1]) throw std::runtime_error("bad binary instance block");
            for (lb2_int v : B.idx)
                if (v < 0 || v >= index_limit) throw std::runtime_error("bad binary instance block");
        };
        for (lb2_int k = 0; k < nblk; ++k) {
            const lb2_int n = S->dims[(size_t)k];
            if (n <= 0 || n > (lb2_int)3000000000LL) throw std::runtime_error("bad binary instance header");
            block(S->blocks[(size_t)k], n * (n + 1) / 2);
        }
        if (S->nLpCols > 0) block(S->lp, S->nLpCols);
        *out = S.release();
    } catch (const std::exception &e) {
        g_reader_err = e.what();
        return LB2_ERR_ARG;
    }
    return LB2_OK;
}

}  // namespace

extern "C" {

const char *lb2_sdpa_last_error(void) { return g_reader_err.c_str(); }

int lb2_sdpa_save(const lb2_sdpa *s, const char *path) {
    if (!s || !path) return LB2_ERR_ARG;
    FILE *f = std::fopen(path, "wb");
    if (!f) { g_reader_err = std::string("cannot create ") + path; return LB2_ERR_ARG; }
    bool ok = true;
    auto put = [&](const void *src, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(src, 1, bytes, f) == bytes); };
    const lb2_int hdr[4] = {s->m, (lb2_int)s->dims.size(), s->nLpCols, s->nElems};
    put(kBinMagic, 8); put(hdr, sizeof(hdr));
    put(s->dims.data(), sizeof(lb2_int) * s->dims.size());
    put(s->rhs.data(), sizeof(double) * s->rhs.size());
    auto block = [&](const lb2_sdpa::Block &B) {
        const lb2_int nnz = (lb2_int)B.idx.size();
        put(&nnz, sizeof(nnz));
        put(B.beg.data(), sizeof(lb2_int) * B.beg.size());
        put(B.idx.data(), sizeof(lb2_int) * B.idx.size());
        put(B.elem.data(), sizeof(double) * B.elem.size());
    };
    for (const lb2_sdpa::Block &B : s->blocks) block(B);
    if (s->nLpCols > 0) block(s->lp);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { g_reader_err = "write failed"; return LB2_ERR_ARG; }
    return LB2_OK;
}

int lb2_read_sdpa(const char *path, lb2_sdpa **out) {
    if (!path || !out) return LB2_ERR_ARG;
    *out = nullptr;
    FILE *f = std::fopen(path, "rb");
    if (!f) { g_reader_err = std::string("cannot open ") + path; return LB2_ERR_ARG; }
    if (std::fseek(f, 0, SEEK_END) != 0) { std::fclose(f); g_reader_err = "cannot seek in the instance file"; return LB2_ERR_ARG; }
    const long sz = std::ftell(f);
    if (sz < 0 || std::fseek(f, 0, SEEK_SET) != 0) { std::fclose(f); g_reader_err = "cannot determine the size of the instance file"; return LB2_ERR_ARG; }
    std::vector<char> buf((size_t)sz + 1);
    if (sz > 0 && std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { std::fclose(f); g_reader_err = "short read"; return LB2_ERR_ARG; }
    std::fclose(f);
    buf[(size_t)sz] = '\0';
    const char *p = buf.data(), *end = buf.data() + sz;
    if (sz >= 8 && std::memcmp(p, kBinMagic, 8) == 0) return load_binary(p, end, out);
    try {
        lb2_sdpa *S = new lb2_sdpa();
        std::unique_ptr<lb2_sdpa> guard(S);
        // comments: lines starting with '*' or '"'
        while (p < end && (*p == '*' || *p == '"')) skip_line(p, end);
        long long v;
        if (!parse_int(p, end, v) || v <= 0) throw std::runtime_error("bad constraint count");
        S->m = v;
        skip_line(p, end);
        if (!parse_int(p, end, v) || v <= 0) throw std::runtime_error("bad block count");
        long long nblk = v;
        skip_line(p, end);
        // block dimensions, possibly decorated with { } ( ) , '
        std::vector<long long> dims;
        while ((long long)dims.size() < nblk && p < end) {
            skip_ws(p, end);
            if (p < end && (is_punct(*p) || *p == '\n')) { ++p; continue; }
            if (!parse_int(p, end, v)) throw std::runtime_error("bad block dimension");
            dims.push_back(v);
        }
        if ((long long)dims.size() != nblk) throw std::runtime_error("missing block dimensions");
        skip_line(p, end);
        for (long long k = 0; k < nblk; ++k) {
            if (dims[k] < 0 && k == nblk - 1) { S->nLpCols = -dims[k]; }
            else if (dims[k] <= 0) throw std::runtime_error("only one diagonal (LP) block is supported and it must be the last one");
            else S->dims.push_back(dims[k]);
        }
        const long long nsdp = (long long)S->dims.size();
        // right-hand side
        S->rhs.resize((size_t)S->m);
        for (lb2_int i = 0; i < S->m; ++i) {
            skip_ws(p, end);
            while (p < end && (is_punct(*p) || *p == '\n')) { ++p; skip_ws(p, end); }
            if (!parse_double(p, end, S->rhs[(size_t)i])) throw std::runtime_error("bad right-hand side");
        }
        skip_line(p, end);
        // entries: the rest of the file is cut at line boundaries into one chunk per thread; every thread collects
        // the triplets of its chunk per block, and the chunks are concatenated in file order afterwards (the reader's
        // output keeps file order inside a constraint column)
        struct Trip { lb2_int con, pos; double val; };
        struct Part {
            std::vector<std::vector<Trip>> trips;
            std::vector<Trip> lp;
            lb2_int nElems = 0;
            bool stop = false;          // BEGIN.COMMENT seen: everything after this chunk position is ignored
            std::string err;
        };
        const size_t body = (size_t)(end - p);
        unsigned nth = 1;
        if (body > (size_t)(8u << 20)) nth = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<const char *> cut(nth + 1);
        cut[0] = p; cut[nth] = end;
        for (unsigned t = 1; t < nth; ++t) {
            const char *q = p + body * t / nth;
            while (q < end && *q != '\n') ++q;
            cut[t] = (q < end) ? q + 1 : end;
        }
        std::vector<Part> parts(nth);
        const lb2_int mrows = S->m, nlp = S->nLpCols;
        const std::vector<lb2_int> &bdims = S->dims;
        auto work = [&](unsigned t) {
            Part &P = parts[t];
            P.trips.resize((size_t)nsdp);
            const char *q = cut[t], *qe = cut[t + 1];
            while (q < qe) {
                skip_ws(q, qe);
                if (q >= qe) break;
                if (*q == '\n') { ++q; continue; }
                long long con, blk, i, j;
                double val;
                const char *line = q;
                if (!parse_int(q, qe, con) || !parse_int(q, qe, blk) || !parse_int(q, qe, i) || !parse_int(q, qe, j) || !parse_double(q, qe, val)) {
                    if (std::strncmp(line, "BEGIN.COMMENT", 13) == 0) { P.stop = true; return; }
                    P.err = "malformed entry line";
                    return;
                }
                skip_line(q, qe);
                if (con < 0 || con > mrows || blk < 1 || blk > nblk) { P.err = "entry index out of range"; return; }
                if (std::fabs(val) < 1e-12) continue;
                if (con == 0) val = -val;
                blk -= 1; i -= 1; j -= 1;
                if (nlp > 0 && blk == nsdp) {
                    if (i < 0 || i >= nlp) { P.err = "LP column index out of range"; return; }
                    P.lp.push_back({(lb2_int)con, (lb2_int)i, val});
                } else {
                    const long long n = bdims[(size_t)blk];
                    const long long hi = i > j ? i : j, lo = i > j ? j : i;
                    if (lo < 0 || hi >= n) { P.err = "matrix index out of range"; return; }
                    P.trips[(size_t)blk].push_back({(lb2_int)con, (lb2_int)((2 * n - lo - 1) * lo / 2 + hi), val});
                }
                P.nElems += 1;
            }
        };
        if (nth == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nth; ++t) th.emplace_back(work, t);
            for (std::thread &x : th) x.join();
        }
        std::vector<std::vector<Trip>> trips((size_t)nsdp);
        std::vector<Trip> lptrips;
        for (unsigned t = 0; t < nth; ++t) {
            if (!parts[t].err.empty()) throw std::runtime_error(parts[t].err);
            for (long long k = 0; k < nsdp; ++k)
                trips[(size_t)k].insert(trips[(size_t)k].end(), parts[t].trips[(size_t)k].begin(), parts[t].trips[(size_t)k].end());
            lptrips.insert(lptrips.end(), parts[t].lp.begin(), parts[t].lp.end());
            S->nElems += parts[t].nElems;
            if (parts[t].stop) break;
        }
        auto compress = [&](const std::vector<Trip> &t, lb2_sdpa::Block &B) {
            B.beg.assign((size_t)S->m + 2, 0);
            for (const Trip &e : t) B.beg[(size_t)e.con + 1]++;
            for (lb2_int c = 0; c <= S->m; ++c) B.beg[(size_t)c + 1] += B.beg[(size_t)c];
            B.idx.resize(t.size()); B.elem.resize(t.size());
            std::vector<lb2_int> cur(B.beg.begin(), B.beg.end() - 1);
            for (const Trip &e : t) { lb2_int q = cur[(size_t)e.con]++; B.idx[(size_t)q] = e.pos; B.elem[(size_t)q] = e.val; }
        };
        S->blocks.resize((size_t)nsdp);
        for (long long k = 0; k < nsdp; ++k) compress(trips[(size_t)k], S->blocks[(size_t)k]);
        if (S->nLpCols > 0) compress(lptrips, S->lp);
        *out = guard.release();
    } catch (const std::exception &e) {
        g_reader_err = e.what();
        return LB2_ERR_ARG;
    }
    return LB2_OK;
}

lb2_int lb2_sdpa_info(const lb2_sdpa *s, int what, lb2_int k) {
    if (!s) return -1;
    switch (what) {
    case 0: return s->m;
    case 1: return (lb2_int)s->dims.size();
    case 4: return s->nLpCols;
    case 5: return s->nElems;
    }
    if (what == 3 && k == (lb2_int)s->dims.size() && s->nLpCols > 0) return (lb2_int)s->lp.idx.size();   // the LP block
    if (k < 0 || k >= (lb2_int)s->dims.size()) return -1;
    if (what == 2) return s->dims[(size_t)k];
    if (what == 3) return (lb2_int)s->blocks[(size_t)k].idx.size();
    return -1;
}

int lb2_sdpa_get(const lb2_sdpa *s, lb2_int k, lb2_int *beg, lb2_int *idx, double *elem, double *rhs) {
    if (!s) return LB2_ERR_ARG;
    if (rhs) std::memcpy(rhs, s->rhs.data(), sizeof(double) * s->rhs.size());
    if (k < 0) return LB2_OK;
    const bool want_lp = (k == (lb2_int)s->dims.size() && s->nLpCols > 0);
    if (k >= (lb2_int)s->dims.size() && !want_lp) return LB2_ERR_ARG;
    const lb2_sdpa::Block &B = want_lp ? s->lp : s->blocks[(size_t)k];
    if (beg) std::memcpy(beg, B.beg.data(), sizeof(lb2_int) * B.beg.size());
    if (idx) std::memcpy(idx, B.idx.data(), sizeof(lb2_int) * B.idx.size());
    if (elem) std::memcpy(elem, B.elem.data(), sizeof(double) * B.elem.size());
    return LB2_OK;
}

void lb2_sdpa_free(lb2_sdpa *s) { delete s; }

}  // extern "C"
