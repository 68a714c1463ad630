// Host-side text I/O either side of the scoring path: the trial-list reader and the
// score-file writer.  The reference parses trial files with np.genfromtxt(dtype='str') and maps ids through
// Python dicts row by row (sv_trials_loaders.py:376-383, 399-406; scorefile_generator.py:26, 45), and writes
// scores with ndarray.astype(str) + np.savetxt (scorefile_generator.py:37-38, 54-55).  With the scoring itself
// at >1 G trials/s that text handling is the wall clock of generate_*_scores; this file does it in one pass
// over the bytes.  Plain C++ (no device code); same file formats, byte for byte.
#include <charconv>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/nplda.h"

struct nplda_trials {
    std::string data;                 // whole file
    std::vector<uint32_t> off, len;   // rows * cols fields (files < 4 GiB)
    int64_t rows = 0;
    int cols = 0;
    std::string_view field(int64_t r, int c) const {
        const size_t i = (size_t)r * cols + c;
        return std::string_view(data.data() + off[i], len[i]);
    }
};

namespace {

// os.path.splitext on the last path component: the extension starts at the last '.', unless that dot leads
// the component (".bashrc" has no extension) -- leading dots are skipped as CPython does.
std::string_view splitext_root(std::string_view s) {
    const size_t slash = s.rfind('/');
    const size_t base = slash == std::string_view::npos ? 0 : slash + 1;
    const size_t dot = s.rfind('.');
    if (dot == std::string_view::npos || dot < base) return s;
    size_t lead = base;
    while (lead < s.size() && s[lead] == '.') ++lead;
    if (dot < lead) return s;
    return s.substr(0, dot);
}
std::string_view basename_of(std::string_view s) {
    const size_t slash = s.rfind('/');
    return slash == std::string_view::npos ? s : s.substr(slash + 1);
}

// str(np.float32(v)) / ndarray.astype(str): shortest digits that round-trip in float32; positional notation when
// 1e-4 <= |v| < 1e6 (numpy's float32 bounds, compared on the value: float32(1e-4) = 9.9999997e-05 prints as
// "1e-04", 999999.94 as "999999.94", 1e6 as "1e+06"; probed against numpy 2.3 over 5e5 values), else scientific
// with a two-digit (at least) exponent; integral values keep one trailing ".0".
int format_f32_numpy(float v, char *out) {
    if (std::isnan(v)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(v)) { const char *s = v < 0 ? "-inf" : "inf"; const int n = (int)strlen(s); memcpy(out, s, n); return n; }
    char *p = out;
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.f) { memcpy(p, "0.0", 3); return (int)(p - out) + 3; }
    char sci[48];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    const std::string_view s(sci, res.ptr - sci);
    const size_t epos = s.find('e');
    char digits[16];
    int nd = 0;
    for (size_t i = 0; i < epos; ++i) if (s[i] != '.') digits[nd++] = s[i];
    const int e10 = atoi(std::string(s.substr(epos + 1)).c_str());
    const bool scientific = (double)v < 1e-4 || (double)v >= 1e6;     // numpy decides on the VALUE; float32 bounds
    if (scientific) {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        *p++ = e10 < 0 ? '-' : '+';
        const int ae = e10 < 0 ? -e10 : e10;
        if (ae < 10) *p++ = '0';
        p = std::to_chars(p, p + 4, ae).ptr;
    } else if (e10 < 0) {
        *p++ = '0'; *p++ = '.';
        for (int z = 0; z < -e10 - 1; ++z) *p++ = '0';
        memcpy(p, digits, nd); p += nd;
    } else {
        const int ip = e10 + 1;                       // digits before the point
        for (int i = 0; i < ip; ++i) *p++ = i < nd ? digits[i] : '0';
        *p++ = '.';
        if (nd > ip) { memcpy(p, digits + ip, nd - ip); p += nd - ip; }
        else *p++ = '0';
    }
    return (int)(p - out);
}

}  // namespace

extern "C" int nplda_format_f32(float v, char *out24) { return format_f32_numpy(v, out24); }

extern "C" int nplda_trials_open(const char *path, nplda_trials **out) {
    if (!path || !out) return NPLDA_ERR_BAD_ARG;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return NPLDA_ERR_IO;
    nplda_trials *t = new nplda_trials();
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0 || (uint64_t)sz >= 0xFFFFFFFFull) { fclose(f); delete t; return NPLDA_ERR_IO; }
    t->data.resize((size_t)sz);
    const size_t got = sz ? fread(&t->data[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) { delete t; return NPLDA_ERR_IO; }
    // np.genfromtxt defaults: fields split on runs of whitespace, '#' starts a comment, empty lines are skipped,
    // every remaining line must have the same number of fields.
    const char *d = t->data.data();
    size_t i = 0;
    const size_t n = t->data.size();
    while (i < n) {
        size_t eol = i;
        while (eol < n && d[eol] != '\n') ++eol;
        size_t end = i;
        while (end < eol && d[end] != '#') ++end;
        int nf = 0;
        size_t j = i;
        while (j < end) {
            while (j < end && (d[j] == ' ' || d[j] == '\t' || d[j] == '\r' || d[j] == '\v' || d[j] == '\f')) ++j;
            if (j >= end) break;
            size_t k = j;
            while (k < end && !(d[k] == ' ' || d[k] == '\t' || d[k] == '\r' || d[k] == '\v' || d[k] == '\f')) ++k;
            t->off.push_back((uint32_t)j);
            t->len.push_back((uint32_t)(k - j));
            ++nf;
            j = k;
        }
        if (nf > 0) {
            if (t->rows == 0) t->cols = nf;
            else if (nf != t->cols) { delete t; return NPLDA_ERR_FORMAT; }
            ++t->rows;
        }
        i = eol + 1;
    }
    *out = t;
    return NPLDA_OK;
}

extern "C" void nplda_trials_close(nplda_trials *t) { delete t; }
extern "C" int64_t nplda_trials_rows(const nplda_trials *t) { return t ? t->rows : NPLDA_ERR_BAD_ARG; }
extern "C" int nplda_trials_cols(const nplda_trials *t) { return t ? t->cols : NPLDA_ERR_BAD_ARG; }

extern "C" int64_t nplda_trials_field(const nplda_trials *t, int64_t row, int col, const char **ptr) {
    if (!t || !ptr || row < 0 || row >= t->rows || col < 0 || col >= t->cols) return NPLDA_ERR_BAD_ARG;
    const std::string_view v = t->field(row, col);
    *ptr = v.data();
    return (int64_t)v.size();
}

extern "C" int nplda_trials_map_ids(const nplda_trials *t, int col, int mode, const char *ids, int64_t ids_bytes,
                                    int64_t n_ids, const int64_t *values, int64_t first_row, int64_t *idx_out) {
    if (!t || col < 0 || col >= t->cols || !idx_out || n_ids < 0 || (n_ids > 0 && !ids) || mode < 0 || mode > 2 ||
        first_row < 0)
        return NPLDA_ERR_BAD_ARG;
    std::unordered_map<std::string_view, int64_t> map;
    map.reserve((size_t)n_ids * 2);
    const char *p = ids, *end = ids + ids_bytes;
    for (int64_t r = 0; r < n_ids; ++r) {          // ids: n_ids strings, each terminated by '\n'
        const char *q = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!q) return NPLDA_ERR_BAD_ARG;
        map[std::string_view(p, (size_t)(q - p))] = values ? values[r] : r;     // later duplicates win, like a dict
        p = q + 1;
    }
    for (int64_t r = first_row; r < t->rows; ++r) {
        std::string_view v = t->field(r, col);
        if (mode == 2) v = basename_of(v);
        if (mode >= 1) v = splitext_root(v);
        const auto it = map.find(v);
        idx_out[r - first_row] = it == map.end() ? -1 : it->second;
    }
    return NPLDA_OK;
}

extern "C" int nplda_trials_col_float(const nplda_trials *t, int col, int64_t first_row, float *out, uint8_t *ok) {
    if (!t || col < 0 || col >= t->cols || !out || !ok || first_row < 0) return NPLDA_ERR_BAD_ARG;
    std::string tmp;
    for (int64_t r = first_row; r < t->rows; ++r) {
        const std::string_view v = t->field(r, col);
        tmp.assign(v.data(), v.size());
        char *endp = nullptr;
        errno = 0;
        const double x = strtod(tmp.c_str(), &endp);
        const bool good = !tmp.empty() && endp == tmp.c_str() + tmp.size() && tmp.find_first_of("xXpP") == std::string::npos;
        out[r - first_row] = good ? (float)x : 0.f;
        ok[r - first_row] = good ? 1 : 0;
    }
    return NPLDA_OK;
}

extern "C" int nplda_scores_write(const char *path, const nplda_trials *t, int64_t first_row, int ncols_keep,
                                  const float *scores, const char *header_line) {
    if (!path || !t || ncols_keep < 0 || ncols_keep > t->cols || first_row < 0 || first_row > t->rows) return NPLDA_ERR_BAD_ARG;
    if (t->rows - first_row > 0 && !scores) return NPLDA_ERR_BAD_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return NPLDA_ERR_IO;
    std::string buf;
    buf.reserve(1 << 22);
    if (header_line) { buf.append(header_line); buf.push_back('\n'); }
    char num[32];
    for (int64_t r = first_row; r < t->rows; ++r) {
        for (int c = 0; c < ncols_keep; ++c) {
            const std::string_view v = t->field(r, c);
            buf.append(v.data(), v.size());
            buf.push_back('\t');
        }
        buf.append(num, (size_t)format_f32_numpy(scores[r - first_row], num));
        buf.push_back('\n');
        if (buf.size() > (1u << 22) - 4096) {
            if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); return NPLDA_ERR_IO; }
            buf.clear();
        }
    }
    const bool okw = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    return (fclose(f) == 0 && okw) ? NPLDA_OK : NPLDA_ERR_IO;
}
