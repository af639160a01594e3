// ptk_csv_reader.h -- host-side columnar reader for rods_df_{color}.csv ("next" row).
//
// Replaces `pd.read_csv(f_in, sep=",", index_col=0)` at match2D.py:601-602 / 1053 and matchND.py:317:
// pandas needs ~60 ms per colour of a 1000-frame experiment (20x the GPU time of the whole
// experiment).  This reader memory-maps the file, finds the record starts, then classifies and
// parses the columns in parallel over row ranges.  It follows pandas' C parser where results depend
// on it:
//   * floats go through the same arithmetic as pandas' default `precise_xstrtod` (at most 17
//     significant digits accumulated in a double, one multiplication or division by a power of ten)
//     so that every value is bit-identical to what the reference reads (tests/test_csv_reader.py
//     checks millions of strings against pandas);
//   * a column is int64 if every field is an integer, float64 if every field is a number or one of
//     pandas' NA strings (-> NaN), bool for True/False, else string (NA strings -> missing);
//   * blank lines are skipped, "\r\n" is accepted, fields may be quoted with '"' ("" escapes).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace ptk {

enum CsvKind { CSV_INT64 = 0, CSV_FLOAT64 = 1, CSV_STRING = 2, CSV_BOOL = 3 };

static inline bool csv_isdigit(char c) { return c >= '0' && c <= '9'; }
static inline bool csv_isspace(char c) { return c == ' ' || c == '\t'; }

// pandas/_libs/src/parser/tokenizer.c: precise_xstrtod (decimal '.', sci 'E', no thousands separator,
// trailing whitespace skipped).  Returns false if [s, e) is not entirely a number.
inline bool csv_parse_double(const char* s, const char* e, double* out)
{
    static const double p10[] = {
        1e0,   1e1,   1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,  1e15,
        1e16,  1e17,  1e18,  1e19,  1e20,  1e21,  1e22,  1e23,  1e24,  1e25,  1e26,  1e27,  1e28,  1e29,  1e30,  1e31,
        1e32,  1e33,  1e34,  1e35,  1e36,  1e37,  1e38,  1e39,  1e40,  1e41,  1e42,  1e43,  1e44,  1e45,  1e46,  1e47,
        1e48,  1e49,  1e50,  1e51,  1e52,  1e53,  1e54,  1e55,  1e56,  1e57,  1e58,  1e59,  1e60,  1e61,  1e62,  1e63,
        1e64,  1e65,  1e66,  1e67,  1e68,  1e69,  1e70,  1e71,  1e72,  1e73,  1e74,  1e75,  1e76,  1e77,  1e78,  1e79,
        1e80,  1e81,  1e82,  1e83,  1e84,  1e85,  1e86,  1e87,  1e88,  1e89,  1e90,  1e91,  1e92,  1e93,  1e94,  1e95,
        1e96,  1e97,  1e98,  1e99,  1e100, 1e101, 1e102, 1e103, 1e104, 1e105, 1e106, 1e107, 1e108, 1e109, 1e110, 1e111,
        1e112, 1e113, 1e114, 1e115, 1e116, 1e117, 1e118, 1e119, 1e120, 1e121, 1e122, 1e123, 1e124, 1e125, 1e126, 1e127,
        1e128, 1e129, 1e130, 1e131, 1e132, 1e133, 1e134, 1e135, 1e136, 1e137, 1e138, 1e139, 1e140, 1e141, 1e142, 1e143,
        1e144, 1e145, 1e146, 1e147, 1e148, 1e149, 1e150, 1e151, 1e152, 1e153, 1e154, 1e155, 1e156, 1e157, 1e158, 1e159,
        1e160, 1e161, 1e162, 1e163, 1e164, 1e165, 1e166, 1e167, 1e168, 1e169, 1e170, 1e171, 1e172, 1e173, 1e174, 1e175,
        1e176, 1e177, 1e178, 1e179, 1e180, 1e181, 1e182, 1e183, 1e184, 1e185, 1e186, 1e187, 1e188, 1e189, 1e190, 1e191,
        1e192, 1e193, 1e194, 1e195, 1e196, 1e197, 1e198, 1e199, 1e200, 1e201, 1e202, 1e203, 1e204, 1e205, 1e206, 1e207,
        1e208, 1e209, 1e210, 1e211, 1e212, 1e213, 1e214, 1e215, 1e216, 1e217, 1e218, 1e219, 1e220, 1e221, 1e222, 1e223,
        1e224, 1e225, 1e226, 1e227, 1e228, 1e229, 1e230, 1e231, 1e232, 1e233, 1e234, 1e235, 1e236, 1e237, 1e238, 1e239,
        1e240, 1e241, 1e242, 1e243, 1e244, 1e245, 1e246, 1e247, 1e248, 1e249, 1e250, 1e251, 1e252, 1e253, 1e254, 1e255,
        1e256, 1e257, 1e258, 1e259, 1e260, 1e261, 1e262, 1e263, 1e264, 1e265, 1e266, 1e267, 1e268, 1e269, 1e270, 1e271,
        1e272, 1e273, 1e274, 1e275, 1e276, 1e277, 1e278, 1e279, 1e280, 1e281, 1e282, 1e283, 1e284, 1e285, 1e286, 1e287,
        1e288, 1e289, 1e290, 1e291, 1e292, 1e293, 1e294, 1e295, 1e296, 1e297, 1e298, 1e299, 1e300, 1e301, 1e302, 1e303,
        1e304, 1e305, 1e306, 1e307, 1e308};
    const char* p = s;
    while (p < e && csv_isspace(*p)) p++;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    // inf / infinity / nan spellings (pandas: to_double's special-casing, case-insensitive)
    if (p < e && !csv_isdigit(*p) && *p != '.') {
        const size_t n = (size_t)(e - p);
        auto ieq = [&](const char* w) {
            const size_t l = strlen(w);
            if (n < l) return false;
            for (size_t k = 0; k < l; k++)
                if ((p[k] | 0x20) != w[k]) return false;
            for (size_t k = l; k < n; k++)
                if (!csv_isspace(p[k])) return false;
            return true;
        };
        if (ieq("infinity") || ieq("inf")) { *out = neg ? -HUGE_VAL : HUGE_VAL; return true; }
        return false;
    }
    // pandas accumulates `number = number * 10. + digit` in a double for at most 17 digits; the first
    // 15 digits stay below 2^53, where that recurrence is exact, so they are gathered in an integer.
    const int max_digits = 17;
    double number = 0.0;
    uint64_t acc = 0;
    int exponent = 0, num_digits = 0, num_decimals = 0;
    auto push = [&](int d) {
        if (num_digits < 15) acc = acc * 10 + (unsigned)d;
        else {
            if (num_digits == 15) number = (double)acc;
            number = number * 10.0 + d;
        }
        num_digits++;
    };
    while (p < e && csv_isdigit(*p)) {
        if (num_digits < max_digits) push(*p - '0');
        else ++exponent;
        p++;
    }
    if (p < e && *p == '.') {
        p++;
        while (num_digits < max_digits && p < e && csv_isdigit(*p)) {
            push(*p - '0');
            p++;
            num_decimals++;
        }
        if (num_digits >= max_digits)
            while (p < e && csv_isdigit(*p)) ++p;
        exponent -= num_decimals;
    }
    if (num_digits <= 15) number = (double)acc;
    if (num_digits == 0) return false;
    if (neg) number = -number;
    if (p < e && (*p == 'e' || *p == 'E')) {
        const char* save = p;
        p++;
        bool eneg = false;
        if (p < e && (*p == '-' || *p == '+')) { eneg = *p == '-'; p++; }
        int n = 0, nd = 0;
        while (p < e && csv_isdigit(*p)) {
            if (n < 100000) n = n * 10 + (*p - '0');
            nd++;
            p++;
        }
        if (nd == 0) p = save;
        else exponent += eneg ? -n : n;
    }
    if (exponent > 308) {
        return false;  // pandas: ERANGE -> the column falls back to object; not a rods_df case
    } else if (exponent > 0) {
        number *= p10[exponent];
    } else if (exponent < -308) {
        if (exponent < -616) number = 0.0;
        else { number /= p10[-308 - exponent]; number /= p10[308]; }
    } else {
        number /= p10[-exponent];
    }
    if (number == HUGE_VAL || number == -HUGE_VAL) return false;
    while (p < e && csv_isspace(*p)) p++;
    if (p != e) return false;
    *out = number;
    return true;
}

inline bool csv_parse_int(const char* s, const char* e, int64_t* out)
{
    const char* p = s;
    while (p < e && csv_isspace(*p)) p++;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    if (p >= e || !csv_isdigit(*p)) return false;
    uint64_t v = 0;
    while (p < e && csv_isdigit(*p)) {
        const unsigned d = (unsigned)(*p - '0');
        if (v > (UINT64_MAX - d) / 10) return false;
        v = v * 10 + d;
        p++;
    }
    while (p < e && csv_isspace(*p)) p++;
    if (p != e) return false;
    if (neg) {
        if (v > (uint64_t)INT64_MAX + 1) return false;
        *out = (int64_t)(0 - v);
    } else {
        if (v > (uint64_t)INT64_MAX) return false;
        *out = (int64_t)v;
    }
    return true;
}

// pandas' default na_values (io/parsers: STR_NA_VALUES)
inline bool csv_is_na(const char* s, const char* e)
{
    const size_t n = (size_t)(e - s);
    if (n == 0) return true;
    if (n > 9 || n < 2) return false;
    switch (*s) {
    case '#': case '-': case '1': case '<': case 'N': case 'n': break;
    default: return false;
    }
    static const char* const na[] = {"#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                                     "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    for (const char* w : na)
        if (strlen(w) == n && memcmp(w, s, n) == 0) return true;
    return false;
}

inline int csv_parse_bool(const char* s, const char* e)  // 1 true, 0 false, -1 neither
{
    const size_t n = (size_t)(e - s);
    if (n == 4 && (!memcmp(s, "True", 4) || !memcmp(s, "TRUE", 4) || !memcmp(s, "true", 4))) return 1;
    if (n == 5 && (!memcmp(s, "False", 5) || !memcmp(s, "FALSE", 5) || !memcmp(s, "false", 5))) return 0;
    return -1;
}

struct CsvField {
    const char* s;
    const char* e;
    bool quoted;  // contains "" escapes or was quoted: s..e is the inside of the quotes
};

// Splits one record [p, end) (no trailing newline) into fields.  Returns the field count.
template <typename F>
inline int csv_split(const char* p, const char* end, F&& emit)
{
    int k = 0;
    while (true) {
        CsvField f;
        if (p < end && *p == '"') {
            const char* q = p + 1;
            f.s = q;
            while (q < end) {
                if (*q == '"') {
                    if (q + 1 < end && q[1] == '"') { q += 2; continue; }
                    break;
                }
                q++;
            }
            f.e = q;
            f.quoted = true;
            p = q < end ? q + 1 : end;
            while (p < end && *p != ',') p++;  // junk after the closing quote is dropped
        } else {
            const char* q = (const char*)memchr(p, ',', (size_t)(end - p));
            if (!q) q = end;
            f.s = p;
            f.e = q;
            f.quoted = false;
            p = q;
        }
        emit(k++, f);
        if (p >= end) break;
        p++;  // the comma
        if (p == end) {  // trailing comma: one more empty field
            CsvField z{end, end, false};
            emit(k++, z);
            break;
        }
    }
    return k;
}

inline std::string csv_unquote(const CsvField& f)
{
    std::string out;
    out.reserve((size_t)(f.e - f.s));
    for (const char* q = f.s; q < f.e; q++) {
        out.push_back(*q);
        if (f.quoted && *q == '"' && q + 1 < f.e && q[1] == '"') q++;
    }
    return out;
}

struct CsvColumn {
    std::string name;
    int kind = CSV_INT64;
    std::vector<int64_t> i64;      // CSV_INT64 / CSV_BOOL (0/1)
    std::vector<double> f64;       // CSV_FLOAT64
    std::vector<int32_t> codes;    // CSV_STRING: index into `dict`, -1 = missing
    std::vector<std::string> dict;
};

struct CsvTable {
    long long n_rows = 0;
    std::vector<CsvColumn> cols;
    std::string error;
};


inline int csv_read_file(const char* path, int n_threads, CsvTable* t)
{
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { t->error = "cannot open file"; return -1; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); t->error = "cannot stat file"; return -1; }
    const size_t size = (size_t)sb.st_size;
    if (size == 0) { close(fd); t->error = "empty file"; return -2; }
    const char* data = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) { t->error = "cannot map file"; return -1; }
    const char* const end = data + size;

    // ---- record boundaries (quote-aware only when the file contains a quote character)
    const bool has_quote = memchr(data, '"', size) != nullptr;
    std::vector<const char*> starts, stops;
    starts.reserve(size / 64 + 16);
    stops.reserve(size / 64 + 16);
    {
        const char* p = data;
        while (p < end) {
            const char* nl;
            if (!has_quote) {
                nl = (const char*)memchr(p, '\n', (size_t)(end - p));
                if (!nl) nl = end;
            } else {
                bool inq = false;
                nl = p;
                while (nl < end && (inq || *nl != '\n')) {
                    if (*nl == '"') inq = !inq;
                    nl++;
                }
            }
            const char* stop = nl;
            if (stop > p && stop[-1] == '\r') stop--;
            if (stop > p) {  // skip_blank_lines=True
                starts.push_back(p);
                stops.push_back(stop);
            }
            p = nl + 1;
        }
    }
    if (starts.empty()) { munmap((void*)data, size); t->error = "no header"; return -2; }

    // ---- header
    std::vector<std::string> names;
    csv_split(starts[0], stops[0], [&](int, const CsvField& f) { names.push_back(csv_unquote(f)); });
    const int nc = (int)names.size();
    const long long nr = (long long)starts.size() - 1;
    t->n_rows = nr;
    t->cols.resize(nc);
    for (int c = 0; c < nc; c++) t->cols[c].name = names[c];

    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 64));
    if (nr < 4096) T = 1;
    T = (int)std::min<long long>(T, std::max<long long>(1, nr / 1024));
    auto range = [&](int w, long long* lo, long long* hi) {
        *lo = nr * w / T;
        *hi = nr * (w + 1) / T;
    };
    auto run = [&](auto&& fn) {
        if (T == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int w = 0; w < T; w++) th.emplace_back(fn, w);
        for (auto& x : th) x.join();
    };

    // ---- column kinds: guessed from the first record, then every record is parsed under the guess;
    // a column that meets a field outside its kind is widened (int -> float -> string) and parsed
    // again.  rods_df files settle in one pass.
    std::vector<int> kind(nc, CSV_FLOAT64);
    std::vector<char> dirty(nc, 1);
    if (nr > 0)
        csv_split(starts[1], stops[1], [&](int c, const CsvField& f) {
            if (c >= nc) return;
            int64_t iv;
            double dv;
            if (f.s == f.e) kind[c] = CSV_FLOAT64;
            else if (csv_parse_int(f.s, f.e, &iv)) kind[c] = CSV_INT64;
            else if (csv_parse_double(f.s, f.e, &dv) || csv_is_na(f.s, f.e)) kind[c] = CSV_FLOAT64;
            else kind[c] = CSV_STRING;
        });
    struct Dict {
        std::vector<std::string> v;
        std::unordered_map<std::string, int> m;
        int last = -1;
        int code(const char* s, size_t len)
        {
            if (last >= 0 && v[last].size() == len && memcmp(v[last].data(), s, len) == 0) return last;
            std::string key(s, len);
            auto it = m.find(key);
            if (it == m.end()) {
                v.push_back(key);
                it = m.emplace(std::move(key), (int)v.size() - 1).first;
            }
            return last = it->second;
        }
    };
    struct Flags {
        char need_float = 0, need_string = 0, any_na = 0;
    };
    std::vector<std::vector<Dict>> tdict(T, std::vector<Dict>(nc));
    std::vector<char> any_na(nc, 0);
    std::atomic<long long> bad_row{-1};
    const double nan = std::nan("");
    for (int pass = 0; pass < 3 && nr > 0; pass++) {
        for (int c = 0; c < nc; c++) {
            if (!dirty[c]) continue;
            CsvColumn& col = t->cols[c];
            col.i64.clear(); col.f64.clear(); col.codes.clear();
            col.i64.shrink_to_fit(); col.f64.shrink_to_fit();
            if (kind[c] == CSV_INT64) col.i64.resize((size_t)nr);
            else if (kind[c] == CSV_FLOAT64) col.f64.resize((size_t)nr);
            else col.codes.resize((size_t)nr);
        }
        std::vector<Flags> flags((size_t)T * nc);
        run([&](int w) {
            long long lo, hi;
            range(w, &lo, &hi);
            Flags* fl = &flags[(size_t)w * nc];
            auto& dicts = tdict[w];
            for (long long r = lo; r < hi; r++) {
                const int k = csv_split(starts[r + 1], stops[r + 1], [&](int c, const CsvField& f) {
                    if (c >= nc || !dirty[c]) return;
                    CsvColumn& col = t->cols[c];
                    switch (kind[c]) {
                    case CSV_INT64: {
                        if (csv_parse_int(f.s, f.e, &col.i64[r])) break;
                        double dv;
                        if (csv_is_na(f.s, f.e) || csv_parse_double(f.s, f.e, &dv)) fl[c].need_float = 1;
                        else fl[c].need_string = 1;
                        break;
                    }
                    case CSV_FLOAT64: {
                        double v = nan;
                        if (f.s != f.e && !csv_parse_double(f.s, f.e, &v)) {
                            v = nan;
                            if (!csv_is_na(f.s, f.e)) fl[c].need_string = 1;
                        }
                        col.f64[r] = v;
                        break;
                    }
                    default: {
                        if (csv_is_na(f.s, f.e)) { col.codes[r] = -1; fl[c].any_na = 1; break; }
                        if (!f.quoted) col.codes[r] = dicts[c].code(f.s, (size_t)(f.e - f.s));
                        else {
                            const std::string u = csv_unquote(f);
                            col.codes[r] = dicts[c].code(u.data(), u.size());
                        }  // thread-local code, remapped below
                    }
                    }
                });
                if (k > nc) { long long exp = -1; bad_row.compare_exchange_strong(exp, r); }
                for (int c = k; c < nc; c++) {  // short record: missing fields are NA
                    if (!dirty[c]) continue;
                    CsvColumn& col = t->cols[c];
                    if (kind[c] == CSV_INT64) fl[c].need_float = 1;
                    else if (kind[c] == CSV_FLOAT64) col.f64[r] = nan;
                    else { col.codes[r] = -1; fl[c].any_na = 1; }
                }
            }
        });
        if (bad_row.load() >= 0) {
            munmap((void*)data, size);
            t->error = "Error tokenizing data: too many fields in data row " + std::to_string(bad_row.load());
            return -3;
        }
        bool again = false;
        for (int c = 0; c < nc; c++) {
            if (!dirty[c]) continue;
            bool nf = false, ns = false;
            for (int w = 0; w < T; w++) {
                nf |= flags[(size_t)w * nc + c].need_float != 0;
                ns |= flags[(size_t)w * nc + c].need_string != 0;
                any_na[c] |= flags[(size_t)w * nc + c].any_na;
            }
            if (ns) kind[c] = CSV_STRING;
            else if (nf) kind[c] = CSV_FLOAT64;
            else dirty[c] = 0;
            again |= dirty[c] != 0;
        }
        if (!again) break;
    }
    for (int c = 0; c < nc; c++) t->cols[c].kind = nr > 0 ? kind[c] : CSV_STRING;  // pandas: empty columns are object
    std::vector<std::unordered_map<std::string, int>> gmap(nc);
    for (int c = 0; c < nc; c++) {
        CsvColumn& col = t->cols[c];
        if (col.kind != CSV_STRING) continue;
        for (int w = 0; w < T; w++) {
            long long lo, hi;
            range(w, &lo, &hi);
            const auto& d = tdict[w][c].v;
            std::vector<int32_t> remap(d.size());
            for (size_t q = 0; q < d.size(); q++) {
                auto it = gmap[c].find(d[q]);
                if (it == gmap[c].end()) {
                    col.dict.push_back(d[q]);
                    it = gmap[c].emplace(d[q], (int)col.dict.size() - 1).first;
                }
                remap[q] = it->second;
            }
            for (long long r = lo; r < hi; r++)
                if (col.codes[r] >= 0) col.codes[r] = remap[col.codes[r]];
        }
    }
    for (int c = 0; c < nc; c++) {  // True/False columns without NA are bool
        CsvColumn& col = t->cols[c];
        if (col.kind != CSV_STRING || any_na[c] || col.dict.empty()) continue;
        std::vector<int> tf(col.dict.size());
        bool all = true;
        for (size_t q = 0; q < col.dict.size(); q++) {
            tf[q] = csv_parse_bool(col.dict[q].data(), col.dict[q].data() + col.dict[q].size());
            all &= tf[q] >= 0;
        }
        if (!all) continue;
        col.kind = CSV_BOOL;
        col.i64.resize((size_t)nr);
        for (long long r = 0; r < nr; r++) col.i64[r] = tf[col.codes[r]];
        col.codes.clear();
        col.dict.clear();
    }
    munmap((void*)data, size);
    return 0;
}

}  // namespace ptk
