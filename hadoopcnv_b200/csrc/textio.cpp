// textio.cpp -- B3: config.txt and the text layouts between the reference's jobs (host code, no CUDA).
//
//   UserConfig.init / getters   UserConfig.java:25-141  (java.util.Properties syntax)
//   bins part files             Reducers.java:191-203   chr \t bin \t start \t end \t median \t mse0..5 \t total \t n
//   results.txt                 Hmm.java:479-490 + Reducers.java:238-241
//   Float.toString / Double.toString number formatting (shortest digits that round-trip, Java notation)
#include "../../include/hcnv.h"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace {

// ---- shortest round-trip digits ---------------------------------------------------------------
// Returns digits (no leading/trailing zeros unless the value needs them) and the decimal exponent of the
// first digit.  Tries increasing precision; at each precision also the neighbours of the nearest decimal,
// because the rounding interval is asymmetric at powers of two.
template <class T> T parse_as(const char* s);
template <> float parse_as<float>(const char* s) { return strtof(s, nullptr); }
template <> double parse_as<double>(const char* s) { return strtod(s, nullptr); }

template <class T>
void shortest_digits(T v, char* digits, int* nd_out, int* e10_out) {
    const int maxp = sizeof(T) == 4 ? 9 : 17;
    char tmp[48];
    for (int p = 1; p <= maxp; ++p) {
        snprintf(tmp, sizeof tmp, "%.*e", p - 1, (double)v);
        unsigned long long n = 0; int nd = 0; const char* c = tmp;
        for (; *c && *c != 'e'; ++c) if (*c != '.') { n = n * 10 + (unsigned)(*c - '0'); ++nd; }
        int e10 = atoi(c + 1);
        unsigned long long lim = 1; for (int i = 0; i < nd; ++i) lim *= 10;
        const long long deltas[3] = {0, -1, 1};
        for (int di = 0; di < 3; ++di) {
            unsigned long long m = n; int e = e10;
            if (deltas[di] < 0) { if (m == lim / 10) continue; m -= 1; }
            else if (deltas[di] > 0) { m += 1; if (m >= lim) { m /= 10; e += 1; } }
            char cand[48], dg[24];
            snprintf(dg, sizeof dg, "%0*llu", nd, m);
            snprintf(cand, sizeof cand, "%c.%se%d", dg[0], dg[1] ? dg + 1 : "0", e);
            if (parse_as<T>(cand) == v) {
                if (p == 1) {
                    // Double.toString (JDK >= 19 wording): when the shortest decimal that rounds to v has ONE digit, the
                    // answer is the closest decimal of length TWO (4.9E-324 for Double.MIN_VALUE, 1.4E-45 for Float.MIN_VALUE;
                    // it differs from d.0 only for subnormals).  The nearest two-digit decimal lies between v and the
                    // one-digit one, so it rounds to v as well.
                    snprintf(tmp, sizeof tmp, "%.1e", (double)v);
                    dg[0] = tmp[0]; dg[1] = tmp[2]; dg[2] = 0;
                    e = atoi(strchr(tmp, 'e') + 1);
                    int k2 = dg[1] == '0' ? 1 : 2;
                    memcpy(digits, dg, (size_t)k2); digits[k2] = 0;
                    *nd_out = k2; *e10_out = e;
                    return;
                }
                int k = nd;
                while (k > 1 && dg[k - 1] == '0') --k;
                memcpy(digits, dg, (size_t)k); digits[k] = 0;
                *nd_out = k; *e10_out = e;
                return;
            }
        }
    }
    digits[0] = '0'; digits[1] = 0; *nd_out = 1; *e10_out = 0;   // not reached for finite v
}

template <class T> int java_to_string_slow(T v, char* buf);

// Float.toString / Double.toString.  The columns of a bins or results file repeat a few values millions of times (scaled depths
// are (integer - mean) / mean, MSEs are mostly 0.0, totals are integers), and the digit search below costs microseconds: integers
// print directly and everything else goes through a small direct-mapped cache of finished strings.
template <class T>
int java_to_string(T v, char* buf) {
    if (std::isnan(v)) { strcpy(buf, "NaN"); return 3; }
    if (std::isinf(v)) { strcpy(buf, v < 0 ? "-Infinity" : "Infinity"); return (int)strlen(buf); }
    if (v == 0) { strcpy(buf, std::signbit(v) ? "-0.0" : "0.0"); return (int)strlen(buf); }
    const double dv = (double)v;
    if (std::fabs(dv) < 1e7 && dv == (double)(long long)dv) return sprintf(buf, "%lld.0", (long long)dv);       // (1e-3 <= |v| < 1e7: plain decimal notation)
    struct Entry { unsigned long long bits; int len; char text[28]; };
    static thread_local Entry cache[4096];
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    bits |= (unsigned long long)sizeof(T) << 60;                      // (float and double share the cache; a double's top bits are its exponent < 0x7ff: no clash with the tag of 8)
    Entry& e = cache[(bits * 0x9E3779B97F4A7C15ull) >> 52];
    if (e.bits == bits && e.len > 0) { memcpy(buf, e.text, (size_t)e.len + 1); return e.len; }
    const int n = java_to_string_slow<T>(v, buf);
    if (n < (int)sizeof e.text) { e.bits = bits; e.len = n; memcpy(e.text, buf, (size_t)n + 1); }
    return n;
}

template <class T>
int java_to_string_slow(T v, char* buf) {
    char dg[24]; int nd, e10;
    T av = v < 0 ? -v : v;
    shortest_digits<T>(av, dg, &nd, &e10);
    char* o = buf;
    if (v < 0) *o++ = '-';
    if ((double)av >= 1e-3 && (double)av < 1e7) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) *o++ = i < nd ? dg[i] : '0';
            *o++ = '.';
            if (nd > e10 + 1) { for (int i = e10 + 1; i < nd; ++i) *o++ = dg[i]; } else *o++ = '0';
        } else {
            *o++ = '0'; *o++ = '.';
            for (int i = 0; i < -e10 - 1; ++i) *o++ = '0';
            for (int i = 0; i < nd; ++i) *o++ = dg[i];
        }
    } else {
        *o++ = dg[0]; *o++ = '.';
        if (nd > 1) { for (int i = 1; i < nd; ++i) *o++ = dg[i]; } else *o++ = '0';
        *o++ = 'E';
        o += sprintf(o, "%d", e10);
    }
    *o = 0;
    return (int)(o - buf);
}

// ---- java.util.Properties.load (the subset config.txt uses: comments, '=' ':' or blank separators,
// backslash escapes and line continuations) ------------------------------------------------------
bool load_properties(const char* path, std::map<std::string, std::string>& kv) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    std::string text; char b[4096]; size_t n;
    while ((n = fread(b, 1, sizeof b, f)) > 0) text.append(b, n);
    fclose(f);
    size_t i = 0, N = text.size();
    auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\f'; };
    while (i < N) {
        // logical line
        std::string line;
        bool first_phys = true;
        while (i < N) {
            size_t e = i;
            while (e < N && text[e] != '\n' && text[e] != '\r') ++e;
            std::string phys = text.substr(i, e - i);
            i = e;
            if (i < N) { if (text[i] == '\r' && i + 1 < N && text[i + 1] == '\n') i += 2; else ++i; }
            size_t s = 0;
            while (s < phys.size() && is_ws(phys[s])) ++s;
            phys = phys.substr(s);
            if (first_phys && (phys.empty() || phys[0] == '#' || phys[0] == '!')) { line.clear(); break; }
            first_phys = false;
            size_t bs = 0;
            for (size_t k = phys.size(); k > 0 && phys[k - 1] == '\\'; --k) ++bs;
            if (bs % 2 == 1) { line += phys.substr(0, phys.size() - 1); continue; }
            line += phys;
            break;
        }
        if (line.empty()) continue;
        std::string key, val;
        size_t p = 0;
        auto unescape = [&](std::string& dst, bool is_key) {
            while (p < line.size()) {
                char c = line[p];
                if (is_key && (c == '=' || c == ':' || is_ws(c))) break;
                if (c == '\\' && p + 1 < line.size()) {
                    char d = line[++p];
                    if (d == 't') dst += '\t'; else if (d == 'n') dst += '\n'; else if (d == 'r') dst += '\r'; else if (d == 'f') dst += '\f';
                    else dst += d;
                    ++p;
                } else { dst += c; ++p; }
            }
        };
        unescape(key, true);
        while (p < line.size() && is_ws(line[p])) ++p;
        if (p < line.size() && (line[p] == '=' || line[p] == ':')) { ++p; while (p < line.size() && is_ws(line[p])) ++p; }
        unescape(val, false);
        kv[key] = val;
    }
    return true;
}

std::string jtrim(const std::string& s) {   // String.trim(): strips chars <= ' '
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') ++a;
    while (b > a && (unsigned char)s[b - 1] <= ' ') --b;
    return s.substr(a, b - a);
}
bool parse_bool(const std::map<std::string, std::string>& kv, const char* k) {   // Boolean.parseBoolean (no trim)
    auto it = kv.find(k);
    if (it == kv.end()) return false;
    const std::string& s = it->second;
    return s.size() == 4 && tolower(s[0]) == 't' && tolower(s[1]) == 'r' && tolower(s[2]) == 'u' && tolower(s[3]) == 'e';
}
bool parse_int(const std::string& s, int32_t* out) {   // Integer.parseInt
    if (s.empty()) return false;
    size_t p = 0; if (s[0] == '-' || s[0] == '+') p = 1;
    if (p == s.size()) return false;
    long long v = 0;
    for (; p < s.size(); ++p) { if (!isdigit((unsigned char)s[p])) return false; v = v * 10 + (s[p] - '0'); if (v > 2147483648LL) return false; }
    if (s[0] == '-') v = -v;
    if (v > 2147483647LL || v < -2147483648LL) return false;
    *out = (int32_t)v; return true;
}
bool parse_float(const std::string& s0, float* out) {   // Float.parseFloat(trimmed): optional f/F/d/D suffix
    std::string s = jtrim(s0);
    if (s.empty()) return false;
    char last = s[s.size() - 1];
    if (last == 'f' || last == 'F' || last == 'd' || last == 'D') s.erase(s.size() - 1);
    if (s.empty()) return false;
    if (s == "NaN" || s == "+NaN" || s == "-NaN") { *out = NAN; return true; }
    if (s == "Infinity" || s == "+Infinity") { *out = INFINITY; return true; }
    if (s == "-Infinity") { *out = -INFINITY; return true; }
    for (char c : s) if (!(isdigit((unsigned char)c) || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-')) return false;
    char* end = nullptr;
    float v = strtof(s.c_str(), &end);
    if (!end || *end) return false;
    *out = v; return true;
}

}  // namespace

extern "C" {

int hcnv_format_float(float v, char* buf) { return buf ? java_to_string<float>(v, buf) : HCNV_E_INVALID; }
int hcnv_format_double(double v, char* buf) { return buf ? java_to_string<double>(v, buf) : HCNV_E_INVALID; }

int hcnv_config_load(const char* path, hcnv_config* out) {
    if (!path || !out) return HCNV_E_INVALID;
    std::map<std::string, std::string> kv;
    if (!load_properties(path, kv)) return HCNV_E_IO;
    memset(out, 0, sizeof *out);
    auto get = [&](const char* k) -> const std::string* { auto it = kv.find(k); return it == kv.end() ? nullptr : &it->second; };
    if (const std::string* s = get("BAM_FILE")) snprintf(out->bam_file, sizeof out->bam_file, "%s", s->c_str());
    if (const std::string* s = get("VCF_FILE")) snprintf(out->vcf_file, sizeof out->vcf_file, "%s", s->c_str());
    out->run_depth_caller = parse_bool(kv, "RUN_DEPTH_CALLER");
    out->run_global_sort = parse_bool(kv, "RUN_GLOBAL_SORT");
    out->init_vcf_lookup = parse_bool(kv, "INIT_VCF_LOOKUP");
    out->run_binner = parse_bool(kv, "RUN_BINNER");
    out->run_cnv_caller = parse_bool(kv, "RUN_CNV_CALLER");
    out->run_sr_pe_extracter = parse_bool(kv, "RUN_SR_PE_EXTRACTER");
    // PennCnvSeq.run reads all of these unconditionally (PennCnvSeq.java:106-118): a missing or malformed one throws
    const std::string* s;
    if (!(s = get("ABBERATION_PENALTY")) || !parse_float(*s, &out->abberation_penalty)) return HCNV_E_PARSE;
    if (!(s = get("TRANSITION_PENALTY")) || !parse_float(*s, &out->transition_penalty)) return HCNV_E_PARSE;
    if (!(s = get("BAM_FILE_REDUCERS")) || !parse_int(*s, &out->bam_file_reducers)) return HCNV_E_PARSE;     // no trim (UserConfig.java:92)
    if (!(s = get("TEXT_FILE_REDUCERS")) || !parse_int(*s, &out->text_file_reducers)) return HCNV_E_PARSE;
    if (!(s = get("REDUCER_TASKS")) || !parse_int(jtrim(*s), &out->reducer_tasks)) return HCNV_E_PARSE;      // trimmed (UserConfig.java:127)
    return HCNV_OK;
}

int hcnv_write_bins_text(const char* path, int append, const char* chr, int64_t n,
                         const int32_t* bin, const int32_t* start, const int32_t* end, const float* median,
                         const double* mse, const double* total, const int32_t* cnt) {
    if (!path || !chr || n < 0) return HCNV_E_INVALID;
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) return HCNV_E_IO;
    char b[64];
    for (int64_t i = 0; i < n; ++i) {
        fprintf(f, "%s\t%d\t%d\t%d\t", chr, bin[i], start[i], end[i]);
        java_to_string<double>((double)median[i], b); fputs(b, f);         // Double.toString(medianDepth), Reducers.java:202
        for (int s = 0; s < 6; ++s) { java_to_string<double>(mse[i * 6 + s], b); fputc('\t', f); fputs(b, f); }
        java_to_string<double>(total[i], b);
        fprintf(f, "\t%s\t%d\n", b, cnt[i]);
    }
    return fclose(f) == 0 ? HCNV_OK : HCNV_E_IO;
}

int hcnv_write_results_text(const char* path, int append, const char* chr, int64_t n,
                            const int32_t* start, const int32_t* end, const float* scaled,
                            const int32_t* state, const int32_t* cn, const float* mse) {
    if (!path || !chr || n < 0) return HCNV_E_INVALID;
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) return HCNV_E_IO;
    char b[64];
    for (int64_t i = 0; i < n; ++i) {
        java_to_string<float>(scaled[i], b);
        static const int cn_of[6] = {0, 1, 2, 2, 3, 4};                      // Hmm.java:35; cn == NULL: the copy number is this look-up of the state
        fprintf(f, "%s\t%d\t%d\t%s\t%d\t%d", chr, start[i], end[i], b, state[i], cn ? cn[i] : cn_of[state[i] >= 0 && state[i] < 6 ? state[i] : 0]);
        for (int s = 0; s < 6; ++s) { java_to_string<float>(mse[i * 6 + s], b); fputc('\t', f); fputs(b, f); }
        fputc('\n', f);
    }
    return fclose(f) == 0 ? HCNV_OK : HCNV_E_IO;
}

// ---- bins part files back in: what BinSortMapper + the (chr, bin) secondary sort + Hmm.init do to them ----------
}  // extern "C"

struct hcnv_bins_text {
    struct Chrom {
        std::string name;
        std::vector<int32_t> bin, start, end, n;
        std::vector<float> median, mse, total;
    };
    std::vector<Chrom> chroms;
    std::map<std::string, int> index;
};

extern "C" {

int hcnv_read_bins_text(const char* path, hcnv_bins_text** out) {
    if (!path || !out) return HCNV_E_INVALID;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return HCNV_E_IO;
    hcnv_bins_text* t = new hcnv_bins_text();
    std::string line; int ch; int rc = HCNV_OK;
    auto handle = [&](const std::string& ln) -> int {
        if (ln.empty()) return HCNV_OK;
        std::vector<std::string> parts; size_t s = 0;
        for (;;) { size_t e = ln.find('\t', s); parts.push_back(ln.substr(s, e == std::string::npos ? e : e - s)); if (e == std::string::npos) break; s = e + 1; }
        if (parts.size() < 13) return HCNV_E_PARSE;            // BinSortMapper reads parts[0..1], Hmm.init parts[0..10] of the value
        int32_t bin, st, en, cnt;
        float med, tot, m[6];
        // Integer.parseInt for bin (Mappers.java:75), start, end, n (Hmm.java:360-373); Float.valueOf for the rest
        if (!parse_int(parts[1], &bin) || !parse_int(parts[2], &st) || !parse_int(parts[3], &en) || !parse_int(parts[12], &cnt)) return HCNV_E_PARSE;
        if (!parse_float(parts[4], &med) || !parse_float(parts[11], &tot)) return HCNV_E_PARSE;
        for (int j = 0; j < 6; ++j) if (!parse_float(parts[5 + j], &m[j])) return HCNV_E_PARSE;
        auto it = t->index.find(parts[0]);
        int ci;
        if (it == t->index.end()) { ci = (int)t->chroms.size(); t->index[parts[0]] = ci; t->chroms.emplace_back(); t->chroms.back().name = parts[0]; }
        else ci = it->second;
        hcnv_bins_text::Chrom& c = t->chroms[(size_t)ci];
        c.bin.push_back(bin); c.start.push_back(st); c.end.push_back(en); c.n.push_back(cnt);
        c.median.push_back(med); c.total.push_back(tot);
        for (int j = 0; j < 6; ++j) c.mse.push_back(m[j]);
        return HCNV_OK;
    };
    while ((ch = fgetc(f)) != EOF) {
        if (ch == '\n') { if (!line.empty() && line.back() == '\r') line.pop_back(); rc = handle(line); line.clear(); if (rc) break; }
        else line.push_back((char)ch);
    }
    if (!rc && !line.empty()) rc = handle(line);
    fclose(f);
    if (rc) { delete t; return rc; }
    // the reducer sees a chromosome's bins in ascending bin order (RefBinKey.compareTo, Keys.java:226-238); part files
    // written by this library already are, files merged from several reducers may not be
    for (auto& c : t->chroms) {
        const size_t n = c.bin.size();
        bool sorted = true;
        for (size_t i = 1; i < n; ++i) if (c.bin[i] < c.bin[i - 1]) { sorted = false; break; }
        if (sorted) continue;
        std::vector<size_t> o(n);
        for (size_t i = 0; i < n; ++i) o[i] = i;
        std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return c.bin[a] < c.bin[b]; });
        hcnv_bins_text::Chrom d; d.name = c.name;
        for (size_t i : o) {
            d.bin.push_back(c.bin[i]); d.start.push_back(c.start[i]); d.end.push_back(c.end[i]); d.n.push_back(c.n[i]);
            d.median.push_back(c.median[i]); d.total.push_back(c.total[i]);
            for (int j = 0; j < 6; ++j) d.mse.push_back(c.mse[6 * i + j]);
        }
        c = d;
    }
    *out = t;
    return HCNV_OK;
}

int32_t hcnv_bins_text_n_chromosomes(const hcnv_bins_text* t) { return t ? (int32_t)t->chroms.size() : 0; }

int hcnv_bins_text_get(const hcnv_bins_text* t, int32_t i, const char** chrom, int64_t* n, const int32_t** bin, const int32_t** start,
                       const int32_t** end, const float** median, const float** mse, const float** total, const int32_t** cnt) {
    if (!t || i < 0 || i >= (int32_t)t->chroms.size()) return HCNV_E_INVALID;
    const hcnv_bins_text::Chrom& c = t->chroms[(size_t)i];
    if (chrom) *chrom = c.name.c_str();
    if (n) *n = (int64_t)c.bin.size();
    if (bin) *bin = c.bin.data();
    if (start) *start = c.start.data();
    if (end) *end = c.end.data();
    if (median) *median = c.median.data();
    if (mse) *mse = c.mse.data();
    if (total) *total = c.total.data();
    if (cnt) *cnt = c.n.data();
    return HCNV_OK;
}

void hcnv_bins_text_free(hcnv_bins_text* t) { delete t; }

}  // extern "C"

// the start / end columns from their implicit form (hcnv_results_download): runs of consecutive full bins + the rows that are not full
extern "C" int hcnv_expand_start_end(int64_t n, int32_t W, int64_t n_runs, const int32_t* run_rows, const int32_t* run_bins,
                                     int64_t n_se, const int32_t* se_rows, const int32_t* se_vals, int32_t* start, int32_t* end) {
    if (n < 0 || W < 1 || n_runs < 0 || n_se < 0 || (n > 0 && (!start || !end)) || (n_runs > 0 && (!run_rows || !run_bins)) || (n_se > 0 && (!se_rows || !se_vals)))
        return HCNV_E_INVALID;
    if (n == 0) return HCNV_OK;
    std::vector<std::pair<int32_t, int32_t>> runs((size_t)n_runs);
    for (int64_t k = 0; k < n_runs; ++k) {
        if (run_rows[k] < 0 || run_rows[k] >= n) return HCNV_E_RANGE;
        runs[(size_t)k] = { run_rows[k], run_bins[k] };
    }
    std::sort(runs.begin(), runs.end());
    if (runs.empty() || runs[0].first != 0) return HCNV_E_RANGE;
    for (size_t k = 0; k < runs.size(); ++k) {
        const int64_t r0 = runs[k].first, r1 = k + 1 < runs.size() ? runs[k + 1].first : n;
        int64_t id = runs[k].second;
        for (int64_t r = r0; r < r1; ++r, id += W) { start[r] = (int32_t)(id > 1 ? id : 1); end[r] = (int32_t)(id + W - 1); }
    }
    for (int64_t k = 0; k < n_se; ++k) {
        if (se_rows[k] < 0 || se_rows[k] >= n) return HCNV_E_RANGE;
        start[se_rows[k]] = se_vals[2 * k]; end[se_rows[k]] = se_vals[2 * k + 1];
    }
    return HCNV_OK;
}
