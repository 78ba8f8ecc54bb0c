// packer.cpp -- host side of B1 (no CUDA): what stays on the host in the north star.
//
//   hcnv_vcf_load   VcfLookup.parseVcf2Text (VcfLookup.java:135-196) + Utils.getZygosity / getVarType
//                   (Utils.java:15-99): VCF -> per-contig SNP tables (ascending unique 1-based positions, zygosity)
//   hcnv_bam_pack   what SAMRecordWindowMapper.map reads from each SAMRecord (Mappers.java:174-212) with htsjdk
//                   1.118's getAlignmentBlocks rules: BAM/BGZF -> sorted hcnv_block stream, long blocks and
//                   per-(block, SNP) base observations, per contig, ready for hcnv_bin_contigs.
#include "../../include/hcnv.h"
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <string>
#include <vector>

struct hcnv_snp_table {
    std::vector<std::string> names;                       // contig names in first-seen order (with the "chr" prefix rule)
    std::map<std::string, int> index;
    std::vector<std::vector<int32_t>> pos;
    std::vector<std::vector<int8_t>> zyg;
    std::string err;
};

struct hcnv_pack {
    struct Contig {
        std::string name; int32_t length = 0; int32_t max_len = 0;
        std::vector<hcnv_block> blocks; std::vector<hcnv_long_block> longs; std::vector<uint32_t> obs;
        std::vector<hcnv_cblock> cstore; std::vector<int32_t> anchors; hcnv_cblock* cblocks = nullptr; int64_t n_cblocks = -1;   // built on demand
        std::vector<hcnv_cobs> cobs; std::vector<int32_t> cobs_anchors; int64_t n_cobs = -1;                                        // built on demand
        const std::vector<int32_t>* snp_pos = nullptr; const std::vector<int8_t>* snp_zyg = nullptr;
    };
    std::vector<Contig> contigs;
    int64_t n_records = 0, n_unmapped = 0;
    std::string err;
};

namespace {

std::string chr_prefix(const std::string& s) {             // Mappers.java:176-180, VcfLookup.java:161-165
    return s.compare(0, 3, "chr") == 0 ? s : "chr" + s;
}

// Utils.getZygosity (Utils.java:15-84); returns 1 for a genotype the reference would System.exit on
int zygosity(std::string g, int8_t* out) {
    size_t c = g.find(':');
    if (c != std::string::npos) g = g.substr(0, c);
    if (g.size() != 3 || (g[1] != '/' && g[1] != '|')) return 1;
    char a = g[0], b = g[2];
    if ((a != '0' && a != '1' && a != '2') || (b != '0' && b != '1' && b != '2')) return 1;
    if (a == '0' && b == '0') *out = 0;
    else if ((a == '1' && b == '1') || (a == '2' && b == '2')) *out = -2;
    else *out = -1;
    return 0;
}

bool all_acgt_dot(const std::string& s) {
    if (s.empty()) return false;
    for (char ch : s) if (!(ch == 'A' || ch == 'G' || ch == 'C' || ch == 'T' || ch == '.')) return false;
    return true;
}
// Utils.getVarType (Utils.java:87-98): "1" only for multi-base ref/alt made of [AGCT.]; everything else "0"
bool var_type_is_zero(const std::string& ref, const std::string& alt) {
    bool ref_is_n = ref.size() == 1 && (ref[0] == 'N' || ref[0] == 'n');
    if (!ref_is_n && all_acgt_dot(ref) && all_acgt_dot(alt)) return ref.size() == 1 && alt.size() == 1;
    return true;
}

// ---- BGZF: concatenated gzip members -----------------------------------------------------------
bool inflate_all(const std::vector<unsigned char>& in, std::vector<unsigned char>& out, std::string& err) {
    size_t pos = 0;
    out.clear();
    out.reserve(in.size() * 4);
    std::vector<unsigned char> buf(1 << 16);
    while (pos < in.size()) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) { err = "inflateInit2 failed"; return false; }
        zs.next_in = const_cast<unsigned char*>(in.data() + pos);
        zs.avail_in = (uInt)std::min<size_t>(in.size() - pos, 1u << 30);
        int rc;
        do {
            zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); err = "corrupt BGZF block"; return false; }
            out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
        } while (rc != Z_STREAM_END);
        pos = (size_t)(zs.next_in - in.data());
        inflateEnd(&zs);
    }
    return true;
}

bool read_file(const char* path, std::vector<unsigned char>& data) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    unsigned char b[1 << 16]; size_t n;
    while ((n = fread(b, 1, sizeof b, f)) > 0) data.insert(data.end(), b, b + n);
    fclose(f);
    return true;
}

inline int32_t rd32(const unsigned char* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rdu32(const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rdu16(const unsigned char* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct PendingBlock { int32_t start0; uint32_t len_mapq; };
struct ByStart { bool operator()(const PendingBlock& a, const PendingBlock& b) const { return a.start0 > b.start0; } };

}  // namespace

extern "C" {

int hcnv_vcf_load(const char* path, hcnv_snp_table** out) {
    if (!path || !out) return HCNV_E_INVALID;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return HCNV_E_IO;
    hcnv_snp_table* t = new hcnv_snp_table();
    std::vector<std::map<int32_t, int8_t>> acc;            // per contig: position -> zygosity (the last record wins)
    std::string line; int ch; int rc = HCNV_OK;
    auto handle = [&](std::string& ln) -> int {
        if (ln.empty()) return HCNV_E_PARSE;                // line.charAt(0) throws -> "Error in parseVcf2Text" -> exit
        if (ln[0] == '#') return HCNV_OK;
        ln.erase(std::remove(ln.begin(), ln.end(), '\r'), ln.end());
        std::vector<std::string> parts; size_t s = 0;
        for (;;) { size_t e = ln.find('\t', s); parts.push_back(ln.substr(s, e == std::string::npos ? e : e - s)); if (e == std::string::npos) break; s = e + 1; }
        while (!parts.empty() && parts.back().empty()) parts.pop_back();     // String.split drops trailing empties
        if (parts.size() < 5) return HCNV_E_PARSE;          // parts[4] throws
        int8_t z = 0;
        if (parts.size() >= 10 && zygosity(parts[9], &z)) return HCNV_E_PARSE;   // Utils.java:77-81 exits
        if (!var_type_is_zero(parts[3], parts[4])) return HCNV_OK;
        char* endp = nullptr;
        long p = strtol(parts[1].c_str(), &endp, 10);        // BinMapper's Integer.parseInt (Mappers.java:52)
        if (parts[1].empty() || *endp || p < -2147483647L || p > 2147483647L) return HCNV_E_PARSE;
        std::string chr = chr_prefix(parts[0]);
        auto it = t->index.find(chr);
        int ci;
        if (it == t->index.end()) { ci = (int)t->names.size(); t->index[chr] = ci; t->names.push_back(chr); acc.emplace_back(); }
        else ci = it->second;
        acc[ci][(int32_t)p] = z;
        return HCNV_OK;
    };
    bool any = false;
    while ((ch = fgetc(f)) != EOF) {
        any = true;
        if (ch == '\n') { rc = handle(line); line.clear(); any = false; if (rc) break; }
        else line.push_back((char)ch);
    }
    if (!rc && any) rc = handle(line);
    fclose(f);
    if (rc) { delete t; return rc; }
    t->pos.resize(acc.size()); t->zyg.resize(acc.size());
    for (size_t i = 0; i < acc.size(); ++i)
        for (auto& kv : acc[i]) { t->pos[i].push_back(kv.first); t->zyg[i].push_back(kv.second); }
    *out = t;
    return HCNV_OK;
}

int32_t hcnv_snp_table_n_contigs(const hcnv_snp_table* t) { return t ? (int32_t)t->names.size() : 0; }

int hcnv_snp_table_get(const hcnv_snp_table* t, int32_t i, const char** name, const int32_t** pos1, const int8_t** zyg, int64_t* n) {
    if (!t || i < 0 || i >= (int32_t)t->names.size()) return HCNV_E_INVALID;
    if (name) *name = t->names[i].c_str();
    if (pos1) *pos1 = t->pos[i].data();
    if (zyg) *zyg = t->zyg[i].data();
    if (n) *n = (int64_t)t->pos[i].size();
    return HCNV_OK;
}

void hcnv_snp_table_free(hcnv_snp_table* t) { delete t; }

int hcnv_bam_pack(const char* bam_path, const hcnv_snp_table* snps, hcnv_pack** out) {
    if (!bam_path || !out) return HCNV_E_INVALID;
    *out = nullptr;
    std::vector<unsigned char> comp, raw;
    if (!read_file(bam_path, comp)) return HCNV_E_IO;
    std::string err;
    if (!inflate_all(comp, raw, err)) return HCNV_E_PARSE;
    comp.clear(); comp.shrink_to_fit();
    const size_t N = raw.size();
    if (N < 12 || memcmp(raw.data(), "BAM\1", 4) != 0) return HCNV_E_PARSE;
    size_t off = 4;
    int32_t l_text = rd32(&raw[off]); off += 4;
    if (l_text < 0 || off + (size_t)l_text + 4 > N) return HCNV_E_PARSE;
    off += (size_t)l_text;
    int32_t n_ref = rd32(&raw[off]); off += 4;
    if (n_ref < 0) return HCNV_E_PARSE;
    hcnv_pack* P = new hcnv_pack();
    P->contigs.resize((size_t)n_ref + 1);                    // + the "chr*" pseudo contig (dropped again if it stays empty)
    for (int32_t i = 0; i < n_ref; ++i) {
        if (off + 4 > N) { delete P; return HCNV_E_PARSE; }
        int32_t l_name = rd32(&raw[off]); off += 4;
        if (l_name < 1 || off + (size_t)l_name + 4 > N) { delete P; return HCNV_E_PARSE; }
        hcnv_pack::Contig& c = P->contigs[(size_t)i];
        c.name = chr_prefix(std::string((const char*)&raw[off], (size_t)l_name - 1)); off += (size_t)l_name;
        c.length = rd32(&raw[off]); off += 4;
        if (snps) {
            auto it = snps->index.find(c.name);
            if (it != snps->index.end()) { c.snp_pos = &snps->pos[(size_t)it->second]; c.snp_zyg = &snps->zyg[(size_t)it->second]; }
        }
    }
    {   // a record without a reference (refID -1) has htsjdk's reference name "*", so the mapper files its alignment
        // blocks, if its CIGAR has any, under "chr*" (Mappers.java:176-180); the reference's own example/hg19.extra.bam
        // holds 241 such records (the unplaced hg19 scaffolds).  Its length is the furthest aligned base.
        hcnv_pack::Contig& c = P->contigs[(size_t)n_ref];
        c.name = "chr*";
        if (snps) {
            auto it = snps->index.find(c.name);
            if (it != snps->index.end()) { c.snp_pos = &snps->pos[(size_t)it->second]; c.snp_zyg = &snps->zyg[(size_t)it->second]; }
        }
    }
    const int32_t n_slots = n_ref + 1;
    // per contig: blocks come out in read order; later blocks of a read (after D/N) are held in a small heap until
    // the read starts catch up, so the stream is emitted sorted by start0 for coordinate-sorted input
    std::vector<std::priority_queue<PendingBlock, std::vector<PendingBlock>, ByStart>> heaps((size_t)n_slots);
    std::vector<char> unsorted((size_t)n_slots, 0);
    std::vector<int32_t> last_emit((size_t)n_slots, -2147483647 - 1);
    auto emit = [&](hcnv_pack::Contig& c, int ci, const PendingBlock& b) {
        if (b.start0 < last_emit[(size_t)ci]) unsorted[(size_t)ci] = 1;
        last_emit[(size_t)ci] = std::max(last_emit[(size_t)ci], b.start0);
        c.blocks.push_back(hcnv_block{b.start0, b.len_mapq});
    };
    int rc = HCNV_OK;
    while (off + 4 <= N) {
        int32_t bs = rd32(&raw[off]);
        if (bs < 32 || off + 4 + (size_t)bs > N) { rc = HCNV_E_PARSE; break; }
        const unsigned char* r = &raw[off + 4];
        off += 4 + (size_t)bs;
        ++P->n_records;
        int32_t ref_id = rd32(r); const int32_t pos0 = rd32(r + 4);
        const uint32_t l_read_name = r[8], mapq = r[9];
        const uint32_t n_cigar = rdu16(r + 12);
        const int32_t l_seq = rd32(r + 16);
        if (ref_id >= n_ref || ref_id < -1) { rc = HCNV_E_PARSE; break; }
        if (ref_id < 0) { ++P->n_unmapped; ref_id = n_ref; }                      // reference name "*" -> "chr*"
        if (l_seq < 0 || 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + (size_t)((l_seq + 1) / 2) > (size_t)bs) { rc = HCNV_E_PARSE; break; }
        const unsigned char* cig = r + 32 + l_read_name;
        const unsigned char* seq = cig + 4 * (size_t)n_cigar;
        hcnv_pack::Contig& c = P->contigs[(size_t)ref_id];
        auto& heap = heaps[(size_t)ref_id];
        while (!heap.empty() && heap.top().start0 <= pos0) { emit(c, ref_id, heap.top()); heap.pop(); }
        // htsjdk SAMRecord.getAlignmentBlocks: M/=/X make a block; I,S advance the read; D,N advance the reference; H,P nothing
        int64_t read_base = 1, ref_base = (int64_t)pos0 + 1;
        bool first_block = true;
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t v = rdu32(cig + 4 * (size_t)k), op = v & 15u; const int64_t len = v >> 4;
            if (op == 0 || op == 7 || op == 8) {
                // observations at the SNP sites this block covers: base = bases[read_base-1+i] or 'A' (Mappers.java:248-256)
                if (c.snp_pos && !c.snp_pos->empty()) {
                    const std::vector<int32_t>& sp = *c.snp_pos;
                    size_t lo = (size_t)(std::lower_bound(sp.begin(), sp.end(), (int32_t)std::min<int64_t>(ref_base, 2147483647)) - sp.begin());
                    for (size_t s = lo; s < sp.size() && (int64_t)sp[s] < ref_base + len; ++s) {
                        const int64_t ridx = read_base - 1 + ((int64_t)sp[s] - ref_base);
                        uint32_t code = 1;                                         // 'A'
                        if (ridx < (int64_t)l_seq) { const unsigned char byte = seq[ridx >> 1]; code = (ridx & 1) ? (byte & 15u) : (byte >> 4); }
                        c.obs.push_back((uint32_t)(s << 4) | code);
                    }
                }
                // split into pieces the packed formats can hold (splitting is results-neutral)
                int64_t s0 = ref_base - 1, left = len;
                if (len > HCNV_SHORT_MAX) {
                    while (left > 0) { int64_t piece = std::min<int64_t>(left, 0x7fffffff); c.longs.push_back(hcnv_long_block{(int32_t)s0, (int32_t)piece, (int32_t)mapq, 0}); s0 += piece; left -= piece; }
                } else if (len > 0) {
                    PendingBlock b{(int32_t)s0, (uint32_t)len | (mapq << 24)};
                    c.max_len = std::max<int32_t>(c.max_len, (int32_t)len);
                    if (first_block) emit(c, ref_id, b); else heap.push(b);
                    first_block = false;
                }
                read_base += len; ref_base += len;
                if (ref_id == n_ref) c.length = (int32_t)std::max<int64_t>(c.length, std::min<int64_t>(ref_base - 1, 2147483647));
            } else if (op == 1 || op == 4) read_base += len;        // I, S
            else if (op == 2 || op == 3) ref_base += len;           // D, N
            // H (5), P (6): nothing
        }
    }
    if (rc) { delete P; return rc; }
    for (int32_t i = 0; i < n_slots; ++i) {
        hcnv_pack::Contig& c = P->contigs[(size_t)i];
        auto& heap = heaps[(size_t)i];
        while (!heap.empty()) { emit(c, i, heap.top()); heap.pop(); }
        if (unsorted[(size_t)i])                                     // input was not coordinate-sorted: sort now
            std::stable_sort(c.blocks.begin(), c.blocks.end(), [](const hcnv_block& a, const hcnv_block& b) { return a.start0 < b.start0; });
    }
    if (P->contigs.back().blocks.empty() && P->contigs.back().longs.empty()) P->contigs.pop_back();
    *out = P;
    return HCNV_OK;
}

int64_t hcnv_compact_blocks(const hcnv_block* blocks, int64_t n, double mapping_quality_threshold,
                            hcnv_cblock* out, int32_t* anchors, int64_t capacity) {
    if (n < 0 || (n > 0 && !blocks) || (out && !anchors)) return HCNV_E_INVALID;
    bool pass[256];
    for (int q = 0; q < 256; ++q) pass[q] = (1. - std::pow(10.0, -q * .1)) >= mapping_quality_threshold;   // Mappers.java:193,200
    int64_t k = 0;                                          // compact records written (or counted) so far
    int64_t prev = 0;                                       // start0 of the last record written
    bool first = true;
    auto put = [&](int64_t start0, uint32_t delta, uint32_t len) -> bool {
        if (out) {
            if (k >= capacity) return false;
            if (k % HCNV_CGROUP == 0) anchors[k / HCNV_CGROUP] = (int32_t)start0;
            out[k] = (hcnv_cblock)(delta | (len << 8));
        }
        ++k;
        return true;
    };
    // one record of at most 255 bases at `pos` (>= every record written so far): fillers over the gap, then the record
    auto emit = [&](int64_t pos, uint32_t piece) -> bool {
        int64_t gap = first ? 0 : pos - prev;
        while (gap > 255) {                                 // fillers: len 0, advance 255
            prev += 255; gap -= 255;
            if (!put(prev, 255u, 0u)) return false;
        }
        if (!put(pos, (uint32_t)gap, piece)) return false;
        prev = pos; first = false;
        return true;
    };
    // the later pieces of a block above 255 bases wait here until the input catches up, so the stream stays sorted
    std::priority_queue<std::pair<int64_t, uint32_t>, std::vector<std::pair<int64_t, uint32_t>>, std::greater<std::pair<int64_t, uint32_t>>> later;
    int32_t last_in = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t s = blocks[i].start0;
        uint32_t len = blocks[i].len_mapq & 0xFFFFFFu; const uint32_t mapq = blocks[i].len_mapq >> 24;
        if (s < 0 || len > (uint32_t)HCNV_SHORT_MAX) return HCNV_E_RANGE;
        if (i > 0 && s < last_in) return HCNV_E_UNSORTED;
        last_in = s;
        if (!pass[mapq]) continue;
        while (!later.empty() && later.top().first <= s) { if (!emit(later.top().first, later.top().second)) return HCNV_E_CAPACITY; later.pop(); }
        if (!emit(s, len > 255u ? 255u : len)) return HCNV_E_CAPACITY;
        for (int64_t pos = (int64_t)s + 255; len > 255u; pos += 255) { len -= 255u; later.push({pos, len > 255u ? 255u : len}); }
    }
    while (!later.empty()) { if (!emit(later.top().first, later.top().second)) return HCNV_E_CAPACITY; later.pop(); }
    while (k % HCNV_CGROUP) if (!put(prev, 0u, 0u)) return HCNV_E_CAPACITY;
    return k;
}

int64_t hcnv_compact_obs(const uint32_t* obs, int64_t n, hcnv_cobs* out, int32_t* anchors, int64_t capacity) {
    if (n < 0 || (n > 0 && !obs) || (out && !anchors)) return HCNV_E_INVALID;
    int64_t k = 0;                                          // entries written (or counted) so far
    int64_t i = 0;
    while (i < n) {
        // one group: as many of the next observations as keep max - min of the indices within 4094, at most HCNV_OGROUP
        uint32_t lo = obs[i] >> 4, hi = lo;
        if (lo >= (1u << 28) - 1) return HCNV_E_RANGE;
        int64_t j = i + 1;
        for (; j < n && j - i < HCNV_OGROUP; ++j) {
            const uint32_t x = obs[j] >> 4;
            const uint32_t nlo = x < lo ? x : lo, nhi = x > hi ? x : hi;
            if (nhi - nlo > 4094u) break;
            lo = nlo; hi = nhi;
        }
        if (out) {
            if (k + HCNV_OGROUP > capacity) return HCNV_E_CAPACITY;
            anchors[k / HCNV_OGROUP] = (int32_t)lo;
            for (int64_t t = i; t < j; ++t) out[k + (t - i)] = (hcnv_cobs)((((obs[t] >> 4) - lo) << 4) | (obs[t] & 15u));
            for (int64_t t = j - i; t < HCNV_OGROUP; ++t) out[k + t] = (hcnv_cobs)HCNV_COBS_FILLER;
        }
        k += HCNV_OGROUP;
        i = j;
    }
    return k;
}

int32_t hcnv_pack_n_contigs(const hcnv_pack* p) { return p ? (int32_t)p->contigs.size() : 0; }
int64_t hcnv_pack_n_records(const hcnv_pack* p) { return p ? p->n_records : 0; }

int hcnv_pack_contig(const hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len) {
    if (!p || !in || i < 0 || i >= (int32_t)p->contigs.size()) return HCNV_E_INVALID;
    const hcnv_pack::Contig& c = p->contigs[(size_t)i];
    if (name) *name = c.name.c_str();
    memset(in, 0, sizeof *in);
    in->length = c.length;
    in->blocks = c.blocks.data(); in->n_blocks = (int64_t)c.blocks.size();
    in->long_blocks = c.longs.data(); in->n_long = (int64_t)c.longs.size();
    if (c.snp_pos) { in->snp_pos1 = c.snp_pos->data(); in->snp_zyg = c.snp_zyg->data(); in->n_snp = (int64_t)c.snp_pos->size(); }
    in->obs = c.obs.data(); in->n_obs = (int64_t)c.obs.size();
    if (max_block_len) *max_block_len = c.max_len;
    return HCNV_OK;
}

int hcnv_pack_contig_compact(hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len) {
    int rc = hcnv_pack_contig(p, i, name, in, max_block_len);
    if (rc) return rc;
    hcnv_pack::Contig& c = p->contigs[(size_t)i];
    if (c.n_cblocks < 0) {                                  // first request: build the 2-byte stream, 16-byte aligned
        const int64_t n = hcnv_compact_blocks(c.blocks.data(), (int64_t)c.blocks.size(), 0.0, nullptr, nullptr, 0);
        if (n < 0) return (int)n;
        c.cstore.resize((size_t)n + 8);
        const uintptr_t a = reinterpret_cast<uintptr_t>(c.cstore.data());
        c.cblocks = c.cstore.data() + ((16 - (a & 15)) & 15) / 2;
        c.anchors.resize((size_t)(n / HCNV_CGROUP) + 1);
        const int64_t m = hcnv_compact_blocks(c.blocks.data(), (int64_t)c.blocks.size(), 0.0, c.cblocks, c.anchors.data(), n);
        if (m != n) return HCNV_E_INTERNAL;
        c.n_cblocks = n;
    }
    if (c.n_cobs < 0) {                                     // and the 2-byte observations
        const int64_t n = hcnv_compact_obs(c.obs.data(), (int64_t)c.obs.size(), nullptr, nullptr, 0);
        if (n < 0) return (int)n;
        c.cobs.resize((size_t)n);
        c.cobs_anchors.resize((size_t)(n / HCNV_OGROUP) + 1);
        const int64_t m = hcnv_compact_obs(c.obs.data(), (int64_t)c.obs.size(), c.cobs.data(), c.cobs_anchors.data(), n);
        if (m != n) return HCNV_E_INTERNAL;
        c.n_cobs = n;
    }
    in->blocks = nullptr; in->n_blocks = 0;
    in->cblocks = c.n_cblocks ? c.cblocks : nullptr; in->anchors = c.n_cblocks ? c.anchors.data() : nullptr; in->n_cblocks = c.n_cblocks;
    in->obs = nullptr; in->n_obs = 0;
    in->cobs = c.n_cobs ? c.cobs.data() : nullptr; in->cobs_anchors = c.n_cobs ? c.cobs_anchors.data() : nullptr; in->n_cobs = c.n_cobs;
    return HCNV_OK;
}

void hcnv_pack_free(hcnv_pack* p) { delete p; }

}  // extern "C"
