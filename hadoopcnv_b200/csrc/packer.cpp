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
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
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
        std::vector<uint8_t> wstore, wside; std::vector<int32_t> wside_offsets; hcnv_wblock* wblocks = nullptr; int64_t n_wside = -1; int32_t wdefault = 0;   // built on demand
        std::vector<hcnv_wobs> wobs; std::vector<int32_t> wobs_anchors; int64_t n_wobs = -1;                                        // built on demand
        std::vector<hcnv_cobs> cobs; std::vector<int32_t> cobs_anchors; int64_t n_cobs = -1;                                        // built on demand
        const std::vector<int32_t>* snp_pos = nullptr; const std::vector<int8_t>* snp_zyg = nullptr;
    };
    std::vector<Contig> contigs;
    int64_t n_records = 0, n_unmapped = 0;
    int64_t bytes_compressed = 0, bytes_inflated = 0;
    double seconds = 0.0;
    int threads = 1;
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

// ---- streaming BGZF reader -------------------------------------------------------------------------------------------
// A 30X BAM is ~50 GB compressed and ~150 GB inflated: the file is never held in memory.  A reader thread cuts the file into
// batches of whole BGZF blocks (the BC extra field gives every block's size without inflating it), a pool of threads inflates
// the batches (raw deflate per block, each at its own offset: the blocks' ISIZE trailers give the layout), and the calling
// thread parses the inflated batches IN ORDER -- BAM records straddle block and batch boundaries freely, so a record that is
// cut at a batch's end is finished from the head of the next one.  The number of batches in flight is bounded.
namespace {

struct Batch {
    std::vector<unsigned char> comp;                      // the compressed blocks, back to back
    std::vector<uint32_t> cdata_off, cdata_len, isize, out_off;
    std::vector<unsigned char> out;
    bool done = false, failed = false;
};

class BgzfPipeline {
public:
    BgzfPipeline(FILE* f, int threads) : f_(f), nthreads_(std::max(1, threads)) {
        limit_ = (size_t)(2 * nthreads_ + 2);
        reader_ = std::thread([this] { read_loop(); });
        for (int i = 0; i < nthreads_; ++i) workers_.emplace_back([this] { inflate_loop(); });
    }
    ~BgzfPipeline() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        if (reader_.joinable()) reader_.join();
        for (auto& t : workers_) if (t.joinable()) t.join();
    }
    // next inflated batch in file order; nullptr at the end of the file; *bad on a corrupt or non-BGZF file
    std::shared_ptr<Batch> next(bool* bad) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return !order_.empty() ? order_.front()->done : eof_; });
        if (order_.empty()) { *bad = reader_failed_; return nullptr; }
        std::shared_ptr<Batch> b = order_.front();
        order_.pop_front();
        --in_flight_;
        lk.unlock();
        cv_.notify_all();
        if (b->failed) { *bad = true; return nullptr; }
        return b;
    }
    int64_t bytes_compressed() const { return bytes_compressed_; }

private:
    void read_loop() {
        std::vector<unsigned char> buf;
        size_t pos = 0;
        bool file_end = false;
        auto fill = [&](size_t need) {                     // make buf[pos .. pos + need) available if the file has it
            if (buf.size() - pos >= need) return true;
            if (pos > 0) { buf.erase(buf.begin(), buf.begin() + (long)pos); pos = 0; }
            while (!file_end && buf.size() < need) {
                const size_t old = buf.size();
                buf.resize(old + (1u << 20));
                const size_t n = fread(buf.data() + old, 1, 1u << 20, f_);
                buf.resize(old + n);
                if (n == 0) file_end = true;
            }
            return buf.size() - pos >= need;
        };
        for (;;) {
            auto b = std::make_shared<Batch>();
            size_t payload = 0, inflated = 0;
            bool bad = false;
            while (payload < (1u << 20)) {
                if (!fill(18)) { if (buf.size() - pos != 0) bad = true; break; }
                const unsigned char* h = buf.data() + pos;
                if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { bad = true; break; }
                const unsigned xlen = rdu16(h + 10);
                if (!fill(12 + (size_t)xlen)) { bad = true; break; }
                h = buf.data() + pos;
                long bsize = -1;
                for (unsigned x = 0; x + 4 <= xlen;) {
                    const unsigned slen = rdu16(h + 12 + x + 2);
                    if (h[12 + x] == 'B' && h[12 + x + 1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (long)rdu16(h + 12 + x + 4) + 1;
                    x += 4 + slen;
                }
                if (bsize < (long)(12 + xlen + 8)) { bad = true; break; }
                if (!fill((size_t)bsize)) { bad = true; break; }
                h = buf.data() + pos;
                const size_t o = b->comp.size();
                b->comp.insert(b->comp.end(), h, h + bsize);
                b->cdata_off.push_back((uint32_t)(o + 12 + xlen));
                b->cdata_len.push_back((uint32_t)(bsize - 12 - xlen - 8));
                const uint32_t isz = rdu32(h + bsize - 4);
                b->isize.push_back(isz);
                b->out_off.push_back((uint32_t)inflated);
                inflated += isz;
                payload += (size_t)bsize;
                pos += (size_t)bsize;
            }
            if (b->comp.empty() && !bad) break;
            b->out.resize(inflated);
            bytes_compressed_ += (int64_t)b->comp.size();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return in_flight_ < limit_ || stop_; });
                if (stop_) return;
                if (bad) { reader_failed_ = true; break; }
                order_.push_back(b); todo_.push_back(b); ++in_flight_;
            }
            cv_.notify_all();
        }
        { std::lock_guard<std::mutex> lk(mu_); eof_ = true; }
        cv_.notify_all();
    }
    void inflate_loop() {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        const bool zok = inflateInit2(&zs, -15) == Z_OK;
        for (;;) {
            std::shared_ptr<Batch> b;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !todo_.empty() || eof_ || stop_; });
                if (stop_ || (todo_.empty() && eof_)) break;
                b = todo_.front(); todo_.pop_front();
            }
            bool ok = zok;
            for (size_t i = 0; ok && i < b->isize.size(); ++i) {
                if (b->isize[i] == 0) continue;
                inflateReset(&zs);
                zs.next_in = b->comp.data() + b->cdata_off[i]; zs.avail_in = b->cdata_len[i];
                zs.next_out = b->out.data() + b->out_off[i]; zs.avail_out = b->isize[i];
                const int rc = inflate(&zs, Z_FINISH);
                ok = rc == Z_STREAM_END && zs.avail_out == 0;
            }
            std::vector<unsigned char>().swap(b->comp);
            { std::lock_guard<std::mutex> lk(mu_); b->done = true; b->failed = !ok; }
            cv_.notify_all();
        }
        if (zok) inflateEnd(&zs);
    }
    FILE* f_;
    int nthreads_;
    size_t limit_ = 4, in_flight_ = 0;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Batch>> order_, todo_;
    bool eof_ = false, stop_ = false, reader_failed_ = false;
    std::thread reader_;
    std::vector<std::thread> workers_;
    std::atomic<int64_t> bytes_compressed_{0};
};

// first SNP site at or after `key`, starting from where the previous block of this contig looked (reads come sorted)
inline size_t snp_seek(const std::vector<int32_t>& sp, size_t hint, int32_t key) {
    const size_t n = sp.size();
    size_t lo = std::min(hint, n);
    if (lo > 0 && sp[lo - 1] >= key) return (size_t)(std::lower_bound(sp.begin(), sp.begin() + (long)lo, key) - sp.begin());
    for (int steps = 0; lo < n && sp[lo] < key; ++lo)
        if (++steps > 8) return (size_t)(std::lower_bound(sp.begin() + (long)lo, sp.end(), key) - sp.begin());
    return lo;
}

}  // namespace

int hcnv_bam_pack_mt(const char* bam_path, const hcnv_snp_table* snps, int32_t threads, hcnv_pack** out) {
    if (!bam_path || !out) return HCNV_E_INVALID;
    *out = nullptr;
    FILE* f = fopen(bam_path, "rb");
    if (!f) return HCNV_E_IO;
    const auto t_start = std::chrono::steady_clock::now();
    int nthreads = threads > 0 ? threads : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    hcnv_pack* P = new hcnv_pack();
    P->threads = nthreads;
    int rc = HCNV_OK;
    {
        BgzfPipeline pipe(f, nthreads);
        // ---- parser state
        std::vector<unsigned char> carry;                  // the bytes of a header / record that is cut at a batch's end
        bool header_done = false;
        int32_t n_ref = 0, n_slots = 0;
        std::vector<std::priority_queue<PendingBlock, std::vector<PendingBlock>, ByStart>> heaps;
        std::vector<char> unsorted;
        std::vector<int32_t> last_emit;
        std::vector<size_t> snp_hint;
        auto emit = [&](hcnv_pack::Contig& c, int ci, const PendingBlock& b) {
            if (b.start0 < last_emit[(size_t)ci]) unsorted[(size_t)ci] = 1;
            last_emit[(size_t)ci] = std::max(last_emit[(size_t)ci], b.start0);
            c.blocks.push_back(hcnv_block{b.start0, b.len_mapq});
        };
        // the BAM header, once all of it is in `d`: returns its length, 0 if more bytes are needed, -1 if malformed
        auto parse_header = [&](const unsigned char* d, size_t N) -> long {
            if (N < 12) return 0;
            if (memcmp(d, "BAM\1", 4) != 0) return -1;
            size_t off = 4;
            const int32_t l_text = rd32(d + off); off += 4;
            if (l_text < 0) return -1;
            if (off + (size_t)l_text + 4 > N) return 0;
            off += (size_t)l_text;
            const int32_t nr = rd32(d + off); off += 4;
            if (nr < 0) return -1;
            std::vector<std::pair<std::string, int32_t>> refs;
            for (int32_t i = 0; i < nr; ++i) {
                if (off + 4 > N) return 0;
                const int32_t l_name = rd32(d + off); off += 4;
                if (l_name < 1) return -1;
                if (off + (size_t)l_name + 4 > N) return 0;
                refs.emplace_back(chr_prefix(std::string((const char*)d + off, (size_t)l_name - 1)), rd32(d + off + (size_t)l_name));
                off += (size_t)l_name + 4;
            }
            n_ref = nr; n_slots = nr + 1;
            P->contigs.resize((size_t)n_slots);              // + the "chr*" pseudo contig (dropped again if it stays empty)
            for (int32_t i = 0; i < n_slots; ++i) {
                hcnv_pack::Contig& c = P->contigs[(size_t)i];
                // a record without a reference (refID -1) has htsjdk's reference name "*", so the mapper files its alignment
                // blocks, if its CIGAR has any, under "chr*" (Mappers.java:176-180); the reference's own example/hg19.extra.bam
                // holds 241 such records (the unplaced hg19 scaffolds).  Its length is the furthest aligned base.
                c.name = i < nr ? refs[(size_t)i].first : std::string("chr*");
                c.length = i < nr ? refs[(size_t)i].second : 0;
                if (snps) {
                    auto it = snps->index.find(c.name);
                    if (it != snps->index.end()) { c.snp_pos = &snps->pos[(size_t)it->second]; c.snp_zyg = &snps->zyg[(size_t)it->second]; }
                }
            }
            heaps.resize((size_t)n_slots); unsorted.assign((size_t)n_slots, 0); last_emit.assign((size_t)n_slots, -2147483647 - 1);
            snp_hint.assign((size_t)n_slots, 0);
            return (long)off;
        };
        // one alignment record (without its block_size word), Mappers.java:174-212 with htsjdk's getAlignmentBlocks rules
        auto record = [&](const unsigned char* r, int32_t bs) -> int {
            ++P->n_records;
            int32_t ref_id = rd32(r); const int32_t pos0 = rd32(r + 4);
            const uint32_t l_read_name = r[8], mapq = r[9];
            const uint32_t n_cigar = rdu16(r + 12);
            const int32_t l_seq = rd32(r + 16);
            if (ref_id >= n_ref || ref_id < -1) return HCNV_E_PARSE;
            if (ref_id < 0) { ++P->n_unmapped; ref_id = n_ref; }                      // reference name "*" -> "chr*"
            if (l_seq < 0 || 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + (size_t)((l_seq + 1) / 2) > (size_t)bs) return HCNV_E_PARSE;
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
                        const size_t lo = snp_seek(sp, snp_hint[(size_t)ref_id], (int32_t)std::min<int64_t>(ref_base, 2147483647));
                        if (first_block) snp_hint[(size_t)ref_id] = lo;
                        for (size_t s_ = lo; s_ < sp.size() && (int64_t)sp[s_] < ref_base + len; ++s_) {
                            const int64_t ridx = read_base - 1 + ((int64_t)sp[s_] - ref_base);
                            uint32_t code = 1;                                         // 'A'
                            if (ridx < (int64_t)l_seq) { const unsigned char byte = seq[ridx >> 1]; code = (ridx & 1) ? (byte & 15u) : (byte >> 4); }
                            c.obs.push_back((uint32_t)(s_ << 4) | code);
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
            return HCNV_OK;
        };
        for (;;) {
            bool bad = false;
            std::shared_ptr<Batch> b = pipe.next(&bad);
            if (!b) { if (bad) rc = HCNV_E_PARSE; break; }
            P->bytes_inflated += (int64_t)b->out.size();
            const unsigned char* d = b->out.data();
            size_t N = b->out.size(), off = 0;
            if (!header_done) {                              // (headers are small: gather bytes until the whole header is there)
                carry.insert(carry.end(), d, d + N);
                const long hl = parse_header(carry.data(), carry.size());
                if (hl < 0) { rc = HCNV_E_PARSE; break; }
                if (hl == 0) continue;
                header_done = true;
                std::vector<unsigned char> rest(carry.begin() + hl, carry.end());
                carry.swap(rest);
                // the rest of the gathered bytes are records: parse them out of `carry`, keep what is cut
                size_t o = 0;
                while (o + 4 <= carry.size()) {
                    const int32_t bs = rd32(carry.data() + o);
                    if (bs < 32) { rc = HCNV_E_PARSE; break; }
                    if (o + 4 + (size_t)bs > carry.size()) break;
                    if ((rc = record(carry.data() + o + 4, bs))) break;
                    o += 4 + (size_t)bs;
                }
                if (rc) break;
                carry.erase(carry.begin(), carry.begin() + (long)o);
                continue;
            }
            if (!carry.empty()) {                            // finish the record that was cut at the previous batch's end
                while (carry.size() < 4 && off < N) carry.push_back(d[off++]);
                if (carry.size() < 4) continue;
                const int32_t bs = rd32(carry.data());
                if (bs < 32) { rc = HCNV_E_PARSE; break; }
                const size_t need = 4 + (size_t)bs - carry.size();
                const size_t take = std::min(need, N - off);
                carry.insert(carry.end(), d + off, d + off + take);
                off += take;
                if (carry.size() < 4 + (size_t)bs) continue;
                if ((rc = record(carry.data() + 4, bs))) break;
                carry.clear();
            }
            while (off + 4 <= N) {
                const int32_t bs = rd32(d + off);
                if (bs < 32) { rc = HCNV_E_PARSE; break; }
                if (off + 4 + (size_t)bs > N) break;
                if ((rc = record(d + off + 4, bs))) break;
                off += 4 + (size_t)bs;
            }
            if (rc) break;
            carry.assign(d + off, d + N);
        }
        if (!rc && (!header_done || !carry.empty())) rc = HCNV_E_PARSE;       // truncated file
        P->bytes_compressed = pipe.bytes_compressed();
        if (!rc) {
            for (int32_t i = 0; i < n_slots; ++i) {
                hcnv_pack::Contig& c = P->contigs[(size_t)i];
                auto& heap = heaps[(size_t)i];
                while (!heap.empty()) { emit(c, i, heap.top()); heap.pop(); }
                if (unsorted[(size_t)i])                                     // input was not coordinate-sorted: sort now
                    std::stable_sort(c.blocks.begin(), c.blocks.end(), [](const hcnv_block& a, const hcnv_block& b2) { return a.start0 < b2.start0; });
            }
            if (!P->contigs.empty() && P->contigs.back().blocks.empty() && P->contigs.back().longs.empty()) P->contigs.pop_back();
        }
    }
    fclose(f);
    if (rc) { delete P; return rc; }
    P->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    *out = P;
    return HCNV_OK;
}

int hcnv_bam_pack(const char* bam_path, const hcnv_snp_table* snps, hcnv_pack** out) { return hcnv_bam_pack_mt(bam_path, snps, 0, out); }

void hcnv_pack_stats(const hcnv_pack* p, double out[5]) {
    if (!p || !out) return;
    out[0] = (double)p->n_records; out[1] = (double)p->bytes_compressed; out[2] = (double)p->bytes_inflated; out[3] = p->seconds; out[4] = (double)p->threads;
}

// ---- synthetic BAM writer (tests and tools/bench_packer.py: the packer needs inputs of real size and there is no network) ------
namespace {
struct Xs { uint64_t s; uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; } double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); } };

int reg2bin(int64_t beg, int64_t end) {                     // SAM specification, section 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// BGZF blocks of `data`, compressed by `threads` threads, appended to the file in order
bool write_bgzf(FILE* f, const std::vector<unsigned char>& data, int threads, int level) {
    const size_t BLK = 0xff00;
    const size_t nb = (data.size() + BLK - 1) / BLK;
    std::vector<std::vector<unsigned char>> outs(nb);
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
        z_stream zs;
        std::vector<unsigned char> tmp(BLK + 1024);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nb) break;
            const size_t o = i * BLK, n = std::min(BLK, data.size() - o);
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; break; }
            zs.next_in = const_cast<unsigned char*>(data.data() + o); zs.avail_in = (uInt)n;
            zs.next_out = tmp.data(); zs.avail_out = (uInt)tmp.size();
            const int rc = deflate(&zs, Z_FINISH);
            const size_t clen = tmp.size() - zs.avail_out;
            deflateEnd(&zs);
            if (rc != Z_STREAM_END || clen + 26 > 65536) { ok = false; break; }
            std::vector<unsigned char>& b = outs[i];
            b.resize(18 + clen + 8);
            const unsigned char hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
            memcpy(b.data(), hdr, 12);
            b[12] = 'B'; b[13] = 'C'; b[14] = 2; b[15] = 0;
            const uint16_t bsize = (uint16_t)(b.size() - 1);
            memcpy(b.data() + 16, &bsize, 2);
            memcpy(b.data() + 18, tmp.data(), clen);
            const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data.data() + o, (uInt)n), isz = (uint32_t)n;
            memcpy(b.data() + 18 + clen, &crc, 4); memcpy(b.data() + 22 + clen, &isz, 4);
        }
    };
    std::vector<std::thread> ts;
    for (int t = 1; t < std::max(1, threads); ++t) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    if (!ok) return false;
    for (auto& b : outs) if (fwrite(b.data(), 1, b.size(), f) != b.size()) return false;
    return true;
}
}  // namespace

// Writes a coordinate-sorted BAM of random reads: per reference n_reads[r] reads of read_len bases with uniform starts, the CIGAR
// mix of SURVEY.md 8(d) (90 % <len>M, 2.5 % one 1-10 bp insertion, 2.5 % one 1-10 bp deletion, 5 % a 5-30 bp soft clip), random
// bases, MAPQ 60 for 90 % of the reads.  stats[0..2] = records, alignment blocks, aligned bases.  Returns 0 or a status.
int hcnv_bam_write_synthetic(const char* path, int32_t n_ref, const char* const* names, const int32_t* lengths, const int64_t* n_reads,
                             int32_t read_len, uint64_t seed, int32_t threads, int64_t stats[3]) {
    if (!path || n_ref <= 0 || !names || !lengths || !n_reads || read_len < 40 || read_len > 1000) return HCNV_E_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return HCNV_E_IO;
    const int nt = threads > 0 ? threads : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<unsigned char> buf;
    auto put32 = [&](int32_t v) { unsigned char b[4]; memcpy(b, &v, 4); buf.insert(buf.end(), b, b + 4); };
    std::string text = "@HD\tVN:1.4\tSO:coordinate\n";
    for (int r = 0; r < n_ref; ++r) text += std::string("@SQ\tSN:") + names[r] + "\tLN:" + std::to_string(lengths[r]) + "\n";
    buf.insert(buf.end(), {'B', 'A', 'M', 1});
    put32((int32_t)text.size()); buf.insert(buf.end(), text.begin(), text.end());
    put32(n_ref);
    for (int r = 0; r < n_ref; ++r) { const size_t l = strlen(names[r]) + 1; put32((int32_t)l); buf.insert(buf.end(), names[r], names[r] + l); put32(lengths[r]); }
    Xs rng{seed * 0x9E3779B97F4A7C15ull + 0x1234567ull};
    int64_t n_rec = 0, n_blk = 0, n_bases = 0;
    bool ok = true;
    for (int r = 0; r < n_ref && ok; ++r) {
        const int64_t span = std::max<int64_t>(1, (int64_t)lengths[r] - read_len - 40);
        std::vector<int32_t> st((size_t)n_reads[r]);
        for (auto& x : st) x = (int32_t)(rng.next() % (uint64_t)span);
        std::sort(st.begin(), st.end());
        for (size_t i = 0; i < st.size() && ok; ++i) {
            const double u = rng.uni();
            uint32_t cig[3]; int nc = 0;
            int64_t ref_span = read_len;
            if (u < 0.90) { cig[nc++] = (uint32_t)read_len << 4; n_blk += 1; n_bases += read_len; }
            else if (u < 0.925) { const int k = 1 + (int)(rng.next() % 10), j = 10 + (int)(rng.next() % (uint64_t)(read_len - 30)); cig[nc++] = (uint32_t)j << 4; cig[nc++] = ((uint32_t)k << 4) | 1; cig[nc++] = (uint32_t)(read_len - j - k) << 4; n_blk += 2; n_bases += read_len - k; ref_span = read_len - k; }
            else if (u < 0.95) { const int k = 1 + (int)(rng.next() % 10), j = 10 + (int)(rng.next() % (uint64_t)(read_len - 30)); cig[nc++] = (uint32_t)j << 4; cig[nc++] = ((uint32_t)k << 4) | 2; cig[nc++] = (uint32_t)(read_len - j) << 4; n_blk += 2; n_bases += read_len; ref_span = read_len + k; }
            else { const int k = 5 + (int)(rng.next() % 26); cig[nc++] = ((uint32_t)k << 4) | 4; cig[nc++] = (uint32_t)(read_len - k) << 4; n_blk += 1; n_bases += read_len - k; ref_span = read_len - k; }
            char name[24];
            const int ln = snprintf(name, sizeof name, "r%lld", (long long)n_rec) + 1;
            const int mapq = rng.uni() < 0.9 ? 60 : (int)(rng.next() % 60);
            const int32_t bs = 32 + ln + 4 * nc + (read_len + 1) / 2 + read_len;
            put32(bs); put32(r); put32(st[i]);
            const uint32_t bin_mq_nl = ((uint32_t)reg2bin(st[i], st[i] + ref_span) << 16) | ((uint32_t)mapq << 8) | (uint32_t)ln;
            put32((int32_t)bin_mq_nl);
            put32((int32_t)(((uint32_t)0 << 16) | (uint32_t)nc));
            put32(read_len); put32(-1); put32(-1); put32(0);
            buf.insert(buf.end(), name, name + ln);
            for (int c = 0; c < nc; ++c) put32((int32_t)cig[c]);
            const size_t so = buf.size();
            buf.resize(so + (size_t)(read_len + 1) / 2 + (size_t)read_len);
            static const unsigned char code[4] = {1, 2, 4, 8};
            for (int b = 0; b < (read_len + 1) / 2; ++b) { const uint64_t x = rng.next(); buf[so + (size_t)b] = (unsigned char)((code[x & 3] << 4) | code[(x >> 2) & 3]); }
            memset(buf.data() + so + (size_t)(read_len + 1) / 2, 30, (size_t)read_len);
            ++n_rec;
            if (buf.size() >= ((size_t)32 << 20)) { ok = write_bgzf(f, buf, nt, 1); buf.clear(); }
        }
    }
    if (ok && !buf.empty()) ok = write_bgzf(f, buf, nt, 1);
    static const unsigned char eof_block[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) ok = fwrite(eof_block, 1, 28, f) == 28;
    ok = (fclose(f) == 0) && ok;
    if (stats) { stats[0] = n_rec; stats[1] = n_blk; stats[2] = n_bases; }
    return ok ? HCNV_OK : HCNV_E_IO;
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

int64_t hcnv_wire_from_cblocks(const hcnv_cblock* cb, int64_t n, int32_t default_len,
                               hcnv_wblock* out, uint8_t* wside, int32_t* wside_offsets, int64_t wside_capacity, int32_t* default_len_used) {
    if (n < 0 || n % HCNV_CGROUP || (n > 0 && !cb) || (out && (!wside_offsets || (wside_capacity > 0 && !wside))) || default_len < 0 || default_len > 255)
        return HCNV_E_INVALID;
    if (default_len == 0) {                                 // the most frequent length (the read length)
        std::vector<int64_t> hist(256, 0);
        for (int64_t i = 0; i < n; ++i) ++hist[cb[i] >> 8];
        int best = 1;
        for (int l = 1; l < 256; ++l) if (hist[(size_t)l] > hist[(size_t)best]) best = l;
        default_len = best;
    }
    if (default_len_used) *default_len_used = default_len;
    int64_t ns = 0;                                         // side-stream bytes written (or counted) so far
    for (int64_t g = 0; g < n / HCNV_CGROUP; ++g) {
        const hcnv_cblock* rec = cb + g * HCNV_CGROUP;
        if (out) { wside_offsets[g] = (int32_t)ns; memset(out + g * HCNV_WGROUP_BYTES, 0, HCNV_WGROUP_BYTES); }
        hcnv_wblock* w = out ? out + g * HCNV_WGROUP_BYTES : nullptr;
        for (int i = 0; i < HCNV_CGROUP; ++i) {
            const uint32_t delta = i ? (rec[i] & 0xFFu) : 0u, len = rec[i] >> 8;       // (the group's first delta is ignored: the anchor says where it starts)
            const bool esc = delta >= 15u, own = len != (uint32_t)default_len;
            if (out) {
                if (ns + (esc ? 1 : 0) + (own ? 1 : 0) > wside_capacity) return HCNV_E_CAPACITY;
                w[i >> 1] |= (hcnv_wblock)((esc ? 15u : delta) << (4 * (i & 1)));
                if (!own) w[64 + (i >> 3)] |= (hcnv_wblock)(1u << (i & 7));
                if (esc) wside[ns] = (uint8_t)delta;
                if (own) wside[ns + (esc ? 1 : 0)] = (uint8_t)len;
            }
            ns += (esc ? 1 : 0) + (own ? 1 : 0);
        }
        if (ns >= (1ll << 31)) return HCNV_E_RANGE;
    }
    if (out) wside_offsets[n / HCNV_CGROUP] = (int32_t)ns;
    return ns;
}

}  // extern "C"

namespace {
// observations -> groups of HCNV_OGROUP entries whose indices lie within `span` of the group's smallest (its anchor)
template <class T, int GROUP>
int64_t group_obs(const uint32_t* obs, int64_t n, T* out, int32_t* anchors, int64_t capacity, uint32_t span, T filler) {
    if (n < 0 || (n > 0 && !obs) || (out && !anchors)) return HCNV_E_INVALID;
    int64_t k = 0;                                          // entries written (or counted) so far
    int64_t i = 0;
    while (i < n) {
        // one group: as many of the next observations as keep max - min of the indices within the span, at most GROUP
        uint32_t lo = obs[i] >> 4, hi = lo;
        if (lo >= (1u << 28) - 1) return HCNV_E_RANGE;
        int64_t j = i + 1;
        for (; j < n && j - i < GROUP; ++j) {
            const uint32_t x = obs[j] >> 4;
            const uint32_t nlo = x < lo ? x : lo, nhi = x > hi ? x : hi;
            if (nhi - nlo > span) break;
            lo = nlo; hi = nhi;
        }
        if (out) {
            if (k + GROUP > capacity) return HCNV_E_CAPACITY;
            anchors[k / GROUP] = (int32_t)lo;
            for (int64_t t = i; t < j; ++t) out[k + (t - i)] = (T)((((obs[t] >> 4) - lo) << 4) | (obs[t] & 15u));
            for (int64_t t = j - i; t < GROUP; ++t) out[k + t] = filler;
        }
        k += GROUP;
        i = j;
    }
    while (k % HCNV_OGROUP) {                               // (groups smaller than a tally span: whole groups of fillers up to one)
        if (out) {
            if (k + GROUP > capacity) return HCNV_E_CAPACITY;
            anchors[k / GROUP] = 0;
            for (int t = 0; t < GROUP; ++t) out[k + t] = filler;
        }
        k += GROUP;
    }
    return k;
}
}  // namespace

extern "C" {

int64_t hcnv_compact_obs(const uint32_t* obs, int64_t n, hcnv_cobs* out, int32_t* anchors, int64_t capacity) {
    return group_obs<hcnv_cobs, HCNV_OGROUP>(obs, n, out, anchors, capacity, 4094u, (hcnv_cobs)HCNV_COBS_FILLER);
}

int64_t hcnv_wire_obs(const uint32_t* obs, int64_t n, hcnv_wobs* out, int32_t* anchors, int64_t capacity) {
    return group_obs<hcnv_wobs, HCNV_WOGROUP>(obs, n, out, anchors, capacity, 14u, (hcnv_wobs)HCNV_WOBS_FILLER);
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

int hcnv_pack_contig_wire(hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len) {
    int rc = hcnv_pack_contig_compact(p, i, name, in, max_block_len);
    if (rc) return rc;
    hcnv_pack::Contig& c = p->contigs[(size_t)i];
    if (c.n_wside < 0) {                                    // first request: the wire form of the 2-byte stream, 16-byte aligned
        int32_t dl = 0;
        const int64_t ns = hcnv_wire_from_cblocks(c.cblocks, c.n_cblocks, 0, nullptr, nullptr, nullptr, 0, &dl);
        if (ns < 0) return (int)ns;
        const size_t groups = (size_t)(c.n_cblocks / HCNV_CGROUP);
        c.wstore.resize(groups * HCNV_WGROUP_BYTES + 16);
        const uintptr_t a = reinterpret_cast<uintptr_t>(c.wstore.data());
        c.wblocks = c.wstore.data() + ((16 - (a & 15)) & 15);
        c.wside.resize((size_t)ns + 1);
        c.wside_offsets.resize(groups + 1);
        const int64_t m = hcnv_wire_from_cblocks(c.cblocks, c.n_cblocks, dl, c.wblocks, c.wside.data(), c.wside_offsets.data(), ns, &c.wdefault);
        if (m != ns) return HCNV_E_INTERNAL;
        c.n_wside = ns;
    }
    if (c.n_cblocks > 0) {
        in->cblocks = nullptr;
        in->wblocks = c.wblocks; in->wside = c.wside.data(); in->wside_offsets = c.wside_offsets.data();
        in->n_wside = c.n_wside; in->wire_default_len = c.wdefault;
    }
    if (c.n_wobs < 0) {                                     // and the 1-byte observations
        const int64_t n = hcnv_wire_obs(c.obs.data(), (int64_t)c.obs.size(), nullptr, nullptr, 0);
        if (n < 0) return (int)n;
        c.wobs.resize((size_t)n);
        c.wobs_anchors.resize((size_t)(n / HCNV_WOGROUP) + 1);
        const int64_t m = hcnv_wire_obs(c.obs.data(), (int64_t)c.obs.size(), c.wobs.data(), c.wobs_anchors.data(), n);
        if (m != n) return HCNV_E_INTERNAL;
        c.n_wobs = n;
    }
    in->cobs = nullptr;
    in->wobs = c.n_wobs ? c.wobs.data() : nullptr; in->cobs_anchors = c.n_wobs ? c.wobs_anchors.data() : nullptr; in->n_cobs = c.n_wobs;
    return HCNV_OK;
}

void hcnv_pack_free(hcnv_pack* p) { delete p; }

}  // extern "C"
