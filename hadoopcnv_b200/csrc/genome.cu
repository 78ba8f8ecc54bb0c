// genome.cu -- one sample end to end behind ONE C-ABI call: what PennCnvSeq.run's three jobs do for a BAM's packed records
// (PennCnvSeq.java:285-350: depth caller -> binner -> one CnvReducer call per chromosome, Reducers.java:225-243), from HOST
// buffers to the results.txt columns (Hmm.java:479-490) in HOST buffers.
//
// The upload bounds this path (0.63 GB per 30X genome in the wire forms, 1.4 GB in the 2-byte forms), so it never pauses: one copy
// stream sends every chromosome's packed records, longest chromosome first, into device staging buffers back to back in pieces of
// at most 64 MB; the chromosomes are grouped into a few tasks of tapering size, and a worker thread (one hcnv_ctx = one stream + its scratch each, like the task
// slots of a Hadoop node) takes a task as soon as its records have landed: hcnv_bin_contigs -> hcnv_bins_to_hmm_input ->
// hcnv_segment_batch on device buffers, then hcnv_download_results per chromosome straight into the caller's arrays -- kernels
// and downloads run behind the uploads of the tasks after it.  Everything the steady state needs (staging buffers, worker
// scratch, streams, events) is kept from run to run.
#include "hcnv_internal.h"
#include <algorithm>
#include <chrono>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

struct GenomeTask { std::vector<int> contigs; };
struct UploadJob { std::vector<int> contigs; std::vector<hcnv_contig_input> host; };   // one pushed group: its host arrays, to be uploaded in order

struct hcnv_genome {
    int device = 0;
    std::vector<hcnv_ctx*> ctx;                   // one per worker
    cudaStream_t copy = nullptr;
    std::vector<DevBuf> stage;                    // ST_N per contig: blocks, long, snp_pos, snp_zyg, obs, cblocks, anchors, cobs, cobs_anchors, wire form (3)
    std::vector<cudaEvent_t> landed;              // one per contig
    // The upload is PACED: pieces of at most piece_bytes, at most depth x piece_bytes queued on the copy stream at any time (small
    // arrays ride along, up to the ring's size).  With everything queued at once (1.4 GB), a copy engine that serves its queue in submission order makes every other copy of the process -- the
    // workers' descriptor tables going up, the results coming down -- wait for the END of the upload: 31 ms per genome in one run,
    // 43 - 66 ms in the next (same binary, same box).  Measured with pacing in round 2's first form (2-byte streams; ms per 30X genome
    // end to end, piece MB x pieces in flight):
    // 64 x 2: 32.3, 32 x 4: 32.6, 16 x 3: 32.6 - 32.7, 8 x 3: 33.6, 4 x 4: 35.4, 16 x 6: 37.4, 8 x 8: 105 (many queued pieces are
    // as bad as one huge one); HCNV_UPLOAD_PIECE_MB / HCNV_UPLOAD_DEPTH override.
    size_t piece_bytes = (size_t)64 << 20;
    int depth = 2;
    double taper = 0.85;                          // hcnv_genome_run's task sizes: group k ~ taper^k of the genome (HCNV_GENOME_TAPER)
    std::vector<cudaEvent_t> ring;                // one event per piece in flight
    size_t ring_pos = 0;
    std::deque<UploadJob> jobs;
    std::condition_variable cv_up;
    std::thread uploader;
    bool jobs_closed = false;
    std::string err;
    std::mutex mu;
    // ---- the sample in flight (hcnv_genome_begin .. hcnv_genome_end)
    bool open = false, closed = false;
    hcnv_bin_params P{};
    float lambda1 = 0.f, lambda2 = 0.f, alpha = 0.5f;
    int n_contigs = 0;
    hcnv_genome_output out{};
    std::vector<long long> cap;
    std::vector<hcnv_contig_input> dev_in;
    long long task_rows_hint = 0;
    std::vector<std::deque<GenomeTask>> queue;   // one per worker: task k goes to worker k mod workers, so that from the second
                                                 // sample on every worker has seen its tasks' sizes (grow-only scratch: no cudaMalloc in the steady state)
    size_t n_tasks = 0;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    std::atomic<int> worst{HCNV_OK};
    bool trace = false;                           // HCNV_GENOME_TRACE=1: per-task wall times to stderr at hcnv_genome_end
    std::chrono::steady_clock::time_point t0;
    std::vector<std::string> trace_lines;
};

namespace {
enum { ST_BLOCKS = 0, ST_LONG, ST_SNPPOS, ST_SNPZYG, ST_OBS, ST_CBLOCKS, ST_ANCHORS, ST_COBS, ST_COBSANC, ST_WBLOCKS, ST_WLENS, ST_WOFFS, ST_N };

struct Field { const void* src; size_t bytes; };

void contig_fields(const hcnv_contig_input& c, Field f[ST_N]) {
    f[ST_BLOCKS]  = { c.blocks, sizeof(hcnv_block) * (size_t)c.n_blocks };
    f[ST_LONG]    = { c.long_blocks, sizeof(hcnv_long_block) * (size_t)c.n_long };
    f[ST_SNPPOS]  = { c.snp_pos1, 4 * (size_t)c.n_snp };
    f[ST_SNPZYG]  = { c.snp_zyg, (size_t)c.n_snp };
    f[ST_OBS]     = { c.obs, 4 * (size_t)c.n_obs };
    const bool wire = c.wblocks != nullptr && c.n_cblocks > 0;                   // the block stream in its wire form (hcnv_wblock)
    f[ST_CBLOCKS] = { c.cblocks, wire ? 0 : sizeof(hcnv_cblock) * (size_t)c.n_cblocks };
    f[ST_WBLOCKS] = { c.wblocks, wire ? (size_t)(c.n_cblocks / HCNV_CGROUP) * HCNV_WGROUP_BYTES : 0 };
    f[ST_WLENS]   = { c.wside, wire ? (size_t)std::max<int64_t>(c.n_wside, 0) : 0 };
    f[ST_WOFFS]   = { c.wside_offsets, wire ? 4 * (size_t)(c.n_cblocks / HCNV_CGROUP + 1) : 0 };
    f[ST_ANCHORS] = { c.anchors, 4 * (size_t)(c.n_cblocks / HCNV_CGROUP) };
    f[ST_COBS]    = { c.wobs ? (const void*)c.wobs : (const void*)c.cobs, (c.wobs ? sizeof(hcnv_wobs) : sizeof(hcnv_cobs)) * (size_t)c.n_cobs };
    f[ST_COBSANC] = { c.cobs_anchors, 4 * (size_t)(c.n_cobs / (c.wobs ? HCNV_WOGROUP : HCNV_OGROUP)) };
}

int genome_fail(hcnv_genome* g, int code, const std::string& msg) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->err.empty()) g->err = msg;
    return code;
}
}  // namespace

extern "C" {

int hcnv_genome_create(int device, int32_t workers, hcnv_genome** out) {
    if (!out) return HCNV_E_INVALID;
    *out = nullptr;
    hcnv_genome* g = new (std::nothrow) hcnv_genome();
    if (!g) return HCNV_E_NOMEM;
    g->device = device;
    const int nw = std::max(1, std::min<int>(workers > 0 ? workers : 4, 16));
    for (int w = 0; w < nw; ++w) {
        hcnv_ctx* c = nullptr;
        int rc = hcnv_ctx_create(device, &c);
        if (rc) { hcnv_genome_destroy(g); return rc; }
        g->ctx.push_back(c);
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&g->copy, cudaStreamNonBlocking) != cudaSuccess) { hcnv_genome_destroy(g); return HCNV_E_CUDA; }
    *out = g;
    return HCNV_OK;
}

void hcnv_genome_destroy(hcnv_genome* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->copy) cudaStreamSynchronize(g->copy);
    for (hcnv_ctx* c : g->ctx) hcnv_ctx_destroy(c);
    for (auto& b : g->stage) b.release();
    for (cudaEvent_t e : g->landed) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ring) if (e) cudaEventDestroy(e);
    if (g->copy) cudaStreamDestroy(g->copy);
    delete g;
}

const char* hcnv_genome_last_error(const hcnv_genome* g) { return g ? g->err.c_str() : "null genome"; }

int64_t hcnv_genome_kernel_launches(const hcnv_genome* g) {
    int64_t n = 0;
    if (g) for (const hcnv_ctx* c : g->ctx) n += c->launches;
    return n;
}

static void genome_uploader(hcnv_genome* g);

// one task = a few chromosomes end to end on worker w
static void genome_task(hcnv_genome* g, int w, const GenomeTask& task) {
    hcnv_ctx* ctx = g->ctx[(size_t)w];
    hcnv_genome_output* out = &g->out;
    auto fail = [&](int rc, const char* what) {
        genome_fail(g, rc, std::string(what) + ": " + hcnv_last_error(ctx));
        int expect = HCNV_OK; g->worst.compare_exchange_strong(expect, rc);
    };
    const std::vector<int>& gr = task.contigs;
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g->t0).count(); };
    const double t_start = g->trace ? now_ms() : 0.0;
    double t_wait = 0, t_bin = 0, t_text = 0, t_seg = 0;
    long long rows_cap = 0;
    for (int c : gr) rows_cap += g->cap[(size_t)c];
    const size_t C = (size_t)std::max<long long>(std::max<long long>(rows_cap, g->task_rows_hint), 1);
    // per-worker device scratch (grows to the largest task): BinReducer's columns, their float32 twins, the calls
    if (ctx->buf[G_BINS_I32].reserve(4 * 4 * C + 64) != cudaSuccess || ctx->buf[G_BINS_F32].reserve(4 * C + 64) != cudaSuccess ||
        ctx->buf[G_BINS_F64].reserve(8 * 7 * C + 64) != cudaSuccess || ctx->buf[G_M32].reserve(4 * 7 * C + 64) != cudaSuccess ||
        ctx->buf[G_SEG_F32].reserve(4 * C + 64) != cudaSuccess || ctx->buf[G_SEG_I32].reserve(4 * 2 * C + 64) != cudaSuccess) {
        cudaGetLastError(); ctx->err = "out of device memory"; fail(HCNV_E_NOMEM, "worker scratch"); return;
    }
    int32_t* bi = ctx->buf[G_BINS_I32].as<int32_t>();
    hcnv_bins bins{};
    bins.capacity = (int64_t)C;
    bins.bin = bi; bins.start = bi + C; bins.end = bi + 2 * C; bins.n = bi + 3 * C;
    bins.median = ctx->buf[G_BINS_F32].as<float>();
    bins.mse = ctx->buf[G_BINS_F64].as<double>(); bins.total = bins.mse + 6 * C;
    float* m32 = ctx->buf[G_M32].as<float>(); float* t32 = m32 + 6 * C;
    float* scaled = ctx->buf[G_SEG_F32].as<float>();
    int32_t* state = ctx->buf[G_SEG_I32].as<int32_t>(); int32_t* cn = state + C;
    // the task's records are on the device once its last chromosome's event has fired (uploads are in order)
    if (cudaStreamWaitEvent(ctx->stream, g->landed[(size_t)gr.back()], 0) != cudaSuccess) { ctx->err = "cudaStreamWaitEvent"; fail(HCNV_E_CUDA, "wait"); return; }
    if (g->trace) { cudaStreamSynchronize(ctx->stream); t_wait = now_ms(); }
    std::vector<hcnv_contig_input> in;
    for (int c : gr) in.push_back(g->dev_in[(size_t)c]);
    std::vector<int64_t> off(gr.size() + 1, 0);
    bins.contig_offset = off.data();
    int rc = hcnv_bin_contigs_impl(ctx, &g->P, (int32_t)gr.size(), in.data(), &bins);
    if (rc) { fail(rc, "hcnv_bin_contigs"); return; }
    const int64_t nrows = off[gr.size()];
    if (g->trace) t_bin = now_ms();
    if (nrows == 0) return;
    rc = hcnv_bins_to_hmm_input_impl(ctx, nrows, bins.mse, bins.total, m32, t32);
    if (rc) { fail(rc, "hcnv_bins_to_hmm_input"); return; }
    if (g->trace) t_text = now_ms();
    std::vector<hcnv_segment_problem> probs;
    std::vector<int> live;
    for (size_t j = 0; j < gr.size(); ++j) {
        const int64_t a = off[j], nb = off[j + 1] - off[j];
        if (nb <= 0) continue;
        hcnv_segment_problem q{};
        q.n_markers = nb; q.median = bins.median + a; q.mse = m32 + 6 * a; q.total = t32 + a; q.n = bins.n + a;
        q.lambda1 = g->lambda1; q.lambda2 = g->lambda2; q.scaled = scaled + a; q.state = state + a; q.cn = cn + a;
        probs.push_back(q); live.push_back((int)j);
    }
    rc = hcnv_segment_batch_impl(ctx, g->alpha, (int32_t)probs.size(), probs.data(), 0);
    if (rc < 0 && rc != HCNV_E_NO_MINIMUM) { fail(rc, "hcnv_segment_batch"); return; }
    if (g->trace) t_seg = now_ms();
    // every chromosome's rows straight into the caller's arrays; two host round trips for the whole task
    std::vector<hcnv_results_download> dl(live.size());
    for (size_t k = 0; k < live.size(); ++k) {
        const size_t j = (size_t)live[k];
        const int c = gr[j];
        const int64_t a = off[j], nb = off[j + 1] - off[j], o = out->row_offset[c];
        hcnv_results_download& d = dl[k];
        d = hcnv_results_download{};
        d.n = nb; d.start = bins.start + a; d.end = bins.end + a; d.scaled = scaled + a; d.state = state + a; d.cn = cn + a; d.mse = m32 + 6 * a;
        d.h_scaled = out->scaled + o; d.h_state = out->state + o;
        d.h_cn = out->cn ? out->cn + o : nullptr;
        d.h_mse_rows = out->mse_rows + o; d.h_mse_vals = out->mse_vals + 6 * o; d.mse_capacity = nb;
        if (out->start) { d.h_start = out->start + o; d.h_end = out->end + o; }
        else {
            d.bin_width = g->P.bin_width;
            d.h_run_rows = out->run_rows + o; d.h_run_bins = out->run_bins + o; d.run_capacity = nb;
            d.h_se_rows = out->se_rows + o; d.h_se_vals = out->se_vals + 2 * o; d.se_capacity = nb;
        }
    }
    rc = hcnv_download_results_many_impl(ctx, (int32_t)dl.size(), dl.data());
    if (rc) { fail(rc, "hcnv_download_results"); return; }
    for (size_t k = 0; k < live.size(); ++k) {
        const int c = gr[(size_t)live[k]];
        out->n_bins[c] = dl[k].n; out->n_mse_rows[c] = dl[k].n_mse_rows;
        if (!out->start) { out->n_runs[c] = dl[k].n_runs; out->n_se_rows[c] = dl[k].n_se_rows; }
        const hcnv_segment_problem& q = probs[k];
        if (out->status) out->status[c] = q.status;
        if (out->mean_coverage) out->mean_coverage[c] = q.mean_coverage;
        if (out->final_loss) out->final_loss[c] = q.final_loss;
        if (out->lambda1_used) out->lambda1_used[c] = q.lambda1_used;
        if (out->lambda2_used) out->lambda2_used[c] = q.lambda2_used;
        if (q.status < 0) { int expect = HCNV_OK; g->worst.compare_exchange_strong(expect, q.status); }
    }
    if (g->trace) {
        char line[256];
        snprintf(line, sizeof line, "task w%d contigs=%zu rows=%lld start=%.2f landed=%.2f bin=%.2f text=%.2f seg=%.2f d2h=%.2f", w, gr.size(), (long long)nrows, t_start, t_wait, t_bin, t_text, t_seg, now_ms());
        std::lock_guard<std::mutex> lk(g->mu);
        g->trace_lines.push_back(line);
    }
}

static void genome_worker(hcnv_genome* g, int w) {
    cudaSetDevice(g->device);
    for (;;) {
        GenomeTask task;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            std::deque<GenomeTask>& q = g->queue[(size_t)w];
            g->cv.wait(lk, [&] { return !q.empty() || g->closed; });
            if (q.empty()) return;
            task = std::move(q.front());
            q.pop_front();
        }
        if (g->worst.load() == HCNV_OK) genome_task(g, w, task);
    }
}

// The sample as a stream of chromosomes: begin (the contig lengths fix the layout of the output), push groups of chromosomes as
// the packer finishes them (each push uploads its records and hands the group to a worker; it returns at once), end (waits).
// This is the per-key push of the reference's map side (SAMRecordWindowMapper.map emits per record, the shuffle groups by
// chromosome, Mappers.java:167-272) at the granularity the GPU wants: the host never has to hold more than the chromosomes
// it is still packing, and the GPU work hides behind the packer.
int hcnv_genome_begin(hcnv_genome* g, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                      int32_t n_contigs, const int32_t* lengths, int64_t max_task_rows, hcnv_genome_output* out) {
    if (!g || n_contigs <= 0 || !lengths || !out) return HCNV_E_INVALID;
    if (g->open) return genome_fail(g, HCNV_E_INVALID, "hcnv_genome_begin: a sample is already in flight");
    if (!out->scaled || !out->state || !out->mse_rows || !out->mse_vals || !out->row_offset || !out->n_bins || !out->n_mse_rows)
        return genome_fail(g, HCNV_E_INVALID, "null output array");
    if (out->start ? !out->end : (!out->run_rows || !out->run_bins || !out->se_rows || !out->se_vals || !out->n_runs || !out->n_se_rows))
        return genome_fail(g, HCNV_E_INVALID, "null start / end arrays (dense: start, end; implicit: run_rows, run_bins, se_rows, se_vals, n_runs, n_se_rows)");
    if (cudaSetDevice(g->device) != cudaSuccess) return genome_fail(g, HCNV_E_CUDA, "cudaSetDevice failed");
    g->err.clear();
    if (params) g->P = *params; else hcnv_bin_params_default(&g->P);
    const int W = g->P.bin_width;
    if (W < 1) return genome_fail(g, HCNV_E_INVALID, "bin_width");
    // rows of contig c start at the sum of the capacities (length / W + 1: what BinReducer can reach) of the contigs before it
    g->cap.assign((size_t)n_contigs, 0);
    long long total_cap = 0;
    for (int c = 0; c < n_contigs; ++c) {
        if (lengths[c] < 0) return genome_fail(g, HCNV_E_INVALID, "negative contig length");
        g->cap[(size_t)c] = lengths[c] / W + 1;
        out->row_offset[c] = total_cap; out->n_bins[c] = 0; out->n_mse_rows[c] = 0;
        if (!out->start) { out->n_runs[c] = 0; out->n_se_rows[c] = 0; }
        if (out->status) out->status[c] = HCNV_OK;
        if (out->mean_coverage) out->mean_coverage[c] = 0.f;
        if (out->final_loss) out->final_loss[c] = 0.f;
        if (out->lambda1_used) out->lambda1_used[c] = lambda1;
        if (out->lambda2_used) out->lambda2_used[c] = lambda2;
        total_cap += g->cap[(size_t)c];
    }
    if (total_cap > out->capacity) return genome_fail(g, HCNV_E_CAPACITY, "output capacity below the sum of length / bin_width + 1");
    if (g->stage.size() < (size_t)n_contigs * ST_N) g->stage.resize((size_t)n_contigs * ST_N);
    while (g->landed.size() < (size_t)n_contigs) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return genome_fail(g, HCNV_E_CUDA, "cudaEventCreate failed");
        g->landed.push_back(e);
    }
    g->dev_in.assign((size_t)n_contigs, hcnv_contig_input{});
    g->lambda1 = lambda1; g->lambda2 = lambda2; g->alpha = alpha; g->n_contigs = n_contigs; g->out = *out;
    g->task_rows_hint = max_task_rows;
    while (g->ring.size() < 32) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return genome_fail(g, HCNV_E_CUDA, "cudaEventCreate failed");
        g->ring.push_back(e);
    }
    g->ring_pos = 0;
    if (const char* e = getenv("HCNV_UPLOAD_PIECE_MB")) g->piece_bytes = (size_t)std::max(1, atoi(e)) << 20;
    if (const char* e = getenv("HCNV_UPLOAD_DEPTH")) g->depth = std::max(1, std::min(16, atoi(e)));
    g->worst = HCNV_OK; g->closed = false; g->queue.assign(g->ctx.size(), {}); g->n_tasks = 0; g->jobs.clear(); g->jobs_closed = false;
    g->trace = getenv("HCNV_GENOME_TRACE") != nullptr; g->t0 = std::chrono::steady_clock::now(); g->trace_lines.clear();
    g->uploader = std::thread(genome_uploader, g);
    for (size_t w = 0; w < g->ctx.size(); ++w) g->workers.emplace_back(genome_worker, g, (int)w);
    g->open = true;
    return HCNV_OK;
}

// the uploader thread: the pushed groups in order, piece by piece, never more than g->depth pieces queued on the copy stream
static void genome_uploader(hcnv_genome* g) {
    cudaSetDevice(g->device);
    auto fail = [&](int rc, const char* what) {
        genome_fail(g, rc, what);
        int expect = HCNV_OK; g->worst.compare_exchange_strong(expect, rc);
    };
    std::deque<std::pair<size_t, size_t>> fly;      // (ring slot, bytes) of the pieces queued on the copy stream, oldest first
    size_t fly_bytes = 0;
    const size_t budget = g->piece_bytes * (size_t)g->depth;
    for (;;) {
        UploadJob job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_up.wait(lk, [g] { return !g->jobs.empty() || g->jobs_closed; });
            if (g->jobs.empty()) break;
            job = std::move(g->jobs.front());
            g->jobs.pop_front();
        }
        if (g->worst.load() != HCNV_OK) continue;
        GenomeTask task;
        bool ok = true;
        for (size_t k = 0; k < job.contigs.size() && ok; ++k) {
            const int c = job.contigs[k];
            Field f[ST_N];
            contig_fields(job.host[k], f);
            void* d[ST_N];
            for (int q = 0; q < ST_N && ok; ++q) {
                DevBuf& b = g->stage[(size_t)c * ST_N + q];
                d[q] = nullptr;
                if (!f[q].bytes) continue;
                if (b.reserve(f[q].bytes) != cudaSuccess) { cudaGetLastError(); fail(HCNV_E_NOMEM, "staging buffer"); ok = false; break; }
                d[q] = b.p;
                for (size_t o = 0; o < f[q].bytes && ok; o += g->piece_bytes) {
                    const size_t nb = std::min(g->piece_bytes, f[q].bytes - o);
                    // wait until the pieces in flight leave room: at most depth x piece_bytes queued (small arrays ride along, up to the ring's size)
                    while (!fly.empty() && (fly_bytes + nb > budget || fly.size() >= g->ring.size())) {
                        cudaEventSynchronize(g->ring[fly.front().first]);
                        fly_bytes -= fly.front().second;
                        fly.pop_front();
                    }
                    const size_t slot = g->ring_pos;
                    g->ring_pos = (g->ring_pos + 1) % g->ring.size();
                    if (cudaMemcpyAsync((char*)b.p + o, (const char*)f[q].src + o, nb, cudaMemcpyHostToDevice, g->copy) != cudaSuccess ||
                        cudaEventRecord(g->ring[slot], g->copy) != cudaSuccess) { fail(HCNV_E_CUDA, "upload failed"); ok = false; break; }
                    fly.emplace_back(slot, nb);
                    fly_bytes += nb;
                }
            }
            if (!ok) break;
            if (cudaEventRecord(g->landed[(size_t)c], g->copy) != cudaSuccess) { fail(HCNV_E_CUDA, "cudaEventRecord failed"); break; }
            hcnv_contig_input& di = g->dev_in[(size_t)c];
            di = job.host[k];
            di.blocks = (const hcnv_block*)d[ST_BLOCKS]; di.long_blocks = (const hcnv_long_block*)d[ST_LONG];
            di.snp_pos1 = (const int32_t*)d[ST_SNPPOS]; di.snp_zyg = (const int8_t*)d[ST_SNPZYG]; di.obs = (const uint32_t*)d[ST_OBS];
            di.cblocks = (const hcnv_cblock*)d[ST_CBLOCKS]; di.anchors = (const int32_t*)d[ST_ANCHORS];
            di.cobs = job.host[k].wobs ? nullptr : (const hcnv_cobs*)d[ST_COBS]; di.wobs = job.host[k].wobs ? (const hcnv_wobs*)d[ST_COBS] : nullptr;
            di.cobs_anchors = (const int32_t*)d[ST_COBSANC];
            di.wblocks = (const hcnv_wblock*)d[ST_WBLOCKS]; di.wside = (const uint8_t*)d[ST_WLENS]; di.wside_offsets = (const int32_t*)d[ST_WOFFS];
            task.contigs.push_back(c);
        }
        if (!ok || task.contigs.empty()) continue;
        { std::lock_guard<std::mutex> lk(g->mu); g->queue[g->n_tasks++ % g->queue.size()].push_back(std::move(task)); }     // (its last chromosome's event is recorded: a worker may wait on it)
        g->cv.notify_all();
        if (g->trace) fprintf(stderr, "group of %zu contigs queued on the copy stream at %.2f ms\n", job.contigs.size(),
                              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g->t0).count());
    }
    { std::lock_guard<std::mutex> lk(g->mu); g->closed = true; }
    g->cv.notify_all();
}

int hcnv_genome_push(hcnv_genome* g, int32_t n, const int32_t* contig_index, const hcnv_contig_input* contigs) {
    if (!g || !g->open || n <= 0 || !contig_index || !contigs) return HCNV_E_INVALID;
    UploadJob job;
    for (int k = 0; k < n; ++k) {
        const int c = contig_index[k];
        if (c < 0 || c >= g->n_contigs) return genome_fail(g, HCNV_E_INVALID, "contig index");
        if (contigs[k].region_first_bin || contigs[k].region_n_bins) return genome_fail(g, HCNV_E_INVALID, "hcnv_genome_push takes whole contigs");
        if ((long long)contigs[k].length / g->P.bin_width + 1 != g->cap[(size_t)c]) return genome_fail(g, HCNV_E_INVALID, "contig length differs from hcnv_genome_begin's");
        Field f[ST_N];
        contig_fields(contigs[k], f);
        for (int q = 0; q < ST_N; ++q) if (f[q].bytes && !f[q].src) return genome_fail(g, HCNV_E_INVALID, "null input array with non-zero count");
        job.contigs.push_back(c); job.host.push_back(contigs[k]);
    }
    { std::lock_guard<std::mutex> lk(g->mu); g->jobs.push_back(std::move(job)); }
    g->cv_up.notify_one();
    return HCNV_OK;
}

// Waits for every pushed group; the caller's host buffers of a group must stay valid until then (the uploads are asynchronous).
int hcnv_genome_end(hcnv_genome* g) {
    if (!g || !g->open) return HCNV_E_INVALID;
    { std::lock_guard<std::mutex> lk(g->mu); g->jobs_closed = true; }
    g->cv_up.notify_all();
    g->uploader.join();                           // (it closes the workers' queue behind the last group)
    for (auto& t : g->workers) t.join();
    g->workers.clear();
    g->open = false;
    if (g->trace) {
        fprintf(stderr, "hcnv_genome_end at %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g->t0).count());
        for (auto& l : g->trace_lines) fprintf(stderr, "  %s\n", l.c_str());
    }
    cudaSetDevice(g->device);
    if (g->worst.load() != HCNV_OK) cudaStreamSynchronize(g->copy);      // a failed sample still drains the copy stream before the caller reuses its arrays
    return g->worst.load();
}

// The same chain for records that are already on the device, on the caller's context: what a task does between its upload and
// its download, as one call (no interpreter between the three stages: on a 3 ms step the gaps are a tenth of it).
int hcnv_sample_run_device(hcnv_ctx* ctx, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                           int32_t n_contigs, const hcnv_contig_input* contigs, hcnv_bins* bins,
                           float* mse32, float* total32, float* scaled, int32_t* state, int32_t* cn, hcnv_segment_problem* probs) {
    if (!ctx) return HCNV_E_INVALID;
    if (n_contigs <= 0 || !contigs || !bins || !mse32 || !total32 || !scaled || !state || !cn || !probs || !bins->contig_offset)
        return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    int rc = hcnv_bin_contigs_impl(ctx, params, n_contigs, contigs, bins);
    if (rc) return rc;
    const int64_t* off = bins->contig_offset;
    const int64_t nrows = off[n_contigs];
    for (int c = 0; c < n_contigs; ++c) {
        hcnv_segment_problem q{};
        const int64_t a = off[c];
        q.n_markers = off[c + 1] - a; q.median = bins->median + a; q.mse = mse32 + 6 * a; q.total = total32 + a; q.n = bins->n + a;
        q.lambda1 = lambda1; q.lambda2 = lambda2; q.lambda1_used = lambda1; q.lambda2_used = lambda2;
        q.scaled = scaled + a; q.state = state + a; q.cn = cn + a;
        probs[c] = q;
    }
    if (nrows == 0) return HCNV_OK;
    rc = hcnv_bins_to_hmm_input_impl(ctx, nrows, bins->mse, bins->total, mse32, total32);
    if (rc) return rc;
    std::vector<hcnv_segment_problem> live;
    std::vector<int> which;
    for (int c = 0; c < n_contigs; ++c) if (probs[c].n_markers > 0) { live.push_back(probs[c]); which.push_back(c); }
    rc = hcnv_segment_batch_impl(ctx, alpha, (int32_t)live.size(), live.data(), 0);
    for (size_t k = 0; k < live.size(); ++k) probs[which[k]] = live[k];
    return rc;
}

int hcnv_genome_run(hcnv_genome* g, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                    int32_t n_contigs, const hcnv_contig_input* contigs, int32_t n_groups, hcnv_genome_output* out) {
    if (!g || n_contigs <= 0 || !contigs || !out) return HCNV_E_INVALID;
    // ---- tasks: the chromosomes longest first, grouped into about n_groups pieces of equal length
    std::vector<int32_t> lengths((size_t)n_contigs);
    long long total_len = 0;
    for (int c = 0; c < n_contigs; ++c) { lengths[(size_t)c] = contigs[c].length; total_len += std::max(0, contigs[c].length); }
    std::vector<int> order((size_t)n_contigs);
    for (int c = 0; c < n_contigs; ++c) order[(size_t)c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return contigs[a].length > contigs[b].length; });
    const int NG = std::max(1, n_groups > 0 ? n_groups : 8);
    const int W = params && params->bin_width > 0 ? params->bin_width : 100;
    std::vector<std::vector<int>> groups;
    {
        // group k's share of the genome ~ taper^k: later groups smaller (taper < 1), so that what is left to do when the upload ends
        // is little, while the early groups stay large (a task's chain kernels cost their longest chromosome whatever else rides along)
        double taper = g->taper;
        if (const char* e = getenv("HCNV_GENOME_TAPER")) taper = std::max(0.3, std::min(1.0, atof(e)));
        std::vector<double> upto((size_t)NG);
        { double s = 0, w = 1; for (int k = 0; k < NG; ++k) { s += w; upto[(size_t)k] = s; w *= taper; } for (double& u : upto) u /= s; }
        std::vector<int> cur; long long acc = 0;
        for (int c : order) {
            cur.push_back(c); acc += contigs[c].length;
            if ((double)acc >= (double)std::max<long long>(total_len, 1) * upto[std::min(groups.size(), (size_t)NG - 1)] - 0.5) { groups.push_back(cur); cur.clear(); }
        }
        if (!cur.empty()) groups.push_back(cur);
        // The LAST task is the only one nothing hides (it starts when the upload ends).  Halving the last group to make it small
        // does not pay: measured 32.5 / 32.5 / 33.3 / 37.4 ms per genome with 0 / 1 / 2 / 3 halvings -- even the smallest
        // chromosomes' task takes ~3 ms of launch and round-trip latency, and more tasks contend for the workers.  Off by default.
        int splits = 0;
        if (const char* e = getenv("HCNV_GENOME_TAIL_SPLITS")) splits = std::max(0, std::min(6, atoi(e)));
        for (int k = 0; k < splits && groups.back().size() >= 2; ++k) {
            std::vector<int> last = groups.back();
            groups.pop_back();
            long long len = 0, half = 0;
            for (int c : last) len += contigs[c].length;
            size_t cutp = 0;
            while (cutp + 1 < last.size() && (half + contigs[last[cutp]].length) * 2 <= len + contigs[last[cutp]].length) half += contigs[last[cutp++]].length;
            cutp = std::max<size_t>(1, std::min(cutp, last.size() - 1));
            groups.emplace_back(last.begin(), last.begin() + (long)cutp);
            groups.emplace_back(last.begin() + (long)cutp, last.end());
        }
    }
    long long max_group_cap = 1;
    for (auto& gr : groups) { long long s = 0; for (int c : gr) s += contigs[c].length / W + 1; max_group_cap = std::max(max_group_cap, s); }
    int rc = hcnv_genome_begin(g, params, lambda1, lambda2, alpha, n_contigs, lengths.data(), max_group_cap, out);
    if (rc) return rc;
    for (auto& gr : groups) {
        std::vector<hcnv_contig_input> in;
        for (int c : gr) in.push_back(contigs[c]);
        rc = hcnv_genome_push(g, (int32_t)gr.size(), gr.data(), in.data());
        if (rc) break;
    }
    const int rc2 = hcnv_genome_end(g);
    return rc ? rc : rc2;
}

}  // extern "C"
