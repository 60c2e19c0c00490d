// extern "C" boundary of libautocycler_gpu.so (include/autocycler_gpu.h).  No exception leaves this file.
#include "../../include/autocycler_gpu.h"

#include <sched.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "backend.h"
#include "host_graph.h"
#include "host_io.h"
#include <immintrin.h>
#include <functional>
#include <thread>
#include "pipeline.h"

namespace {
thread_local std::string g_error;

struct NoDevice { std::string msg; };

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct PinnedBytes {     // host staging buffer for the concatenated padded strands (pinned, so the H2D copy is a DMA)
    uint8_t* p = nullptr; size_t size = 0, cap = 0;
    void append(const uint8_t* src, size_t n) {
        if (size + n > cap) {
            size_t ncap = std::max<size_t>(cap * 2, size + n + (1 << 20));
            uint8_t* q = (uint8_t*)ac_host_alloc(ncap);
            if (size) memcpy(q, p, size);
            ac_host_free(p); p = q; cap = ncap;
        }
        memcpy(p + size, src, n); size += n;
    }
    void clear() { size = 0; }
    ~PinnedBytes() { ac_host_free(p); }
};
}  // namespace

struct ac_handle {
    ac_config cfg{};
    mutable std::string err;
    std::vector<HostSeq> seqs;
    std::vector<SeqInfo> infos;
    PinnedBytes ascii;
    LoadedInput loaded;                    // only when filled by ac_load_sequences (keeps the YAML details)
    std::unique_ptr<DevicePipeline> pipe;                   // the pipeline that finishes the graph (devices[0])
    std::vector<std::unique_ptr<DevicePipeline>> peers;     // ac_config.n_devices > 1: the pipelines of devices[1..]
    const char* path_lines = nullptr; uint64_t path_lines_len = 0;     // ac_path_lines_render: this rank's P lines (pinned, owned by the pipeline)
    std::vector<int32_t> devices;
    PipelineResult res;
    HostGraph graph;
    std::string gfa;
    const char* gfa_ptr = nullptr; uint64_t gfa_len = 0;     // the finished file: h->gfa, or the pinned buffer the device wrote the S and L lines into
    bool device_text_ok = false;                             // the device-written text describes the graph as it is now
    bool uploaded = false, built = false, gfa_ready = false;
    bool fused = false;                                      // built by ac_compress: simplified on the device; the host graph is adopted on first use
    bool graph_ready = false;                                // h->graph describes the current graph
    ac_timings t{};
    uint64_t links_now = 0;
};

static int set_error(const ac_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    g_error = msg;
    return code;
}

static int ok(const ac_handle* h) {      // a call that succeeded leaves no stale message behind (ac_last_error)
    if (h) h->err.clear();
    g_error.clear();
    return AC_OK;
}

#define AC_GUARD_BEGIN try {
#define AC_GUARD_END(h) } catch (const InputError& e) { return set_error(h, AC_EINPUT, e.msg); } \
    catch (const NoDevice& e) { return set_error(h, AC_ENODEVICE, e.msg); } \
    catch (const std::bad_alloc&) { return set_error(h, AC_ERANGE, "out of host memory"); } \
    catch (const std::exception& e) { std::string m = e.what(); \
        int code = m.find("no CUDA device") != std::string::npos ? AC_ENODEVICE : (m.find("cuda") != std::string::npos ? AC_ECUDA : AC_EINVAL); \
        return set_error(h, code, m); } \
    catch (...) { return set_error(h, AC_EINVAL, "unknown error"); }

// ac_config.n_devices > 1: the stages of SURVEY.md 8e driven from this one process.  The assemblies are sharded by file (contiguous
// blocks in the order of ac_add_sequence = sorted file order); every device builds the table of its block, then folds the other
// devices' deduplicated entries in by reading them where they lie (peer memory over NVLink: the merge kernel is the collective),
// computes adjacency and its own occurrences; devices[0] reads all occurrences the same way and finishes the graph.
static void build_on_devices(ac_handle* h, bool fused) {
    std::vector<DevicePipeline*> pipes{h->pipe.get()};
    for (auto& p : h->peers) pipes.push_back(p.get());
    const size_t N = pipes.size();
    std::vector<uint32_t> first_of_file;                     // index of the first sequence of every file
    for (size_t i = 0; i < h->seqs.size(); ++i) if (i == 0 || h->seqs[i].filename != h->seqs[i - 1].filename) first_of_file.push_back((uint32_t)i);
    first_of_file.push_back((uint32_t)h->seqs.size());
    const size_t F = first_of_file.size() - 1;
    auto on_all = [&](const std::function<void(size_t)>& work) {
        std::vector<std::thread> th; std::vector<std::exception_ptr> err(N);
        for (size_t d = 1; d < N; ++d) th.emplace_back([&, d] { try { work(d); } catch (...) { err[d] = std::current_exception(); } });
        try { work(0); } catch (...) { err[0] = std::current_exception(); }
        for (auto& t : th) t.join();
        for (auto& e : err) if (e) std::rethrow_exception(e);
    };
    std::vector<const void*> entries(N), runs(N); std::vector<uint64_t> n_entries(N), n_runs(N);
    on_all([&](size_t d) {
        pipes[d]->build_local(first_of_file[F * d / N], first_of_file[F * (d + 1) / N], true);
        entries[d] = pipes[d]->export_entries_own(&n_entries[d]);
    });
    on_all([&](size_t d) {
        for (size_t r = 0; r < N; ++r) if (r != d && n_entries[r]) pipes[d]->merge_entries(entries[r], n_entries[r]);
        pipes[d]->runs_local();
        runs[d] = pipes[d]->export_runs_own(&n_runs[d]);
    });
    pipes[0]->import_runs_from(runs.data(), n_runs.data(), (uint32_t)N);
    pipes[0]->finish(h->res, h->cfg.keep_positions != 0, fused);
}


extern "C" {

const char* ac_last_error(const ac_handle* h) { return h ? h->err.c_str() : g_error.c_str(); }

const char* ac_version(void) {
#ifdef AC_EMULATE
    return "autocycler_gpu 0.1 (host emulation build: tests only)";
#else
    return "autocycler_gpu 0.1 (sm_100a)";
#endif
}

int ac_create(ac_handle** out, const ac_config* cfg) {
    ac_handle* h = nullptr;
    AC_GUARD_BEGIN
    if (!out || !cfg) return set_error(nullptr, AC_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->k < 3 || (cfg->k & 1) == 0) return set_error(nullptr, AC_EINVAL, "--kmer must be odd");          // compress.rs:58
    if (cfg->k > AC_MAX_K)
        return set_error(nullptr, AC_EINVAL, "k-mer sizes above " + std::to_string(AC_MAX_K) + " are not supported by the GPU path (there is no CPU fallback)");
    h = new ac_handle;
    h->cfg = *cfg;
    if (cfg->n_devices > 1) {
        if (!cfg->devices) { delete h; return set_error(nullptr, AC_EINVAL, "ac_config.devices is null"); }
        if (cfg->n_devices > 16) { delete h; return set_error(nullptr, AC_EINVAL, "at most 16 devices"); }
        h->devices.assign(cfg->devices, cfg->devices + cfg->n_devices);
        for (size_t a = 0; a < h->devices.size(); ++a) for (size_t b = 0; b < a; ++b) if (h->devices[a] == h->devices[b]) { delete h; return set_error(nullptr, AC_EINVAL, "ac_config.devices names a device twice"); }
        h->cfg.device = h->devices[0]; h->cfg.stream = nullptr; h->cfg.devices = nullptr;
        h->pipe.reset(new DevicePipeline(h->devices[0], nullptr));
        for (size_t d = 1; d < h->devices.size(); ++d) h->peers.emplace_back(new DevicePipeline(h->devices[d], nullptr));
        DevicePipeline::enable_peer_access(h->devices.data(), (int)h->devices.size());
    } else {
        h->cfg.n_devices = 1; h->cfg.devices = nullptr;
        h->pipe.reset(new DevicePipeline(cfg->device, cfg->stream));
    }
    *out = h;
    return ok(h);
    AC_GUARD_END(((delete h), (ac_handle*)nullptr))
}

void ac_destroy(ac_handle* h) { delete h; }

int ac_clear_sequences(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    h->seqs.clear(); h->infos.clear(); h->ascii.clear(); h->loaded = LoadedInput();
    h->uploaded = h->built = h->gfa_ready = h->fused = h->graph_ready = false;
    return ok(h);
}

int ac_add_sequence(ac_handle* h, uint16_t seq_id, const uint8_t* fwd, uint64_t n, const char* filename, const char* header) {
    if (!h || !fwd) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    const uint32_t k = h->cfg.k, half = k / 2;
    if (n < 2ull * k - 1) return set_error(h, AC_EINVAL, "padded sequence shorter than 2k-1 (contigs shorter than k are skipped by the caller, compress.rs:109)");
    if (seq_id == 0 || seq_id > 32767) return set_error(h, AC_EINVAL, "sequence id must be in 1..32767 (position.rs:21)");
    uint64_t lead = 0, trail = 0;
    while (lead < n && fwd[lead] == '.') ++lead;
    while (trail < n && fwd[n - 1 - trail] == '.') ++trail;
    if (lead > half || trail > half) return set_error(h, AC_EINVAL, "more than k/2 padding dots at a sequence end");
    for (uint64_t i = lead; i < n - trail; ++i) {
        const uint8_t c = fwd[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return set_error(h, AC_EINVAL, "sequence bytes must be ACGT with dots only at the ends");
    }
    if (n - (k - 1) > 0xFFFFFFFFull) return set_error(h, AC_ERANGE, "contig longer than Position.pos allows (position.rs:20)");
    HostSeq s; s.id = seq_id; s.filename = filename ? filename : ""; s.contig_header = header ? header : "";
    s.length = n - (k - 1); s.start = h->ascii.size;
    SeqInfo info{}; info.start = s.start; info.len = (uint32_t)s.length; info.lead = (uint16_t)lead; info.trail = (uint16_t)trail; info.id = seq_id;
    h->ascii.append(fwd, n);
    h->seqs.push_back(std::move(s)); h->infos.push_back(info);
    h->uploaded = h->built = h->gfa_ready = false;
    return ok(h);
    AC_GUARD_END(h)
}

static int upload_block(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (h->seqs.empty()) return set_error(h, AC_EINPUT, "no sequences found in input assemblies");
    if (h->infos.size() != h->seqs.size()) return set_error(h, AC_EINVAL, "this handle holds a loaded graph: add sequences (after ac_clear_sequences) before ac_upload");
    {   // what save_gfa prints around every path's unitig list (unitig_graph.rs:352-360), for the device writer of the P lines
        std::string blob; std::vector<uint32_t> pre, suf;
        for (const HostSeq& q : h->seqs) {
            const std::string a = "P\t" + std::to_string(q.id) + "\t";
            std::string b = "\t*\tLN:i:" + std::to_string(q.length) + "\tFN:Z:" + q.filename + "\tHD:Z:" + q.contig_header;
            if (q.cluster > 0) b += "\tCL:i:" + std::to_string(q.cluster);
            b += "\n";
            blob += a; blob += b; pre.push_back((uint32_t)a.size()); suf.push_back((uint32_t)b.size());
        }
        h->pipe->set_path_line_texts(blob.data(), pre.data(), suf.data(), (uint32_t)h->seqs.size());
    }
    if (!h->peers.empty() && !(seq_lo == 0 && seq_hi >= h->infos.size())) return set_error(h, AC_EINVAL, "ac_upload_shard is for one device per process");
    h->pipe->upload(h->ascii.p, h->ascii.size, h->infos.data(), (uint32_t)h->infos.size(), h->cfg.k, seq_lo, seq_hi);
    for (auto& peer : h->peers) peer->upload(h->ascii.p, h->ascii.size, h->infos.data(), (uint32_t)h->infos.size(), h->cfg.k);      // every device holds every sequence (end k-mers of foreign occurrences are read from them)
    h->uploaded = true; h->built = h->gfa_ready = h->fused = h->graph_ready = false;
    return ok(h);
    AC_GUARD_END(h)
}
int ac_upload(ac_handle* h) { return upload_block(h, 0, 0xFFFFFFFFu); }
int ac_upload_shard(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi) { return upload_block(h, seq_lo, seq_hi); }
int ac_strand_block(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi, void** dev_ptr, uint64_t* n_bytes) {
    if (!h || !dev_ptr || !n_bytes) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (!h->uploaded) return set_error(h, AC_EINVAL, "ac_upload_shard must precede ac_strand_block");
    *dev_ptr = h->pipe->strand_block(seq_lo, seq_hi, n_bytes);
    return ok(h);
    AC_GUARD_END(h)
}

// The host stages edit the pinned result buffers in place, so after a graph has been processed their lines sit (dirty) in the
// CPU caches, and the next build's device->host copies onto the same lines run at a fraction of the link rate (measured:
// 56 GB/s onto untouched pinned memory, 33 GB/s and less onto lines the CPU wrote; profiles/r1k_pcie_probe2.log).  When a handle
// is reused, the previous result is therefore flushed from the caches while the GPU is busy building the next one.
namespace {
bool cpu_has_clflushopt() { unsigned a, b, c2, d; __asm__("cpuid" : "=a"(a), "=b"(b), "=c"(c2), "=d"(d) : "a"(7), "c"(0)); return (b >> 23) & 1; }
__attribute__((target("clflushopt"))) void flush_lines_opt(const char* p, size_t n) { for (size_t i = 0; i < n; i += 64) _mm_clflushopt((void*)(p + i)); _mm_sfence(); }
void flush_lines(const char* p, size_t n) {
    static const bool opt = cpu_has_clflushopt();
    if (opt) { flush_lines_opt(p, n); return; }
    for (size_t i = 0; i < n; i += 64) _mm_clflush(p + i);
    _mm_sfence();
}
struct ResultFlusher {
    std::vector<std::thread> threads;
    void start(const PipelineResult& r) {
        static const bool off = getenv("AC_NO_RESULT_FLUSH") != nullptr;
        if (off || !r.rec) return;
        const size_t U = r.n_unitigs, strands = 2 * U + 1;
        std::vector<std::pair<const char*, size_t>> ranges = {
            {(const char*)r.rec, U * sizeof(UnitigRec)}, {(const char*)r.depth, U * 4}, {(const char*)r.order, U * 4}, {r.arena, (size_t)r.arena_cap},
            {(const char*)r.next_off, strands * 4}, {(const char*)r.prev_off, strands * 4}, {(const char*)r.next, (size_t)r.n_links * 4}, {(const char*)r.prev, (size_t)r.n_links * 4},
            {(const char*)r.path, (size_t)r.n_runs * 4}, {(const char*)r.path_off, ((size_t)r.n_seqs + 1) * 8}, {(const char*)r.cands, (size_t)r.n_cands * sizeof(ExpandCandidate)},
            {(const char*)r.deps, U * sizeof(ExpandDeps)}, {(const char*)r.spec_len, (size_t)r.n_cands * 4}, {(const char*)r.fixed_start, 2 * U}};
        size_t total = 0; for (auto& x : ranges) if (x.first) total += x.second;
        const size_t T = std::max<size_t>(1, std::min<size_t>(8, total >> 20));
        for (size_t t = 0; t < T; ++t)
            threads.emplace_back([ranges, t, T] {
                for (auto& x : ranges) {
                    if (!x.first || !x.second) continue;
                    const size_t lines = (x.second + 63) / 64, a = lines * t / T, b = lines * (t + 1) / T;
                    if (b > a) flush_lines(x.first + a * 64, (b - a) * 64);
                }
            });
    }
    void join() { for (auto& th : threads) th.join(); threads.clear(); }
    ~ResultFlusher() { join(); }
};
struct CallbackScope {      // DevicePipeline::before_results holds a reference to a stack object: never let it outlive the call
    DevicePipeline* pipe;
    CallbackScope(DevicePipeline* p, std::function<void()> f) : pipe(p) { pipe->before_results = std::move(f); }
    ~CallbackScope() { pipe->before_results = nullptr; }
};
}  // namespace

static void record_timings(ac_handle* h) {
    const PipelineTimings& pt = h->res.t;
    ac_timings& t = h->t;
    t.h2d = pt.h2d; t.pack = pt.pack; t.insert = pt.insert; t.adjacency = pt.adjacency; t.boundaries = pt.boundaries; t.runs = pt.runs;
    t.unitigs = pt.unitigs; t.links = pt.links; t.seed_sort = pt.seed_sort; t.emit = pt.emit; t.d2h = pt.d2h; t.device_total = pt.total;
    t.sample = pt.sample; t.device_simplify = pt.simplify; t.device_gfa = pt.gfa;
    t.insert_kernel = pt.insert_kernel; t.reserved0 = 0;
    uint64_t windows = 0; for (auto& s : h->seqs) windows += s.length;
    t.insert_occurrences = windows; t.table_capacity = h->res.capacity; t.table_used = h->res.n_slots_used;
    t.kernel_launches = h->pipe->kernel_launches();
    t.h2d_bytes = h->res.h2d_bytes; t.d2h_bytes = h->res.d2h_bytes;
}

static void adopt_result(ac_handle* h) {   // host graph over the device result + bookkeeping shared by ac_build / ac_build_finish / the first use after ac_compress
    const double t0 = now_ms();
    h->graph.build(h->res, h->seqs, h->cfg.k, h->cfg.keep_positions != 0);
    if (!h->graph.device_sort) { DevicePipeline* pipe = h->pipe.get(); h->graph.device_sort = [pipe](const NumberKey* k, uint32_t n, uint32_t* out) { pipe->sort_number_keys(k, n, out); }; }
    const double ta = now_ms();
    h->graph.check_links();
    const double tb = now_ms();
    if (!h->graph.adopt_candidates(h->res)) h->graph.prepare_simplify();   // the expand_repeats work list: made on the device, or (needing links and paths only) here while the sequences are still being copied
    const double tc = now_ms();
    h->pipe->complete(h->res);
    const double t1 = now_ms();
    if (getenv("AC_HOST_PROFILE")) fprintf(stderr, "[host] adopt: graph %.2f, check_links %.2f, work list %.2f, wait for sequences %.2f ms\n", ta - t0, tb - ta, tc - tb, t1 - tc);
    if (!h->fused) { record_timings(h); h->t.host_graph = (float)(t1 - t0); h->t.host_simplify = 0; h->t.host_gfa = 0; }
    h->graph_ready = true;
}

// After ac_compress the graph arrays are still in HBM: bring them over and adopt them the first time a call needs the host graph.
static void ensure_graph(const ac_handle* ch) {
    ac_handle* h = const_cast<ac_handle*>(ch);
    if (!h->built || h->graph_ready) return;
    if (!h->fused) throw std::runtime_error("no graph on this handle");
    h->pipe->fetch_graph(h->res, h->cfg.keep_positions != 0);
    adopt_result(h);
    h->graph.simplify_structure();           // everything was done on the device: this adopts its numbering (nothing is recomputed)
}

int ac_build(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->uploaded) return set_error(h, AC_EINVAL, "ac_upload must precede ac_build");
    ResultFlusher flusher;
    if (h->built) flusher.start(h->res);
    {
        CallbackScope scope(h->pipe.get(), [&flusher] { flusher.join(); });      // cleared again on every way out, exceptions included
        if (h->peers.empty()) h->pipe->build(h->res, h->cfg.keep_positions != 0);
        else build_on_devices(h, false);
    }
    h->fused = false; h->graph_ready = false;
    adopt_result(h);
    h->built = true; h->gfa_ready = false; h->device_text_ok = false;
    return ok(h);
    AC_GUARD_END(h)
}

// compress.rs:42-47 in one call: build_kmer_graph, build_unitig_graph, simplify_unitig_graph and the text save_gfa writes, all on the
// device; only the text (and the counts compress prints) come back.  The graph itself is fetched when a later call asks for it.
int ac_compress(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->uploaded) return set_error(h, AC_EINVAL, "ac_upload must precede ac_compress");
    static const bool host_tail = getenv("AC_HOST_SIMPLIFY") != nullptr;       // comparison only: device graph, then expand_repeats and the text on the host
    if (host_tail) {
        int rc = ac_build(h); if (rc != AC_OK) return rc;
        if ((rc = ac_simplify(h)) != AC_OK) return rc;
        uint64_t n = 0; return ac_gfa_size(h, &n);
    }
    ResultFlusher flusher;
    if (h->built && h->graph_ready) flusher.start(h->res);                     // the host only ever writes to the graph arrays, and only once they were fetched
    {
        CallbackScope scope(h->pipe.get(), [&flusher] { flusher.join(); });
        if (h->peers.empty()) h->pipe->build(h->res, h->cfg.keep_positions != 0, true);
        else build_on_devices(h, true);
    }
    h->pipe->complete(h->res);
    h->fused = true; h->graph_ready = false; h->built = true;
    h->gfa_ptr = h->res.gfa_text; h->gfa_len = h->res.gfa_bytes; h->gfa_ready = true; h->device_text_ok = true;
    record_timings(h);
    h->t.host_graph = h->t.host_simplify = h->t.host_gfa = 0;
    return ok(h);
    AC_GUARD_END(h)
}

// ---- multi-GPU stages: one process per GPU, the caller runs the collectives between them on device buffers ----
int ac_build_local(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi, uint32_t multi) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->uploaded) return set_error(h, AC_EINVAL, "ac_upload must precede ac_build_local");
    h->built = false;
    h->pipe->build_local(seq_lo, seq_hi, multi != 0);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_entries_count(ac_handle* h, uint64_t* n) {
    if (!h || !n) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    *n = h->pipe->count_entries();
    return ok(h);
    AC_GUARD_END(h)
}
int ac_entries_export(ac_handle* h, void* dst, uint64_t cap_records) {
    if (!h || !dst) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->export_entries(dst, cap_records);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_entries_merge(ac_handle* h, const void* src, uint64_t n) {
    if (!h || (!src && n)) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->merge_entries(src, n);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_runs_local(ac_handle* h, uint64_t* n_runs) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    h->pipe->runs_local();
    if (n_runs) *n_runs = h->pipe->local_runs();
    return ok(h);
    AC_GUARD_END(h)
}
int ac_runs_export(ac_handle* h, void* dst, uint64_t cap_records) {
    if (!h || !dst) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->export_runs(dst, cap_records);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_runs_import(ac_handle* h, const void* src, uint64_t n) {
    if (!h || !src) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->import_runs(src, n);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_runs_import_padded(ac_handle* h, const void* src, uint64_t stride_records, const uint64_t* counts, uint32_t n_ranks) {
    if (!h || !src || !counts) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->import_runs_padded(src, stride_records, counts, n_ranks);
    return ok(h);
    AC_GUARD_END(h)
}
// ac_compress on the rank that imported every rank's occurrences: simplify_structure and the GFA text on the device as well
static int compress_finish(ac_handle* h, bool split_paths) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    static const bool host_tail = getenv("AC_HOST_SIMPLIFY") != nullptr;
    if (host_tail && !split_paths) {
        int rc = ac_build_finish(h); if (rc != AC_OK) return rc;
        if ((rc = ac_simplify(h)) != AC_OK) return rc;
        uint64_t n = 0; return ac_gfa_size(h, &n);
    }
    ResultFlusher flusher;
    if (h->built && h->graph_ready) flusher.start(h->res);
    {
        CallbackScope scope(h->pipe.get(), [&flusher] { flusher.join(); });
        h->pipe->finish(h->res, h->cfg.keep_positions != 0, true, split_paths);
    }
    h->pipe->complete(h->res);
    h->fused = true; h->graph_ready = false; h->built = true;
    h->gfa_ptr = h->res.gfa_text; h->gfa_len = h->res.gfa_bytes; h->gfa_ready = true; h->device_text_ok = true;
    record_timings(h);
    h->t.host_graph = h->t.host_simplify = h->t.host_gfa = 0;
    return ok(h);
    AC_GUARD_END(h)
}
int ac_compress_finish(ac_handle* h) { return compress_finish(h, false); }
int ac_compress_finish_split(ac_handle* h) { return compress_finish(h, true); }
int ac_path_tokens_export(ac_handle* h, void* dst, uint64_t stride_tokens, const uint64_t* counts, uint32_t n_ranks) {
    if (!h || !dst || !counts) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->pipe->export_path_tokens(dst, stride_tokens, counts, n_ranks);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_path_lines_render(ac_handle* h, const void* tokens, uint64_t n_tokens) {
    if (!h || (!tokens && n_tokens)) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->path_lines = nullptr; h->path_lines_len = 0;
    h->pipe->render_path_lines(tokens, n_tokens, &h->path_lines, &h->path_lines_len);
    return ok(h);
    AC_GUARD_END(h)
}
int ac_path_lines_data(ac_handle* h, const char** data, uint64_t* n_bytes) {
    if (!h || !data || !n_bytes) return set_error(h, AC_EINVAL, "null argument");
    *data = h->path_lines; *n_bytes = h->path_lines_len;
    return ok(h);
}
int ac_build_finish(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    ResultFlusher flusher;
    if (h->built) flusher.start(h->res);
    {
        CallbackScope scope(h->pipe.get(), [&flusher] { flusher.join(); });
        h->pipe->finish(h->res, h->cfg.keep_positions != 0);
    }
    h->fused = false; h->graph_ready = false;
    adopt_result(h);
    h->built = true; h->gfa_ready = false; h->device_text_ok = false;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_simplify(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_simplify");
    if (h->fused && !h->graph_ready) return ok(h);        // ac_compress has simplified the graph already
    const double t0 = now_ms();
    h->graph.simplify_structure();
    h->t.host_simplify = (float)(now_ms() - t0);
    h->gfa_ready = false;
    h->device_text_ok = !h->fused && h->res.gfa_text != nullptr && h->graph.last_simplify_on_device;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_merge_linear_paths(ac_handle* h, int use_paths) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_merge_linear_paths");
    ensure_graph(h);
    h->graph.merge_linear_paths(use_paths != 0);
    h->gfa_ready = false; h->device_text_ok = false;
    return ok(h);
    AC_GUARD_END(h)
}

// One process per GPU, kept on the CPU socket that GPU hangs off (what `numactl --cpunodebind` does): pinned buffers allocated
// afterwards are local to the DMA engine and to the host threads that edit them.  Measured on the two-socket B200 host: 11-14 %
// per step (profiles/r1s_*).  Changes the calling thread's affinity mask (threads created later inherit it); opt-in for that reason.
int ac_bind_host_to_device(int32_t device) {
    AC_GUARD_BEGIN
#ifdef AC_EMULATE
    (void)device;
    return set_error(nullptr, AC_ENODEVICE, "no CUDA device in the emulation build");
#else
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return set_error(nullptr, AC_ENODEVICE, "cannot query the PCI bus id of the device"); }
    std::string id(bus);
    for (char& ch : id) ch = (char)tolower((unsigned char)ch);
    if (id.size() > 12 && id.find(':') == 8) id = id.substr(4);                      // an 8-digit PCI domain where sysfs has 4
    FILE* f = fopen(("/sys/bus/pci/devices/" + id + "/numa_node").c_str(), "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return set_error(nullptr, AC_EINVAL, "the device reports no NUMA node");
    f = fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r");
    if (!f) return set_error(nullptr, AC_EINVAL, "cannot read the CPU list of the device's NUMA node");
    char list[4096] = {0};
    const bool got = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!got) return set_error(nullptr, AC_EINVAL, "cannot read the CPU list of the device's NUMA node");
    cpu_set_t now, want; CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof now, &now) != 0) return set_error(nullptr, AC_EINVAL, "sched_getaffinity failed");
    int count = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int n = sscanf(tok, "%d-%d", &a, &b);
        if (n < 1) continue;
        if (n == 1) b = a;
        for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; ++cpu) if (CPU_ISSET(cpu, &now)) { CPU_SET(cpu, &want); ++count; }
    }
    if (count < 8) return set_error(nullptr, AC_EINVAL, "fewer than 8 usable CPUs on the device's NUMA node: affinity left alone");
    if (sched_setaffinity(0, sizeof want, &want) != 0) return set_error(nullptr, AC_EINVAL, "sched_setaffinity failed");
    return node;
#endif
    AC_GUARD_END(nullptr)
}

int ac_load_gfa(ac_handle* h, const char* gfa_text, uint64_t length) {
    if (!h || !gfa_text) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    h->built = false; h->gfa_ready = false; h->uploaded = false; h->device_text_ok = false; h->fused = false; h->graph_ready = false;
    h->seqs.clear(); h->infos.clear(); h->ascii.clear(); h->loaded = LoadedInput(); h->res = PipelineResult(); h->t = ac_timings{};   // a loaded graph has no sequence bytes: ac_upload / ac_build need ac_add_sequence again
    h->graph.device_sort = nullptr;
    h->graph.load_gfa(gfa_text, (size_t)length, h->seqs);
    h->cfg.k = h->graph.k;
    h->built = true; h->graph_ready = true;
    return ok(h);
    AC_GUARD_END(h)
}

// cluster.rs:132-151: distances[a][b] = 1 - (length of the unitigs shared by a's and b's paths) / (length of a's unitigs), in the order
// of the handle's sequences.  The shared lengths are whole numbers found on the device; the one division per pair is done here in
// f64 exactly as the reference does it.
int ac_pairwise_distances(ac_handle* h, double* out, uint64_t cap) {
    if (!h || !out) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "a graph must be built or loaded before ac_pairwise_distances");
    ensure_graph(h);
    const HostGraph& g = h->graph;
    const uint64_t S = h->seqs.size();
    if (cap < S * S) return set_error(h, AC_ERANGE, "buffer too small");
    std::vector<uint32_t> len(g.U);
    for (uint32_t u = 0; u < g.U; ++u) len[u] = g.rec[u].len;
    std::vector<uint64_t> shared(S * S);
    h->pipe->pair_shared_lengths(g.path, g.path_off, (uint32_t)S, len.data(), g.U, shared.data());
    for (uint64_t a = 0; a < S; ++a) {
        const double a_len = (double)(uint32_t)shared[a * S + a];                    // the reference sums u32 lengths, then converts
        for (uint64_t b = 0; b < S; ++b) out[a * S + b] = 1.0 - ((double)shared[a * S + b] / a_len);
    }
    return ok(h);
    AC_GUARD_END(h)
}

// save_distance_matrix (cluster.rs:160-176): count, then one row per sequence: its Display form (sequence.rs:112-135) and the
// distances with eight decimals.
int ac_distance_matrix_text(ac_handle* h, char* out, uint64_t cap, uint64_t* length) {
    if (!h || !length) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    const uint64_t S = h->seqs.size();
    std::vector<double> d(std::max<uint64_t>(1, S * S));
    const int rc = ac_pairwise_distances(h, d.data(), S * S);
    if (rc != AC_OK) return rc;
    auto lower = [](std::string s) { for (char& c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; };
    auto weight = [&](const std::string& header, const std::string& key) -> uint64_t {   // sequence.rs:96-108
        const std::string low = lower(header);
        for (size_t a = 0; a < low.size();) {
            while (a < low.size() && isspace((unsigned char)low[a])) ++a;
            size_t b = a; while (b < low.size() && !isspace((unsigned char)low[b])) ++b;
            if (b > a && low.compare(a, key.size(), key) == 0 && b - a > key.size()) {
                const std::string v = low.substr(a + key.size(), b - a - key.size());
                const size_t first = v[0] == '+' ? 1 : 0;
                if (v.size() > first && v.find_first_not_of("0123456789", first) == std::string::npos) return strtoull(v.c_str(), nullptr, 10);
            }
            a = b;
        }
        return 1;
    };
    std::string text = std::to_string(S) + "\n";
    for (uint64_t a = 0; a < S; ++a) {
        const HostSeq& s = h->seqs[a];
        const std::string low = lower(s.contig_header);
        std::vector<std::string> extras;
        if (low.find("autocycler_trusted") != std::string::npos) extras.push_back("trusted");
        if (low.find("autocycler_ignore") != std::string::npos) extras.push_back("ignored");
        const uint64_t cw = weight(s.contig_header, "autocycler_cluster_weight="), nw = weight(s.contig_header, "autocycler_consensus_weight=");
        if (cw != 1) extras.push_back("cluster weight = " + std::to_string(cw));
        if (nw != 1) extras.push_back("consensus weight = " + std::to_string(nw));
        text += s.filename + " " + s.contig_header.substr(0, s.contig_header.find(' ')) + " (" + std::to_string(s.length) + " bp)";
        if (!extras.empty()) { text += " ["; for (size_t i = 0; i < extras.size(); ++i) { if (i) text += ", "; text += extras[i]; } text += "]"; }
        for (uint64_t b = 0; b < S; ++b) { char buf[64]; snprintf(buf, sizeof buf, "\t%.8f", d[a * S + b]); text += buf; }
        text += "\n";
    }
    *length = text.size();
    if (!out) return ok(h);
    if (cap < text.size()) return set_error(h, AC_ERANGE, "buffer too small");
    memcpy(out, text.data(), text.size());
    return ok(h);
    AC_GUARD_END(h)
}

int ac_renumber_unitigs(ac_handle* h) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_renumber_unitigs");
    ensure_graph(h);
    h->graph.renumber();
    h->gfa_ready = false; h->device_text_ok = false;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_counts_get(const ac_handle* h, ac_counts* out) {
    if (!h || !out) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_counts_get");
    memset(out, 0, sizeof *out);
    if (h->fused && !h->graph_ready) {       // straight from the device: what compress prints (compress.rs:152, unitig_graph.rs:509-516)
        const PipelineResult& r = h->res;
        out->n_kmers = 2 * r.n_slots_used; out->n_unitigs = r.n_unitigs; out->n_links = r.links_single;
        out->total_length = r.length_after; out->seq_bytes = r.length_after; out->length_before_simplify = r.length_before;
        out->n_next = r.n_links; out->n_sequences = h->seqs.size(); out->n_path_steps = r.n_runs;
        if (h->cfg.keep_positions) { out->n_fwd_pos = r.n_runs; out->n_rev_pos = r.n_runs; }
        return ok(h);
    }
    const HostGraph& g = h->graph;
    out->length_before_simplify = h->res.length_before;
    out->n_kmers = 2 * h->res.n_slots_used;
    out->n_unitigs = g.U;
    out->n_links = g.link_count_single();
    out->total_length = g.total_length();
    out->seq_bytes = out->total_length;
    out->n_fwd_pos = g.fpos.size(); out->n_rev_pos = g.rpos.size();
    out->n_next = g.n_links;
    out->n_sequences = h->seqs.size();
    out->n_path_steps = g.n_path;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_unitigs_copy(const ac_handle* h, ac_unitigs* o) {
    if (!h || !o) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_unitigs_copy");
    ensure_graph(h);
    const HostGraph& g = h->graph;
    const bool have_pos = !g.fpos_off.empty();
    if ((o->fpos_off || o->fpos || o->rpos_off || o->rpos) && !have_pos) return set_error(h, AC_EINVAL, "positions need ac_config.keep_positions");
    uint64_t so = 0, fo = 0, ro = 0, no = 0;
    for (uint32_t n = 0; n < g.U; ++n) {
        const uint32_t u = g.order[n];
        if (o->number) o->number[n] = g.number[u];
        if (o->depth) o->depth[n] = g.depth_of(u);
        if (o->seq_off) o->seq_off[n] = so;
        if (o->seq) memcpy(o->seq + so, g.seq_ptr(u), g.rec[u].len);
        so += g.rec[u].len;
        if (o->fpos_off) o->fpos_off[n] = fo;
        if (o->rpos_off) o->rpos_off[n] = ro;
        if (have_pos) {
            for (uint64_t x = g.fpos_off[u]; x < g.fpos_off[u + 1]; ++x) { if (o->fpos) o->fpos[fo] = (uint32_t)(g.fpos[x] >> 16); if (o->fpos_id_strand) o->fpos_id_strand[fo] = (uint16_t)(g.fpos[x] & 0xFFFF); ++fo; }
            for (uint64_t x = g.rpos_off[u]; x < g.rpos_off[u + 1]; ++x) { if (o->rpos) o->rpos[ro] = (uint32_t)(g.rpos[x] >> 16); if (o->rpos_id_strand) o->rpos_id_strand[ro] = (uint16_t)(g.rpos[x] & 0xFFFF); ++ro; }
        }
        for (uint32_t rev = 0; rev < 2; ++rev) {
            if (o->next_off) o->next_off[2 * (size_t)n + rev] = no;
            const UStrand from = us_make(u, rev != 0);
            for (uint32_t x = g.next_off[from]; x < g.next_off[from + 1]; ++x) {
                if (o->next) { const int32_t num = (int32_t)g.number[us_index(g.next[x])]; o->next[no] = us_reverse(g.next[x]) ? -num : num; }
                ++no;
            }
        }
    }
    if (o->seq_off) o->seq_off[g.U] = so;
    if (o->fpos_off) o->fpos_off[g.U] = fo;
    if (o->rpos_off) o->rpos_off[g.U] = ro;
    if (o->next_off) o->next_off[2 * (size_t)g.U] = no;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_path_copy(const ac_handle* h, uint64_t seq_index, int32_t* out, uint64_t cap, uint64_t* n) {
    if (!h || !n) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (h->built) ensure_graph(h);
    const HostGraph& g = h->graph;
    if (!h->built || seq_index >= g.n_seqs) return set_error(h, AC_EINVAL, "no such sequence");
    const uint64_t a = g.path_off[seq_index], b = g.path_off[seq_index + 1];
    *n = b - a;
    if (!out) return ok(h);
    if (cap < b - a) return set_error(h, AC_ERANGE, "path buffer too small");
    for (uint64_t x = a; x < b; ++x) { const int32_t num = (int32_t)g.number[us_index(g.path[x])]; out[x - a] = us_reverse(g.path[x]) ? -num : num; }
    return ok(h);
    AC_GUARD_END(h)
}

int ac_gfa_size(ac_handle* h, uint64_t* n_bytes) {
    if (!h || !n_bytes) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_gfa_size");
    if (!h->gfa_ready) {
        const double t0 = now_ms();
        if (h->device_text_ok) {      // the whole file was rendered on the device (AC_DEVICE_SIMPLIFY + AC_DEVICE_GFA in a plain build)
            h->gfa_ptr = h->res.gfa_text; h->gfa_len = h->res.gfa_bytes;
        } else {
            ensure_graph(h);
            h->graph.gfa_text(h->seqs, h->gfa);
            h->gfa_ptr = h->gfa.data(); h->gfa_len = h->gfa.size();
        }
        h->t.host_gfa = (float)(now_ms() - t0);
        h->gfa_ready = true;
        if (getenv("AC_HOST_PROFILE")) {
            const HostProfile& p = h->graph.prof;
            fprintf(stderr, "[host] adopt %.1f renumber(total) %.1f expand %.1f (%d passes; candidates %.1f, +compare %.1f, pass1 %.1f) gfa %.1f ms; U=%u\n",
                    p.adopt, p.renumber, p.expand, p.passes, p.candidates, p.compare, p.pass1, (double)h->t.host_gfa, h->graph.U);
        }
    }
    *n_bytes = h->gfa_len;
    return ok(h);
    AC_GUARD_END(h)
}

int ac_gfa_data(ac_handle* h, const char** data, uint64_t* n_bytes) {   // borrowed pointer, valid until the next call on this handle
    if (!data) return set_error(h, AC_EINVAL, "null argument");
    int rc = ac_gfa_size(h, n_bytes);
    if (rc != AC_OK) return rc;
    *data = h->gfa_ptr;
    return ok(h);
}

int ac_gfa_copy(ac_handle* h, char* buf, uint64_t cap) {
    uint64_t n = 0;
    int rc = ac_gfa_size(h, &n);
    if (rc != AC_OK) return rc;
    if (!buf || cap < n) return set_error(h, AC_ERANGE, "GFA buffer too small");
    memcpy(buf, h->gfa_ptr, n);
    return ok(h);
}

int ac_timings_get(const ac_handle* h, ac_timings* out) {
    if (!h || !out) return set_error(h, AC_EINVAL, "null argument");
    *out = h->t;
    out->kernel_launches = h->pipe->kernel_launches();
    return ok(h);
}

int ac_load_sequences(ac_handle* h, const char* dir, uint32_t max_contigs, uint32_t threads, uint64_t* assembly_count) {
    if (!h || !dir) return set_error(h, AC_EINVAL, "null argument");
    AC_GUARD_BEGIN
    ac_clear_sequences(h);
    // end repair runs on the device unless its k/2-base literals are longer than the scan kernel's two key words (k > 129) or AC_HOST_END_REPAIR is set
    const bool device_repair = !getenv("AC_HOST_END_REPAIR") && h->cfg.k / 2 <= 64;
    LoadedInput in = load_sequences(dir, h->cfg.k, max_contigs, threads ? threads : 1, false, device_repair ? h->pipe.get() : nullptr);
    for (size_t i = 0; i < in.seqs.size(); ++i) {
        int rc = ac_add_sequence(h, in.seqs[i].id, (const uint8_t*)in.padded[i].data(), in.padded[i].size(),
                                 in.seqs[i].filename.c_str(), in.seqs[i].contig_header.c_str());
        if (rc != AC_OK) return rc;
    }
    if (assembly_count) *assembly_count = in.assembly_count;
    in.padded.clear(); in.padded.shrink_to_fit();
    h->loaded = std::move(in);
    return ok(h);
    AC_GUARD_END(h)
}

int ac_sequence_get(const ac_handle* h, uint64_t index, uint16_t* seq_id, uint64_t* length, char* fwd, uint64_t cap_fwd,
                    char* filename, uint64_t cap_fn, char* header, uint64_t cap_hd) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    if (index >= h->seqs.size()) return set_error(h, AC_EINVAL, "no such sequence");
    const HostSeq& s = h->seqs[index];
    const uint64_t padded = s.length + h->cfg.k - 1;
    if (seq_id) *seq_id = s.id;
    if (length) *length = s.length;
    if (fwd) {
        if (!h->ascii.p || h->infos.empty()) return set_error(h, AC_EINVAL, "this handle holds a loaded graph: its sequences have no bytes (use ac_sequence_reconstruct)");
        if (cap_fwd < padded + 1) return set_error(h, AC_ERANGE, "buffer too small");
        memcpy(fwd, h->ascii.p + s.start, padded); fwd[padded] = 0;
    }
    if (filename) { if (cap_fn < s.filename.size() + 1) return set_error(h, AC_ERANGE, "buffer too small"); memcpy(filename, s.filename.c_str(), s.filename.size() + 1); }
    if (header) { if (cap_hd < s.contig_header.size() + 1) return set_error(h, AC_ERANGE, "buffer too small"); memcpy(header, s.contig_header.c_str(), s.contig_header.size() + 1); }
    return ok(h);
}

// reconstruct_original_sequence / get_sequence_from_path (unitig_graph.rs:383-400): the unitig strands of the sequence's
// path, concatenated; equals the input contig (tests.rs:114-127).
int ac_sequence_reconstruct(const ac_handle* h, uint64_t index, char* out, uint64_t cap, uint64_t* length) {
    if (!h) return set_error(nullptr, AC_EINVAL, "null handle");
    AC_GUARD_BEGIN
    if (!h->built) return set_error(h, AC_EINVAL, "ac_build must precede ac_sequence_reconstruct");
    if (index >= h->seqs.size()) return set_error(h, AC_EINVAL, "no such sequence");
    ensure_graph(h);
    const HostGraph& g = h->graph;
    uint64_t total = 0;
    for (uint64_t x = g.path_off[index]; x < g.path_off[index + 1]; ++x) total += g.rec[us_index(g.path[x])].len;
    if (length) *length = total;
    if (!out) return ok(h);
    if (cap < total) return set_error(h, AC_ERANGE, "buffer too small");
    char* p = out;
    for (uint64_t x = g.path_off[index]; x < g.path_off[index + 1]; ++x) {
        const UStrand s = g.path[x]; const uint32_t n = g.rec[us_index(s)].len; const char* src = g.seq_ptr(us_index(s));
        if (!us_reverse(s)) memcpy(p, src, n);
        else for (uint32_t j = 0; j < n; ++j) { const char b = src[n - 1 - j]; p[j] = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b; }
        p += n;
    }
    return ok(h);
    AC_GUARD_END(h)
}

int ac_compress_dir(const char* assemblies_dir, const char* autocycler_dir, uint32_t k, uint32_t max_contigs, uint32_t threads,
                    int32_t device, int32_t verbose) {
    return ac_compress_dir_devices(assemblies_dir, autocycler_dir, k, max_contigs, threads, &device, 1, verbose);
}

int ac_compress_dir_devices(const char* assemblies_dir, const char* autocycler_dir, uint32_t k, uint32_t max_contigs, uint32_t threads,
                            const int32_t* devices, int32_t n_devices, int32_t verbose) {
    if (!assemblies_dir || !autocycler_dir || !devices || n_devices < 1) return set_error(nullptr, AC_EINVAL, "null argument");
    ac_handle* h = nullptr;
    AC_GUARD_BEGIN
    // check_settings, compress.rs:53-62
    struct stat st;
    if (stat(assemblies_dir, &st) != 0) return set_error(nullptr, AC_EINPUT, std::string("directory does not exist: ") + assemblies_dir);
    if (!S_ISDIR(st.st_mode)) return set_error(nullptr, AC_EINPUT, std::string(assemblies_dir) + " is not a directory");
    if (stat(autocycler_dir, &st) == 0 && !S_ISDIR(st.st_mode)) return set_error(nullptr, AC_EINPUT, std::string(autocycler_dir) + " exists but is not a directory");
    if (k < 11) return set_error(nullptr, AC_EINPUT, "--kmer cannot be less than 11");
    if (k > 501) return set_error(nullptr, AC_EINPUT, "--kmer cannot be greater than 501");
    if (k % 2 == 0) return set_error(nullptr, AC_EINPUT, "--kmer must be odd");
    if (threads < 1) return set_error(nullptr, AC_EINPUT, "--threads cannot be less than 1");
    if (threads > 100) return set_error(nullptr, AC_EINPUT, "--threads cannot be greater than 100");
    if (k > AC_MAX_K) return set_error(nullptr, AC_EINPUT, "--kmer above " + std::to_string(AC_MAX_K) + " is not supported by this build of the GPU path (there is no CPU fallback)");
    ac_config cfg{}; cfg.k = k; cfg.device = devices[0]; cfg.stream = nullptr; cfg.keep_positions = 0; cfg.n_devices = n_devices; cfg.devices = devices;
    int rc = ac_create(&h, &cfg);
    if (rc != AC_OK) return rc;
    std::unique_ptr<ac_handle, void (*)(ac_handle*)> guard(h, ac_destroy);
    { std::string d = autocycler_dir; for (size_t i = 1; i <= d.size(); ++i) if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0777); }   // create_dir_all
    uint64_t assemblies = 0;
    const double t0 = now_ms();
    if ((rc = ac_load_sequences(h, assemblies_dir, max_contigs, threads, &assemblies)) != AC_OK) { g_error = h->err; return rc; }
    const double t1 = now_ms();
    if (verbose) fprintf(stderr, "%zu sequence%s loaded from %llu assembl%s\n\n", h->seqs.size(), h->seqs.size() == 1 ? "" : "s",
                         (unsigned long long)assemblies, assemblies == 1 ? "y" : "ies");
    if ((rc = ac_upload(h)) != AC_OK || (rc = ac_compress(h)) != AC_OK) { g_error = h->err; return rc; }
    ac_counts c{};
    ac_counts_get(h, &c);
    // simplify_structure moves bases between unitigs: the counts stay, the total length shrinks (print_basic_graph_info, unitig_graph.rs:509-516)
    if (verbose) fprintf(stderr, "Graph contains %llu k-mers\n\n%llu unitig%s, %llu link%s\ntotal length: %llu bp\n\n", (unsigned long long)c.n_kmers,
                         (unsigned long long)c.n_unitigs, c.n_unitigs == 1 ? "" : "s", (unsigned long long)c.n_links, c.n_links == 1 ? "" : "s",
                         (unsigned long long)c.length_before_simplify);
    if (verbose) fprintf(stderr, "%llu unitig%s, %llu link%s\ntotal length: %llu bp\n\n", (unsigned long long)c.n_unitigs, c.n_unitigs == 1 ? "" : "s",
                         (unsigned long long)c.n_links, c.n_links == 1 ? "" : "s", (unsigned long long)c.total_length);
    uint64_t n = 0;
    if ((rc = ac_gfa_size(h, &n)) != AC_OK) { g_error = h->err; return rc; }
    const std::string out_gfa = std::string(autocycler_dir) + "/input_assemblies.gfa", out_yaml = std::string(autocycler_dir) + "/input_assemblies.yaml";
    FILE* f = fopen(out_gfa.c_str(), "wb");
    if (!f || fwrite(h->gfa_ptr, 1, h->gfa_len, f) != h->gfa_len) { if (f) fclose(f); return set_error(nullptr, AC_EIO, "cannot write " + out_gfa); }
    fclose(f);
    const std::string yaml = metrics_yaml(h->loaded, c.n_unitigs, c.total_length);
    f = fopen(out_yaml.c_str(), "wb");
    if (!f || fwrite(yaml.data(), 1, yaml.size(), f) != yaml.size()) { if (f) fclose(f); return set_error(nullptr, AC_EIO, "cannot write " + out_yaml); }
    fclose(f);
    if (verbose) {
        const ac_timings& t = h->t;
        fprintf(stderr, "Compressed unitig graph: %s\nInput assembly stats:    %s\n", out_gfa.c_str(), out_yaml.c_str());
        fprintf(stderr, "load+repair %.1f ms | h2d %.2f pack %.2f insert %.2f adjacency %.2f boundaries %.2f runs %.2f unitigs %.2f links %.2f d2h %.2f ms"
                        " simplify %.2f gfa %.2f | host graph %.1f simplify %.1f gfa %.1f ms | total %.1f ms\n\n",
                t1 - t0, t.h2d, t.pack, t.insert, t.adjacency, t.boundaries, t.runs, t.unitigs, t.links, t.d2h, t.device_simplify, t.device_gfa, t.host_graph, t.host_simplify, t.host_gfa, now_ms() - t0);
    }
    return ok(h);
    AC_GUARD_END(nullptr)
}

// `autocycler decompress` (decompress.rs:27-114): reconstructs every input contig from its path and writes them back per
// original file (FASTA, gzip when the name ends in .gz) and/or into one FASTA file.
int ac_decompress_gfa(const char* in_gfa, const char* out_dir, const char* out_file, int32_t device, int32_t verbose) {
    if (!in_gfa) return set_error(nullptr, AC_EINVAL, "null argument");
    ac_handle* h = nullptr;
    AC_GUARD_BEGIN
    struct stat st;
    if (stat(in_gfa, &st) != 0) return set_error(nullptr, AC_EINPUT, std::string("file does not exist: ") + in_gfa);            // misc.rs:98-107
    if (!S_ISREG(st.st_mode)) return set_error(nullptr, AC_EINPUT, std::string(in_gfa) + " is not a file");
    const bool to_dir = out_dir && *out_dir, to_file = out_file && *out_file;
    if (!to_dir && !to_file) return set_error(nullptr, AC_EINPUT, "either --out_dir or --out_file is required");                  // decompress.rs:45-47
    if (to_dir && stat(out_dir, &st) == 0 && !S_ISDIR(st.st_mode)) return set_error(nullptr, AC_EINPUT, std::string(out_dir) + " exists but is not a directory");
    std::string text;
    {
        FILE* f = fopen(in_gfa, "rb");
        if (!f) return set_error(nullptr, AC_EIO, std::string("cannot read ") + in_gfa);
        text.resize((size_t)st.st_size);
        const size_t got = text.empty() ? 0 : fread(&text[0], 1, text.size(), f);
        fclose(f);
        if (got != text.size()) return set_error(nullptr, AC_EIO, std::string("cannot read ") + in_gfa);
    }
    ac_config cfg{}; cfg.k = 51; cfg.device = device;
    int rc = ac_create(&h, &cfg);
    if (rc != AC_OK) return rc;
    std::unique_ptr<ac_handle, void (*)(ac_handle*)> guard(h, ac_destroy);
    if ((rc = ac_load_gfa(h, text.data(), text.size())) != AC_OK) { g_error = h->err; return rc; }
    if (verbose) {
        ac_counts c{}; ac_counts_get(h, &c);
        fprintf(stderr, "%llu unitig%s, %llu link%s\ntotal length: %llu bp\n\n", (unsigned long long)c.n_unitigs, c.n_unitigs == 1 ? "" : "s",
                (unsigned long long)c.n_links, c.n_links == 1 ? "" : "s", (unsigned long long)c.total_length);
    }
    // reconstruct_original_sequences (unitig_graph.rs:362-370): per filename, in sequence order; filenames sorted when written
    std::vector<std::string> seqs(h->seqs.size());
    for (size_t i = 0; i < h->seqs.size(); ++i) {
        uint64_t n = 0;
        if ((rc = ac_sequence_reconstruct(h, i, nullptr, 0, &n)) != AC_OK) { g_error = h->err; return rc; }
        if (n != h->seqs[i].length) return set_error(nullptr, AC_EINPUT, "reconstructed sequence does not have expected length");   // unitig_graph.rs:386
        seqs[i].resize(n);
        if (n && (rc = ac_sequence_reconstruct(h, i, &seqs[i][0], n, &n)) != AC_OK) { g_error = h->err; return rc; }
    }
    std::vector<std::string> names;
    for (auto& s : h->seqs) names.push_back(s.filename);
    std::sort(names.begin(), names.end()); names.erase(std::unique(names.begin(), names.end()), names.end());
    auto first_word = [](const std::string& hd) { return hd.substr(0, hd.find(' ')); };
    if (to_dir) {
        { std::string d = out_dir; for (size_t i = 1; i <= d.size(); ++i) if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0777); }
        for (const std::string& name : names) {
            const std::string path = std::string(out_dir) + "/" + name;
            if (verbose) fprintf(stderr, "%s:\n", path.c_str());
            std::string body;
            for (size_t i = 0; i < h->seqs.size(); ++i)
                if (h->seqs[i].filename == name) {
                    if (verbose) fprintf(stderr, "  %s (%zu bp)\n", first_word(h->seqs[i].contig_header).c_str(), seqs[i].size());
                    body += ">" + h->seqs[i].contig_header + "\n" + seqs[i] + "\n";
                }
            const bool gz = path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0;       // decompress.rs:96
            if (gz) {
                gzFile g = gzopen(path.c_str(), "wb");
                if (!g || (body.size() && gzwrite(g, body.data(), (unsigned)body.size()) != (int)body.size())) { if (g) gzclose(g); return set_error(nullptr, AC_EIO, "cannot write " + path); }
                gzclose(g);
            } else {
                FILE* f = fopen(path.c_str(), "wb");
                if (!f || fwrite(body.data(), 1, body.size(), f) != body.size()) { if (f) fclose(f); return set_error(nullptr, AC_EIO, "cannot write " + path); }
                fclose(f);
            }
            if (verbose) fprintf(stderr, "\n");
        }
    }
    if (to_file) {   // decompress.rs:116-137
        if (verbose) fprintf(stderr, "%s:\n", out_file);
        std::string body;
        for (const std::string& name : names) {
            std::string clean = name; for (char& ch : clean) if (ch == ' ') ch = '_';
            for (size_t i = 0; i < h->seqs.size(); ++i)
                if (h->seqs[i].filename == name) {
                    if (verbose) fprintf(stderr, "  %s__%s (%zu bp)\n", name.c_str(), first_word(h->seqs[i].contig_header).c_str(), seqs[i].size());
                    body += ">" + clean + "__" + h->seqs[i].contig_header + "\n" + seqs[i] + "\n";
                }
        }
        FILE* f = fopen(out_file, "wb");
        if (!f || fwrite(body.data(), 1, body.size(), f) != body.size()) { if (f) fclose(f); return set_error(nullptr, AC_EIO, std::string("cannot write ") + out_file); }
        fclose(f);
        if (verbose) fprintf(stderr, "\n");
    }
    return ok(h);
    AC_GUARD_END(nullptr)
}

}  // extern "C"
