// Device pipeline interface (host side).  Replaces the reference's KmerGraph build + unitig walk
// (compress.rs:42-43 -> kmer_graph.rs:86-134, unitig_graph.rs:176-226, unitig.rs:112-155) with an
// order-free formulation that runs as data-parallel kernels; see DESIGN.md.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "kmer_key.h"

#define AC_MAX_W 4            // k <= 127
#define AC_MAX_LINKS 5        // successors over the 5-letter alphabet (kmer_graph.rs:142)

struct PipelineTimings {      // milliseconds, CUDA events on the pipeline's stream (0 under emulation)
    float h2d = 0, pack = 0, insert = 0, adjacency = 0, boundaries = 0, runs = 0, unitigs = 0, links = 0, d2h = 0, total = 0;
};

struct DeviceUnitig {
    uint64_t start;           // global coordinate of the first window of the representative occurrence
    uint32_t len;             // number of k-mers == trimmed length (unitig.rs:157-165)
    uint32_t depth;           // Kmer::depth() of every k-mer in the chain (unitig.rs:148-155)
    uint32_t flip;            // 1: the unitig's forward strand is the reverse complement of the representative occurrence
    int32_t min_d;            // smallest k-mer of both strands (kmer_graph.rs:168-173 order) = the walk's seed
    uint64_t min_w[AC_MAX_W];
};

struct PipelineResult {
    uint32_t W = 0;
    uint64_t n_slots_used = 0;                 // distinct canonical k-mers; KmerGraph.kmers.len() == 2x this
    uint64_t capacity = 0;
    uint64_t n_dotted = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;     // bytes copied host->device by upload() and device->host by build()
    std::vector<DeviceUnitig> unitigs;         // in representative-occurrence order (not yet seed order)
    std::vector<uint32_t> link_count;          // [2*U]   index 2j+e, e=0: strand of the representative occurrence, e=1: its reverse
    std::vector<uint32_t> links;               // [2*U*AC_MAX_LINKS] targets as 2j'+e'
    std::vector<uint64_t> run_start;           // [R] global coordinate of each unitig occurrence along the input sequences
    std::vector<uint32_t> run_len;             // [R]
    std::vector<uint32_t> run_unitig;          // [R] (unitig << 1) | same_direction_as_representative
    PipelineTimings t;
};

class DevicePipeline {
public:
    DevicePipeline(int device, void* stream);
    ~DevicePipeline();
    // ascii: all padded, end-repaired forward strands concatenated (bytes in "ACGT."); seqs: their layout.
    // host_pinned: ascii lives in pinned host memory (H2D copy can be async).
    void upload(const uint8_t* ascii, uint64_t total, const SeqInfo* seqs, uint32_t n_seqs, uint32_t k);
    void build(PipelineResult& out);           // kernels + D2H of the (small) results
    unsigned long long kernel_launches() const;
    struct Impl;
private:
    Impl* impl;
};
