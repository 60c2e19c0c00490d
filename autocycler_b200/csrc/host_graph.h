// Host side of the compress path: everything that operates on unitigs (10^2..10^6 objects, order
// dependent) — link ordering, renumbering, repeat expansion and GFA text.  Input is the device
// pipeline's result; output is byte-identical to the reference's UnitigGraph::save_gfa.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pipeline.h"

struct HostSeq {              // sequence.rs:19-28 minus the bytes (they live in one concatenated buffer)
    uint16_t id;
    std::string filename, contig_header;
    uint64_t length;          // L
    uint64_t start;           // global coordinate of padded byte 0
};

// A unitig strand: (index << 1) | reverse.  Indices are seed order (the order in which the
// reference's walk would have created the unitigs, unitig_graph.rs:179-225) and never change.
typedef uint32_t UStrand;
static inline uint32_t us_index(UStrand s) { return s >> 1; }
static inline bool us_reverse(UStrand s) { return s & 1; }
static inline UStrand us_make(uint32_t idx, bool reverse) { return (idx << 1) | (reverse ? 1u : 0u); }
static inline UStrand us_flip(UStrand s) { return s ^ 1u; }

struct HostUnitig {           // unitig.rs:30-45
    uint32_t number = 0;
    std::string seq;          // forward_seq (trimmed)
    uint32_t depth = 0;       // integral on this path: every k-mer of a chain has the same depth
    uint32_t min_fpos = 0xFFFFFFFFu, min_rpos = 0xFFFFFFFFu;   // min over forward_positions / reverse_positions
    std::vector<UStrand> next[2], prev[2];                     // [0] forward strand, [1] reverse strand
    std::vector<uint64_t> fpos, rpos;                          // optional full position lists (pos << 16 | id_and_strand)
};

struct GraphStats {
    uint64_t n_kmers = 0;                     // both strands, == KmerGraph.kmers.len() (compress.rs:152)
    uint64_t unitigs_before = 0, links_before = 0, length_before = 0;
    uint64_t unitigs_after = 0, links_after = 0, length_after = 0;
    double ms_build = 0, ms_simplify = 0, ms_gfa = 0;
};

class HostGraph {
public:
    uint32_t k = 0;
    std::vector<HostUnitig> units;            // seed order
    std::vector<uint32_t> order;              // current numbering order: order[n-1] = index of unitig number n
    std::vector<std::vector<UStrand>> paths;  // per sequence, its unitig path (unitig_graph.rs:447-465)

    // unitig_graph.rs:36-48 from the device result (build, simplify_seqs, create_links, trim_overlaps, renumber, check)
    void build(const PipelineResult& r, const std::vector<HostSeq>& seqs, const uint8_t* ascii, uint32_t k,
               bool keep_positions);
    void renumber();                          // unitig_graph.rs:295-315
    void check_links() const;                 // unitig_graph.rs:752-793
    void simplify_structure();                // graph_simplification.rs:26-40
    size_t expand_repeats();                  // graph_simplification.rs:43-86
    std::string gfa_text(const std::vector<HostSeq>& seqs) const;   // unitig_graph.rs:317-360
    uint64_t total_length() const;
    uint64_t link_count_single() const;       // unitig_graph.rs:478-507 (.1)
private:
    std::vector<uint8_t> fixed_start, fixed_end;
    bool fixed_ready = false;
    void compute_fixed();
};
