// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement, for tests/ (ctypes),
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Every returned char*
// is malloc'd; free with orc_free().  NULL / non-zero means failure; see orc_last_error().
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "autocycler_oracle.hpp"

using namespace orc;

static thread_local std::string g_err;

static char* dup_out(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p;
}

static std::vector<std::string> split_lines(const char* text) {
    std::vector<std::string> v; std::string s(text); size_t i = 0;
    while (i < s.size()) { size_t j = s.find('\n', i); if (j == std::string::npos) j = s.size(); v.push_back(s.substr(i, j - i)); i = j + 1; }
    return v;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const OracleError& e) { g_err = e.msg; return ret; } \
                         catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

struct orc_stats {
    uint64_t n_kmers;                 // KmerGraph.kmers.len() (both strands), compress.rs:152
    uint64_t unitigs_before, links_before, length_before;   // print_basic_graph_info after from_kmer_graph
    uint64_t unitigs_after, links_after, length_after;      // ... after simplify_structure
    uint64_t input_bases;             // sum of Sequence.length
    double t_load, t_kmer_graph, t_unitig_graph, t_simplify, t_gfa;   // seconds per reference stage
};

const char* orc_last_error() { return g_err.c_str(); }
void orc_free(char* p) { free(p); }

static Sequence one_seq(const char* seq, int k) {
    std::string s(seq);
    return Sequence::new_with_seq(1, s, "assembly.fasta", "contig_1", s.size(), (uint32_t)k / 2);
}

// kmer_graph.rs:266-282 style listing: one "KMER:positions" line per k-mer in iterate_kmers order
char* orc_kmers_sorted(const char* seq, int k) {
    ORC_TRY
    Sequence s = one_seq(seq, k); KmerGraph g((uint32_t)k); g.add_sequence(s, 1);
    std::string out; for (Kmer* km : g.iterate_kmers()) { out += km->display(); out += '\n'; }
    return dup_out(out);
    ORC_CATCH(nullptr)
}

char* orc_neighbour_kmers(const char* seq, int k, const char* kmer, int next) {
    ORC_TRY
    Sequence s = one_seq(seq, k); KmerGraph g((uint32_t)k); g.add_sequence(s, 1);
    auto v = next ? g.next_kmers(kmer) : g.prev_kmers(kmer);
    std::string out; for (Kmer* km : v) { out += km->seq(); out += '\n'; }
    return dup_out(out);
    ORC_CATCH(nullptr)
}

char* orc_find_best_match(const char* newline_joined) {
    ORC_TRY
    return dup_out(find_best_match(split_lines(newline_joined)));
    ORC_CATCH(nullptr)
}

// "id\tfilename\theader\tlength\tforward_seq\n" per kept sequence; first line "assembly_count\tN"
char* orc_load_sequences(const char* dir, int k, int max_contigs, int threads) {
    ORC_TRY
    InputAssemblyMetrics m;
    auto r = load_sequences(dir, (uint32_t)k, m, (uint32_t)max_contigs, threads, false);
    std::string out = "assembly_count\t" + std::to_string(r.second) + "\n";
    for (auto& s : r.first)
        out += std::to_string(s.id) + "\t" + s.filename + "\t" + s.contig_header + "\t" + std::to_string(s.length) + "\t" + s.forward_seq + "\n";
    return dup_out(out);
    ORC_CATCH(nullptr)
}

static std::string run_graph_stages(std::vector<Sequence>& sequences, size_t assembly_count, int k, orc_stats* st,
                                    std::string* seed_dump) {
    double t0 = now_s();
    KmerGraph kg((uint32_t)k);
    kg.add_sequences(sequences, assembly_count);                       // compress.rs:144-155
    double t1 = now_s();
    UnitigGraph ug;
    if (seed_dump) {   // same stages as from_kmer_graph, with a dump of the seed-order state in between
        ug.k_size = kg.k_size;
        ug.build_unitigs_from_kmer_graph(kg);
        ug.simplify_seqs();
        std::string& d = *seed_dump;
        for (auto& u : ug.unitigs) {
            char dp[64]; snprintf(dp, sizeof dp, "%.6f", u->depth);
            d += u->forward_seq; d += '\t'; d += dp; d += '\t';
            for (size_t i = 0; i < u->forward_positions.size(); ++i) { if (i) d += ','; d += u->forward_positions[i].to_string(); }
            d += '\t';
            for (size_t i = 0; i < u->reverse_positions.size(); ++i) { if (i) d += ','; d += u->reverse_positions[i].to_string(); }
            d += '\n';
        }
        ug.create_links(); ug.trim_overlaps(); ug.renumber_unitigs(); ug.check_links();
    } else {
        ug = UnitigGraph::from_kmer_graph(kg);                         // compress.rs:158-167
    }
    double t2 = now_s();
    if (st) { st->n_kmers = kg.len(); st->unitigs_before = ug.unitigs.size(); st->links_before = ug.link_count().second; st->length_before = ug.total_length(); }
    double t2b = now_s();
    simplify_structure(ug, sequences);                                 // compress.rs:170-178
    double t3 = now_s();
    if (st) { st->unitigs_after = ug.unitigs.size(); st->links_after = ug.link_count().second; st->length_after = ug.total_length(); }
    double t3b = now_s();
    std::string gfa = ug.gfa_text(sequences);                          // compress.rs:47
    double t4 = now_s();
    if (st) {
        st->t_kmer_graph = t1 - t0; st->t_unitig_graph = t2 - t1; st->t_simplify = t3 - t2b; st->t_gfa = t4 - t3b;
        st->input_bases = 0; for (auto& s : sequences) st->input_bases += s.length;
    }
    return gfa;
}

// Whole `compress` from a directory of FASTA files: returns the GFA text; *yaml_out gets the sidecar.
char* orc_compress_dir(const char* dir, int k, int max_contigs, int threads, char** yaml_out, orc_stats* st) {
    ORC_TRY
    InputAssemblyMetrics m;
    double t0 = now_s();
    auto r = load_sequences(dir, (uint32_t)k, m, (uint32_t)max_contigs, threads, false);
    if (st) st->t_load = now_s() - t0;
    std::string gfa = run_graph_stages(r.first, r.second, k, st, nullptr);
    if (yaml_out) {   // save_metrics, compress.rs:181-189
        m.input_assemblies_count = (uint32_t)r.second;
        m.input_assemblies_total_contigs = (uint32_t)r.first.size();
        m.input_assemblies_total_length = 0; for (auto& s : r.first) m.input_assemblies_total_length += s.length;
        m.compressed_unitig_count = st ? (uint32_t)st->unitigs_after : 0;
        m.compressed_unitig_total_length = st ? st->length_after : 0;
        *yaml_out = dup_out(m.to_yaml());
    }
    return dup_out(gfa);
    ORC_CATCH(nullptr)
}

// The graph stages only (compress.rs:42-47) on sequences that are already padded and end-repaired:
// the same inputs the GPU library's C ABI takes.  seed_dump_out (optional) receives the pre-link,
// seed-order unitigs: "untrimmed_fwd_seq\tdepth\tfwd_positions\trev_positions".
char* orc_compress_seqs(int n, const char* const* padded_fwd, const uint16_t* ids, const char* const* filenames,
                        const char* const* headers, int assembly_count, int k, orc_stats* st, char** seed_dump_out) {
    ORC_TRY
    std::vector<Sequence> sequences;
    for (int i = 0; i < n; ++i) {
        Sequence s; s.id = ids[i]; s.forward_seq = padded_fwd[i]; s.reverse_seq = reverse_complement(s.forward_seq);
        s.filename = filenames[i]; s.contig_header = headers[i]; s.length = s.forward_seq.size() - 2 * (size_t)(k / 2);
        sequences.push_back(std::move(s));
    }
    std::string seed;
    std::string gfa = run_graph_stages(sequences, (size_t)assembly_count, k, st, seed_dump_out ? &seed : nullptr);
    if (seed_dump_out) *seed_dump_out = dup_out(seed);
    return dup_out(gfa);
    ORC_CATCH(nullptr)
}

// tests.rs:108-112: load a GFA and save it again
char* orc_gfa_roundtrip(const char* gfa_text) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    return dup_out(r.first.gfa_text(r.second));
    ORC_CATCH(nullptr)
}

// tests.rs:114-127 / decompress.rs:83-105
int orc_decompress(const char* gfa_text, const char* out_dir) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    save_original_seqs_to_dir(out_dir, r.first, r.second);
    return 0;
    ORC_CATCH(1)
}

// cluster.rs:132-176: the distance matrix file for the sequences of a GFA
char* orc_pairwise_distances(const char* gfa_text) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    return dup_out(pairwise_distance_matrix(r.first, r.second));
    ORC_CATCH(nullptr)
}

// graph_simplification.rs:743-803: from_gfa_lines, merge_linear_paths (with or without the paths), save again
char* orc_gfa_merge_linear_paths(const char* gfa_text, int use_paths, int renumber) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    std::vector<Sequence> none;
    merge_linear_paths(r.first, use_paths ? r.second : none);
    if (renumber) r.first.renumber_unitigs();      // trim.rs:266-268
    return dup_out(r.first.gfa_text(use_paths ? r.second : none));   // without the paths the merged graph no longer carries them
    ORC_CATCH(nullptr)
}

// graph_simplification.rs:702-707: the fixed starts / ends merge_linear_paths works from, "s s s\ne e e\n" ascending
char* orc_gfa_merge_fixed_sets(const char* gfa_text) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    std::unordered_set<uint32_t> fs, fe;
    merge_fixed_sets(r.first, r.second, fs, fe);
    std::string out;
    for (auto* set : {&fs, &fe}) {
        std::vector<uint32_t> v(set->begin(), set->end()); std::sort(v.begin(), v.end());
        for (size_t i = 0; i < v.size(); ++i) { if (i) out += ' '; out += std::to_string(v[i]); }
        out += '\n';
    }
    return dup_out(out);
    ORC_CATCH(nullptr)
}

// graph_simplification.rs:627-671: from_gfa_lines, optionally simplify_structure, dump "number\tseq" in graph order
char* orc_gfa_unitig_seqs(const char* gfa_text, int simplify, int use_paths) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    std::vector<Sequence> none;
    if (simplify) simplify_structure(r.first, use_paths ? r.second : none);
    std::string out;
    for (auto& u : r.first.unitigs) out += std::to_string(u->number) + "\t" + u->forward_seq + "\n";
    return dup_out(out);
    ORC_CATCH(nullptr)
}

static std::string strand_list(std::vector<UnitigStrand> v) {
    std::sort(v.begin(), v.end(), [](const UnitigStrand& a, const UnitigStrand& b) { return a.number() != b.number() ? a.number() < b.number() : a.strand < b.strand; });
    std::string s; for (size_t i = 0; i < v.size(); ++i) { if (i) s += ','; s += std::to_string(v[i].number()); s += v[i].strand ? '+' : '-'; }
    return s;
}

// graph_simplification.rs:582-625: per unitig "number\tinputs\toutputs"
char* orc_gfa_exclusive(const char* gfa_text) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    std::string out;
    for (auto& u : r.first.unitigs)
        out += std::to_string(u->number) + "\t" + strand_list(get_exclusive_inputs(u.get())) + "\t" + strand_list(get_exclusive_outputs(u.get())) + "\n";
    return dup_out(out);
    ORC_CATCH(nullptr)
}

// graph_simplification.rs:540-580: spec = "1+,2+,3-"; which = 0 common start, 1 common end
char* orc_gfa_common_seq(const char* gfa_text, const char* spec, int which) {
    ORC_TRY
    auto r = UnitigGraph::from_gfa_lines(split_lines(gfa_text));
    std::vector<UnitigStrand> v; std::string s(spec); size_t i = 0;
    while (i < s.size()) { size_t j = s.find(',', i); if (j == std::string::npos) j = s.size();
        std::string t = s.substr(i, j - i); v.push_back({r.first.unitig_index.at((uint32_t)atoi(t.c_str())), t.back() == '+'}); i = j + 1; }
    return dup_out(which ? get_common_end_seq(v) : get_common_start_seq(v));
    ORC_CATCH(nullptr)
}

// unitig.rs:457-555: apply one shift op to a one-segment unitig carrying test positions; returns
// "fwd\trev\tfpos0,fpos1\trpos0,rpos1"
char* orc_unitig_shift(const char* segment_line, int op, int amount, const char* seq) {
    ORC_TRY
    Unitig u = Unitig::from_segment_line(segment_line);
    u.forward_positions.emplace_back(1, strand::FORWARD, 100); u.reverse_positions.emplace_back(2, strand::REVERSE, 890);
    u.forward_positions.emplace_back(2, strand::REVERSE, 200); u.reverse_positions.emplace_back(2, strand::FORWARD, 790);
    if (op == 0) u.remove_seq_from_start((size_t)amount);
    else if (op == 1) u.remove_seq_from_end((size_t)amount);
    else if (op == 2) u.add_seq_to_start(seq);
    else u.add_seq_to_end(seq);
    std::string out = u.forward_seq + "\t" + u.reverse_seq + "\t" + std::to_string(u.forward_positions[0].pos) + "," + std::to_string(u.forward_positions[1].pos) +
                      "\t" + std::to_string(u.reverse_positions[0].pos) + "," + std::to_string(u.reverse_positions[1].pos);
    return dup_out(out);
    ORC_CATCH(nullptr)
}

// unitig.rs:410-440: build a unitig from three consecutive k-mers of a sequence; returns "fwd\trev\ttrimmed_fwd\ttrimmed_rev"
char* orc_unitig_from_kmers(const char* seq, int k, int first_fwd_start) {
    ORC_TRY
    Sequence s = one_seq(seq, k);
    std::vector<Kmer> f(3), r(3);
    size_t plen = s.forward_seq.size();
    for (int i = 0; i < 3; ++i) {
        f[i].pointer = s.forward_seq.data() + first_fwd_start + i; f[i].length = (size_t)k;
        r[i].pointer = s.reverse_seq.data() + (plen - (size_t)k - (size_t)(first_fwd_start + i)); r[i].length = (size_t)k;
        f[i].positions.emplace_back(1, true, 0); r[i].positions.emplace_back(1, false, 0);
    }
    Unitig u = Unitig::from_kmers(123, &f[1], &r[1]);
    u.add_kmer_to_start(&f[0], &r[0]); u.add_kmer_to_end(&f[2], &r[2]);
    u.simplify_seqs();
    std::string out = u.forward_seq + "\t" + u.reverse_seq;
    u.trim_overlaps((size_t)k);
    out += "\t" + u.forward_seq + "\t" + u.reverse_seq;
    return dup_out(out);
    ORC_CATCH(nullptr)
}

char* orc_position_display(int seq_id, int strand, uint64_t pos) { return dup_out(Position((uint16_t)seq_id, strand != 0, pos).to_string()); }
char* orc_reverse_complement(const char* s) { return dup_out(reverse_complement(s)); }

}  // extern "C"

#ifdef ORC_MAIN
// autocycler-oracle compress -i <assemblies_dir> -a <autocycler_dir> [--kmer 51] [--max_contigs 25] [-t 8]
#include <fstream>
#include <sys/stat.h>
int main(int argc, char** argv) {
    std::string in, out; int k = 51, maxc = 25, threads = 8;
    if (argc < 2 || std::string(argv[1]) != "compress") { fprintf(stderr, "usage: autocycler-oracle compress -i DIR -a DIR [--kmer K] [--max_contigs N] [-t T]\n"); return 2; }
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
        if (a == "-i" || a == "--assemblies_dir") in = val();
        else if (a == "-a" || a == "--autocycler_dir") out = val();
        else if (a == "--kmer") k = atoi(val().c_str());
        else if (a == "--max_contigs") maxc = atoi(val().c_str());
        else if (a == "-t" || a == "--threads") threads = atoi(val().c_str());
    }
    mkdir(out.c_str(), 0777);
    orc_stats st{}; char* yaml = nullptr;
    char* gfa = orc_compress_dir(in.c_str(), k, maxc, threads, &yaml, &st);
    if (!gfa) { fprintf(stderr, "\nError: %s\n", orc_last_error()); return 1; }
    { std::ofstream f(out + "/input_assemblies.gfa", std::ios::binary); f << gfa; }
    { std::ofstream f(out + "/input_assemblies.yaml", std::ios::binary); f << yaml; }
    fprintf(stderr, "Graph contains %llu k-mers\n%llu unitigs, %llu links -> %llu unitigs, %llu links\n"
            "load %.3fs  kmer_graph %.3fs  unitig_graph %.3fs  simplify %.3fs  gfa %.3fs\n",
            (unsigned long long)st.n_kmers, (unsigned long long)st.unitigs_before, (unsigned long long)st.links_before,
            (unsigned long long)st.unitigs_after, (unsigned long long)st.links_after,
            st.t_load, st.t_kmer_graph, st.t_unitig_graph, st.t_simplify, st.t_gfa);
    return 0;
}
#endif
