// Host side of the compress path (see host_graph.h).  Citations are file:line in the reference's src/.
#include "host_graph.h"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstring>
#include <stdexcept>

static inline char comp(char c) {   // misc.rs:324-333 (unitig sequences hold only ACGT after trimming)
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c == '.' ? '.' : 'N'; }
}

static inline void append_uint(std::string& s, uint64_t v) {
    char buf[24]; auto r = std::to_chars(buf, buf + sizeof buf, v); s.append(buf, r.ptr - buf);
}

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
namespace {
struct SeedKey { uint64_t w[AC_MAX_W]; int32_t d; uint32_t dev; };
inline bool seed_less(const SeedKey& a, const SeedKey& b) {   // byte order of the k-mer text, '.' < A < C < G < T (kmer_key.h key_less5)
    int la = a.d > 0 ? a.d : 0, lb = b.d > 0 ? b.d : 0;
    if (la != lb) return la > lb;
    for (int j = 0; j < AC_MAX_W; ++j) if (a.w[j] != b.w[j]) return a.w[j] < b.w[j];
    int ta = a.d < 0 ? -a.d : 0, tb = b.d < 0 ? -b.d : 0;
    return ta > tb;
}
}  // namespace

void HostGraph::build(const PipelineResult& r, const std::vector<HostSeq>& seqs, const uint8_t* ascii, uint32_t k_size,
                      bool keep_positions) {
    k = k_size;
    const uint32_t h = k / 2;
    const uint32_t U = (uint32_t)r.unitigs.size();
    units.clear(); units.resize(U); order.resize(U); paths.clear(); fixed_ready = false;

    // Seed order (kmer_graph.rs:168-173 + unitig_graph.rs:179-185): ascending smallest k-mer of both strands.
    // min_w holds W significant words; the unused high slots are zero for every unitig, so comparing all is safe.
    std::vector<SeedKey> keys(U);
    for (uint32_t j = 0; j < U; ++j) {
        SeedKey& sk = keys[j];
        for (int w = 0; w < AC_MAX_W; ++w) sk.w[w] = r.unitigs[j].min_w[w];
        sk.d = r.unitigs[j].min_d; sk.dev = j;
    }
    std::sort(keys.begin(), keys.end(), seed_less);
    std::vector<uint32_t> rank(U);
    for (uint32_t s = 0; s < U; ++s) rank[keys[s].dev] = s;

    // Sequences (unitig.rs:120-133 + 157-165): the trimmed unitig is the centre base of each of its k-mers,
    // i.e. a slice of the padded input shifted by k/2.
    for (uint32_t j = 0; j < U; ++j) {
        const DeviceUnitig& d = r.unitigs[j];
        HostUnitig& u = units[rank[j]];
        const char* src = (const char*)ascii + d.start + h;
        u.seq.resize(d.len);
        if (!d.flip) memcpy(&u.seq[0], src, d.len);
        else for (uint32_t i = 0; i < d.len; ++i) u.seq[i] = comp(src[d.len - 1 - i]);
        u.depth = d.depth;
    }

    // Links.  Device strand e (0 = direction of the representative occurrence) -> unitig strand.
    auto to_strand = [&](uint32_t dev_strand) -> UStrand {
        const uint32_t j = dev_strand >> 1, e = dev_strand & 1;
        return us_make(rank[j], (e ^ r.unitigs[j].flip) != 0);
    };
    std::vector<UStrand> tmp;
    for (uint32_t j = 0; j < U; ++j) {
        for (uint32_t e = 0; e < 2; ++e) {
            const uint32_t i = 2 * j + e, n = r.link_count[i];
            if (n > AC_MAX_LINKS) throw std::runtime_error("link overflow");
            const UStrand from = to_strand(i);
            const uint32_t a = us_index(from);
            tmp.clear();
            for (uint32_t x = 0; x < n; ++x) tmp.push_back(to_strand(r.links[(size_t)i * AC_MAX_LINKS + x]));
            std::vector<UStrand>& out = units[a].next[us_reverse(from) ? 1 : 0];
            if (!us_reverse(from)) {
                // forward_next: all b+ ascending, then all b- ascending (unitig_graph.rs:255-275 blocks 1 and 2 of iteration a)
                std::sort(tmp.begin(), tmp.end(), [](UStrand x, UStrand y) {
                    if (us_reverse(x) != us_reverse(y)) return !us_reverse(x);
                    return us_index(x) < us_index(y); });
                out = tmp;
            } else {
                // reverse_next: x- pushed by block 1 of iteration x (x < a), then iteration a: a- (self loop) and all b+
                // ascending (block 3), then x- for x > a (unitig_graph.rs:262-264, 277-285)
                auto phase = [a](UStrand t) { if (us_reverse(t)) return us_index(t) < a ? 0 : (us_index(t) == a ? 1 : 3); return 2; };
                std::sort(tmp.begin(), tmp.end(), [&](UStrand x, UStrand y) {
                    int px = phase(x), py = phase(y);
                    if (px != py) return px < py;
                    return us_index(x) < us_index(y); });
                out = tmp;
            }
        }
    }
    // prev lists are the mirror image: (a,s)->(b,t)  <=>  (b,t).prev has (a,s).  Push order as in create_links.
    for (uint32_t a = 0; a < U; ++a) {
        // forward_prev(b) receives a+ (block 1) then a- (block 3) during iteration a
        for (UStrand t : units[a].next[0]) if (!us_reverse(t)) units[us_index(t)].prev[0].push_back(us_make(a, false));
        // reverse_prev(a) receives b- for every a+ -> b+ (block 1), then block 2 pushes a+ into reverse_prev(b) for a+ -> b-
        for (UStrand t : units[a].next[0]) if (!us_reverse(t)) units[a].prev[1].push_back(us_make(us_index(t), true));
        for (UStrand t : units[a].next[0]) if (us_reverse(t)) units[us_index(t)].prev[1].push_back(us_make(a, false));
        for (UStrand t : units[a].next[1]) if (!us_reverse(t)) units[us_index(t)].prev[0].push_back(us_make(a, true));
    }

    // Paths and positions from the occurrences.  An occurrence [fs, fs+n) on the forward strand of sequence i is
    // also an occurrence of the opposite unitig strand at L-fs-n on its reverse strand (kmer_graph.rs:103-108);
    // forward_positions / reverse_positions are those of the first k-mer of each unitig strand (unitig.rs:135-146).
    paths.resize(seqs.size());
    size_t si = 0;
    const size_t R = r.run_start.size();
    for (size_t x = 0; x < R; ++x) {
        const uint64_t g = r.run_start[x];
        while (si + 1 < seqs.size() && seqs[si + 1].start <= g) ++si;
        const HostSeq& s = seqs[si];
        const uint32_t dev = r.run_unitig[x] >> 1, same = r.run_unitig[x] & 1;
        const uint32_t idx = rank[dev];
        const bool plus = (same ^ r.unitigs[dev].flip) != 0;
        paths[si].push_back(us_make(idx, !plus));
        const uint32_t fs = (uint32_t)(g - s.start), n = r.run_len[x];
        const uint32_t mirrored = (uint32_t)(s.length - fs - n);
        HostUnitig& u = units[idx];
        const uint32_t on_fwd_strand_pos = plus ? fs : mirrored;     // where the unitig's forward strand starts (on seq + if plus, seq - otherwise)
        const uint32_t on_rev_strand_pos = plus ? mirrored : fs;
        u.min_fpos = std::min(u.min_fpos, on_fwd_strand_pos);
        u.min_rpos = std::min(u.min_rpos, on_rev_strand_pos);
        if (keep_positions) {
            const uint64_t fwd_tag = (uint64_t)(s.id | (plus ? 0x8000u : 0u)), rev_tag = (uint64_t)(s.id | (plus ? 0u : 0x8000u));
            u.fpos.push_back(((uint64_t)on_fwd_strand_pos << 16) | fwd_tag);
            u.rpos.push_back(((uint64_t)on_rev_strand_pos << 16) | rev_tag);
        }
    }

    for (uint32_t s = 0; s < U; ++s) { order[s] = s; units[s].number = s + 1; }
    renumber();
}

// ------------------------------------------------------------------------------------------------
// renumber / checks / counts
// ------------------------------------------------------------------------------------------------
void HostGraph::renumber() {   // unitig_graph.rs:295-315: stable sort by length desc, sequence asc, depth desc
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        const HostUnitig& a = units[x]; const HostUnitig& b = units[y];
        if (a.seq.size() != b.seq.size()) return a.seq.size() > b.seq.size();
        int c = a.seq.compare(b.seq);
        if (c != 0) return c < 0;
        return a.depth > b.depth;
    });
    for (uint32_t n = 0; n < order.size(); ++n) units[order[n]].number = n + 1;
}

void HostGraph::check_links() const {   // unitig_graph.rs:752-793: every link has its mirror and its prev entry
    auto has = [](const std::vector<UStrand>& v, UStrand x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    for (uint32_t a = 0; a < units.size(); ++a)
        for (int s = 0; s < 2; ++s) {
            const UStrand from = us_make(a, s == 1);
            for (UStrand to : units[a].next[s]) {
                if (!has(units[us_index(to)].prev[us_reverse(to) ? 1 : 0], from)) throw std::runtime_error("missing prev link");
                if (!has(units[us_index(to)].next[us_reverse(to) ? 0 : 1], us_flip(from))) throw std::runtime_error("missing next link");
            }
            for (UStrand p : units[a].prev[s])
                if (!has(units[us_index(p)].next[us_reverse(p) ? 1 : 0], from)) throw std::runtime_error("missing next link");
        }
}

uint64_t HostGraph::total_length() const { uint64_t t = 0; for (auto& u : units) t += u.seq.size(); return t; }

uint64_t HostGraph::link_count_single() const {   // unitig_graph.rs:478-507: a link and its mirror count once; hairpins are their own mirror
    uint64_t all = 0, hairpins = 0;
    for (uint32_t a = 0; a < units.size(); ++a)
        for (int s = 0; s < 2; ++s)
            for (UStrand to : units[a].next[s]) { ++all; if (us_index(to) == a && us_reverse(to) != (s == 1)) ++hairpins; }
    return (all - hairpins) / 2 + hairpins;
}

// ------------------------------------------------------------------------------------------------
// graph_simplification.rs:26-312
// ------------------------------------------------------------------------------------------------
namespace {
struct StrandView {   // UnitigStrand::get_seq without materialising the reverse complement
    const std::string* s; bool rev;
    size_t size() const { return s->size(); }
    char from_start(size_t i) const { return rev ? comp((*s)[s->size() - 1 - i]) : (*s)[i]; }
    char from_end(size_t i) const { return rev ? comp((*s)[i]) : (*s)[s->size() - 1 - i]; }
};
}  // namespace

void HostGraph::compute_fixed() {   // graph_simplification.rs:190-230; paths and links never change during simplification
    const size_t U = units.size();
    fixed_start.assign(U, 0); fixed_end.assign(U, 0);
    for (auto& path : paths) {
        if (path.empty()) continue;
        const UStrand first = path.front(), last = path.back();
        if (!us_reverse(first)) fixed_start[us_index(first)] = 1; else fixed_end[us_index(first)] = 1;
        if (!us_reverse(last)) fixed_end[us_index(last)] = 1; else fixed_start[us_index(last)] = 1;
    }
    const std::vector<uint8_t> starts_copy = fixed_start, ends_copy = fixed_end;
    for (size_t u = 0; u < U; ++u) {
        if (starts_copy[u])
            for (UStrand up : units[u].prev[0]) { if (!us_reverse(up)) fixed_end[us_index(up)] = 1; else fixed_start[us_index(up)] = 1; }
        if (ends_copy[u])
            for (UStrand down : units[u].next[0]) { if (!us_reverse(down)) fixed_start[us_index(down)] = 1; else fixed_end[us_index(down)] = 1; }
    }
    fixed_ready = true;
}

size_t HostGraph::expand_repeats() {   // graph_simplification.rs:43-86
    if (!fixed_ready) compute_fixed();
    size_t total_shifted = 0;
    std::vector<UStrand> group;
    std::string common;
    for (uint32_t idx : order) {
        HostUnitig& u = units[idx];
        const UStrand self_fwd = us_make(idx, false);

        // get_exclusive_inputs, :233-255
        group.clear();
        bool ok = true;
        for (UStrand prev : u.prev[0]) {
            const std::vector<UStrand>& nx = units[us_index(prev)].next[us_reverse(prev) ? 1 : 0];
            if (!(nx.size() == 1 && nx[0] == self_fwd)) { ok = false; break; }
            group.push_back(prev);
        }
        if (ok) for (UStrand g : group) if (us_index(g) == idx) { ok = false; break; }
        if (ok && group.size() >= 2 && !fixed_start[idx]) {
            bool can_shift = true;
            for (UStrand in : group)
                if ((!us_reverse(in) && fixed_end[us_index(in)]) || (us_reverse(in) && fixed_start[us_index(in)])) { can_shift = false; break; }
            if (can_shift) {   // shift_sequence_1, :89-116
                // get_common_end_seq, :298-312
                StrandView first{&units[us_index(group[0])].seq, us_reverse(group[0])};
                size_t len = first.size();
                bool dup = false; uint32_t min_len = 0xFFFFFFFFu;
                for (size_t a = 0; a < group.size(); ++a) {
                    StrandView v{&units[us_index(group[a])].seq, us_reverse(group[a])};
                    len = std::min(len, v.size());
                    size_t m = 0; while (m < len && v.from_end(m) == first.from_end(m)) ++m;
                    len = m;
                    min_len = std::min<uint32_t>(min_len, (uint32_t)v.size());
                    for (size_t b = 0; b < a; ++b) if (us_index(group[a]) == us_index(group[b])) dup = true;
                }
                // avoid_zero_len_unitigs (:141-158) and avoid_start_of_path (:161-181) both trim from the start
                if (len > 0) len = std::min<size_t>(len, (min_len - 1) / (dup ? 2 : 1));
                if (len > 0) len = u.min_fpos == 0 ? 0 : std::min<size_t>(len, u.min_fpos - 1);
                if (len > 0) {
                    common.resize(len);
                    for (size_t i = 0; i < len; ++i) common[len - 1 - i] = first.from_end(i);
                    for (UStrand src : group) {
                        HostUnitig& s = units[us_index(src)];
                        if (!us_reverse(src)) { s.min_rpos += (uint32_t)len; s.seq.resize(s.seq.size() - len); }      // remove_seq_from_end, unitig.rs:225-232
                        else { s.min_fpos += (uint32_t)len; s.seq.erase(0, len); }                                     // remove_seq_from_start, unitig.rs:216-223
                    }
                    u.min_fpos -= (uint32_t)len; u.seq.insert(0, common);                                               // add_seq_to_start, unitig.rs:234-240
                    total_shifted += len;
                }
            }
        }

        // get_exclusive_outputs, :258-280
        group.clear();
        ok = true;
        for (UStrand next : u.next[0]) {
            const std::vector<UStrand>& pv = units[us_index(next)].prev[us_reverse(next) ? 1 : 0];
            if (!(pv.size() == 1 && pv[0] == self_fwd)) { ok = false; break; }
            group.push_back(next);
        }
        if (ok) for (UStrand g : group) if (us_index(g) == idx) { ok = false; break; }
        if (ok && group.size() >= 2 && !fixed_end[idx]) {
            bool can_shift = true;
            for (UStrand o : group)
                if ((!us_reverse(o) && fixed_start[us_index(o)]) || (us_reverse(o) && fixed_end[us_index(o)])) { can_shift = false; break; }
            if (can_shift) {   // shift_sequence_2, :119-138
                StrandView first{&units[us_index(group[0])].seq, us_reverse(group[0])};
                size_t len = first.size();
                bool dup = false; uint32_t min_len = 0xFFFFFFFFu;
                for (size_t a = 0; a < group.size(); ++a) {
                    StrandView v{&units[us_index(group[a])].seq, us_reverse(group[a])};
                    len = std::min(len, v.size());
                    size_t m = 0; while (m < len && v.from_start(m) == first.from_start(m)) ++m;
                    len = m;
                    min_len = std::min<uint32_t>(min_len, (uint32_t)v.size());
                    for (size_t b = 0; b < a; ++b) if (us_index(group[a]) == us_index(group[b])) dup = true;
                }
                if (len > 0) len = std::min<size_t>(len, (min_len - 1) / (dup ? 2 : 1));
                if (len > 0) len = u.min_rpos == 0 ? 0 : std::min<size_t>(len, u.min_rpos - 1);
                if (len > 0) {
                    common.resize(len);
                    for (size_t i = 0; i < len; ++i) common[i] = first.from_start(i);
                    for (UStrand src : group) {
                        HostUnitig& s = units[us_index(src)];
                        if (!us_reverse(src)) { s.min_fpos += (uint32_t)len; s.seq.erase(0, len); }
                        else { s.min_rpos += (uint32_t)len; s.seq.resize(s.seq.size() - len); }
                    }
                    u.min_rpos -= (uint32_t)len; u.seq.append(common);                                                  // add_seq_to_end, unitig.rs:242-248
                    total_shifted += len;
                }
            }
        }
    }
    return total_shifted;
}

void HostGraph::simplify_structure() {   // graph_simplification.rs:26-40
    while (expand_repeats() > 0) {}
    renumber();
}

// ------------------------------------------------------------------------------------------------
// unitig_graph.rs:317-360 save_gfa
// ------------------------------------------------------------------------------------------------
std::string HostGraph::gfa_text(const std::vector<HostSeq>& seqs) const {
    std::string out;
    size_t est = 64 + total_length();
    for (auto& u : units) est += 32 + 16 * (u.next[0].size() + u.next[1].size());
    for (auto& p : paths) est += 128 + 10 * p.size();
    out.reserve(est);
    out += "H\tVN:Z:1.0\tKM:i:"; append_uint(out, k); out += '\n';
    for (uint32_t idx : order) {   // unitig.rs:167-171; depth is integral here so {:.2} renders as N.00
        const HostUnitig& u = units[idx];
        out += "S\t"; append_uint(out, u.number); out += '\t'; out += u.seq; out += "\tDP:f:"; append_uint(out, u.depth); out += ".00\n";
    }
    for (uint32_t idx : order) {   // get_links_for_gfa, :333-350
        const HostUnitig& a = units[idx];
        for (int s = 0; s < 2; ++s)
            for (UStrand b : a.next[s]) {
                out += "L\t"; append_uint(out, a.number); out += s == 0 ? "\t+\t" : "\t-\t";
                append_uint(out, units[us_index(b)].number); out += us_reverse(b) ? "\t-\t0M\n" : "\t+\t0M\n";
            }
    }
    for (size_t i = 0; i < seqs.size(); ++i) {   // get_gfa_path_line, :352-360
        const HostSeq& s = seqs[i];
        out += "P\t"; append_uint(out, s.id); out += '\t';
        const std::vector<UStrand>& path = paths[i];
        for (size_t x = 0; x < path.size(); ++x) {
            if (x) out += ',';
            append_uint(out, units[us_index(path[x])].number); out += us_reverse(path[x]) ? '-' : '+';
        }
        out += "\t*\tLN:i:"; append_uint(out, s.length); out += "\tFN:Z:"; out += s.filename; out += "\tHD:Z:"; out += s.contig_header; out += '\n';
    }
    return out;
}
