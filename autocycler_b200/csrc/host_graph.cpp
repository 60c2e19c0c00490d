// Host side of the compress path (see host_graph.h).  Citations are file:line in the reference's src/.
#include "host_graph.h"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace {
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline char comp(char c) {   // misc.rs:324-333 (unitig sequences hold only ACGT after trimming)
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c == '.' ? '.' : 'N'; }
}

struct SeedKey { uint64_t w[AC_MAX_W]; int32_t d; uint32_t dev; };
inline bool seed_less(const SeedKey& a, const SeedKey& b) {   // byte order of the k-mer text, '.' < A < C < G < T (kmer_key.h key_less5)
    const int la = a.d > 0 ? a.d : 0, lb = b.d > 0 ? b.d : 0;
    if (la != lb) return la > lb;
    for (int j = 0; j < AC_MAX_W; ++j) if (a.w[j] != b.w[j]) return a.w[j] < b.w[j];
    const int ta = a.d < 0 ? -a.d : 0, tb = b.d < 0 ? -b.d : 0;
    return ta > tb;
}

const uint32_t SLACK = 32;   // spare bytes on each side of every unitig in the arena
}  // namespace

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
void HostGraph::build(const PipelineResult& r, const std::vector<HostSeq>& seqs, const uint8_t* ascii, uint32_t k_size,
                      bool keep_positions) {
    prof = HostProfile();
    double t0 = now_ms();
    k = k_size;
    const uint32_t h = k / 2;
    U = (uint32_t)r.unitigs.size();
    fixed_ready = false;

    // Seed order (kmer_graph.rs:168-173 + unitig_graph.rs:179-185): ascending smallest k-mer of both strands.
    // min_w holds W significant words; the unused slots are zero for every unitig, so comparing all is safe.
    std::vector<SeedKey> keys(U);
    for (uint32_t j = 0; j < U; ++j) {
        SeedKey& sk = keys[j];
        for (int w = 0; w < AC_MAX_W; ++w) sk.w[w] = r.unitigs[j].min_w[w];
        sk.d = r.unitigs[j].min_d; sk.dev = j;
    }
    std::sort(keys.begin(), keys.end(), seed_less);
    std::vector<uint32_t> rank(U);
    for (uint32_t s = 0; s < U; ++s) rank[keys[s].dev] = s;
    double t1 = now_ms(); prof.seed_sort = t1 - t0;

    // Sequences (unitig.rs:120-133 + 157-165): the trimmed unitig is the centre base of each of its k-mers, i.e. a
    // slice of the padded input shifted by k/2 (reverse-complemented when the seed k-mer lies on the other strand).
    number.assign(U, 0); depth.resize(U); len.resize(U); seq_off.resize(U);
    room_before.assign(U, SLACK); room_after.assign(U, SLACK);
    min_fpos.assign(U, 0xFFFFFFFFu); min_rpos.assign(U, 0xFFFFFFFFu);
    uint64_t arena_bytes = 0;
    for (uint32_t s = 0; s < U; ++s) {
        const DeviceUnitig& d = r.unitigs[keys[s].dev];
        seq_off[s] = arena_bytes + SLACK; len[s] = d.len; depth[s] = d.depth;
        arena_bytes += (uint64_t)d.len + 2 * SLACK;
    }
    arena.resize(arena_bytes);
    for (uint32_t s = 0; s < U; ++s) {
        const DeviceUnitig& d = r.unitigs[keys[s].dev];
        const char* src = (const char*)ascii + d.start + h;
        char* dst = arena.data() + seq_off[s];
        if (!d.flip) memcpy(dst, src, d.len);
        else for (uint32_t i = 0; i < d.len; ++i) dst[i] = comp(src[d.len - 1 - i]);
    }
    double t2 = now_ms(); prof.seqs = t2 - t1;

    // Links.  Device strand e (0 = direction of the representative occurrence) -> unitig strand.
    std::vector<uint8_t> flip(U);
    for (uint32_t j = 0; j < U; ++j) flip[j] = (uint8_t)r.unitigs[j].flip;
    auto to_strand = [&](uint32_t dev_strand) -> UStrand {
        const uint32_t j = dev_strand >> 1, e = dev_strand & 1;
        return us_make(rank[j], (e ^ flip[j]) != 0);
    };
    next_off.assign(2 * (size_t)U + 1, 0); prev_off.assign(2 * (size_t)U + 1, 0);
    for (uint32_t i = 0; i < 2 * U; ++i) {
        if (r.link_count[i] > AC_MAX_LINKS) throw std::runtime_error("link overflow");
        next_off[to_strand(i) + 1] = r.link_count[i];
    }
    for (size_t s = 0; s < 2 * (size_t)U; ++s) next_off[s + 1] += next_off[s];
    next.resize(next_off[2 * (size_t)U]);
    for (uint32_t i = 0; i < 2 * U; ++i) {
        const UStrand from = to_strand(i);
        const uint32_t n = r.link_count[i], a = us_index(from);
        UStrand* out = next.data() + next_off[from];
        for (uint32_t x = 0; x < n; ++x) out[x] = to_strand(r.links[(size_t)i * AC_MAX_LINKS + x]);
        if (n < 2) continue;
        if (!us_reverse(from)) {
            // forward_next: all b+ ascending, then all b- ascending (unitig_graph.rs:255-275, blocks 1 and 2 of iteration a)
            std::sort(out, out + n, [](UStrand x, UStrand y) {
                if (us_reverse(x) != us_reverse(y)) return !us_reverse(x);
                return us_index(x) < us_index(y); });
        } else {
            // reverse_next: x- pushed by block 1 of iteration x (x < a); iteration a pushes a- (self loop) then all b+
            // ascending (block 3); then x- for x > a (unitig_graph.rs:262-264, 277-285)
            auto phase = [a](UStrand t) { if (us_reverse(t)) return us_index(t) < a ? 0 : (us_index(t) == a ? 1 : 3); return 2; };
            std::sort(out, out + n, [&](UStrand x, UStrand y) {
                const int px = phase(x), py = phase(y);
                if (px != py) return px < py;
                return us_index(x) < us_index(y); });
        }
    }
    // prev lists mirror next lists: (a,s)->(b,t) puts (a,s) into prev(b,t).  Only membership matters downstream.
    for (UStrand t : next) prev_off[t + 1] += 1;
    for (size_t s = 0; s < 2 * (size_t)U; ++s) prev_off[s + 1] += prev_off[s];
    prev.resize(next.size());
    {
        std::vector<uint32_t> cursor(prev_off.begin(), prev_off.end() - 1);
        for (UStrand from = 0; from < 2 * U; ++from)
            for (uint32_t x = next_off[from]; x < next_off[from + 1]; ++x) prev[cursor[next[x]]++] = from;
    }
    double t3 = now_ms(); prof.links = t3 - t2;

    // Paths and positions from the occurrences.  An occurrence [fs, fs+n) on the forward strand of sequence i is also
    // an occurrence of the opposite unitig strand at L-fs-n on its reverse strand (kmer_graph.rs:103-108);
    // forward_positions / reverse_positions are those of the first k-mer of each unitig strand (unitig.rs:135-146).
    const size_t R = r.run_start.size(), S = seqs.size();
    path_off.assign(S + 1, 0); path.resize(R);
    std::vector<uint32_t> pos_count;
    if (keep_positions) pos_count.assign(U, 0);
    size_t si = 0;
    for (size_t x = 0; x < R; ++x) {
        const uint64_t g = r.run_start[x];
        while (si + 1 < S && seqs[si + 1].start <= g) ++si;
        const HostSeq& s = seqs[si];
        const uint32_t dev = r.run_unitig[x] >> 1, same = r.run_unitig[x] & 1;
        const uint32_t idx = rank[dev];
        const bool plus = (same ^ flip[dev]) != 0;
        path[x] = us_make(idx, !plus);
        path_off[si + 1] = x + 1;
        const uint32_t fs = (uint32_t)(g - s.start), n = r.run_len[x];
        const uint32_t mirrored = (uint32_t)(s.length - fs - n);
        const uint32_t f = plus ? fs : mirrored, rv = plus ? mirrored : fs;
        if (f < min_fpos[idx]) min_fpos[idx] = f;
        if (rv < min_rpos[idx]) min_rpos[idx] = rv;
        if (keep_positions) pos_count[idx] += 1;
    }
    for (size_t i = 1; i <= S; ++i) if (path_off[i] < path_off[i - 1]) path_off[i] = path_off[i - 1];
    fpos_off.clear(); rpos_off.clear(); fpos.clear(); rpos.clear();
    if (keep_positions) {
        fpos_off.assign((size_t)U + 1, 0);
        for (uint32_t u = 0; u < U; ++u) fpos_off[u + 1] = fpos_off[u] + pos_count[u];
        rpos_off = fpos_off; fpos.resize(R); rpos.resize(R);
        std::vector<uint64_t> cursor(fpos_off.begin(), fpos_off.end() - 1);
        si = 0;
        for (size_t x = 0; x < R; ++x) {
            const uint64_t g = r.run_start[x];
            while (si + 1 < S && seqs[si + 1].start <= g) ++si;
            const HostSeq& s = seqs[si];
            const UStrand p = path[x]; const bool plus = !us_reverse(p); const uint32_t idx = us_index(p);
            const uint32_t fs = (uint32_t)(g - s.start), n = r.run_len[x], mirrored = (uint32_t)(s.length - fs - n);
            const uint64_t at = cursor[idx]++;
            fpos[at] = ((uint64_t)(plus ? fs : mirrored) << 16) | (uint64_t)(s.id | (plus ? 0x8000u : 0u));
            rpos[at] = ((uint64_t)(plus ? mirrored : fs) << 16) | (uint64_t)(s.id | (plus ? 0u : 0x8000u));
        }
    }
    double t4 = now_ms(); prof.paths = t4 - t3;

    order.resize(U);
    for (uint32_t s = 0; s < U; ++s) order[s] = s;
    renumber();
}

// ------------------------------------------------------------------------------------------------
// renumber / checks / counts
// ------------------------------------------------------------------------------------------------
void HostGraph::renumber() {   // unitig_graph.rs:295-315: stable sort by length desc, sequence asc, depth desc
    const double t0 = now_ms();
    struct Key { uint32_t len; uint32_t pos; uint64_t prefix; uint32_t idx; uint32_t depth; };
    std::vector<Key> keys(U);
    for (uint32_t n = 0; n < U; ++n) {
        const uint32_t idx = order[n];
        Key& key = keys[n];
        key.len = len[idx]; key.pos = n; key.idx = idx; key.depth = depth[idx];
        const unsigned char* p = (const unsigned char*)seq_ptr(idx);
        uint64_t v = 0;
        const uint32_t m = key.len < 8 ? key.len : 8;
        for (uint32_t i = 0; i < m; ++i) v |= (uint64_t)p[i] << (56 - 8 * i);
        key.prefix = v;
    }
    std::sort(keys.begin(), keys.end(), [&](const Key& a, const Key& b) {
        if (a.len != b.len) return a.len > b.len;
        if (a.prefix != b.prefix) return a.prefix < b.prefix;
        if (a.len > 8) { const int c = memcmp(seq_ptr(a.idx) + 8, seq_ptr(b.idx) + 8, a.len - 8); if (c != 0) return c < 0; }
        if (a.depth != b.depth) return a.depth > b.depth;
        return a.pos < b.pos;   // ties keep their previous order: slice::sort_by is stable
    });
    for (uint32_t n = 0; n < U; ++n) { order[n] = keys[n].idx; number[keys[n].idx] = n + 1; }
    prof.renumber += now_ms() - t0;
}

void HostGraph::check_links() const {   // unitig_graph.rs:752-793: every link has its mirror and its prev entry
    auto has = [](const UStrand* b, uint32_t n, UStrand x) { for (uint32_t i = 0; i < n; ++i) if (b[i] == x) return true; return false; };
    for (UStrand from = 0; from < 2 * U; ++from) {
        for (uint32_t x = next_off[from]; x < next_off[from + 1]; ++x) {
            const UStrand to = next[x];
            if (!has(prev_begin(to), prev_size(to), from)) throw std::runtime_error("missing prev link");
            if (!has(next_begin(us_flip(to)), next_size(us_flip(to)), us_flip(from))) throw std::runtime_error("missing next link");
        }
        for (uint32_t x = prev_off[from]; x < prev_off[from + 1]; ++x)
            if (!has(next_begin(prev[x]), next_size(prev[x]), from)) throw std::runtime_error("missing next link");
    }
}

uint64_t HostGraph::total_length() const { uint64_t t = 0; for (uint32_t u = 0; u < U; ++u) t += len[u]; return t; }

uint64_t HostGraph::link_count_single() const {   // unitig_graph.rs:478-507: a link and its mirror count once; hairpins are their own mirror
    uint64_t all = next.size(), hairpins = 0;
    for (UStrand from = 0; from < 2 * U; ++from)
        for (uint32_t x = next_off[from]; x < next_off[from + 1]; ++x) if (next[x] == us_flip(from)) ++hairpins;
    return (all - hairpins) / 2 + hairpins;
}

// ------------------------------------------------------------------------------------------------
// arena growth
// ------------------------------------------------------------------------------------------------
void HostGraph::relocate(uint32_t idx, uint32_t before, uint32_t after) {
    const uint64_t old_off = seq_off[idx], at = arena.size();
    arena.resize(at + before + len[idx] + after);
    memcpy(arena.data() + at + before, arena.data() + old_off, len[idx]);
    seq_off[idx] = at + before; room_before[idx] = before; room_after[idx] = after;
}
void HostGraph::grow_front(uint32_t idx, uint32_t need) { if (room_before[idx] < need) relocate(idx, need + 4 * SLACK, room_after[idx] < SLACK ? SLACK : room_after[idx]); }
void HostGraph::grow_back(uint32_t idx, uint32_t need) { if (room_after[idx] < need) relocate(idx, room_before[idx] < SLACK ? SLACK : room_before[idx], need + 4 * SLACK); }

// ------------------------------------------------------------------------------------------------
// graph_simplification.rs:26-312
// ------------------------------------------------------------------------------------------------
void HostGraph::compute_fixed() {   // graph_simplification.rs:190-230; paths and links never change during simplification
    fixed_start.assign(U, 0); fixed_end.assign(U, 0);
    for (size_t i = 0; i + 1 < path_off.size(); ++i) {
        if (path_off[i + 1] == path_off[i]) continue;
        const UStrand first = path[path_off[i]], last = path[path_off[i + 1] - 1];
        if (!us_reverse(first)) fixed_start[us_index(first)] = 1; else fixed_end[us_index(first)] = 1;
        if (!us_reverse(last)) fixed_end[us_index(last)] = 1; else fixed_start[us_index(last)] = 1;
    }
    const std::vector<uint8_t> starts_copy = fixed_start, ends_copy = fixed_end;
    for (uint32_t u = 0; u < U; ++u) {
        if (starts_copy[u]) {
            const UStrand s = us_make(u, false);
            for (uint32_t x = prev_off[s]; x < prev_off[s + 1]; ++x) { const UStrand up = prev[x]; if (!us_reverse(up)) fixed_end[us_index(up)] = 1; else fixed_start[us_index(up)] = 1; }
        }
        if (ends_copy[u]) {
            const UStrand s = us_make(u, false);
            for (uint32_t x = next_off[s]; x < next_off[s + 1]; ++x) { const UStrand down = next[x]; if (!us_reverse(down)) fixed_start[us_index(down)] = 1; else fixed_end[us_index(down)] = 1; }
        }
    }
    fixed_ready = true;
}

size_t HostGraph::expand_repeats() {   // graph_simplification.rs:43-86
    const double t0 = now_ms();
    if (!fixed_ready) compute_fixed();
    size_t total_shifted = 0;
    std::string common;
    // character i counted from the start / from the end of a unitig strand (UnitigStrand::get_seq without the copy)
    auto from_start = [&](UStrand s, size_t i) { const uint32_t u = us_index(s); const char* p = seq_ptr(u); return us_reverse(s) ? comp(p[len[u] - 1 - i]) : p[i]; };
    auto from_end = [&](UStrand s, size_t i) { const uint32_t u = us_index(s); const char* p = seq_ptr(u); return us_reverse(s) ? comp(p[i]) : p[len[u] - 1 - i]; };

    for (uint32_t n = 0; n < U; ++n) {
        const uint32_t idx = order[n];
        const UStrand self_fwd = us_make(idx, false);

        // ---- get_exclusive_inputs (:233-255) + shift_sequence_1 (:89-116) ----
        {
            const UStrand* grp = prev_begin(self_fwd); const uint32_t gn = prev_size(self_fwd);
            bool ok = gn >= 2 && !fixed_start[idx];
            for (uint32_t a = 0; ok && a < gn; ++a) {
                const UStrand p = grp[a];
                if (!(next_size(p) == 1 && next_begin(p)[0] == self_fwd) || us_index(p) == idx) ok = false;
                else if ((!us_reverse(p) && fixed_end[us_index(p)]) || (us_reverse(p) && fixed_start[us_index(p)])) ok = false;
            }
            if (ok) {
                size_t c = len[us_index(grp[0])];               // get_common_end_seq, :298-312
                bool dup = false; uint32_t min_len = 0xFFFFFFFFu;
                for (uint32_t a = 0; a < gn; ++a) {
                    const uint32_t la = len[us_index(grp[a])];
                    if (la < c) c = la;
                    size_t m = 0; while (m < c && from_end(grp[a], m) == from_end(grp[0], m)) ++m;
                    c = m;
                    if (la < min_len) min_len = la;
                    for (uint32_t b = 0; b < a; ++b) if (us_index(grp[a]) == us_index(grp[b])) dup = true;
                }
                // avoid_zero_len_unitigs (:141-158) and avoid_start_of_path (:161-181): both trim from the start
                if (c > 0) c = std::min<size_t>(c, (min_len - 1) / (dup ? 2 : 1));
                if (c > 0) c = min_fpos[idx] == 0 ? 0 : std::min<size_t>(c, min_fpos[idx] - 1);
                if (c > 0) {
                    common.resize(c);
                    for (size_t i = 0; i < c; ++i) common[c - 1 - i] = from_end(grp[0], i);
                    for (uint32_t a = 0; a < gn; ++a) {
                        const uint32_t s = us_index(grp[a]);
                        if (!us_reverse(grp[a])) { min_rpos[s] += (uint32_t)c; len[s] -= (uint32_t)c; room_after[s] += (uint32_t)c; }                       // remove_seq_from_end, unitig.rs:225-232
                        else { min_fpos[s] += (uint32_t)c; len[s] -= (uint32_t)c; seq_off[s] += c; room_before[s] += (uint32_t)c; }                      // remove_seq_from_start, unitig.rs:216-223
                    }
                    grow_front(idx, (uint32_t)c);                                                                                                          // add_seq_to_start, unitig.rs:234-240
                    seq_off[idx] -= c; room_before[idx] -= (uint32_t)c; len[idx] += (uint32_t)c; min_fpos[idx] -= (uint32_t)c;
                    memcpy(arena.data() + seq_off[idx], common.data(), c);
                    total_shifted += c;
                }
            }
        }

        // ---- get_exclusive_outputs (:258-280) + shift_sequence_2 (:119-138) ----
        {
            const UStrand* grp = next_begin(self_fwd); const uint32_t gn = next_size(self_fwd);
            bool ok = gn >= 2 && !fixed_end[idx];
            for (uint32_t a = 0; ok && a < gn; ++a) {
                const UStrand q = grp[a];
                if (!(prev_size(q) == 1 && prev_begin(q)[0] == self_fwd) || us_index(q) == idx) ok = false;
                else if ((!us_reverse(q) && fixed_start[us_index(q)]) || (us_reverse(q) && fixed_end[us_index(q)])) ok = false;
            }
            if (ok) {
                size_t c = len[us_index(grp[0])];               // get_common_start_seq, :283-295
                bool dup = false; uint32_t min_len = 0xFFFFFFFFu;
                for (uint32_t a = 0; a < gn; ++a) {
                    const uint32_t la = len[us_index(grp[a])];
                    if (la < c) c = la;
                    size_t m = 0; while (m < c && from_start(grp[a], m) == from_start(grp[0], m)) ++m;
                    c = m;
                    if (la < min_len) min_len = la;
                    for (uint32_t b = 0; b < a; ++b) if (us_index(grp[a]) == us_index(grp[b])) dup = true;
                }
                if (c > 0) c = std::min<size_t>(c, (min_len - 1) / (dup ? 2 : 1));
                if (c > 0) c = min_rpos[idx] == 0 ? 0 : std::min<size_t>(c, min_rpos[idx] - 1);
                if (c > 0) {
                    common.resize(c);
                    for (size_t i = 0; i < c; ++i) common[i] = from_start(grp[0], i);
                    for (uint32_t a = 0; a < gn; ++a) {
                        const uint32_t s = us_index(grp[a]);
                        if (!us_reverse(grp[a])) { min_fpos[s] += (uint32_t)c; len[s] -= (uint32_t)c; seq_off[s] += c; room_before[s] += (uint32_t)c; }
                        else { min_rpos[s] += (uint32_t)c; len[s] -= (uint32_t)c; room_after[s] += (uint32_t)c; }
                    }
                    grow_back(idx, (uint32_t)c);                                                                                                           // add_seq_to_end, unitig.rs:242-248
                    memcpy(arena.data() + seq_off[idx] + len[idx], common.data(), c);
                    room_after[idx] -= (uint32_t)c; len[idx] += (uint32_t)c; min_rpos[idx] -= (uint32_t)c;
                    total_shifted += c;
                }
            }
        }
    }
    prof.expand += now_ms() - t0; prof.passes += 1;
    return total_shifted;
}

void HostGraph::simplify_structure() {   // graph_simplification.rs:26-40
    while (expand_repeats() > 0) {}
    renumber();
}

// ------------------------------------------------------------------------------------------------
// unitig_graph.rs:317-360 save_gfa
// ------------------------------------------------------------------------------------------------
namespace {
inline char* put_uint(char* p, uint64_t v) { auto r = std::to_chars(p, p + 24, v); return r.ptr; }
inline char* put_str(char* p, const char* s, size_t n) { memcpy(p, s, n); return p + n; }
}  // namespace

void HostGraph::gfa_text(const std::vector<HostSeq>& seqs, std::string& out) const {
    size_t est = 64 + total_length() + (size_t)U * 40 + next.size() * 40 + path.size() * 12;
    for (auto& s : seqs) est += 96 + s.filename.size() + s.contig_header.size();
    out.resize(est);
    char* const base = &out[0];
    char* p = base;
    p = put_str(p, "H\tVN:Z:1.0\tKM:i:", 16); p = put_uint(p, k); *p++ = '\n';
    for (uint32_t n = 0; n < U; ++n) {   // unitig.rs:167-171; depth is integral here so {:.2} renders as N.00
        const uint32_t idx = order[n];
        *p++ = 'S'; *p++ = '\t'; p = put_uint(p, n + 1); *p++ = '\t';
        p = put_str(p, seq_ptr(idx), len[idx]);
        p = put_str(p, "\tDP:f:", 6); p = put_uint(p, depth[idx]); p = put_str(p, ".00\n", 4);
    }
    for (uint32_t n = 0; n < U; ++n) {   // get_links_for_gfa, :333-350: forward_next then reverse_next of each unitig
        const uint32_t idx = order[n];
        for (uint32_t rev = 0; rev < 2; ++rev) {
            const UStrand from = us_make(idx, rev != 0);
            for (uint32_t x = next_off[from]; x < next_off[from + 1]; ++x) {
                const UStrand to = next[x];
                *p++ = 'L'; *p++ = '\t'; p = put_uint(p, n + 1); *p++ = '\t'; *p++ = rev ? '-' : '+'; *p++ = '\t';
                p = put_uint(p, number[us_index(to)]); *p++ = '\t'; *p++ = us_reverse(to) ? '-' : '+'; p = put_str(p, "\t0M\n", 4);
            }
        }
    }
    for (size_t i = 0; i < seqs.size(); ++i) {   // get_gfa_path_line, :352-360
        const HostSeq& s = seqs[i];
        *p++ = 'P'; *p++ = '\t'; p = put_uint(p, s.id); *p++ = '\t';
        for (uint64_t x = path_off[i]; x < path_off[i + 1]; ++x) {
            if (x != path_off[i]) *p++ = ',';
            p = put_uint(p, number[us_index(path[x])]); *p++ = us_reverse(path[x]) ? '-' : '+';
        }
        p = put_str(p, "\t*\tLN:i:", 8); p = put_uint(p, s.length);
        p = put_str(p, "\tFN:Z:", 6); p = put_str(p, s.filename.data(), s.filename.size());
        p = put_str(p, "\tHD:Z:", 6); p = put_str(p, s.contig_header.data(), s.contig_header.size()); *p++ = '\n';
    }
    if ((size_t)(p - base) > est) throw std::runtime_error("GFA size estimate too small");
    out.resize((size_t)(p - base));
}
