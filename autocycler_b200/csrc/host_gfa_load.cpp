// UnitigGraph::from_gfa_lines (unitig_graph.rs:55-174) into the host graph: lets the library pick up an
// `input_assemblies.gfa` written earlier (by this library or by the reference) for what follows compress — sequence
// reconstruction (decompress.rs:83-105), merge_linear_paths, simplify_structure, renumbering, saving again.
//
// Depths are kept as f64 and the segment colour tags as UnitigType (unitig.rs:72-86), so graphs that trim / resolve have
// re-weighted or coloured load and save back byte for byte.  Lines are split the way BufRead::lines does it (misc.rs:51-61):
// at '\n', with one trailing '\r' dropped.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

#include "host_graph.h"

namespace {
struct Span { const char* p; size_t n; bool starts(const char* s) const { const size_t m = strlen(s); return n >= m && memcmp(p, s, m) == 0; }
              bool is(const char* s) const { return n == strlen(s) && memcmp(p, s, n) == 0; } std::string str() const { return std::string(p, n); } };

std::vector<Span> split(Span line, char sep) {
    std::vector<Span> parts;
    size_t a = 0;
    for (size_t i = 0; i <= line.n; ++i)
        if (i == line.n || line.p[i] == sep) { parts.push_back({line.p + a, i - a}); a = i + 1; }
    return parts;
}
bool parse_u32(Span s, uint32_t& out) {   // str::parse::<u32>: optional '+', digits only, no overflow
    size_t i = 0;
    if (s.n > 0 && s.p[0] == '+') i = 1;
    if (i == s.n) return false;
    uint64_t v = 0;
    for (; i < s.n; ++i) { if (s.p[i] < '0' || s.p[i] > '9') return false; v = v * 10 + (uint64_t)(s.p[i] - '0'); if (v > 0xFFFFFFFFull) return false; }
    out = (uint32_t)v;
    return true;
}
[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }
}  // namespace

void HostGraph::load_gfa(const char* text, size_t len, std::vector<HostSeq>& seqs) {
    std::vector<Span> seg_lines, link_lines, path_lines;
    k = 0;
    for (size_t a = 0; a < len;) {
        const char* nl = (const char*)memchr(text + a, '\n', len - a);
        const size_t b = nl ? (size_t)(nl - text) : len;
        Span line{text + a, b - a};
        if (line.n > 0 && nl && line.p[line.n - 1] == '\r') line.n -= 1;         // "\r\n" ends a line too (BufRead::lines)
        if (line.n > 0) {
            const size_t tab = std::min<size_t>(line.n, (size_t)((const char*)memchr(line.p, '\t', line.n) ? (const char*)memchr(line.p, '\t', line.n) - line.p : line.n));
            const Span tag{line.p, tab};
            if (tag.is("H")) { for (const Span& part : split(line, '\t')) if (part.starts("KM:i:")) { uint32_t kk; if (parse_u32({part.p + 5, part.n - 5}, kk)) { k = kk; break; } } }   // :80-89
            else if (tag.is("S")) seg_lines.push_back(line);
            else if (tag.is("L")) link_lines.push_back(line);
            else if (tag.is("P")) path_lines.push_back(line);
        }
        a = b + 1;
    }

    // ---- segments (unitig.rs:62-91), in file order ----
    const uint32_t n_seg = (uint32_t)seg_lines.size();
    own_rec.assign(n_seg, UnitigRec{}); own_depth.assign(n_seg, 0); own_depth_f.assign(n_seg, 0.0); own_type.assign(n_seg, 0); number.assign(n_seg, 0); order.resize(n_seg);
    std::unordered_map<uint32_t, uint32_t> index;                       // build_unitig_index (:76-78): the last segment with a number wins
    uint64_t bytes = 0;
    std::vector<Span> seg_seq(n_seg);
    for (uint32_t i = 0; i < n_seg; ++i) {
        const std::vector<Span> parts = split(seg_lines[i], '\t');
        if (parts.size() < 3) fail("Segment line does not have enough parts.");
        if (!parse_u32(parts[1], number[i])) fail("Unable to parse unitig number.");
        seg_seq[i] = parts[2];
        bool found = false;
        for (const Span& part : parts)
            if (part.starts("DP:f:")) {
                const std::string v(part.p + 5, part.n - 5);              // str::parse::<f64>: decimal or exponent form, inf / infinity / nan; no blanks, no hex
                bool plain = !v.empty();
                for (char c : v) if (!(isalnum((unsigned char)c) || c == '.' || c == '+' || c == '-') || c == 'x' || c == 'X' || c == 'p' || c == 'P' || c == '(') plain = false;
                char* end = nullptr; const double d = plain ? strtod(v.c_str(), &end) : 0.0;
                if (plain && end && *end == 0) {
                    own_depth_f[i] = d; found = true;
                    own_depth[i] = (d >= 0 && d <= 4294967295.0) ? (uint32_t)d : 0;     // whole-number view for callers that want one; depth_f is what is saved
                }
                break;                                                   // Iterator::find stops at the first DP:f: part
            }
        if (!found) fail("Could not find a depth tag (e.g. DP:f:10.00) in the GFA segment line.\nAre you sure this is an Autocycler-generated GFA file?");
        {   // unitig.rs:78-86: consentig, else anchor, else bridge, else other — whichever tags the line carries
            bool consentig = false, anchor = false, bridge = false;
            for (const Span& part : parts) { consentig |= part.is("CL:Z:steelblue"); anchor |= part.is("CL:Z:forestgreen"); bridge |= part.is("CL:Z:pink"); }
            own_type[i] = consentig ? 3 : anchor ? 1 : bridge ? 2 : 0;
        }
        index[number[i]] = i;
        order[i] = i;
        bytes += parts[2].n + 2 * AC_SEQ_SLACK;
    }
    U = n_seg;
    arena_overflow.assign(bytes + bytes / 4 + (1u << 16), 0);
    arena = arena_overflow.data(); arena_cap = arena_overflow.size(); arena_used = 0;
    for (uint32_t i = 0; i < n_seg; ++i) {
        UnitigRec& r = own_rec[i];
        r.seq_off = arena_used + AC_SEQ_SLACK; r.len = (uint32_t)seg_seq[i].n; r.room_before = r.room_after = AC_SEQ_SLACK; r.flags = 0;
        r.min_fpos = r.min_rpos = 0xFFFFFFFFu;
        memcpy(arena + r.seq_off, seg_seq[i].p, seg_seq[i].n);
        arena_used = r.seq_off + r.len + AC_SEQ_SLACK;
    }

    // ---- links (:91-115): next lists in file order, prev lists as the loader fills them ----
    std::vector<std::vector<UStrand>> nx(2 * (size_t)U), pv(2 * (size_t)U);
    for (const Span& line : link_lines) {
        const std::vector<Span> parts = split(line, '\t');
        if (parts.size() < 6 || !parts[5].is("0M")) fail("non-zero overlap found on the GFA link line.\nAre you sure this is an Autocycler-generated GFA file?");
        uint32_t a = 0, b = 0;
        if (!parse_u32(parts[1], a)) fail("Error parsing segment 1 as integer");
        if (!parse_u32(parts[3], b)) fail("Error parsing segment 2 as integer");
        const auto ia = index.find(a); if (ia == index.end()) fail("link refers to nonexistent unitig: " + std::to_string(a));
        const auto ib = index.find(b); if (ib == index.end()) fail("link refers to nonexistent unitig: " + std::to_string(b));
        const UStrand from = us_make(ia->second, !parts[2].is("+")), to = us_make(ib->second, !parts[4].is("+"));
        nx[from].push_back(to); pv[to].push_back(from);
    }
    own_next_off.assign(2 * (size_t)U + 1, 0); own_prev_off.assign(2 * (size_t)U + 1, 0); own_next.clear(); own_prev.clear();
    for (size_t s = 0; s < 2 * (size_t)U; ++s) {
        own_next.insert(own_next.end(), nx[s].begin(), nx[s].end()); own_next_off[s + 1] = (uint32_t)own_next.size();
        own_prev.insert(own_prev.end(), pv[s].begin(), pv[s].end()); own_prev_off[s + 1] = (uint32_t)own_prev.size();
    }

    // ---- paths (:117-174): sequences without bytes, and the positions they imply on both strands ----
    seqs.clear();
    own_path.clear(); own_path_off.assign(1, 0);
    for (const Span& line : path_lines) {
        const std::vector<Span> parts = split(line, '\t');
        uint32_t id = 0;
        if (parts.size() < 3 || !parse_u32(parts[1], id) || id > 0xFFFF) fail("Error parsing sequence ID as integer");
        bool has_len = false, has_fn = false, has_hd = false; uint32_t length = 0, cluster = 0;
        HostSeq s; s.id = (uint16_t)id; s.start = 0; s.length = 0; s.cluster = 0;
        for (size_t x = 2; x < parts.size(); ++x) {
            const Span& p = parts[x];
            if (p.starts("LN:i:")) { if (!parse_u32({p.p + 5, p.n - 5}, length)) fail("Error parsing length"); has_len = true; }
            else if (p.starts("FN:Z:")) { s.filename.assign(p.p + 5, p.n - 5); has_fn = true; }
            else if (p.starts("HD:Z:")) { s.contig_header.assign(p.p + 5, p.n - 5); has_hd = true; }
            else if (p.starts("CL:i:")) { if (!parse_u32({p.p + 5, p.n - 5}, cluster) || cluster > 0xFFFF) fail("Error parsing cluster"); }
        }
        if (!has_len || !has_fn || !has_hd) fail("missing required tag in GFA path line.");
        s.length = length; s.cluster = (uint16_t)cluster;
        uint64_t pos = 0;
        if (parts[2].n > 0)
            for (const Span& step : split(parts[2], ',')) {              // parse_unitig_path (:971-984)
                if (step.n == 0 || (step.p[step.n - 1] != '+' && step.p[step.n - 1] != '-')) fail("Invalid path strand");
                uint32_t num = 0;
                if (!parse_u32({step.p, step.n - 1}, num)) fail("Error parsing unitig number in path");
                const auto it = index.find(num);
                if (it == index.end()) fail("unitig " + std::to_string(num) + " not found in unitig index");
                own_path.push_back(us_make(it->second, step.p[step.n - 1] == '-'));
                pos += own_rec[it->second].len;
            }
        if (pos != length) fail("Position calculation mismatch");
        // add_positions_from_path (:160-174) for the path and its reverse: a forward step at offset p holds (seq, +, p) in its
        // forward positions and (seq, -, L - p - len) in its reverse positions; a reverse step the other way round
        uint64_t at = 0;
        for (uint64_t x = own_path_off.back(); x < own_path.size(); ++x) {
            UnitigRec& r = own_rec[us_index(own_path[x])];
            const uint32_t here = (uint32_t)at, mirrored = (uint32_t)(length - at - r.len);
            if (!us_reverse(own_path[x])) { r.min_fpos = std::min(r.min_fpos, here); r.min_rpos = std::min(r.min_rpos, mirrored); }
            else { r.min_rpos = std::min(r.min_rpos, here); r.min_fpos = std::min(r.min_fpos, mirrored); }
            at += r.len;
        }
        own_path_off.push_back(own_path.size());
        seqs.push_back(std::move(s));
    }

    rec = own_rec.data(); depth = own_depth.data(); depth_f = own_depth_f.data(); utype = own_type.data();
    next_off = own_next_off.data(); prev_off = own_prev_off.data(); next = own_next.data(); prev = own_prev.data(); n_links = own_next.size();
    path_off = own_path_off.data(); path = own_path.data(); n_path = own_path.size(); n_seqs = (uint32_t)seqs.size();
    fpos_off.clear(); rpos_off.clear(); fpos.clear(); rpos.clear();
    fixed_ready = false; cands_ready = false; first_pass = true; spec_from_device = false;
    prof = HostProfile();
    check_links();
}
