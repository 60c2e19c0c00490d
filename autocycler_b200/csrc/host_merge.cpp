// merge_linear_paths on the host graph (graph_simplification.rs:315-526): the first thing every command downstream of
// compress does with input_assemblies.gfa (cluster.rs:804, trim.rs:266, resolve.rs:255, clean.rs:114).
//
// The reference edits Rc-linked unitigs in place; here the decisions are taken on the CSR graph that compress left
// behind, the link edits are replayed on the few lists they touch, and one compaction pass writes the new graph
// (surviving unitigs in their current order, then the merged ones in creation order, like `graph.unitigs`).
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "host_graph.h"

namespace {
inline char complement(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
const uint32_t GONE = 0xFFFFFFFFu;
}

void HostGraph::merge_linear_paths(bool use_paths) {
    // ---- fixed starts / ends (:330, :190-230), plus one start per simple circular component (:374-384) ----
    if (use_paths) compute_fixed(); else { fixed_start.assign(U, 0); fixed_end.assign(U, 0); }
    std::vector<uint8_t> fs = fixed_start, fe = fixed_end;
    {
        std::vector<uint32_t> comp(U, GONE), stack, stamp(U, 0);
        uint32_t round = 0;
        for (uint32_t n = 0; n < U; ++n) {
            const uint32_t seed = order[n];
            if (comp[seed] != GONE) continue;
            stack.assign(1, seed); comp[seed] = seed;
            bool all_single = true; uint32_t lowest = seed; size_t members = 0;
            while (!stack.empty()) {   // unitig_graph.rs:905-947 connected_components; component[0] is its smallest number
                const uint32_t u = stack.back(); stack.pop_back();
                ++members;
                if (number[u] < number[lowest]) lowest = u;
                for (uint32_t r = 0; r < 2; ++r) {
                    const UStrand s = us_make(u, r);
                    if (next_size(s) != 1 || prev_size(s) != 1) all_single = false;
                    for (uint32_t x = 0; x < next_size(s); ++x) { const uint32_t v = us_index(next_begin(s)[x]); if (comp[v] == GONE) { comp[v] = seed; stack.push_back(v); } }
                    for (uint32_t x = 0; x < prev_size(s); ++x) { const uint32_t v = us_index(prev_begin(s)[x]); if (comp[v] == GONE) { comp[v] = seed; stack.push_back(v); } }
                }
            }
            if (!all_single) continue;   // component_is_circular_loop (unitig_graph.rs:949-967) needs four single links on every member
            ++round;
            uint32_t num = lowest; bool rev = false; size_t visited = 0; bool loop = true;
            while (num != lowest || visited == 0) {
                if (stamp[num] == round) { loop = false; break; }
                stamp[num] = round; ++visited;
                const UStrand nx = next_begin(us_make(num, rev))[0];
                num = us_index(nx); rev = us_reverse(nx);
            }
            if (loop && visited == members) fs[lowest] = 1;
        }
    }

    // ---- the mergeable paths (:332-362), found in graph order exactly as the reference's loop finds them ----
    auto exclusive_inputs = [&](uint32_t u, UStrand& only) -> uint32_t {    // :233-255; 0 unless every input leads only here
        const UStrand me = us_make(u, false);
        const uint32_t n = prev_size(me); const UStrand* p = prev_begin(me);
        for (uint32_t i = 0; i < n; ++i) if (next_size(p[i]) != 1 || next_begin(p[i])[0] != me) return 0;
        for (uint32_t i = 0; i < n; ++i) if (us_index(p[i]) == u) return 0;
        if (n) only = p[0];
        return n;
    };
    auto exclusive_outputs = [&](uint32_t u, UStrand& only) -> uint32_t {   // :258-280
        const UStrand me = us_make(u, false);
        const uint32_t n = next_size(me); const UStrand* p = next_begin(me);
        for (uint32_t i = 0; i < n; ++i) if (prev_size(p[i]) != 1 || prev_begin(p[i])[0] != me) return 0;
        for (uint32_t i = 0; i < n; ++i) if (us_index(p[i]) == u) return 0;
        if (n) only = p[0];
        return n;
    };
    auto cannot_merge_start = [&](UStrand s) { return us_reverse(s) ? fe[us_index(s)] : fs[us_index(s)]; };   // :387-390
    auto cannot_merge_end = [&](UStrand s) { return us_reverse(s) ? fs[us_index(s)] : fe[us_index(s)]; };     // :398-401
    std::vector<uint8_t> used(U, 0);
    std::vector<UStrand> flat;                 // the merge paths, one after another
    std::vector<size_t> path_begin;
    for (uint32_t n = 0; n < U; ++n) {
        const uint32_t u = order[n];
        for (uint32_t r = 0; r < 2; ++r) {
            if (used[u]) continue;
            const UStrand start = us_make(u, r);
            UStrand only = 0;
            const uint32_t inputs = r ? exclusive_outputs(u, only) : exclusive_inputs(u, only);
            if (inputs == 1 && !cannot_merge_start(start)) continue;      // something upstream will pick this unitig up
            const size_t begin = flat.size();
            flat.push_back(start); used[u] = 1;
            for (;;) {
                const UStrand cur = flat.back();
                if (cannot_merge_end(cur)) break;
                UStrand out = 0;
                const uint32_t outputs = us_reverse(cur) ? exclusive_inputs(us_index(cur), out) : exclusive_outputs(us_index(cur), out);
                if (outputs != 1) break;
                if (us_reverse(cur)) out = us_flip(out);
                if (used[us_index(out)]) break;
                if (cannot_merge_start(out)) break;
                flat.push_back(out); used[us_index(out)] = 1;
            }
            if (flat.size() - begin > 1) path_begin.push_back(begin); else flat.resize(begin);
        }
    }
    const uint32_t M = (uint32_t)path_begin.size();
    fixed_ready = false; cands_ready = false; first_pass = true;
    if (M == 0) return;
    path_begin.push_back(flat.size());

    // ---- merge_path (:410-487) replayed on link lists: the old lists only ever get entries appended (pruning by number
    // is delete_dangling_links' job at the end, unitig_graph.rs:547-564), so CSR + a sparse "appended" map is enough ----
    struct NewUnitig { std::vector<UStrand> next[2], prev[2]; };          // [0] forward_*, [1] reverse_*
    std::vector<NewUnitig> made(M);
    std::unordered_map<UStrand, std::vector<UStrand>> more_next, more_prev;
    auto current = [&](bool want_next, UStrand s, std::vector<UStrand>& out) {
        const uint32_t idx = us_index(s);
        if (idx >= U) { out = want_next ? made[idx - U].next[us_reverse(s)] : made[idx - U].prev[us_reverse(s)]; return; }
        if (want_next) out.assign(next_begin(s), next_begin(s) + next_size(s)); else out.assign(prev_begin(s), prev_begin(s) + prev_size(s));
        auto& extra = want_next ? more_next : more_prev;
        auto it = extra.find(s);
        if (it != extra.end()) out.insert(out.end(), it->second.begin(), it->second.end());
    };
    auto append = [&](bool to_next, UStrand s, UStrand value) {
        const uint32_t idx = us_index(s);
        if (idx >= U) (to_next ? made[idx - U].next[us_reverse(s)] : made[idx - U].prev[us_reverse(s)]).push_back(value);
        else (to_next ? more_next : more_prev)[s].push_back(value);
    };
    std::vector<UStrand> scratch;
    auto link_exists = [&](UStrand a, UStrand b) {                        // unitig_graph.rs:723-735
        current(true, a, scratch);
        return std::find(scratch.begin(), scratch.end(), b) != scratch.end();
    };
    for (uint32_t i = 0; i < M; ++i) {
        const UStrand first = flat[path_begin[i]], last = flat[path_begin[i + 1] - 1];
        const UStrand fwd = us_make(U + i, false), rev = us_make(U + i, true);
        const bool end_to_start = link_exists(last, first), start_flip = link_exists(us_flip(first), first), end_flip = link_exists(last, us_flip(last));
        NewUnitig& nu = made[i];
        current(false, first, nu.prev[0]);               // forward_prev
        current(true, us_flip(first), nu.next[1]);       // reverse_next
        current(true, last, nu.next[0]);                 // forward_next
        current(false, us_flip(last), nu.prev[1]);       // reverse_prev
        for (UStrand u : nu.next[0]) append(false, u, fwd);
        for (UStrand u : nu.prev[0]) append(true, u, fwd);
        for (UStrand u : nu.next[1]) append(false, u, rev);
        for (UStrand u : nu.prev[1]) append(true, u, rev);
        if (end_to_start) { nu.next[0].push_back(fwd); nu.prev[0].push_back(fwd); nu.next[1].push_back(rev); nu.prev[1].push_back(rev); }
        if (start_flip) { nu.next[1].push_back(fwd); nu.prev[0].push_back(rev); }
        if (end_flip) { nu.next[0].push_back(rev); nu.prev[1].push_back(fwd); }
    }

    // ---- compaction: new index = position in the new `unitigs` order ----
    std::vector<uint32_t> member_of(U, GONE);             // merge path of an old unitig
    for (uint32_t i = 0; i < M; ++i) for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) member_of[us_index(flat[x])] = i;
    std::vector<uint32_t> new_index((size_t)U + M, GONE);
    uint32_t kept = 0, max_number = 0;
    for (uint32_t n = 0; n < U; ++n) { const uint32_t u = order[n]; max_number = std::max(max_number, number[u]); if (member_of[u] == GONE) new_index[u] = kept++; }
    for (uint32_t i = 0; i < M; ++i) new_index[U + i] = kept + i;
    const uint32_t U2 = kept + M;

    uint64_t merged_bytes = 0;
    for (uint32_t i = 0; i < M; ++i) { merged_bytes += 2 * AC_SEQ_SLACK; for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) merged_bytes += rec[us_index(flat[x])].len; }
    reserve_arena(merged_bytes);

    std::vector<UnitigRec> rec2(U2);
    std::vector<uint32_t> depth2(U2), number2(U2), order2(U2);
    // Loaded graphs: f64 depths and unitig types travel along; the merged unitig's depth is get_merge_path_depth (:503-526).
    const bool general = depth_f != nullptr;
    std::vector<double> depth_f2(general ? U2 : 0); std::vector<uint8_t> type2(general ? U2 : 0);
    std::vector<uint32_t> visits;              // forward_positions.len() of every unitig = the path steps that pass through it (unitig_graph.rs:160-174)
    if (general) { visits.assign(U, 0); for (uint64_t x = 0; x < n_path; ++x) visits[us_index(path[x])] += 1; }
    for (uint32_t n = 0; n < U; ++n) {
        const uint32_t u = order[n], v = new_index[u];
        if (v == GONE) continue;
        rec2[v] = rec[u]; depth2[v] = depth[u]; number2[v] = number[u];
        if (general) { depth_f2[v] = depth_f[u]; type2[v] = type_of(u); }
    }
    for (uint32_t i = 0; i < M; ++i) {
        const uint32_t v = kept + i;
        const UStrand first = flat[path_begin[i]], last = flat[path_begin[i + 1] - 1];
        UnitigRec& r = rec2[v];
        r.seq_off = arena_used + AC_SEQ_SLACK; r.room_before = r.room_after = AC_SEQ_SLACK; r.flags = 0;
        char* out = arena + r.seq_off;
        for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) {       // merge_unitig_seqs, :490-500
            const UStrand m = flat[x]; const UnitigRec& mr = rec[us_index(m)];
            const char* src = arena + mr.seq_off;
            if (!us_reverse(m)) memcpy(out, src, mr.len);
            else for (uint32_t j = 0; j < mr.len; ++j) out[j] = complement(src[mr.len - 1 - j]);
            out += mr.len;
        }
        r.len = (uint32_t)(out - (arena + r.seq_off));
        arena_used = r.seq_off + r.len + AC_SEQ_SLACK;
        // forward_positions come from the first unitig, reverse_positions from the last (:413-414); depth = their count (:503-507)
        r.min_fpos = us_reverse(first) ? rec[us_index(first)].min_rpos : rec[us_index(first)].min_fpos;
        r.min_rpos = us_reverse(last) ? rec[us_index(last)].min_fpos : rec[us_index(last)].min_rpos;
        depth2[v] = depth[us_index(first)];
        if (general) {
            double d; bool has_anchor = false, consentig = false; double anchor_depth = 0;
            for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) {
                const uint32_t m = us_index(flat[x]);
                if (type_of(m) == 1 && !has_anchor) { has_anchor = true; anchor_depth = depth_f[m]; }
                if (type_of(m) == 1 || type_of(m) == 3) consentig = true;
            }
            if (visits[us_index(first)] > 0) d = (double)visits[us_index(first)];          // positions exist: their count
            else if (has_anchor) d = anchor_depth;                                         // the first anchor's depth
            else {                                                                         // weighted_mean_depth (:517-525)
                uint32_t total = 0; double sum = 0.0;
                for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) total += rec[us_index(flat[x])].len;
                for (size_t x = path_begin[i]; x < path_begin[i + 1]; ++x) { const uint32_t m = us_index(flat[x]); sum += depth_f[m] * (double)rec[m].len; }
                d = sum / (double)total;
            }
            depth_f2[v] = d; type2[v] = consentig ? 3 : 0;                                 // :439-441
            depth2[v] = (d >= 0 && d <= 4294967295.0) ? (uint32_t)d : 0;
        }
        number2[v] = max_number + 1 + i;
    }
    for (uint32_t v = 0; v < U2; ++v) order2[v] = v;

    // links: survivors keep their entries in order, minus the merged-away unitigs, plus what the merges appended
    std::vector<uint32_t> next_off2(2 * (size_t)U2 + 1, 0), prev_off2(2 * (size_t)U2 + 1, 0);
    std::vector<UStrand> next2, prev2, list;
    next2.reserve(n_links); prev2.reserve(n_links);
    auto old_strand = [&](uint32_t v, uint32_t r) -> UStrand { return us_make(v, r); };
    std::vector<uint32_t> old_of(U2);
    for (uint32_t u = 0; u < U; ++u) if (new_index[u] != GONE) old_of[new_index[u]] = u;
    for (uint32_t i = 0; i < M; ++i) old_of[kept + i] = U + i;
    for (uint32_t v = 0; v < U2; ++v)
        for (uint32_t r = 0; r < 2; ++r) {
            const UStrand s = old_strand(old_of[v], r);
            current(true, s, list);
            for (UStrand t : list) { const uint32_t w = new_index[us_index(t)]; if (w != GONE) next2.push_back(us_make(w, us_reverse(t))); }
            next_off2[2 * (size_t)v + r + 1] = (uint32_t)next2.size();
            current(false, s, list);
            for (UStrand t : list) { const uint32_t w = new_index[us_index(t)]; if (w != GONE) prev2.push_back(us_make(w, us_reverse(t))); }
            prev_off2[2 * (size_t)v + r + 1] = (uint32_t)prev2.size();
        }

    // paths: a sequence enters a merged run at one of its ends (anything else was a fixed start or end), so the run
    // collapses to the new unitig at its entry element and the other members drop out
    std::vector<uint64_t> path_off2((size_t)n_seqs + 1, 0);
    std::vector<UStrand> path2; path2.reserve(n_path);
    for (uint32_t q = 0; use_paths && q < n_seqs; ++q) {
        for (uint64_t x = path_off[q]; x < path_off[q + 1]; ++x) {
            const UStrand p = path[x]; const uint32_t u = us_index(p), i = member_of[u];
            if (i == GONE) { path2.push_back(us_make(new_index[u], us_reverse(p))); continue; }
            const UStrand first = flat[path_begin[i]], last = flat[path_begin[i + 1] - 1];
            if (p == first) path2.push_back(us_make(kept + i, false));
            else if (p == us_flip(last)) path2.push_back(us_make(kept + i, true));
        }
        path_off2[q + 1] = path2.size();
    }

    // full position lists, if they were asked for: the merged unitig inherits the first unitig's forward positions and
    // the last one's reverse positions
    if (!fpos_off.empty()) {
        std::vector<uint64_t> fo((size_t)U2 + 1, 0), ro((size_t)U2 + 1, 0), f, rr;
        auto take = [&](const std::vector<uint64_t>& off, const std::vector<uint64_t>& val, uint32_t u, std::vector<uint64_t>& out) { out.insert(out.end(), val.begin() + off[u], val.begin() + off[u + 1]); };
        for (uint32_t v = 0; v < U2; ++v) {
            if (v < kept) { take(fpos_off, fpos, old_of[v], f); take(rpos_off, rpos, old_of[v], rr); }
            else {
                const uint32_t i = v - kept; const UStrand first = flat[path_begin[i]], last = flat[path_begin[i + 1] - 1];
                if (us_reverse(first)) take(rpos_off, rpos, us_index(first), f); else take(fpos_off, fpos, us_index(first), f);
                if (us_reverse(last)) take(fpos_off, fpos, us_index(last), rr); else take(rpos_off, rpos, us_index(last), rr);
            }
            fo[v + 1] = f.size(); ro[v + 1] = rr.size();
        }
        fpos_off.swap(fo); rpos_off.swap(ro); fpos.swap(f); rpos.swap(rr);
    }

    // ---- adopt the new graph (graph.build_unitig_index / check_links, :368-370) ----
    own_rec.swap(rec2); own_depth.swap(depth2); number.swap(number2); order.swap(order2);
    if (general) { own_depth_f.swap(depth_f2); own_type.swap(type2); depth_f = own_depth_f.data(); utype = own_type.data(); }
    own_next_off.swap(next_off2); own_prev_off.swap(prev_off2); own_next.swap(next2); own_prev.swap(prev2);
    own_path_off.swap(path_off2); own_path.swap(path2);
    U = U2; rec = own_rec.data(); depth = own_depth.data();
    next_off = own_next_off.data(); prev_off = own_prev_off.data(); next = own_next.data(); prev = own_prev.data(); n_links = own_next.size();
    path_off = own_path_off.data(); path = own_path.data(); n_path = own_path.size();
    check_links();
}
