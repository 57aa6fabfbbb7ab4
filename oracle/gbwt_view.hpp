// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under vg_b200/ may
// include, link or call this.  CPU restatement of the GBWT / GBWTGraph primitives the
// Giraffe hot path uses, over the flat index format of include/giraffe_b200.h.
//
// The arithmetic lives in jltsiren/gbwt @ c2e0199 and jltsiren/gbwtgraph @ e27bc43, both
// ABSENT from /root/reference (deps/ is empty).  Restated from the published algorithm
// (Siren et al., "Haplotype-aware graph indexes", Bioinformatics 2020) and anchored on
// vg's call sites:
//   get_bd_state       gbwt_extender.cpp:577
//   follow_paths       gbwt_extender.cpp:608, :652 ; minimizer_mapper.cpp:5983
//   bd_find            gbwt_extender.cpp:514
//   get_sequence_view  gbwt_extender.cpp:582, :616, :661
#pragma once
#include "../include/giraffe_b200.h"
#include <cstdint>
#include <functional>
#include <string_view>
#include <vector>

namespace oracle {

struct SearchState {
    uint32_t node = 0;
    int64_t lo = 0, hi = -1;           // closed range; empty when lo > hi
    bool empty() const { return lo > hi; }
    uint64_t size() const { return empty() ? 0 : (uint64_t)(hi - lo + 1); }
    bool operator==(const SearchState& o) const { return node == o.node && lo == o.lo && hi == o.hi; }
};

struct BidirectionalState {
    SearchState forward, backward;
    bool empty() const { return forward.empty(); }
    uint64_t size() const { return forward.size(); }
    void flip() { std::swap(forward, backward); }
    bool operator==(const BidirectionalState& o) const { return forward == o.forward && backward == o.backward; }
};

struct Graph {
    const gb_flat_index* ix;
    explicit Graph(const gb_flat_index* i) : ix(i) {}

    bool has_node(uint32_t v) const { return v >= 2 && v < ix->n_nodes && ix->nodes[v].len > 0; }
    uint32_t get_length(uint32_t v) const { return ix->nodes[v].len; }
    std::string_view get_sequence_view(uint32_t v) const {
        return std::string_view((const char*)ix->seq + ix->nodes[v].seq_off, ix->nodes[v].len);
    }
    uint32_t record_size(uint32_t v) const { return ix->nodes[v].size; }

    // decoded record
    struct Record {
        uint32_t n_edges = 0, n_runs = 0;
        const uint32_t* edges = nullptr;   // (to, offset) pairs
        const uint32_t* runs = nullptr;
        uint32_t successor(uint32_t r) const { return edges[2 * r]; }
        uint32_t offset(uint32_t r) const { return edges[2 * r + 1]; }
        // occurrences of outrank r in body[0..i)
        uint64_t rank(uint64_t i, uint32_t r) const {
            uint64_t pos = 0, cnt = 0;
            for (uint32_t j = 0; j < n_runs && pos < i; j++) {
                uint32_t len = runs[j] >> 10, rr = runs[j] & 1023u;
                uint64_t take = (pos + len <= i) ? len : (i - pos);
                if (rr == r) cnt += take;
                pos += len;
            }
            return cnt;
        }
    };
    Record record(uint32_t v) const {
        Record rec;
        if (!has_node(v) || ix->nodes[v].size == 0) return rec;
        const uint32_t* p = ix->gbwt + ix->nodes[v].rec_off;
        rec.n_edges = p[0]; rec.n_runs = p[1];
        rec.edges = p + 2; rec.runs = p + 2 + 2 * rec.n_edges;
        return rec;
    }

    BidirectionalState get_bd_state(uint32_t v) const {
        BidirectionalState s;
        int64_t sz = record_size(v);
        s.forward.node = v; s.forward.lo = 0; s.forward.hi = sz - 1;
        s.backward.node = v ^ 1u; s.backward.lo = 0; s.backward.hi = sz - 1;
        return s;
    }

    // gbwt::GBWT::bdExtendForward(state, record, outrank)
    BidirectionalState bd_extend_forward(const BidirectionalState& state, const Record& rec, uint32_t outrank) const {
        BidirectionalState next;
        if (state.empty()) return next;
        uint32_t to = rec.successor(outrank);
        uint64_t c_lo = rec.rank((uint64_t)state.forward.lo, outrank);
        uint64_t c_hi = rec.rank((uint64_t)state.forward.hi + 1, outrank);
        next.forward.node = to;
        next.forward.lo = (int64_t)rec.offset(outrank) + (int64_t)c_lo;
        next.forward.hi = (int64_t)rec.offset(outrank) + (int64_t)c_hi - 1;
        // reverse offset: visits in the range whose successor sorts before `to` on the
        // reverse strand
        uint64_t reverse_offset = 0;
        uint32_t reverse_to = to ^ 1u;
        for (uint32_t r = 0; r < rec.n_edges; r++) {
            if ((rec.successor(r) ^ 1u) < reverse_to) {
                reverse_offset += rec.rank((uint64_t)state.forward.hi + 1, r) - rec.rank((uint64_t)state.forward.lo, r);
            }
        }
        next.backward.node = state.backward.node;
        next.backward.lo = state.backward.lo + (int64_t)reverse_offset;
        next.backward.hi = next.backward.lo + (int64_t)next.forward.size() - 1;
        if (next.forward.empty()) { next.backward.lo = 0; next.backward.hi = -1; }
        return next;
    }

    // gbwtgraph::GBWTGraph::follow_paths(BidirectionalState, bool backward, callback)
    bool follow_paths(BidirectionalState state, bool backward,
                      const std::function<bool(const BidirectionalState&)>& iteratee) const {
        if (backward) state.flip();
        Record rec = record(state.forward.node);
        for (uint32_t r = 0; r < rec.n_edges; r++) {
            if (rec.successor(r) == 0) continue;   // endmarker
            BidirectionalState next = bd_extend_forward(state, rec, r);
            if (backward) next.flip();
            if (!next.empty()) {
                if (!iteratee(next)) return false;
            }
        }
        return true;
    }

    // unidirectional follow_paths (SearchState), minimizer_mapper.cpp:5983
    bool follow_paths(const SearchState& state, const std::function<bool(const SearchState&)>& iteratee) const {
        Record rec = record(state.node);
        for (uint32_t r = 0; r < rec.n_edges; r++) {
            if (rec.successor(r) == 0) continue;
            SearchState next;
            next.node = rec.successor(r);
            next.lo = (int64_t)rec.offset(r) + (int64_t)rec.rank((uint64_t)state.lo, r);
            next.hi = (int64_t)rec.offset(r) + (int64_t)rec.rank((uint64_t)state.hi + 1, r) - 1;
            if (!next.empty()) {
                if (!iteratee(next)) return false;
            }
        }
        return true;
    }

    SearchState get_state(uint32_t v) const {
        SearchState s; s.node = v; s.lo = 0; s.hi = (int64_t)record_size(v) - 1; return s;
    }

    // gbwtgraph::GBWTGraph::bd_find(path)
    BidirectionalState bd_find(const std::vector<uint32_t>& path) const {
        if (path.empty()) return BidirectionalState();
        BidirectionalState state = get_bd_state(path[0]);
        for (size_t i = 1; i < path.size() && !state.empty(); i++) {
            Record rec = record(state.forward.node);
            bool found = false;
            for (uint32_t r = 0; r < rec.n_edges; r++) {
                if (rec.successor(r) == path[i]) { state = bd_extend_forward(state, rec, r); found = true; break; }
            }
            if (!found) return BidirectionalState();
        }
        return state;
    }
};

} // namespace oracle
