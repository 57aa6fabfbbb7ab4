// giraffe_b200.hpp — C++ mirror of the reference's seams over the C ABI (header only).
//
// The C ABI (giraffe_b200.h) is batch-shaped; vg's call sites are object-shaped.  These classes keep vg's names,
// argument meaning and error behaviour so that a call site reads the same on either side:
//   giraffe_b200::MinimizerMapper        minimizer_mapper.hpp:44-100, :108-521 (public parameters), :538-549
//   giraffe_b200::FragmentLengthDistribution   mapper.hpp:83-139
//   giraffe_b200::Alignment / Path / Mapping / Position / Edit   the fields of vg.proto this path reads or writes
// Errors: the reference throws std::runtime_error for per-read failures (caught at giraffe_main.cpp:2456) and
// exits on fatal input errors; here every C-ABI status other than GB_OK becomes std::runtime_error(gb_last_error()),
// and a per-read capacity status (GB_ITEM_*) becomes an exception naming the read.
// map() / map_paired() on single reads are one-element batches: correct but slow; the *_batch forms are the product path.
#pragma once
#include "giraffe_b200.h"

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace giraffe_b200 {

struct Edit { uint32_t from_length = 0, to_length = 0; std::string sequence; };
struct Position { int64_t node_id = 0; uint64_t offset = 0; bool is_reverse = false; };
struct Mapping { Position position; std::vector<Edit> edit; int64_t rank = 0; };
struct Path { std::vector<Mapping> mapping; };
struct Alignment {
    std::string sequence, quality, name;          // quality: raw phred bytes, as vg keeps them
    Path path;
    int32_t score = 0, mapping_quality = 0;
    bool is_secondary = false;                    // mappings 1 .. max_multimaps - 1 (minimizer_mapper.cpp:1205, :2555)
    double identity = 0.0;
    std::map<std::string, double> annotation;     // mapq_uncapped, mapq_explored_cap (minimizer_mapper.cpp:1173-1174), rescued
    std::string fragment_prev, fragment_next;     // mate names, set by map_paired (pair_all, minimizer_mapper.cpp:1280-1300)
};

class FragmentLengthDistribution {
public:
    FragmentLengthDistribution(size_t maximum_sample_size, size_t reestimation_frequency, double robust_estimation_fraction)
        : h(gb_fragment_create(maximum_sample_size, reestimation_frequency, robust_estimation_fraction)) {
        if (!h) throw std::runtime_error(gb_last_error());
    }
    FragmentLengthDistribution() : FragmentLengthDistribution(1000, 1000, 0.95) {}           // minimizer_mapper.cpp:72
    ~FragmentLengthDistribution() { gb_fragment_destroy(h); }
    FragmentLengthDistribution(const FragmentLengthDistribution&) = delete;
    FragmentLengthDistribution& operator=(const FragmentLengthDistribution&) = delete;
    void force_parameters(double mean, double stddev) { gb_fragment_force(h, mean, stddev); }
    void register_fragment_length(int64_t length) { gb_fragment_register(h, length); }
    double mean() const { return gb_fragment_mean(h); }
    double std_dev() const { return gb_fragment_stdev(h); }
    bool is_finalized() const { return gb_fragment_is_finalized(h) != 0; }
    size_t curr_sample_size() const { return (size_t)gb_fragment_sample_size(h); }
    gb_fragment_distribution* handle() { return h; }
private:
    gb_fragment_distribution* h;
};

class MinimizerMapper {
public:
    // the parameters of minimizer_mapper.hpp:108-521 that this path reads, under their own names
    gb_map_params params;
    uint32_t& hit_cap = params.hit_cap; uint32_t& hard_hit_cap = params.hard_hit_cap;
    double& minimizer_score_fraction = params.minimizer_score_fraction;
    uint32_t& max_extensions = params.max_extensions; uint32_t& max_alignments = params.max_alignments;
    uint32_t& max_rescue_attempts = params.max_rescue_attempts; uint32_t& max_fragment_length = params.max_fragment_length;
    uint32_t& rescue_seed_limit = params.rescue_seed_limit; double& rescue_subgraph_stdevs = params.rescue_subgraph_stdevs;

    explicit MinimizerMapper(gb_device* device) : dev(device) { gb_map_params_default(&params); }
    MinimizerMapper(const MinimizerMapper&) = delete;
    MinimizerMapper& operator=(const MinimizerMapper&) = delete;

    // set_alignment_scores(match, mismatch, gap_open, gap_extend, full_length_bonus), minimizer_mapper.cpp:77
    void set_alignment_scores(int8_t match, int8_t mismatch, int8_t gap_open, int8_t gap_extend, int8_t full_length_bonus) {
        gb_scores s{match, mismatch, gap_open, gap_extend, full_length_bonus};
        check(gb_set_scores(dev, &s));
    }
    // fragment length distribution (minimizer_mapper.hpp:538-549)
    bool fragment_distr_is_finalized() { return fragment_length_distr.is_finalized(); }
    void finalize_fragment_length_distr() { gb_fragment_finalize(fragment_length_distr.handle()); }
    void force_fragment_length_distr(double mean, double stdev) { fragment_length_distr.force_parameters(mean, stdev); }
    double get_fragment_length_mean() const { return fragment_length_distr.mean(); }
    double get_fragment_length_stdev() const { return fragment_length_distr.std_dev(); }
    size_t get_fragment_length_sample_size() const { return fragment_length_distr.curr_sample_size(); }

    uint32_t& max_multimaps = params.max_multimaps;

    // vector<Alignment> map(Alignment& aln), minimizer_mapper.hpp:55: the winner first, then up to max_multimaps - 1 secondaries
    std::vector<Alignment> map(Alignment& aln) {
        Packed in; in.add(aln);
        Outputs out(1, params);
        uint64_t nm = 0, ne = 0;
        check(gb_map_batch(dev, &params, 1, in.reads.data(), in.quality_or_null(), in.off.data(), out.aln.data(), out.maps.data(), out.maps.size(),
                           out.edits.data(), out.edits.size(), out.status.data(), &nm, &ne));
        return ranks_of(aln, 0, 1, out);
    }
    // pair<vector<Alignment>, vector<Alignment>> map_paired(Alignment& aln1, Alignment& aln2), minimizer_mapper.hpp:100;
    // needs a finalized distribution, like the reference's overload without the ambiguous-pair buffer
    std::pair<std::vector<Alignment>, std::vector<Alignment>> map_paired(Alignment& aln1, Alignment& aln2) {
        if (!fragment_distr_is_finalized()) throw std::runtime_error("map_paired: the fragment length distribution is not finalized");
        Packed in; in.add(aln1); in.add(aln2);
        Outputs out(2, params);
        gb_map_params p = params; p.fragment_mean = get_fragment_length_mean(); p.fragment_stdev = get_fragment_length_stdev();
        uint64_t nm = 0, ne = 0;
        check(gb_map_paired_batch(dev, &p, 2, in.reads.data(), in.quality_or_null(), in.off.data(), out.aln.data(), out.maps.data(), out.maps.size(),
                                  out.edits.data(), out.edits.size(), out.status.data(), &nm, &ne));
        std::pair<std::vector<Alignment>, std::vector<Alignment>> res{ranks_of(aln1, 0, 2, out), ranks_of(aln2, 1, 2, out)};
        for (size_t j = 0; j < res.first.size() && j < res.second.size(); j++) { res.first[j].fragment_next = aln2.name; res.second[j].fragment_prev = aln1.name; }
        return res;
    }

    // batch forms: the alignments are filled in place with the PRIMARY mapping (path, score, mapping_quality, identity, annotations)
    void map_batch(std::vector<Alignment>& batch) {
        Packed in; for (Alignment& a : batch) in.add(a);
        Outputs out(in.n(), params);
        uint64_t nm = 0, ne = 0;
        check(gb_map_batch(dev, &params, in.n(), in.reads.data(), in.quality_or_null(), in.off.data(), out.aln.data(), out.maps.data(), out.maps.size(),
                           out.edits.data(), out.edits.size(), out.status.data(), &nm, &ne));
        for (uint32_t i = 0; i < in.n(); i++) fill(batch[i], i, out);
    }
    // the whole paired job of giraffe_main.cpp:2246-2400: trains the distribution on the head of the batch when it is not
    // finalized yet (gb_map_paired_job), then maps everything else paired; route[i] receives GB_PAIR_* when given
    void map_paired_batch(std::vector<std::pair<Alignment, Alignment>>& batch, std::vector<uint8_t>* route = nullptr) {
        Packed in; for (auto& p : batch) { in.add(p.first); in.add(p.second); }
        Outputs out(in.n(), params);
        std::vector<uint8_t> r(batch.size());
        uint64_t nm = 0, ne = 0;
        check(gb_map_paired_job(dev, &params, fragment_length_distr.handle(), 0, in.n(), in.reads.data(), in.quality_or_null(), in.off.data(), out.aln.data(),
                                out.maps.data(), out.maps.size(), out.edits.data(), out.edits.size(), out.status.data(), r.data(), &nm, &ne));
        for (size_t i = 0; i < batch.size(); i++) {
            fill(batch[i].first, (uint32_t)(2 * i), out); fill(batch[i].second, (uint32_t)(2 * i + 1), out);
            batch[i].first.fragment_next = batch[i].second.name; batch[i].second.fragment_prev = batch[i].first.name;      // pair_all, also for training pairs (:1350)
        }
        if (route) *route = r;
    }

private:
    gb_device* dev;
    FragmentLengthDistribution fragment_length_distr;

    static void check(int rc) { if (rc != GB_OK) throw std::runtime_error(gb_last_error()); }

    struct Packed {
        std::vector<uint8_t> reads, quals; std::vector<uint64_t> off{0}; bool any_quality = false, all_quality = true;
        void add(const Alignment& a) {
            reads.insert(reads.end(), a.sequence.begin(), a.sequence.end());
            if (a.quality.size() == a.sequence.size() && !a.quality.empty()) { quals.insert(quals.end(), a.quality.begin(), a.quality.end()); any_quality = true; }
            else { quals.insert(quals.end(), a.sequence.size(), 0); if (!a.sequence.empty()) all_quality = false; }
            off.push_back(reads.size());
        }
        uint32_t n() const { return (uint32_t)(off.size() - 1); }
        // check_quality_length (giraffe_main.cpp:2317): qualities are all there or not used
        const uint8_t* quality_or_null() const { return any_quality && all_quality ? quals.data() : nullptr; }
    };
    struct Outputs {
        std::vector<gb_alignment> aln; std::vector<gb_mapping> maps; std::vector<uint32_t> edits; std::vector<uint8_t> status;
        // n * max_multimaps records, rank-major: record j * n + read (gb_map_batch)
        Outputs(uint32_t n, const gb_map_params& p) : aln((size_t)n * p.max_multimaps), maps((size_t)n * p.max_multimaps * p.mapping_cap_per_read + 1),
                                                      edits((size_t)n * p.max_multimaps * p.edit_cap_per_read + 1), status(n) {}
    };
    // the mappings of read i (of n) that exist, primary first
    std::vector<Alignment> ranks_of(const Alignment& in, uint32_t i, uint32_t n, const Outputs& out) const {
        std::vector<Alignment> res;
        for (uint32_t j = 0; j < params.max_multimaps; j++) {
            const gb_alignment& r = out.aln[(size_t)j * n + i];
            if (r.flags & GB_ALN_ABSENT) break;
            res.push_back(in);
            fill(res.back(), i, out, (size_t)j * n + i);
        }
        return res;
    }
    // one record -> vg's Alignment fields (the inverse of what map_from_extensions sets, minimizer_mapper.cpp:1146-1216)
    static void fill(Alignment& a, uint32_t i, const Outputs& out, size_t record = (size_t)-1) {
        if (record == (size_t)-1) record = i;
        if (out.status[i] != GB_ITEM_OK) throw std::runtime_error("read " + (a.name.empty() ? std::to_string(i) : a.name) + ": per-read capacity exceeded (status " + std::to_string(out.status[i]) + ")");
        const gb_alignment& r = out.aln[record];
        a.path.mapping.clear();
        a.score = r.score; a.mapping_quality = r.mapq; a.is_secondary = (r.flags & GB_ALN_SECONDARY) != 0;
        a.annotation["mapq_uncapped"] = r.mapq_uncapped; a.annotation["mapq_explored_cap"] = r.mapq_explored_cap;
        if (r.flags & GB_ALN_RESCUED) a.annotation["rescued"] = 1.0;
        uint64_t q = 0, matches = 0; uint32_t e = r.edit_off;
        for (uint32_t m = 0; m < r.n_mappings; m++) {
            const gb_mapping& gm = out.maps[r.mapping_off + m];
            Mapping mp; mp.position.node_id = gm.node >> 1; mp.position.is_reverse = gm.node & 1u; mp.position.offset = gm.offset; mp.rank = m + 1;
            for (uint32_t j = 0; j < gm.n_edits; j++, e++) {
                const uint32_t wd = out.edits[e], op = wd & 3u, len = op == GB_EDIT_SUB ? 1u : wd >> 4;
                Edit ed;
                if (op == GB_EDIT_MATCH) { ed.from_length = ed.to_length = len; matches += len; }
                else if (op == GB_EDIT_SUB) { ed.from_length = ed.to_length = 1; ed.sequence = a.sequence.substr(q, 1); }
                else if (op == GB_EDIT_INS) { ed.to_length = len; ed.sequence = a.sequence.substr(q, len); }
                else ed.from_length = len;
                q += ed.to_length;
                mp.edit.push_back(std::move(ed));
            }
            a.path.mapping.push_back(std::move(mp));
        }
        a.identity = a.sequence.empty() || r.n_mappings == 0 ? 0.0 : (double)matches / (double)a.sequence.size();
    }
};

// ---- chaining route seams (algorithms/chain_items.hpp): anchors, the transition candidates, find_best_chains ------------------------
namespace algorithms {

struct ChainScoringScheme { int item_bonus = 0; double gap_scale = 1.0; int recombination_penalty = 0; int consistency_bonus = 0; };   // chain_items.hpp:407-418
using Anchor = gb_chain_anchor;                 // the fields chaining reads of algorithms::Anchor (chain_items.hpp:48-276)
using transition_candidate = gb_chain_candidate;

// MinimizerMapper::to_anchor for a whole seed list (minimizer_mapper_from_chains.cpp:3969-4038): seed i = (oriented node, offset) of
// the minimizer (pin offset, is_reverse, length) at the same index
inline std::vector<Anchor> to_anchors(const gb_flat_index& index, const gb_scores& scores, const std::vector<std::pair<uint32_t, uint32_t>>& seeds,
                                      const std::vector<uint32_t>& pin_offset, const std::vector<uint8_t>& is_reverse, const std::vector<uint32_t>& length) {
    if (pin_offset.size() != seeds.size() || is_reverse.size() != seeds.size() || length.size() != seeds.size()) throw std::runtime_error("to_anchors: one minimizer per seed");
    std::vector<uint32_t> pos; pos.reserve(2 * seeds.size() + 2);
    for (const auto& s : seeds) { pos.push_back(s.first); pos.push_back(s.second); }
    pos.push_back(0); pos.push_back(0);
    std::vector<Anchor> out(seeds.size() + 1);
    if (gb_chain_anchors(&index, &scores, (uint32_t)seeds.size(), pos.data(), pin_offset.data(), is_reverse.data(), length.data(), nullptr, out.data()) != GB_OK)
        throw std::runtime_error(gb_last_error());
    out.resize(seeds.size());
    return out;
}

// what zip_tree_transition_iterator enumerates for one read (chain_items.cpp:116-260), from the library's distance model
inline std::vector<transition_candidate> transition_candidates(gb_device* dev, const std::vector<std::pair<uint32_t, uint32_t>>& seeds, uint64_t max_graph_lookback_bases) {
    std::vector<uint32_t> pos; pos.reserve(2 * seeds.size() + 2);
    for (const auto& s : seeds) { pos.push_back(s.first); pos.push_back(s.second); }
    pos.push_back(0); pos.push_back(0);
    const uint64_t seed_off[2] = {0, seeds.size()};
    uint64_t cand_off[2] = {0, 0};
    int rc = gb_chain_candidates_batch(dev, 1, pos.data(), seed_off, max_graph_lookback_bases, nullptr, 0, cand_off);      // sizes first
    if (rc != GB_OK && rc != GB_ERR_CAPACITY) throw std::runtime_error(gb_last_error());
    std::vector<transition_candidate> out(cand_off[1] + 1);
    if (gb_chain_candidates_batch(dev, 1, pos.data(), seed_off, max_graph_lookback_bases, out.data(), cand_off[1], cand_off) != GB_OK) throw std::runtime_error(gb_last_error());
    out.resize(cand_off[1]);
    return out;
}

// find_best_chains(to_chain, ..., for_each_transition, scheme, max_chains, max_indel_bases), chain_items.hpp:576: (score, anchor
// indices left to right) per chain, best first.  to_chain in read order (the reference's VectorView), candidates over its indices.
inline std::vector<std::pair<int, std::vector<size_t>>> find_best_chains(gb_device* dev, const std::vector<Anchor>& to_chain,
                                                                         const std::vector<transition_candidate>& candidates,
                                                                         const ChainScoringScheme& scheme = ChainScoringScheme(), size_t max_chains = 1,
                                                                         size_t max_indel_bases = 100, size_t max_read_lookback_bases = ~(size_t)0) {
    std::vector<std::pair<int, std::vector<size_t>>> result;
    if (to_chain.empty()) { result.push_back({0, {}}); return result; }                     // :748-755
    gb_chain_params P; gb_chain_params_default(&P);
    P.item_bonus = scheme.item_bonus; P.gap_scale = scheme.gap_scale; P.recombination_penalty = scheme.recombination_penalty;
    P.consistency_bonus = scheme.consistency_bonus; P.max_chains = (uint32_t)max_chains; P.max_indel_bases = max_indel_bases;
    P.max_read_lookback_bases = max_read_lookback_bases;
    const uint64_t aoff[2] = {0, to_chain.size()}, coff[2] = {0, candidates.size()};
    const size_t n = to_chain.size();
    std::vector<int32_t> dp_score(n), chain_score(max_chains); std::vector<uint32_t> dp_source(n), dp_rec(n), chain_begin(max_chains), chain_count(max_chains), items(n);
    std::vector<uint64_t> dp_paths(n); uint32_t n_chains = 0;
    static const transition_candidate none{0, 0, 0};
    if (gb_chain_batch(dev, &P, 1, to_chain.data(), aoff, candidates.empty() ? &none : candidates.data(), coff, dp_score.data(), dp_source.data(), dp_paths.data(),
                       dp_rec.data(), &n_chains, chain_score.data(), chain_begin.data(), chain_count.data(), items.data()) != GB_OK)
        throw std::runtime_error(gb_last_error());
    for (uint32_t c = 0; c < n_chains; c++)
        result.push_back({chain_score[c], std::vector<size_t>(items.begin() + chain_begin[c], items.begin() + chain_begin[c] + chain_count[c])});
    return result;
}

} // namespace algorithms

} // namespace giraffe_b200
