// ORACLE — TEST INFRASTRUCTURE ONLY.
// chain.cpp — CPU restatement of vg's anchor chaining (algorithms/chain_items.cpp @ fd49b9a9), statement by
// statement and in the reference's own order of operations:
//   add_transition_if_legal        chain_items.cpp:262-355
//   the transition sort            :143-148 (by the destination's read start)
//   score_chain_gap                :364-372
//   check_recombination            :376-382
//   TracedScore                    chain_items.hpp:296-354, chain_items.cpp:47-96
//   chain_items_dp                 :384-640
//   chain_items_traceback          :642-735
//   find_best_chains               :737-800 (score = best - penalty)
// The zip-code tree (the transition iterator's source of candidate pairs) is the caller's input here.
// Pinned by the four find_best_chain cases of src/unittest/chain_items.cpp:96-155 (tests/test_chain_golden.py).
// Where the reference's std::sort leaves equal keys in unspecified order this file uses std::stable_sort.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

constexpr size_t NOWHERE = std::numeric_limits<size_t>::max();

struct Anchor {
    gb_chain_anchor a;
    size_t read_start() const { return a.read_start; }
    size_t length() const { return a.length; }
    size_t read_end() const { return (size_t)a.read_start + a.length; }
    size_t read_exclusion_start() const { return (size_t)a.read_start - a.margin_before; }
    size_t read_exclusion_end() const { return read_end() + a.margin_after; }
    int score() const { return a.score; }
    size_t start_hint_offset() const { return a.start_hint_offset; }
    size_t end_hint_offset() const { return a.end_hint_offset; }
    size_t base_seed_length() const { return a.base_seed_length; }
    uint64_t anchor_start_paths() const { return a.start_paths; }
    uint64_t anchor_end_paths() const { return a.end_paths; }
};

struct TracedScore {                       // chain_items.hpp:296-354
    int score; size_t source; uint64_t paths; size_t rec_num = 0;
    bool operator>(const TracedScore& o) const { return score > o.score || (score == o.score && source > o.source); }
    TracedScore add_points(int adj) const { return {score + adj, source, paths, rec_num}; }
    TracedScore set_shared_paths(uint64_t first, uint64_t second) const {          // chain_items.cpp:69-96
        uint64_t updated; size_t rec = rec_num;
        if (first == second) {
            if ((paths & first) == 0) { updated = first; rec++; }
            else updated = paths & first;
        } else updated = second;
        return {score, source, updated, rec};
    }
};

struct Transition { size_t from_anchor, to_anchor, indel_size; };

size_t get_read_distance(const Anchor& from, const Anchor& to) {                   // :984-989
    if (to.read_start() < from.read_end()) return NOWHERE;
    return to.read_start() - from.read_end();
}

void add_transition_if_legal(std::vector<Transition>& transitions, const std::vector<Anchor>& to_chain,
                             size_t max_read_lookback_bases, size_t max_indel_bases,
                             size_t from_anchor, size_t to_anchor, size_t graph_distance) {    // :262-355
    const Anchor& source_anchor = to_chain[from_anchor];
    const Anchor& dest_anchor = to_chain[to_anchor];
    size_t read_distance = get_read_distance(source_anchor, dest_anchor);
    if (read_distance == NOWHERE) return;
    if (read_distance > max_read_lookback_bases) return;
    if (source_anchor.read_exclusion_end() > dest_anchor.read_exclusion_start()) return;
    size_t distance_to_remove = dest_anchor.start_hint_offset() + source_anchor.end_hint_offset();
    if (distance_to_remove > graph_distance) return;
    graph_distance -= distance_to_remove;
    size_t indel_size = (read_distance > graph_distance) ? read_distance - graph_distance : graph_distance - read_distance;
    if (indel_size > max_indel_bases) return;
    transitions.push_back({from_anchor, to_anchor, indel_size});
}

int score_chain_gap(size_t distance_difference, size_t base_seed_length) {        // :364-372
    if (distance_difference == 0) return 0;
    // product and sum as two IEEE operations (the volatile keeps -march=native builds from contracting them into an FMA)
    volatile double product = 0.01 * base_seed_length * distance_difference;
    return product + 0.5 * log2(distance_difference);
}

int check_recombination(const TracedScore& from, const Anchor& to) {              // :376-382
    return (from.paths & to.anchor_start_paths()) == 0 ? 1 : 0;
}

}  // namespace

extern "C" int oracle_chain(const gb_chain_params* P, uint32_t n_anchors, const gb_chain_anchor* anchors,
                            uint64_t n_candidates, const gb_chain_candidate* candidates,
                            int32_t* dp_score, uint32_t* dp_source, uint64_t* dp_paths, uint32_t* dp_rec,
                            uint32_t* n_chains, int32_t* chain_score, uint32_t* chain_begin, uint32_t* chain_count,
                            uint32_t* chain_items, uint32_t* candidate_indel) {
    *n_chains = 0;
    if (n_anchors == 0) return 0;                          // find_best_chains :748-755: one empty chain of score 0
    std::vector<Anchor> to_chain(n_anchors);
    for (uint32_t i = 0; i < n_anchors; i++) to_chain[i].a = anchors[i];

    // ---- the transition iterator: legal candidates, sorted by the destination's read start (:143-152) ----
    std::vector<Transition> all_transitions;
    for (uint64_t c = 0; c < n_candidates; c++) {
        if (candidates[c].from >= n_anchors || candidates[c].to >= n_anchors) return -1;
        const size_t before = all_transitions.size();
        add_transition_if_legal(all_transitions, to_chain, P->max_read_lookback_bases, P->max_indel_bases,
                                candidates[c].from, candidates[c].to, candidates[c].graph_distance);
        if (candidate_indel) candidate_indel[c] = all_transitions.size() > before ? (uint32_t)all_transitions.back().indel_size : 0xffffffffu;
    }
    std::stable_sort(all_transitions.begin(), all_transitions.end(), [&](const Transition& a, const Transition& b) {
        return to_chain[a.to_anchor].read_start() < to_chain[b.to_anchor].read_start();
    });

    // ---- chain_items_dp (:384-640) ----
    size_t base_seed_length = 0;
    for (auto& anchor : to_chain) base_seed_length += anchor.base_seed_length();
    base_seed_length /= to_chain.size();
    std::vector<TracedScore> chain_scores(n_anchors);
    std::vector<int> eval_bonuses(n_anchors, P->consistency_bonus);
    for (size_t i = 0; i < n_anchors; i++) chain_scores[i] = {(int)(to_chain[i].score() + P->item_bonus), NOWHERE, to_chain[i].anchor_end_paths()};
    for (const Transition& transition : all_transitions) {
        const Anchor& here = to_chain[transition.to_anchor];
        auto item_points = here.score() + P->item_bonus;
        {
            TracedScore from_nowhere = {(int)item_points, NOWHERE, here.anchor_end_paths()};
            int nowhere_bonus = P->consistency_bonus;
            int eval_nowhere = from_nowhere.score + nowhere_bonus;
            int eval_current = chain_scores[transition.to_anchor].score + eval_bonuses[transition.to_anchor];
            if (eval_nowhere > eval_current) { chain_scores[transition.to_anchor] = from_nowhere; eval_bonuses[transition.to_anchor] = nowhere_bonus; }
            else if (eval_nowhere == eval_current && from_nowhere > chain_scores[transition.to_anchor]) { chain_scores[transition.to_anchor] = from_nowhere; eval_bonuses[transition.to_anchor] = nowhere_bonus; }
        }
        int jump_points = -score_chain_gap(transition.indel_size, base_seed_length) * P->gap_scale;
        jump_points -= check_recombination(chain_scores[transition.from_anchor], here) * P->recombination_penalty;
        TracedScore source_score = chain_scores[transition.from_anchor];          // TracedScore::score_from :58-64
        source_score.source = transition.from_anchor;
        TracedScore from_source_score = source_score.add_points(jump_points + item_points)
                                                    .set_shared_paths(here.anchor_start_paths(), here.anchor_end_paths());
        int eval_bonus_from = 0;
        if (P->consistency_bonus > 0) {
            int pre_count = __builtin_popcountll(source_score.paths);
            if (pre_count > 0 && (source_score.paths & here.anchor_start_paths()) != 0) {
                int post_count = __builtin_popcountll(from_source_score.paths);
                eval_bonus_from = (P->consistency_bonus * post_count) / pre_count;
            }
        }
        TracedScore& current_best = chain_scores[transition.to_anchor];
        int eval_from = from_source_score.score + eval_bonus_from;
        int eval_best = current_best.score + eval_bonuses[transition.to_anchor];
        if (eval_from > eval_best || (eval_from == eval_best && from_source_score > current_best)) {
            current_best = from_source_score;
            eval_bonuses[transition.to_anchor] = eval_bonus_from;
        }
    }
    TracedScore best_score = {0, NOWHERE, 0};                                      // TracedScore::unset, max_in :47-56
    for (size_t to_anchor = 0; to_anchor < n_anchors; ++to_anchor) {
        const TracedScore& option = chain_scores[to_anchor];
        if (option.score > best_score.score || best_score.source == NOWHERE) { best_score = option; best_score.source = to_anchor; }
    }
    for (size_t i = 0; i < n_anchors; i++) {
        dp_score[i] = chain_scores[i].score;
        dp_source[i] = chain_scores[i].source == NOWHERE ? 0xffffffffu : (uint32_t)chain_scores[i].source;
        dp_paths[i] = chain_scores[i].paths; dp_rec[i] = (uint32_t)chain_scores[i].rec_num;
    }

    // ---- chain_items_traceback (:642-735) ----
    std::vector<std::pair<std::vector<size_t>, int>> tracebacks;
    std::vector<size_t> starts_in_score_order(n_anchors);
    for (size_t i = 0; i < n_anchors; i++) starts_in_score_order[i] = i;
    std::stable_sort(starts_in_score_order.begin(), starts_in_score_order.end(), [&](const size_t& a, const size_t& b) { return chain_scores[a] > chain_scores[b]; });
    std::vector<bool> item_is_used(n_anchors, false);
    for (auto& trace_from : starts_in_score_order) {
        if (item_is_used[trace_from]) continue;
        std::vector<size_t> traceback;
        traceback.push_back(trace_from);
        int penalty = best_score.score - chain_scores[trace_from].score;
        size_t here = trace_from;
        while (here != NOWHERE) {
            item_is_used[here] = true;
            size_t next = chain_scores[here].source;
            if (next != NOWHERE) {
                if (item_is_used[next]) {
                    penalty += chain_scores[here].score;
                    penalty -= (to_chain[here].score() + P->item_bonus);
                    break;
                } else traceback.push_back(next);
            }
            here = next;
        }
        tracebacks.emplace_back();
        tracebacks.back().second = penalty;
        std::copy(traceback.rbegin(), traceback.rend(), std::back_inserter(tracebacks.back().first));
    }
    std::stable_sort(tracebacks.begin(), tracebacks.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
    if (tracebacks.size() > P->max_chains) tracebacks.resize(P->max_chains);

    // ---- find_best_chains (:780-800): score = best - penalty ----
    uint32_t w = 0;
    for (size_t c = 0; c < tracebacks.size(); c++) {
        chain_score[c] = best_score.score - tracebacks[c].second;
        chain_begin[c] = w; chain_count[c] = (uint32_t)tracebacks[c].first.size();
        for (size_t x : tracebacks[c].first) chain_items[w++] = (uint32_t)x;
    }
    *n_chains = (uint32_t)tracebacks.size();
    return 0;
}

// MinimizerMapper::to_anchor, minimizer_mapper_from_chains.cpp:3978-4038 (test entry; node_len = length of the seed's node)
extern "C" void oracle_to_anchor(const gb_scores* scores, uint32_t node_len, uint32_t seed_offset, uint32_t min_offset, int min_is_reverse,
                                 uint32_t min_length, uint64_t paths, gb_chain_anchor* out) {
    size_t length, read_start, hint_start, margin_left, margin_right;
    if (min_is_reverse) {
        size_t graph_end_offset = (size_t)seed_offset + 1;
        length = std::min((size_t)min_length, graph_end_offset);
        margin_left = (size_t)min_length - length;
        margin_right = 0;
        read_start = (size_t)min_offset + 1 - length;
        hint_start = length - 1;
    } else {
        length = std::min((size_t)min_length, (size_t)node_len - seed_offset);
        margin_left = 0;
        margin_right = (size_t)min_length - length;
        read_start = min_offset;
        hint_start = 0;
    }
    int score = scores->match * (int)(margin_left + length + margin_right);         // score_exact_match(aln, read_start - margin_left, ...)
    // Anchor::Anchor(read_start, graph_start, length, margin_before, margin_after, score, seed, hint, hint_start, skippable, paths), chain_items.hpp:231-244
    out->read_start = (uint32_t)read_start; out->length = (uint32_t)length; out->margin_before = (uint32_t)margin_left; out->margin_after = (uint32_t)margin_right;
    out->score = score; out->start_hint_offset = (uint32_t)hint_start; out->end_hint_offset = (uint32_t)(length - hint_start);
    out->base_seed_length = (uint32_t)(margin_left + length + margin_right); out->start_paths = out->end_paths = paths;
}
