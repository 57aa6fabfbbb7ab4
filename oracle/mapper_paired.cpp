// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of MinimizerMapper::map_paired with a
// finalized (forced) fragment length distribution, src/minimizer_mapper.cpp:1462-2942 @ fd49b9a9,
// including the rescue branch (:2288-2457; attempt_rescue itself is oracle/rescue.cpp) and the
// no-rescue branch for max_rescue_attempts = 0 (:2238-2287).  Not restated: supplementary
// alignments (find_supplementaries, off by default).
//   joint clustering: SnarlDistanceIndexClusterer::cluster_seeds snarl_seed_clusterer.cpp:65-145
//   pair score:       score_alignment_pair :6017-6028, distance_between :3879-3903
#include "fragment.hpp"
#include "mapper_common.hpp"
#include "tail_align.hpp"

#include <algorithm>
#include <array>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>

namespace oracle {

// from mapper.cpp
double recover_log_base(const gb_scores& s);
int32_t compute_max_mapping_quality(const std::vector<double>& scores, double log_base);
int32_t compute_max_mapping_quality(const std::vector<double>& scores, double log_base, const std::vector<double>* multiplicities);
Alignment attempt_rescue(const gb_flat_index* ix, const gb_scores& sc, const gb_map_params& P, const Alignment& anchor,
                         const std::string& sequence, const std::vector<Minimizer>& minimizers, bool rescue_forward, MapCounters* counters);
std::vector<Minimizer> find_minimizers(const gb_flat_index* ix, const gb_map_params& P, const std::string& sequence);
std::vector<size_t> sort_minimizers_by_score(const std::vector<Minimizer>& minimizers, LazyRNG& rng);
std::vector<Seed> find_seeds(const gb_flat_index* ix, const gb_map_params& P, const std::vector<Minimizer>& minimizers, size_t read_len);
int64_t unoriented_distance(const gb_flat_index* ix, const Seed& a, const Seed& b);
void score_cluster(Cluster& cluster, const std::vector<Minimizer>& minimizers, const std::vector<Seed>& seeds, size_t seq_length);
int score_extension_group(size_t seq_len, const std::vector<GaplessExtension>& ext, int max_mismatches, int gap_open_penalty, int gap_extend_penalty);
double faster_cap(const std::vector<Minimizer>& minimizers, std::vector<size_t>& explored, const std::string& sequence, const std::string& quality);
Alignment map_from_extensions(const gb_flat_index* ix, const gb_scores& scores, const gb_map_params& P,
                              const std::string& sequence, const std::string& quality, MapCounters* counters);
int pack_secondaries(const Alignment& primary, const gb_map_params& P, uint32_t n_reads, uint32_t r, uint32_t extra_flags,
                     gb_alignment* aln, gb_mapping* mappings, uint32_t* edits);
int pack_alignment(const Alignment& a, uint32_t read_id, gb_alignment* rec, gb_mapping* mappings, uint32_t mapping_cap,
                   uint32_t* edits, uint32_t edit_cap, uint32_t mapping_base, uint32_t edit_base);

namespace {

char complement(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }
std::string reverse_complement(const std::string& s) { std::string r(s.rbegin(), s.rend()); for (char& c : r) c = complement(c); return r; }

// reverse_complement_alignment_in_place (alignment.cpp:3338) on the path
void reverse_complement_path(std::vector<Mapping>& path, const Graph& g) {
    std::vector<Mapping> reversed;
    for (int64_t i = (int64_t)path.size() - 1; i >= 0; i--) {
        const Mapping& m = path[i];
        Mapping r; r.node = m.node ^ 1u;
        size_t used = 0; for (const Edit& e : m.edits) used += e.from_length;
        r.offset = (uint32_t)(g.get_length(m.node) - used - m.offset);
        for (int64_t k = (int64_t)m.edits.size() - 1; k >= 0; k--) { Edit e = m.edits[k]; e.sequence = reverse_complement(e.sequence); r.edits.push_back(std::move(e)); }
        reversed.push_back(std::move(r));
    }
    path.swap(reversed);
}

struct UnionFind {
    std::vector<size_t> p;
    explicit UnionFind(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    size_t find(size_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(size_t a, size_t b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

// minimum_distance(pos1, pos2) with cut-style offsets (offset may equal the node length),
// both positions oriented; unreachable = size_t max, which vg stores into int64_t (= -1)
// (minimizer_mapper.cpp:3879-3884).
int64_t oriented_distance(const gb_flat_index* ix, uint32_t node_a, uint32_t off_a, uint32_t node_b, uint32_t off_b) {
    const int64_t UNREACHABLE = (int64_t)std::numeric_limits<size_t>::max();
    if ((node_a & 1) != (node_b & 1)) return UNREACHABLE;
    uint32_t src_id = node_a >> 1, dst_id = node_b >> 1;
    int64_t src_len = ix->nodes[node_a].len, dst_len = ix->nodes[node_b].len;
    int64_t src_off = off_a, dst_off = off_b;
    if (node_a & 1) {
        // reverse strand: the same walk read on the forward strand goes from b' to a'
        src_id = node_b >> 1; src_len = ix->nodes[node_b].len; src_off = src_len - (int64_t)off_b;
        dst_id = node_a >> 1; dst_len = ix->nodes[node_a].len; dst_off = dst_len - (int64_t)off_a;
    }
    (void)dst_len;
    const gb_dist_payload& ps = ix->dist[src_id]; const gb_dist_payload& pd = ix->dist[dst_id];
    if (ps.component != pd.component) return UNREACHABLE;
    if (src_id == dst_id) return dst_off >= src_off ? dst_off - src_off : UNREACHABLE;
    if (ps.slot < pd.slot) return (src_len - src_off) + ((int64_t)(int32_t)pd.x_in - (int64_t)(int32_t)ps.x_out) + dst_off;
    if (ps.slot == pd.slot) { const int64_t t = site_distance(ix, ps, pd); if (t >= 0) return (src_len - src_off) + t + dst_off; }
    return UNREACHABLE;
}

} // namespace

struct PairResult { Alignment aln[2]; };

PairResult map_paired(const gb_flat_index* ix, const gb_scores& scores, const gb_map_params& P,
                      const std::string& seq1, const std::string& qual1, const std::string& seq2_in, const std::string& qual2_in,
                      MapCounters* counters) {
    Graph g(ix);
    PairResult result;
    const double log_base = recover_log_base(scores);
    const double frag_mean = P.fragment_mean, frag_sd = P.fragment_stdev;
    // :1470
    int64_t fragment_distance_limit = (int64_t)(frag_mean + P.paired_distance_stdevs * frag_sd);
    std::array<std::string, 2> seqs{seq1, reverse_complement(seq2_in)};
    std::array<std::string, 2> quals{qual1, std::string(qual2_in.rbegin(), qual2_in.rend())};
    size_t read_limit = std::max<size_t>(P.distance_limit, seqs[0].size() + 50);
    if (fragment_distance_limit < (int64_t)std::max<size_t>(P.distance_limit, seq1.size() + 50)) {
        // a distribution the clusterer cannot use: both ends single-ended, emitted as a pair (:1471-1496)
        PairResult fallback;
        fallback.aln[0] = map_from_extensions(ix, scores, P, seq1, qual1, counters);
        fallback.aln[1] = map_from_extensions(ix, scores, P, seq2_in, qual2_in, counters);
        return fallback;
    }

    LazyRNG rng([&]() { return seqs[0] + seqs[1]; });

    std::array<std::vector<Minimizer>, 2> minimizers_by_read;
    for (int r : {0, 1}) {
        std::vector<Minimizer> in_read = find_minimizers(ix, P, seqs[r]);
        std::vector<size_t> order = sort_minimizers_by_score(in_read, rng);
        for (size_t i : order) minimizers_by_read[r].push_back(in_read[i]);
    }
    std::vector<std::vector<Seed>> seeds_by_read(2);
    for (int r : {0, 1}) seeds_by_read[r] = find_seeds(ix, P, minimizers_by_read[r], seqs[r].size());

    // ---- joint clustering (snarl_seed_clusterer.cpp:65-145) ---------------------------------
    std::vector<std::vector<Cluster>> all_clusters(2);
    {
        const size_t n0 = seeds_by_read[0].size(), n1 = seeds_by_read[1].size();
        UnionFind frag(n0 + n1);
        auto seed_at = [&](size_t i) -> const Seed& { return i < n0 ? seeds_by_read[0][i] : seeds_by_read[1][i - n0]; };
        for (size_t i = 0; i < n0 + n1; i++) for (size_t j = 0; j < i; j++) {
            int64_t d = unoriented_distance(ix, seed_at(i), seed_at(j));
            if (d != std::numeric_limits<int64_t>::max() && d <= fragment_distance_limit) frag.unite(i, j);
        }
        for (int r : {0, 1}) {
            const auto& seeds = seeds_by_read[r];
            UnionFind uf(seeds.size());
            for (size_t i = 0; i < seeds.size(); i++) for (size_t j = 0; j < i; j++) {
                int64_t d = unoriented_distance(ix, seeds[i], seeds[j]);
                if (d != std::numeric_limits<int64_t>::max() && d <= (int64_t)read_limit) uf.unite(i, j);
            }
            std::map<size_t, size_t> root_to_cluster;
            for (size_t i = 0; i < seeds.size(); i++) {
                size_t root = uf.find(i);
                auto it = root_to_cluster.find(root);
                if (it == root_to_cluster.end()) { it = root_to_cluster.emplace(root, all_clusters[r].size()).first; all_clusters[r].emplace_back(); }
                all_clusters[r][it->second].seeds.push_back(i);
            }
            std::sort(all_clusters[r].begin(), all_clusters[r].end(), [](const Cluster& a, const Cluster& b) { return a.seeds.front() < b.seeds.front(); });
        }
        std::map<size_t, size_t> old_to_new; size_t curr_index = 0, offset = 0;
        for (int r : {0, 1}) {
            for (Cluster& c : all_clusters[r]) {
                size_t head = frag.find(offset + c.seeds[0]);
                auto it = old_to_new.find(head);
                if (it == old_to_new.end()) it = old_to_new.emplace(head, curr_index++).first;
                c.fragment = it->second;
            }
            offset += seeds_by_read[r].size();
        }
    }
    if (counters) for (int r : {0, 1}) { counters->minimizers += minimizers_by_read[r].size(); counters->seeds += seeds_by_read[r].size(); counters->clusters += all_clusters[r].size(); }

    size_t max_fragment_num = 0;
    for (int r : {0, 1}) for (auto& c : all_clusters[r]) max_fragment_num = std::max(max_fragment_num, c.fragment);
    std::vector<bool> has_first_read(max_fragment_num + 1, false), fragment_cluster_has_pair(max_fragment_num + 1, false);
    bool found_paired_cluster = false;
    for (auto& c : all_clusters[0]) has_first_read[c.fragment] = true;
    for (auto& c : all_clusters[1]) { fragment_cluster_has_pair[c.fragment] = has_first_read[c.fragment]; if (has_first_read[c.fragment]) found_paired_cluster = true; }

    std::array<std::vector<double>, 2> cluster_score_by_fragment, cluster_coverage_by_fragment;
    for (int r : {0, 1}) { cluster_score_by_fragment[r].assign(max_fragment_num + 1, 0.0); cluster_coverage_by_fragment[r].assign(max_fragment_num + 1, 0.0); }
    auto total_score = [&](size_t f) { return cluster_score_by_fragment[0][f] + cluster_score_by_fragment[1][f]; };
    auto total_coverage = [&](size_t f) { return cluster_coverage_by_fragment[0][f] + cluster_coverage_by_fragment[1][f]; };
    for (int r : {0, 1}) for (Cluster& c : all_clusters[r]) {
        score_cluster(c, minimizers_by_read[r], seeds_by_read[r], seqs[r].length());
        cluster_score_by_fragment[r][c.fragment] = std::max(cluster_score_by_fragment[r][c.fragment], c.score);
        cluster_coverage_by_fragment[r][c.fragment] = std::max(cluster_coverage_by_fragment[r][c.fragment], c.coverage);
    }
    // better_cluster_count (:1657-1690)
    std::vector<size_t> frag_order(max_fragment_num + 1);
    std::iota(frag_order.begin(), frag_order.end(), 0);
    sort_shuffling_ties(frag_order.begin(), frag_order.end(), [&](size_t a, size_t b) {
        return total_coverage(a) + total_score(a) > total_coverage(b) + total_score(b);
    }, rng);
    std::vector<size_t> better_cluster_count(max_fragment_num + 1);
    double prev_score_sum = 0.0;
    for (int rank = (int)frag_order.size() - 1; rank >= 0; rank--) {
        size_t f = frag_order[rank];
        if (rank == (int)frag_order.size() - 1) better_cluster_count[f] = rank + 1;
        else {
            size_t prev = frag_order[rank + 1];
            double curr = total_coverage(f) + total_score(f);
            if (curr == prev_score_sum) better_cluster_count[f] = better_cluster_count[prev];
            else { better_cluster_count[f] = rank + 1; prev_score_sum = curr; }
        }
    }

    if (g_stage_trace) for (int r : {0, 1}) { auto& t = g_stage_trace->reads[r]; t.minimizers = minimizers_by_read[r]; t.seeds = seeds_by_read[r]; t.clusters = all_clusters[r]; }

    std::array<std::vector<bool>, 2> minimizer_explored_by_read;
    std::vector<std::array<std::vector<Alignment>, 2>> alignments(max_fragment_num + 2);
    std::array<int, 2> best_alignment_scores{0, 0};

    for (size_t read_num = 0; read_num < 2; read_num++) {
        const std::string& sequence = seqs[read_num];
        std::vector<Cluster>& clusters = all_clusters[read_num];
        const std::vector<Minimizer>& minimizers = minimizers_by_read[read_num];
        const std::vector<Seed>& seeds = seeds_by_read[read_num];
        double cluster_score_cutoff = 0.0, cluster_coverage_cutoff = 0.0, second_best_cluster_score = 0.0;
        std::pair<double, double> best_cluster_coverage_score(0.0, 0.0);
        for (auto& c : clusters) {
            if (c.coverage > best_cluster_coverage_score.first) { best_cluster_coverage_score.first = c.coverage; best_cluster_coverage_score.second = c.score; }
            else if (c.coverage == best_cluster_coverage_score.first) best_cluster_coverage_score.second = std::max(best_cluster_coverage_score.second, c.score);
            cluster_coverage_cutoff = std::max(cluster_coverage_cutoff, c.coverage);
            if (c.score > cluster_score_cutoff) { second_best_cluster_score = cluster_score_cutoff; cluster_score_cutoff = c.score; }
            else if (c.score > second_best_cluster_score) second_best_cluster_score = c.score;
        }
        cluster_score_cutoff -= P.cluster_score_threshold;
        cluster_coverage_cutoff -= P.cluster_coverage_threshold;
        if (cluster_score_cutoff - P.pad_cluster_score_threshold < second_best_cluster_score)
            cluster_score_cutoff = std::min(cluster_score_cutoff, second_best_cluster_score);

        std::vector<std::pair<std::vector<GaplessExtension>, size_t>> cluster_extensions;
        std::vector<std::vector<size_t>> minimizer_kept_cluster_count;
        minimizer_explored_by_read[read_num].assign(minimizers.size(), false);
        size_t kept_cluster_count = 0;

        process_until_threshold_e<double>(clusters.size(),
            [&](size_t i) -> double { return clusters[i].coverage; },
            [&](size_t a, size_t b) -> bool {
                size_t fa = clusters[a].fragment, fb = clusters[b].fragment;
                double coverage_a = total_coverage(fa), coverage_b = total_coverage(fb);
                double score_a = total_score(fa), score_b = total_score(fb);
                if (fragment_cluster_has_pair[fa] != fragment_cluster_has_pair[fb]) return (bool)fragment_cluster_has_pair[fa];
                else if (coverage_a != coverage_b) return coverage_a > coverage_b;
                else if (score_a != score_b) return score_a > score_b;
                else if (clusters[a].coverage != clusters[b].coverage) return clusters[a].coverage > clusters[b].coverage;
                else return clusters[a].score > clusters[b].score;
            },
            [&](size_t) -> bool { return false; },
            0, P.min_extensions, P.max_extensions, rng,
            [&](size_t cluster_num, size_t, bool) -> bool {
                Cluster& cluster = clusters[cluster_num];
                if (!found_paired_cluster || fragment_cluster_has_pair[cluster.fragment] ||
                    (cluster.coverage == best_cluster_coverage_score.first && cluster.score == best_cluster_coverage_score.second)) {
                    if (P.cluster_coverage_threshold != 0 && cluster.coverage < cluster_coverage_cutoff && kept_cluster_count >= P.min_extensions) return false;
                    if (P.cluster_score_threshold != 0 && cluster.score < cluster_score_cutoff && kept_cluster_count >= P.min_extensions) return false;
                    minimizer_kept_cluster_count.emplace_back(minimizers.size(), 0);
                    std::vector<std::pair<uint32_t, int64_t>> seed_matchings;
                    for (size_t si : cluster.seeds) {
                        const Seed& seed = seeds[si];
                        seed_matchings.emplace_back(seed.node, (int64_t)minimizers[seed.source].offset - (int64_t)seed.offset);
                        minimizer_kept_cluster_count.back()[seed.source]++;
                    }
                    if (g_stage_trace) g_stage_trace->reads[read_num].items.push_back(StageTrace::Item{cluster_num, cluster.fragment, seed_matchings});
                    cluster_extensions.emplace_back(extend(g, scores, seed_matchings, sequence, P.max_extension_mismatches, 0.8, true), cluster.fragment);
                    if (counters) counters->extend_calls++;
                    kept_cluster_count++;
                    return true;
                }
                return false;
            },
            [&](size_t) {}, [&](size_t) {});

        std::vector<int> estimates(cluster_extensions.size(), 0);
        for (size_t i = 0; i < cluster_extensions.size(); i++)
            estimates[i] = score_extension_group(sequence.size(), cluster_extensions[i].first, 4, scores.gap_open, scores.gap_extend);

        // process_until_threshold_b (:1905): min 2, no score floor (unlike single-end)
        process_until_threshold_e<int>(estimates.size(),
            [&](size_t i) -> int { return estimates[i]; },
            [&](size_t a, size_t b) -> bool { return estimates[a] > estimates[b]; },
            [&](size_t) -> bool { return false; },
            P.extension_set_score_threshold, 2, P.max_alignments, rng,
            [&](size_t processed_num, size_t, bool) -> bool {
                auto& extensions = cluster_extensions[processed_num].first;
                std::vector<Alignment> best_alignments(1);
                if (!extensions.empty() && extensions.front().full() && extensions.front().mismatch_positions.size() <= 4) {
                    auto fill = [&](const GaplessExtension& e, Alignment& a) {
                        a.path = extension_to_path(g, e, sequence); a.score = e.score;
                        a.identity = sequence.empty() ? 0.0 : (sequence.length() - e.mismatch_positions.size()) / (double)sequence.length();
                    };
                    fill(extensions.front(), best_alignments.front());
                    for (auto it = extensions.begin() + 1; it != extensions.end() && it->full(); ++it) { best_alignments.emplace_back(); fill(*it, best_alignments.back()); }
                    if (counters) counters->direct++;
                } else if (P.do_dp) {
                    best_alignments.emplace_back();
                    find_optimal_tail_alignments(g, scores, P, sequence, extensions, rng, best_alignments[0], best_alignments[1], counters);
                }
                size_t fragment_num = cluster_extensions[processed_num].second;
                for (auto it = best_alignments.begin(); it != best_alignments.end() && it->score != 0 && it->score >= best_alignments[0].score * 0.8; ++it) {
                    best_alignment_scores[read_num] = std::max(best_alignment_scores[read_num], it->score);
                    alignments[fragment_num][read_num].emplace_back(std::move(*it));
                }
                for (size_t i = 0; i < minimizer_kept_cluster_count[processed_num].size(); i++)
                    if (minimizer_kept_cluster_count[processed_num][i] > 0) minimizer_explored_by_read[read_num][i] = true;
                return true;
            },
            [&](size_t) {}, [&](size_t) {});
    }

    // ---- pairing (:2046-2208) ------------------------------------------------------------------
    struct Idx { size_t fragment, index; };
    std::vector<std::array<Idx, 2>> paired_alignments;
    std::vector<double> paired_scores; std::vector<int64_t> fragment_distances; std::vector<size_t> better_cluster_count_by_pairs;
    bool found_pair = false;
    struct UIdx { size_t fragment, index; int read; };
    std::vector<UIdx> unpaired_alignments;
    auto score_alignment_pair = [&](const Alignment& a1, const Alignment& a2, int64_t fragment_distance) {
        double dev = fragment_distance - frag_mean;
        double ll = (-dev * dev / (2.0 * frag_sd * frag_sd)) / log_base;
        double score = a1.score + a2.score + ll;
        double worse = std::min(a1.score, a2.score);
        return std::max(score, worse);
    };
    enum PairType { PT_PAIRED, PT_UNPAIRED, PT_RESCUED_FROM_FIRST, PT_RESCUED_FROM_SECOND };
    std::vector<PairType> pair_types;
    std::array<size_t, 2> unpaired_count{0, 0}, rescued_count{0, 0};
    // distance_between(aln1, aln2): initial_position(aln1) -> final_position(aln2) (:3895-3903)
    auto distance_between = [&](const Alignment& a0, const Alignment& a1) -> int64_t {
        const Mapping& last = a1.path.back();
        uint32_t used = 0; for (const Edit& e : last.edits) used += e.from_length;
        return oriented_distance(ix, a0.path.front().node, a0.path.front().offset, last.node, last.offset + used);
    };
    const size_t n_cluster_fragments = alignments.size() - 1;        // the last entry collects rescued alignments
    for (size_t f = 0; f < n_cluster_fragments; f++) {
        auto& fa = alignments[f];
        if (!fa[0].empty() && !fa[1].empty()) {
            found_pair = true;
            for (size_t i0 = 0; i0 < fa[0].size(); i0++) for (size_t i1 = 0; i1 < fa[1].size(); i1++) {
                const Alignment& a0 = fa[0][i0]; const Alignment& a1 = fa[1][i1];
                int64_t dist = distance_between(a0, a1);
                paired_alignments.push_back({Idx{f, i0}, Idx{f, i1}});
                paired_scores.push_back(score_alignment_pair(a0, a1, dist));
                fragment_distances.push_back(dist);
                better_cluster_count_by_pairs.push_back(better_cluster_count[f]);
                pair_types.push_back(PT_PAIRED);
            }
        } else {
            for (int r : {0, 1}) for (size_t i = 0; i < fa[r].size(); i++) { unpaired_alignments.push_back(UIdx{f, i, r}); unpaired_count[r]++; }
        }
    }
    auto finish_read2 = [&](Alignment& a) { reverse_complement_path(a.path, g); };
    std::array<std::vector<double>, 2> unpaired_scores;

    if (!unpaired_alignments.empty()) {
        if (!found_pair) {
            std::array<int64_t, 2> best_index{-1, -1}; std::array<int32_t, 2> best_score{0, 0};
            for (size_t u = 0; u < unpaired_alignments.size(); u++) {
                const UIdx& index = unpaired_alignments[u];
                const Alignment& alignment = alignments[index.fragment][index.read][index.index];
                unpaired_scores[index.read].push_back(alignment.score);
                if (deterministic_beats(alignment.score, best_score[index.read], rng)) { best_index[index.read] = (int64_t)u; best_score[index.read] = alignment.score; }
            }
            if (P.max_rescue_attempts == 0) {
                // best alignment of each end, MAPQ 1 (:2238-2287)
                for (int r : {0, 1}) {
                    if (best_index[r] >= 0) { const UIdx& index = unpaired_alignments[(size_t)best_index[r]]; result.aln[r] = alignments[index.fragment][index.read][index.index]; }
                    else result.aln[r] = Alignment();
                }
                finish_read2(result.aln[1]);
                for (int r : {0, 1}) result.aln[r].mapq = 1;
                return result;
            } else if (best_score[0] != 0 && best_score[1] != 0) {
                // keep the best alignments as a potential (unpaired) pair, distance "infinite" (:2288-2334)
                const UIdx& u0 = unpaired_alignments[(size_t)best_index[0]]; const UIdx& u1 = unpaired_alignments[(size_t)best_index[1]];
                paired_alignments.push_back({Idx{u0.fragment, u0.index}, Idx{u1.fragment, u1.index}});
                paired_scores.push_back(score_alignment_pair(alignments[u0.fragment][0][u0.index], alignments[u1.fragment][1][u1.index], std::numeric_limits<int64_t>::max()));
                fragment_distances.push_back(std::numeric_limits<int64_t>::max());
                better_cluster_count_by_pairs.push_back(0);
                pair_types.push_back(PT_UNPAIRED);
            }
        }
        if (P.max_rescue_attempts != 0) {
            // rescue from the best unpaired alignments (:2338-2457)
            process_until_threshold_e<double>(unpaired_alignments.size(),
                [&](size_t i) -> double { const UIdx& u = unpaired_alignments[i]; return (double)alignments[u.fragment][u.read][u.index].score; },
                [&](size_t a, size_t b) -> bool {
                    const UIdx& ua = unpaired_alignments[a]; const UIdx& ub = unpaired_alignments[b];
                    return alignments[ua.fragment][ua.read][ua.index].score > alignments[ub.fragment][ub.read][ub.index].score;
                },
                [&](size_t) -> bool { return false; },
                0, 1, P.max_rescue_attempts, rng,
                [&](size_t i, size_t, bool) -> bool {
                    const UIdx index = unpaired_alignments[i];
                    const Alignment mapped_aln = alignments[index.fragment][index.read][index.index];
                    if (found_pair && (double)mapped_aln.score < (double)best_alignment_scores[index.read] * P.paired_rescue_score_limit) return true;
                    const int other = 1 - index.read;
                    Alignment rescued_aln = attempt_rescue(ix, scores, P, mapped_aln, seqs[other], minimizers_by_read[other], index.read == 0, counters);
                    rescued_aln.rescued = true;
                    int64_t fragment_dist; double score;
                    if (!rescued_aln.path.empty()) {
                        fragment_dist = index.read == 0 ? distance_between(mapped_aln, rescued_aln) : distance_between(rescued_aln, mapped_aln);
                        score = score_alignment_pair(mapped_aln, rescued_aln, fragment_dist);
                    } else { score = mapped_aln.score; fragment_dist = std::numeric_limits<int64_t>::max(); }
                    const size_t rf = alignments.size() - 1;
                    std::array<Idx, 2> index_pair;
                    index_pair[index.read] = Idx{index.fragment, index.index};
                    index_pair[other] = Idx{rf, alignments[rf][other].size()};
                    alignments[rf][other].push_back(std::move(rescued_aln));
                    rescued_count[index.read]++;
                    paired_alignments.push_back(index_pair);
                    fragment_distances.push_back(fragment_dist);
                    paired_scores.push_back(score);
                    pair_types.push_back(index.read == 0 ? PT_RESCUED_FROM_FIRST : PT_RESCUED_FROM_SECOND);
                    better_cluster_count_by_pairs.push_back(better_cluster_count[index.fragment]);
                    return true;
                },
                [&](size_t) {}, [&](size_t) {});
        }
    }

    // ---- winner (:2505-2598) and MAPQ (:2606-2777) --------------------------------------------------
    std::vector<double> out_scores; std::vector<int64_t> distances; std::vector<size_t> better_cluster_count_by_mappings; std::vector<PairType> types;
    std::array<std::vector<Alignment>, 2> mappings;
    process_until_threshold_e<double>(paired_alignments.size(),
        [&](size_t i) -> double { return paired_scores[i]; },
        [&](size_t a, size_t b) -> bool { return paired_scores[a] > paired_scores[b]; },
        [&](size_t) -> bool { return false; },
        0, 1, P.max_multimaps, rng,
        [&](size_t n, size_t, bool) -> bool {
            out_scores.push_back(paired_scores[n]); distances.push_back(fragment_distances[n]); types.push_back(pair_types[n]);
            better_cluster_count_by_mappings.push_back(better_cluster_count_by_pairs[n]);
            for (int r : {0, 1}) mappings[r].push_back(alignments[paired_alignments[n][r].fragment][r][paired_alignments[n][r].index]);
            finish_read2(mappings[1].back());
            return true;
        },
        [&](size_t n) { out_scores.push_back(paired_scores[n]); distances.push_back(fragment_distances[n]); types.push_back(pair_types[n]); better_cluster_count_by_mappings.push_back(better_cluster_count_by_pairs[n]); },
        [&](size_t) {});

    if (mappings[0].empty()) { result.aln[0] = Alignment(); result.aln[1] = Alignment(); return result; }

    // multiplicities when every pair was found by rescue (:2661-2690)
    std::array<double, 2> estimated_multiplicity_from;
    for (int r : {0, 1}) estimated_multiplicity_from[r] = unpaired_count[r] > 0 ? (double)unpaired_count[r] / (double)std::min<size_t>(rescued_count[r], P.max_rescue_attempts) : 1.0;
    bool all_rescued = true; std::vector<double> paired_multiplicities;
    for (PairType t : types) {
        switch (t) {
        case PT_PAIRED: paired_multiplicities.push_back(1.0); all_rescued = false; break;
        case PT_UNPAIRED: paired_multiplicities.push_back(1.0); break;
        case PT_RESCUED_FROM_FIRST: paired_multiplicities.push_back(estimated_multiplicity_from[0]); break;
        case PT_RESCUED_FROM_SECOND: paired_multiplicities.push_back(estimated_multiplicity_from[1]); break;
        }
    }
    const std::vector<double>* multiplicities = all_rescued ? &paired_multiplicities : nullptr;
    double uncapped_mapq = out_scores[0] == 0 ? 0 : compute_max_mapping_quality(out_scores, log_base, multiplicities);
    double fragment_cluster_cap = std::numeric_limits<float>::infinity();
    if (better_cluster_count_by_mappings.front() > 1)
        fragment_cluster_cap = -10.0 * std::log10(1.0 - (1.0 / (double)better_cluster_count_by_mappings.front()));   // prob_to_phred
    std::array<double, 2> mapq_explored_caps;
    for (int r : {0, 1}) {
        std::vector<size_t> explored;
        for (size_t i = 0; i < minimizers_by_read[r].size(); i++) if (minimizer_explored_by_read[r][i]) explored.push_back(i);
        mapq_explored_caps[r] = faster_cap(minimizers_by_read[r], explored, seqs[r], quals[r]);
    }
    for (int r : {0, 1}) {
        double escape_bonus = uncapped_mapq < std::numeric_limits<int32_t>::max() ? 1.0 : 2.0;
        double mapq_cap = std::min(fragment_cluster_cap, ((mapq_explored_caps[0] + mapq_explored_caps[1]) * escape_bonus));
        if (types.front() == PT_UNPAIRED) mapq_cap = std::min(mapq_cap, (double)compute_max_mapping_quality(unpaired_scores[r], log_base));   // :2735-2739
        double read_mapq = uncapped_mapq;
        double capped_mapq = std::min(mapq_cap, read_mapq);
        if (distances.front() == std::numeric_limits<int64_t>::max()) capped_mapq = capped_mapq / 2.0;
        read_mapq = std::max(std::min(capped_mapq, 120.0) / 2.0, 0.0);
        Alignment& out = mappings[r].front();
        if (out.path.empty()) read_mapq = 0;
        out.mapq = (double)(int32_t)read_mapq;          // Alignment.mapping_quality is int32
        out.mapq_uncapped = uncapped_mapq; out.mapq_explored_cap = mapq_cap;
        result.aln[r] = out;
        // pairs 1 .. max_multimaps - 1 are secondary for both reads (:2552-2557)
        for (size_t i = 1; i < mappings[r].size(); i++) { mappings[r][i].secondary = true; result.aln[r].secondaries.push_back(mappings[r][i]); }
    }
    return result;
}

} // namespace oracle

// reads are interleaved: read 2i = mate 1, read 2i+1 = mate 2 (input orientation, i.e. inward)
extern "C" int oracle_map_paired_batch(const gb_flat_index* ix, const gb_scores* scores, const gb_map_params* p,
                                       uint32_t n_reads, const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off,
                                       gb_alignment* aln, gb_mapping* mappings, uint32_t* edits, uint8_t* status,
                                       int n_threads, uint64_t* counters_out) {
    if (n_reads % 2 != 0) return -2;
    oracle::MapCounters total;
    int failed = 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        oracle::MapCounters local;
#pragma omp for schedule(dynamic, 128)
        for (int64_t pi = 0; pi < (int64_t)n_reads / 2; pi++) {
            std::string s[2], q[2];
            for (int r = 0; r < 2; r++) {
                const uint64_t b = read_off[2 * pi + r], e = read_off[2 * pi + r + 1];
                s[r].assign((const char*)reads + b, (size_t)(e - b));
                if (quals) q[r].assign((const char*)quals + b, (size_t)(e - b));
            }
            oracle::PairResult res = oracle::map_paired(ix, *scores, *p, s[0], q[0], s[1], q[1], &local);
            for (int r = 0; r < 2; r++) {
                const int64_t ri = 2 * pi + r;
                int rc = oracle::pack_alignment(res.aln[r], (uint32_t)ri, aln + ri, mappings + (size_t)ri * p->mapping_cap_per_read,
                                                p->mapping_cap_per_read, edits + (size_t)ri * p->edit_cap_per_read, p->edit_cap_per_read,
                                                (uint32_t)(ri * p->mapping_cap_per_read), (uint32_t)(ri * p->edit_cap_per_read));
                aln[ri].flags |= GB_ALN_PAIRED;
                rc |= oracle::pack_secondaries(res.aln[r], *p, n_reads, (uint32_t)ri, GB_ALN_PAIRED, aln, mappings, edits);
                status[ri] = rc == 0 ? GB_ITEM_OK : GB_ITEM_OUT_FULL;
                if (rc) {
#pragma omp atomic
                    failed++;
                }
            }
        }
#pragma omp critical
        total.add(local);
    }
    if (counters_out) total.store(counters_out);
    return failed ? -1 : 0;
}


// ---- the whole paired job: fragment-length training, then map_paired, then the ambiguous buffer -----------------
// giraffe_main.cpp:2246-2400 drives MinimizerMapper::map_paired(aln1, aln2, ambiguous_pair_buffer)
// (minimizer_mapper.cpp:1303-1395) single-threaded until the distribution is finalized, maps the rest with the
// finalized distribution, finalizes by force at the end of input (finalize_fragment_length_distr,
// minimizer_mapper.hpp:539-543) and maps the buffered pairs last.  route[pair]: GB_PAIR_TRAINING (both ends mapped
// single-ended, distance registered), GB_PAIR_PAIRED, GB_PAIR_BUFFERED (ambiguous during training, mapped paired
// at the end).  frag_out = {mean, stdev, samples registered}.
extern "C" int oracle_map_paired_job(const gb_flat_index* ix, const gb_scores* scores, const gb_map_params* p_in,
                                     uint64_t maximum_sample_size, uint64_t reestimation_frequency, double robust_estimation_fraction,
                                     uint32_t n_reads, const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off,
                                     gb_alignment* aln, gb_mapping* mappings, uint32_t* edits, uint8_t* status, uint8_t* route,
                                     int n_threads, double* frag_out) {
    if (n_reads % 2 != 0 || p_in->max_multimaps != 1) return -2;
    using namespace oracle;
    gb_map_params P = *p_in;
    FragmentLengthDistribution distr(maximum_sample_size, reestimation_frequency, robust_estimation_fraction);
    if (P.fragment_stdev > 0) distr.force_parameters(P.fragment_mean, P.fragment_stdev);
    Graph g(ix);
    MapCounters local;
    const int64_t n_pairs = n_reads / 2;
    int failed = 0;
    auto read_of = [&](int64_t ri, std::string& s, std::string& q) {
        const uint64_t b = read_off[ri], e = read_off[ri + 1];
        s.assign((const char*)reads + b, (size_t)(e - b));
        if (quals) q.assign((const char*)quals + b, (size_t)(e - b)); else q.clear();
    };
    auto emit = [&](const Alignment& a, int64_t ri, bool paired) {
        int rc = pack_alignment(a, (uint32_t)ri, aln + ri, mappings + (size_t)ri * P.mapping_cap_per_read, P.mapping_cap_per_read,
                                edits + (size_t)ri * P.edit_cap_per_read, P.edit_cap_per_read,
                                (uint32_t)(ri * P.mapping_cap_per_read), (uint32_t)(ri * P.edit_cap_per_read));
        if (paired) aln[ri].flags |= GB_ALN_PAIRED;
        status[ri] = rc == 0 ? GB_ITEM_OK : GB_ITEM_OUT_FULL;
        if (rc) {
#pragma omp atomic
            failed++;
        }
    };
    std::vector<int64_t> ambiguous_pair_buffer;
    int64_t pi = 0;
    for (; pi < n_pairs && !distr.is_finalized(); pi++) {
        std::string s[2], q[2];
        Alignment single[2];
        bool both_perfect_unique = true;
        for (int r = 0; r < 2; r++) {
            read_of(2 * pi + r, s[r], q[r]);
            single[r] = map_from_extensions(ix, *scores, P, s[r], q[r], &local);
            const int32_t max_score_aln = scores->match * (int32_t)s[r].size();                    // score_exact_match
            both_perfect_unique = both_perfect_unique && !single[r].path.empty() && single[r].mapq == 60 && single[r].score >= max_score_aln * 0.85;
        }
        bool keep = false;
        if (both_perfect_unique) {
            Alignment flipped = single[1];
            reverse_complement_path(flipped.path, g);
            const Mapping& last = flipped.path.back();
            uint32_t used = 0; for (const Edit& e : last.edits) used += e.from_length;
            const int64_t dist = oriented_distance(ix, single[0].path.front().node, single[0].path.front().offset, last.node, last.offset + used);
            if (!(dist == std::numeric_limits<int64_t>::max() || dist >= (int64_t)P.max_fragment_length)) {
                distr.register_fragment_length(dist);
                keep = true;
            }
        }
        if (keep) { route[pi] = GB_PAIR_TRAINING; for (int r = 0; r < 2; r++) emit(single[r], 2 * pi + r, true); }   // pair_all(mapped_pair), minimizer_mapper.cpp:1345-1350
        else { route[pi] = GB_PAIR_BUFFERED; ambiguous_pair_buffer.push_back(pi); }
    }
    const int64_t first_paired = pi;
    if (!distr.is_finalized()) distr.force_parameters(distr.mean(), distr.std_dev());
    P.fragment_mean = distr.mean(); P.fragment_stdev = distr.std_dev();
    frag_out[0] = distr.mean(); frag_out[1] = distr.std_dev(); frag_out[2] = (double)distr.curr_sample_size();
    std::vector<int64_t> todo;
    for (int64_t x = first_paired; x < n_pairs; x++) { route[x] = GB_PAIR_PAIRED; todo.push_back(x); }
    todo.insert(todo.end(), ambiguous_pair_buffer.begin(), ambiguous_pair_buffer.end());
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t t = 0; t < (int64_t)todo.size(); t++) {
        const int64_t x = todo[t];
        std::string s[2], q[2];
        for (int r = 0; r < 2; r++) read_of(2 * x + r, s[r], q[r]);
        MapCounters mine;
        PairResult res = map_paired(ix, *scores, P, s[0], q[0], s[1], q[1], &mine);
        for (int r = 0; r < 2; r++) emit(res.aln[r], 2 * x + r, true);
    }
    return failed ? -1 : 0;
}


// ---- stage dump for the stage-level parity tests: everything the mapper computes before the first extension call,
// per read, in the layout of gb_debug_seed_stage (include/giraffe_b200.h).  Sequential (the trace is thread-local).
extern "C" int oracle_seed_stage(const gb_flat_index* ix, const gb_scores* scores, const gb_map_params* p, int paired,
                                 uint32_t n_reads, const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off,
                                 gb_stage_read* out_reads, gb_stage_minimizer* mins, uint64_t min_cap, gb_stage_seed* seeds, uint64_t seed_cap,
                                 gb_stage_cluster* clusters, uint64_t cluster_cap, gb_stage_item* items, uint64_t item_cap,
                                 gb_seed* item_seeds, uint64_t item_seed_cap) {
    using namespace oracle;
    uint64_t nm = 0, ns = 0, nc = 0, ni = 0, ne = 0;
    auto read_of = [&](uint64_t ri, std::string& s, std::string& q) {
        const uint64_t b = read_off[ri], e = read_off[ri + 1];
        s.assign((const char*)reads + b, (size_t)(e - b));
        if (quals) q.assign((const char*)quals + b, (size_t)(e - b)); else q.clear();
    };
    const uint32_t step = paired ? 2 : 1;
    for (uint32_t u = 0; u + step <= n_reads; u += step) {
        StageTrace trace;
        g_stage_trace = &trace;
        std::string s[2], q[2];
        read_of(u, s[0], q[0]);
        if (paired) { read_of(u + 1, s[1], q[1]); (void)map_paired(ix, *scores, *p, s[0], q[0], s[1], q[1], nullptr); }
        else (void)map_from_extensions(ix, *scores, *p, s[0], q[0], nullptr);
        g_stage_trace = nullptr;
        for (uint32_t r = 0; r < step; r++) {
            const StageTrace::Read& t = trace.reads[r];
            gb_stage_read& o = out_reads[u + r];
            memset(&o, 0, sizeof o);
            if (nm + t.minimizers.size() > min_cap || ns + t.seeds.size() > seed_cap || nc + t.clusters.size() > cluster_cap || ni + t.items.size() > item_cap) return -1;
            o.min_off = (uint32_t)nm; o.seed_off = (uint32_t)ns; o.cluster_off = (uint32_t)nc; o.item_off = (uint32_t)ni;
            o.min_cnt = (uint32_t)t.minimizers.size(); o.seed_cnt = (uint32_t)t.seeds.size(); o.cluster_cnt = (uint32_t)t.clusters.size(); o.item_cnt = (uint32_t)t.items.size();
            for (const Minimizer& m : t.minimizers)
                mins[nm++] = gb_stage_minimizer{m.hash, m.score, (uint32_t)m.forward_offset(), (uint32_t)m.agglomeration_start, (uint32_t)m.agglomeration_length, m.is_reverse ? 1u : 0u, m.hit_cnt, 0};
            std::vector<uint32_t> cluster_of(t.seeds.size(), 0xffffffffu);
            for (size_t c = 0; c < t.clusters.size(); c++) for (size_t si : t.clusters[c].seeds) cluster_of[si] = (uint32_t)c;
            for (size_t i = 0; i < t.seeds.size(); i++) seeds[ns++] = gb_stage_seed{t.seeds[i].node, t.seeds[i].offset, (uint32_t)t.seeds[i].source, cluster_of[i]};
            for (size_t c = 0; c < t.clusters.size(); c++) {
                uint32_t rank = 0xffffffffu;
                for (size_t x = 0; x < t.items.size(); x++) if (t.items[x].cluster == c) rank = (uint32_t)x;
                clusters[nc++] = gb_stage_cluster{t.clusters[c].score, t.clusters[c].coverage, (uint32_t)t.clusters[c].seeds.front(), (uint32_t)t.clusters[c].seeds.size(), (uint32_t)t.clusters[c].fragment, rank};
            }
            for (const StageTrace::Item& it : t.items) {
                if (ne + it.seeds.size() > item_seed_cap) return -1;
                items[ni++] = gb_stage_item{(uint32_t)it.cluster, (uint32_t)it.fragment, (uint32_t)ne, (uint32_t)it.seeds.size()};
                for (const auto& sd : it.seeds) item_seeds[ne++] = gb_seed{sd.first, (int32_t)sd.second};
            }
        }
    }
    return 0;
}

// The candidate side of find_best_chains (test entry for gb_chain_candidates_batch): every ordered pair of seeds (from, to)
// whose minimum graph distance from -> to exists and is <= limit, sorted by (to, from) — what zip_tree_transition_iterator
// offers (chain_items.cpp:116-260), with the distances taken from the distance payload instead of the zip-code tree.
// Returns the number of candidates (written up to cap).
extern "C" uint64_t oracle_chain_candidates(const gb_flat_index* ix, uint32_t n_seeds, const uint32_t* seed_pos, uint64_t limit,
                                            gb_chain_candidate* out, uint64_t cap) {
    const int64_t UNREACHABLE = (int64_t)std::numeric_limits<size_t>::max();
    uint64_t n = 0;
    for (uint32_t j = 0; j < n_seeds; j++)
        for (uint32_t i = 0; i < n_seeds; i++) {
            if (i == j) continue;
            const int64_t d = oracle::oriented_distance(ix, seed_pos[2 * i], seed_pos[2 * i + 1], seed_pos[2 * j], seed_pos[2 * j + 1]);
            if (d == UNREACHABLE || d < 0 || (uint64_t)d > limit) continue;
            if (n < cap) { out[n].from = i; out[n].to = j; out[n].graph_distance = (uint64_t)d; }
            n++;
        }
    return n;
}
