// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of MinimizerMapper::map_from_extensions
// (src/minimizer_mapper.cpp:608-1284 @ fd49b9a9) and the functions it calls:
//   find_minimizers :3918-3974, sort_minimizers_by_score :4074-4107, find_seeds :4109-4517,
//   cluster_seeds (definition: snarl_seed_clusterer.cpp:256-315), score_cluster :4738-4780,
//   extend_seed_group :4784-5018, score_extension_group :5022-5203,
//   extension_to_alignment :3905-3914, faster_cap :2946-3260,
//   MappingQualityCalculator::compute_max_mapping_quality mapping_quality_calculator.cpp:26-67,355-364.
//
// Third-party pieces restated from their published algorithms (absent under deps/):
//   gbwtgraph minimizer_regions / find (see vg_b200/csrc/minimizer_common.h for the
//   definition; this file restates it independently by brute force over windows);
//   libbdsg minimum_distance via the per-node distance payload of include/giraffe_b200.h.
//
// Canonicalisations (the reference leaves these to std::sort / hash order):
//   * sort_permutation in sort_minimizers_by_score and the sort in faster_cap are stable.
#include "mapper_common.hpp"
#include "tail_align.hpp"

#include <algorithm>
#include <cassert>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <set>

namespace oracle {

// ---------------------------------------------------------------------------------------
// scoring helpers
// ---------------------------------------------------------------------------------------

// AlignmentScorer::recover_log_base, alignment_scorer.cpp:30-99 (gc_content 0.5, tol 1e-12)
double recover_log_base(const gb_scores& s) {
    double matrix[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) matrix[i * 4 + j] = (i == j) ? s.match : -s.mismatch;
    const double nt[4] = {0.25, 0.25, 0.25, 0.25};
    auto partition = [&](double lambda) {
        double p = 0;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) p += nt[i] * nt[j] * std::exp(lambda * matrix[i * 4 + j]);
        return p;
    };
    const double tol = 1e-12;
    double lower, upper, lambda = 1.0;
    double part = partition(lambda);
    if (part < 1.0) {
        lower = lambda;
        while (part <= 1.0) { lower = lambda; lambda *= 2.0; part = partition(lambda); }
        upper = lambda;
    } else {
        upper = lambda;
        while (part >= 1.0) { upper = lambda; lambda /= 2.0; part = partition(lambda); }
        lower = lambda;
    }
    while (upper / lower - 1.0 > tol) {
        lambda = 0.5 * (lower + upper);
        if (partition(lambda) < 1.0) lower = lambda; else upper = lambda;
    }
    return 0.5 * (lower + upper);
}

static inline double add_log(double x, double y) { return x > y ? x + log1p(exp(y - x)) : y + log1p(exp(x - y)); }
static inline double subtract_log(double x, double y) { return x + log1p(-exp(y - x)); }

// mapping_quality_calculator.cpp:26-67 (no multiplicities, max_idx_out given)
double maximum_mapping_quality_exact(const std::vector<double>& scaled) {
    const double quality_scale_factor = 10.0 / std::log(10.0);
    double log_sum_exp = std::numeric_limits<double>::lowest();
    double to_score = std::numeric_limits<double>::lowest();
    for (int64_t i = (int64_t)scaled.size() - 1; i >= 0; --i) {
        double score = scaled[i];
        if (score >= to_score) to_score = score;
        log_sum_exp = add_log(log_sum_exp, score);
    }
    if (scaled.size() == 1) log_sum_exp = add_log(log_sum_exp, 0.0);
    double direct = -quality_scale_factor * subtract_log(0.0, to_score - log_sum_exp);
    return std::isinf(direct) ? (double)std::numeric_limits<int32_t>::max() : direct;
}

// maximum_mapping_quality_exact with multiplicities (:26-67), used when every pair was found by rescue
double maximum_mapping_quality_exact(const std::vector<double>& scaled, const std::vector<double>* multiplicities) {
    const double quality_scale_factor = 10.0 / std::log(10.0);
    double log_sum_exp = std::numeric_limits<double>::lowest();
    double to_score = std::numeric_limits<double>::lowest();
    for (int64_t i = (int64_t)scaled.size() - 1; i >= 0; --i) {
        double score = scaled[i];
        if (score >= to_score) to_score = score;
        if (multiplicities && (*multiplicities)[i] > 1.0) score += std::log((*multiplicities)[i]);
        log_sum_exp = add_log(log_sum_exp, score);
    }
    if (scaled.size() == 1) {
        if (multiplicities && (*multiplicities)[0] <= 1.0) log_sum_exp = add_log(log_sum_exp, 0.0);
        else if (!multiplicities) log_sum_exp = add_log(log_sum_exp, 0.0);
    }
    double direct = -quality_scale_factor * subtract_log(0.0, to_score - log_sum_exp);
    return std::isinf(direct) ? (double)std::numeric_limits<int32_t>::max() : direct;
}
int32_t compute_max_mapping_quality(const std::vector<double>& scores, double log_base, const std::vector<double>* multiplicities) {
    std::vector<double> scaled(scores.size());
    for (size_t i = 0; i < scores.size(); i++) scaled[i] = log_base * scores[i];
    return (int32_t)maximum_mapping_quality_exact(scaled, multiplicities);
}

// compute_max_mapping_quality :355-364 (returns int32_t)
int32_t compute_max_mapping_quality(const std::vector<double>& scores, double log_base) {
    std::vector<double> scaled(scores.size());
    for (size_t i = 0; i < scores.size(); i++) scaled[i] = log_base * scores[i];
    return (int32_t)maximum_mapping_quality_exact(scaled);
}

// statistics.cpp:525-560 with MAX_AT_LEAST_ONE_EVENTS 32, AT_LEAST_ONE_PRECISION 8
static double prob_for_at_least_one(uint64_t p, size_t n) {
    static std::vector<double> table = [] {
        std::vector<double> t((32 + 1) * 256, 0.0);
        for (size_t nn = 1; nn <= 32; nn++)
            for (size_t pp = 0; pp < 256; pp++) {
                double probability = (2 * pp + 1) / (2.0 * 256);
                t[(nn << 8) + pp] = 1.0 - std::pow(1.0 - probability, nn);
            }
        return t;
    }();
    assert(n <= 32);
    p >>= 64 - 8;
    return table[(n << 8) + p];
}
static double phred_to_prob(uint8_t phred) { return std::pow(10, -((double)phred) / 10); }

// ---------------------------------------------------------------------------------------
// minimizers
// ---------------------------------------------------------------------------------------
static inline uint64_t wang_hash_64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

// minimizer_regions: brute force over windows (independent of the product's deque version)
std::vector<Minimizer> minimizer_regions(const std::string& seq, uint32_t k, uint32_t w) {
    std::vector<Minimizer> out;
    const size_t L = seq.size(), window_bp = (size_t)k + w - 1;
    if (L < window_bp) return out;
    const size_t nk = L - k + 1;
    struct Cand { bool valid; uint64_t key, hash; bool rev; };
    std::vector<Cand> cand(nk);
    auto code = [](char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } };
    for (size_t s = 0; s < nk; s++) {
        uint64_t f = 0, r = 0; bool ok = true;
        for (uint32_t i = 0; i < k; i++) {
            int c = code(seq[s + i]);
            if (c > 3) { ok = false; break; }
            f = (f << 2) | (uint64_t)c;
            r |= (uint64_t)(3 - c) << (2 * i);
        }
        cand[s].valid = ok;
        if (!ok) continue;
        uint64_t hf = wang_hash_64(f), hr = wang_hash_64(r);
        if (hr < hf) cand[s] = {true, r, hr, true}; else cand[s] = {true, f, hf, false};
    }
    std::vector<long> first(nk, -1), last(nk, -1);
    for (size_t ws = 0; ws + window_bp <= L; ws++) {
        bool any = false; uint64_t best = 0;
        for (size_t s = ws; s < ws + w; s++) if (cand[s].valid && (!any || cand[s].hash < best)) { any = true; best = cand[s].hash; }
        if (!any) continue;
        for (size_t s = ws; s < ws + w; s++) if (cand[s].valid && cand[s].hash == best) { if (first[s] < 0) first[s] = (long)ws; last[s] = (long)ws; }
    }
    for (size_t s = 0; s < nk; s++) {
        if (first[s] < 0) continue;
        Minimizer m;
        m.key = cand[s].key; m.hash = cand[s].hash; m.is_reverse = cand[s].rev;
        m.offset = cand[s].rev ? (uint32_t)(s + k - 1) : (uint32_t)s;
        m.agglomeration_start = (size_t)first[s];
        m.agglomeration_length = (size_t)(last[s] - first[s]) + window_bp;
        m.length = (int32_t)k; m.candidates_per_window = (int32_t)w;
        out.push_back(m);
    }
    return out;
}

static void index_find(const gb_flat_index* ix, uint64_t key, uint32_t& off, uint32_t& cnt) {
    off = 0; cnt = 0;
    uint64_t mask = ix->table_cells - 1;
    uint64_t h = wang_hash_64(key) & mask;
    while (ix->table[h].key != GB_NO_KEY) {
        if (ix->table[h].key == key) { off = ix->table[h].hit_off; cnt = ix->table[h].hit_cnt; return; }
        h = (h + 1) & mask;
    }
}

// find_minimizers, minimizer_mapper.cpp:3918-3974
std::vector<Minimizer> find_minimizers(const gb_flat_index* ix, const gb_map_params& P, const std::string& sequence) {
    std::vector<Minimizer> result = minimizer_regions(sequence, ix->k, ix->w);
    double base_score = 1.0 + std::log((double)P.hard_hit_cap);
    for (auto& m : result) {
        index_find(ix, m.key, m.hit_off, m.hit_cnt);
        double score = 0.0;
        if (m.hit_cnt > 0) {
            if (m.hit_cnt <= P.hard_hit_cap) score = base_score - std::log((double)m.hit_cnt);
            else score = 1.0;
        }
        m.score = score;
    }
    std::stable_sort(result.begin(), result.end(), [](const Minimizer& a, const Minimizer& b) { return a.forward_offset() < b.forward_offset(); });
    return result;
}

// sort_minimizers_by_score, minimizer_mapper.cpp:4074-4107
std::vector<size_t> sort_minimizers_by_score(const std::vector<Minimizer>& minimizers, LazyRNG& rng) {
    std::vector<size_t> order(minimizers.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return minimizers[a] < minimizers[b]; });
    std::vector<size_t> run_sort_order;
    for (size_t i = 0; i < order.size(); i++)
        if (i == 0 || minimizers[order[i - 1]].key != minimizers[order[i]].key) run_sort_order.push_back(i);
    sort_shuffling_ties(run_sort_order.begin(), run_sort_order.end(), [&](const size_t& a, const size_t& b) {
        return minimizers[order[a]].score > minimizers[order[b]].score;
    }, rng);
    std::vector<size_t> by_key;
    by_key.reserve(minimizers.size());
    for (size_t i : run_sort_order) {
        uint64_t key = minimizers[order[i]].key;
        size_t j = i;
        while (j < order.size() && minimizers[order[j]].key == key) { by_key.push_back(order[j]); j++; }
    }
    return by_key;
}

// find_seeds, minimizer_mapper.cpp:4109-4517 (defaults: downsampling off, exclude_overlapping_min off)
std::vector<Seed> find_seeds(const gb_flat_index* ix, const gb_map_params& P,
                             const std::vector<Minimizer>& minimizers /* score order */, size_t read_len) {
    double base_target_score = 0.0, target_score = 0.0, selected_score = 0.0;
    if (P.hit_cap != 0 || P.minimizer_score_fraction != 1.0) {
        for (const Minimizer& m : minimizers) base_target_score += m.score;
        target_score = (base_target_score * P.minimizer_score_fraction) + 0.000001;
    }
    size_t start = 0, limit = 0, run_hits = 0;
    bool taking_run = false;
    size_t num_minimizers = 0;
    size_t num_min_by_read_len = read_len / P.num_bp_per_min;
    std::vector<bool> read_coverage(read_len, false);
    size_t worst_kept_hits = 0;
    std::vector<Seed> seeds;
    (void)start;

    for (size_t i = 0; i < minimizers.size(); i++) {
        if (i >= limit) {
            start = i; limit = i + 1;
            run_hits = minimizers[i].hits();
            for (size_t j = i + 1; j < minimizers.size() && minimizers[j].key == minimizers[i].key; j++) { limit++; run_hits += minimizers[j].hits(); }
            taking_run = false;
        }
        const Minimizer& m = minimizers[i];
        bool passing = true;
        // "any-hits"
        if (passing) passing = m.hits() > 0;
        // "hard-hit-cap"
        if (passing) passing = run_hits <= P.hard_hit_cap;
        // "max-min||num-bp-per-min"
        if (passing && P.max_unique_min != 0) {
            size_t cs = m.forward_offset() < P.minimizer_coverage_flank ? 0 : m.forward_offset() - P.minimizer_coverage_flank;
            size_t ce = std::min(read_coverage.size(), m.forward_offset() + m.length + P.minimizer_coverage_flank);
            if (num_minimizers < std::max<size_t>(P.max_unique_min, num_min_by_read_len)) {
                for (size_t x = cs; x < ce; x++) read_coverage[x] = true;
                worst_kept_hits = std::max(m.hits(), worst_kept_hits);
                passing = true;
            } else if (m.hits() > worst_kept_hits) {
                passing = false;
            } else {
                bool covered = false;
                for (size_t x = cs; x < ce; x++) if (read_coverage[x]) { covered = true; break; }
                if (covered) passing = false;
                else { for (size_t x = cs; x < ce; x++) read_coverage[x] = true; passing = true; }
            }
        }
        // "hit-cap||score-fraction"
        if (passing && (P.hit_cap != 0 || P.minimizer_score_fraction != 1.0)) {
            passing = (m.hits() <= P.hit_cap) || (run_hits <= P.hard_hit_cap && selected_score + m.score <= target_score) || taking_run;
            if (passing) selected_score += m.score;
            else target_score = selected_score;
        }
        if (passing) {
            taking_run = true;
            num_minimizers++;
            for (size_t j = 0; j < m.hits(); j++) {
                const gb_hit& occ = ix->hits[m.hit_off + j];
                uint32_t node = (uint32_t)(occ.pos >> 10), off = (uint32_t)(occ.pos & 1023u);
                if (m.is_reverse) {
                    // reverse_base_pos, minimizer_mapper.cpp:4462-4465
                    uint32_t node_length = ix->nodes[node].len;
                    node ^= 1u; off = node_length - off - 1;
                }
                seeds.push_back(Seed{node, off, i, occ.payload});
            }
        }
    }
    return seeds;
}

// ---------------------------------------------------------------------------------------
// clustering: connected components under unoriented minimum distance <= limit
// (definition: snarl_seed_clusterer.hpp:15-50, exhaustive checker snarl_seed_clusterer.cpp:256-315)
// ---------------------------------------------------------------------------------------
static const int64_t INF_DIST = std::numeric_limits<int64_t>::max();

// directed minimum distance between two positions on the forward strand of the chain
static int64_t directed_distance(const gb_flat_index* ix, uint32_t id_a, uint32_t off_a, const gb_dist_payload& pa,
                                 uint32_t id_b, uint32_t off_b, const gb_dist_payload& pb) {
    if (pa.component != pb.component) return INF_DIST;
    if (id_a == id_b) return off_b >= off_a ? (int64_t)off_b - off_a : INF_DIST;
    if (pa.slot < pb.slot) {
        int64_t len_a = ix->nodes[2 * id_a].len;
        return (len_a - (int64_t)off_a) + ((int64_t)(int32_t)pb.x_in - (int64_t)(int32_t)pa.x_out) + (int64_t)off_b;
    }
    if (pa.slot == pb.slot) {
        const int64_t t = site_distance(ix, pa, pb);           // inside one site: its all-pairs table
        if (t >= 0) return ((int64_t)ix->nodes[2 * id_a].len - (int64_t)off_a) + t + (int64_t)off_b;
    }
    return INF_DIST;   // unreachable inside the site, or b is upstream of a
}

int64_t unoriented_distance(const gb_flat_index* ix, const Seed& a, const Seed& b) {
    // project both seeds onto the forward strand (min over the four strand combinations
    // reduces to this on a graph without reversing edges)
    uint32_t ida = a.node >> 1, idb = b.node >> 1;
    uint32_t oa = (a.node & 1) ? ix->nodes[a.node].len - 1 - a.offset : a.offset;
    uint32_t ob = (b.node & 1) ? ix->nodes[b.node].len - 1 - b.offset : b.offset;
    int64_t d1 = directed_distance(ix, ida, oa, a.payload, idb, ob, b.payload);
    int64_t d2 = directed_distance(ix, idb, ob, b.payload, ida, oa, a.payload);
    return std::min(d1, d2);
}

struct UnionFind {
    std::vector<size_t> p;
    explicit UnionFind(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    size_t find(size_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(size_t a, size_t b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

// cluster_seeds (single-end), snarl_seed_clusterer.cpp:28-63
std::vector<Cluster> cluster_seeds(const gb_flat_index* ix, const std::vector<Seed>& seeds, size_t limit) {
    UnionFind uf(seeds.size());
    for (size_t i = 0; i < seeds.size(); i++)
        for (size_t j = 0; j < i; j++) {
            int64_t d = unoriented_distance(ix, seeds[i], seeds[j]);
            if (d != INF_DIST && d <= (int64_t)limit) uf.unite(i, j);
        }
    std::map<size_t, size_t> root_to_cluster;
    std::vector<Cluster> result;
    for (size_t i = 0; i < seeds.size(); i++) {
        size_t r = uf.find(i);
        auto it = root_to_cluster.find(r);
        if (it == root_to_cluster.end()) { it = root_to_cluster.emplace(r, result.size()).first; result.emplace_back(); }
        result[it->second].seeds.push_back(i);
    }
    std::sort(result.begin(), result.end(), [](const Cluster& a, const Cluster& b) { return a.seeds.front() < b.seeds.front(); });
    return result;
}

// score_cluster, minimizer_mapper.cpp:4738-4780
void score_cluster(Cluster& cluster, const std::vector<Minimizer>& minimizers, const std::vector<Seed>& seeds, size_t seq_length) {
    cluster.score = 0.0; cluster.coverage = 0.0;
    std::vector<bool> present(minimizers.size(), false);
    for (size_t hit : cluster.seeds) present[seeds[hit].source] = true;
    std::vector<bool> covered(seq_length, false);
    for (size_t j = 0; j < minimizers.size(); j++) {
        if (!present[j]) continue;
        cluster.score += minimizers[j].score;
        size_t s = minimizers[j].forward_offset();
        for (size_t x = s; x < s + (size_t)minimizers[j].length && x < seq_length; x++) covered[x] = true;
    }
    size_t cnt = 0; for (bool b : covered) cnt += b;
    cluster.coverage = cnt / (double)seq_length;
}

// score_extension_group, minimizer_mapper.cpp:5022-5203
int score_extension_group(size_t seq_len, const std::vector<GaplessExtension>& ext, int max_mismatches,
                          int gap_open_penalty, int gap_extend_penalty) {
    if (ext.empty()) return 0;
    if (ext.front().full() && ext.front().mismatch_positions.size() <= (size_t)max_mismatches) return ext.front().score;
    if (seq_len == 0) return 0;
    int64_t sweep_line = 0, last_sweep_line = 0;
    size_t unentered = 0;
    std::vector<std::pair<size_t, size_t>> end_heap;
    auto min_heap_on_first = [](const std::pair<size_t, size_t>& a, const std::pair<size_t, size_t>& b) { return a.first > b.first; };
    int best_gap_score = 0;
    std::vector<int> best_chain_score(ext.size(), 0);
    int best_past_ending_score_ever = 0;
    std::vector<std::pair<int, size_t>> overlap_heap;
    int overlap_score_offset = 0;
    while (last_sweep_line <= (int64_t)seq_len) {
        int64_t next_seed_start = std::numeric_limits<int64_t>::max();
        if (unentered < ext.size()) next_seed_start = ext[unentered].read_interval.first;
        int64_t next_seed_end = std::numeric_limits<int64_t>::max();
        if (!end_heap.empty()) next_seed_end = end_heap.front().first;
        sweep_line = std::min(std::min(next_seed_end, next_seed_start), (int64_t)seq_len);
        int sweep_distance = (int)(sweep_line - last_sweep_line + 1);
        int best_past_ending_score_here = 0;
        while (!end_heap.empty() && (int64_t)end_heap.front().first == sweep_line) {
            size_t past_ending = end_heap.front().second;
            best_past_ending_score_here = std::max(best_past_ending_score_here, best_chain_score[past_ending]);
            std::pop_heap(end_heap.begin(), end_heap.end(), min_heap_on_first);
            end_heap.pop_back();
        }
        best_past_ending_score_ever = std::max(best_past_ending_score_ever, best_past_ending_score_here);
        if (sweep_line == (int64_t)seq_len) break;
        overlap_score_offset += sweep_distance * gap_extend_penalty;
        int best_overlap_score = 0;
        while (!overlap_heap.empty()) {
            if ((int64_t)overlap_heap.front().second <= sweep_line) { std::pop_heap(overlap_heap.begin(), overlap_heap.end()); overlap_heap.pop_back(); }
            else { best_overlap_score = overlap_heap.front().first + overlap_score_offset; break; }
        }
        if (best_gap_score != 0) best_gap_score -= sweep_distance * gap_extend_penalty;
        best_gap_score = std::max(0, std::max(best_gap_score, best_past_ending_score_here - (gap_open_penalty - gap_extend_penalty)));
        while (unentered < ext.size() && (int64_t)ext[unentered].read_interval.first == sweep_line) {
            best_chain_score[unentered] = std::max(best_overlap_score, std::max(best_gap_score, best_past_ending_score_here)) + ext[unentered].score;
            size_t extension_length = ext[unentered].read_interval.second - ext[unentered].read_interval.first;
            int raw_overlap_score = best_chain_score[unentered] - gap_open_penalty - gap_extend_penalty * (int)extension_length;
            int encoded_overlap_score = raw_overlap_score - overlap_score_offset;
            overlap_heap.emplace_back(encoded_overlap_score, ext[unentered].read_interval.second);
            std::push_heap(overlap_heap.begin(), overlap_heap.end());
            end_heap.emplace_back(ext[unentered].read_interval.second, unentered);
            std::push_heap(end_heap.begin(), end_heap.end(), min_heap_on_first);
            unentered++;
        }
        last_sweep_line = sweep_line + 1;
    }
    return best_past_ending_score_ever;
}

// GaplessExtension::to_path, gbwt_extender.cpp:119-156
std::vector<Mapping> extension_to_path(const Graph& g, const GaplessExtension& e, const std::string& sequence) {
    std::vector<Mapping> result;
    auto mismatch = e.mismatch_positions.begin();
    size_t read_offset = e.read_interval.first, node_offset = e.offset;
    for (size_t i = 0; i < e.path.size(); i++) {
        size_t limit = std::min(read_offset + g.get_length(e.path[i]) - node_offset, e.read_interval.second);
        Mapping m; m.node = e.path[i]; m.offset = (uint32_t)node_offset;
        while (mismatch != e.mismatch_positions.end() && *mismatch < limit) {
            if (read_offset < *mismatch) m.edits.push_back(Edit{(uint32_t)(*mismatch - read_offset), (uint32_t)(*mismatch - read_offset), ""});
            m.edits.push_back(Edit{1, 1, std::string(1, sequence[*mismatch])});
            read_offset = *mismatch + 1;
            ++mismatch;
        }
        if (read_offset < limit) { m.edits.push_back(Edit{(uint32_t)(limit - read_offset), (uint32_t)(limit - read_offset), ""}); read_offset = limit; }
        result.push_back(std::move(m));
        node_offset = 0;
    }
    return result;
}

// ---------------------------------------------------------------------------------------
// faster_cap, minimizer_mapper.cpp:2946-3260
// ---------------------------------------------------------------------------------------
static double get_prob_of_disruption_in_column(const std::vector<Minimizer>& minimizers, const std::string& quality,
                                               const std::vector<size_t>& explored, size_t begin, size_t end, size_t index) {
    double p = phred_to_prob((uint8_t)quality[index]);
    for (size_t it = begin; it != end; ++it) {
        const Minimizer& m = minimizers[explored[it]];
        if (!(m.forward_offset() <= index && index < m.forward_offset() + m.length)) {
            size_t possible = std::min((size_t)m.length, std::min(index - m.agglomeration_start + 1,
                                                                   (m.agglomeration_start + m.agglomeration_length) - index));
            p *= prob_for_at_least_one(m.hash, possible);
        }
    }
    return p;
}

static double get_log10_prob_of_disruption_in_interval(const std::vector<Minimizer>& minimizers, const std::string& quality,
                                                       const std::vector<size_t>& explored, size_t begin, size_t end,
                                                       size_t left, size_t right) {
    if (left == right) return 0;
    double p = get_prob_of_disruption_in_column(minimizers, quality, explored, begin, end, left);
    for (size_t i = left + 1; i < right; i++) {
        double col_p = get_prob_of_disruption_in_column(minimizers, quality, explored, begin, end, i);
        p = (p + col_p - (p * col_p));
    }
    return std::log10(p);
}

double faster_cap(const std::vector<Minimizer>& minimizers, std::vector<size_t>& explored,
                  const std::string& sequence, const std::string& quality) {
    if (quality.empty()) return std::numeric_limits<double>::infinity();
    std::stable_sort(explored.begin(), explored.end(), [&](size_t a, size_t b) {
        size_t a_end = minimizers[a].agglomeration_start + minimizers[a].agglomeration_length;
        size_t b_end = minimizers[b].agglomeration_start + minimizers[b].agglomeration_length;
        return a_end < b_end || (a_end == b_end && minimizers[a].agglomeration_start < minimizers[b].agglomeration_start);
    });
    std::vector<double> c(explored.size() + 1, -std::numeric_limits<double>::infinity());
    c[0] = 0.0;
    auto iteratee = [&](size_t left, size_t right, size_t bottom, size_t top) {
        double p_here = get_log10_prob_of_disruption_in_interval(minimizers, quality, explored, bottom, top, left, right);
        double p = c[bottom] + p_here;
        for (size_t i = bottom + 1; i < top + 1; i++) if (c[i] < p) c[i] = p;
    };
    // for_each_agglomeration_interval :3088-3161
    if (!explored.empty()) {
        std::vector<const Minimizer*> stack = {&minimizers[explored.front()]};
        size_t stack_front = 0;
        size_t left = stack[0]->agglomeration_start;
        size_t bottom = 0;
        auto emit_preceding_intervals = [&](size_t right) {
            while (left < right) {
                size_t stack_size = stack.size() - stack_front;
                size_t stack_top_end = stack[stack_front]->agglomeration_start + stack[stack_front]->agglomeration_length;
                if (stack_top_end <= right) {
                    iteratee(left, stack_top_end, bottom, bottom + stack_size);
                    left = stack_size == 1 ? right : stack_top_end;
                    bottom += 1;
                    stack_front++;
                } else {
                    iteratee(left, right, bottom, bottom + stack_size);
                    left = right;
                }
            }
        };
        for (size_t it = 1; it < explored.size(); ++it) {
            const Minimizer& item = minimizers[explored[it]];
            emit_preceding_intervals(item.agglomeration_start);
            stack.push_back(&item);
        }
        emit_preceding_intervals(sequence.size());
    }
    return -c.back() * 10;
}

// ---------------------------------------------------------------------------------------
// map_from_extensions, minimizer_mapper.cpp:608-1284 (max_multimaps 1, find_supplementaries off)
// ---------------------------------------------------------------------------------------
thread_local StageTrace* g_stage_trace = nullptr;

Alignment map_from_extensions(const gb_flat_index* ix, const gb_scores& scores, const gb_map_params& P,
                              const std::string& sequence, const std::string& quality, MapCounters* counters) {
    Graph g(ix);
    LazyRNG rng([&]() { return sequence; });
    const double log_base = recover_log_base(scores);

    std::vector<Minimizer> minimizers_in_read = find_minimizers(ix, P, sequence);
    std::vector<size_t> score_order = sort_minimizers_by_score(minimizers_in_read, rng);
    std::vector<Minimizer> minimizers;
    minimizers.reserve(score_order.size());
    for (size_t i : score_order) minimizers.push_back(minimizers_in_read[i]);

    std::vector<Seed> seeds = find_seeds(ix, P, minimizers, sequence.size());
    size_t distance_limit = std::max<size_t>(P.distance_limit, sequence.size() + 50);
    std::vector<Cluster> clusters = cluster_seeds(ix, seeds, distance_limit);
    if (counters) { counters->minimizers += minimizers.size(); counters->seeds += seeds.size(); counters->clusters += clusters.size(); }

    double best_cluster_score = 0.0, second_best_cluster_score = 0.0;
    for (size_t i = 0; i < clusters.size(); i++) {
        score_cluster(clusters[i], minimizers, seeds, sequence.length());
        if (clusters[i].score > best_cluster_score) { second_best_cluster_score = best_cluster_score; best_cluster_score = clusters[i].score; }
        else if (clusters[i].score > second_best_cluster_score) second_best_cluster_score = clusters[i].score;
    }
    double cluster_score_cutoff = best_cluster_score - P.cluster_score_threshold;
    if (cluster_score_cutoff - P.pad_cluster_score_threshold < second_best_cluster_score)
        cluster_score_cutoff = std::min(cluster_score_cutoff, second_best_cluster_score);
    if (g_stage_trace) { auto& t = g_stage_trace->reads[0]; t.minimizers = minimizers; t.seeds = seeds; t.clusters = clusters; }

    std::vector<std::vector<GaplessExtension>> cluster_extensions;
    std::vector<std::vector<size_t>> minimizer_extended_cluster_count;
    std::vector<bool> minimizer_explored(minimizers.size(), false);
    size_t kept_cluster_count = 0;

    process_until_threshold_e<double>(clusters.size(),
        [&](size_t i) -> double { return clusters[i].coverage; },
        [&](size_t a, size_t b) -> bool {
            return (clusters[a].coverage > clusters[b].coverage) ||
                   (clusters[a].coverage == clusters[b].coverage && clusters[a].score > clusters[b].score);
        },
        [&](size_t) -> bool { return false; },
        P.cluster_coverage_threshold, P.min_extensions, P.max_extensions, rng,
        [&](size_t cluster_num, size_t, bool escaped) -> bool {
            Cluster& cluster = clusters[cluster_num];
            if (P.cluster_score_threshold != 0 && cluster.score < cluster_score_cutoff &&
                kept_cluster_count >= P.min_extensions && !escaped) return false;
            // extend_seed_group :4784-5018
            minimizer_extended_cluster_count.emplace_back(minimizers.size(), 0);
            std::vector<std::pair<uint32_t, int64_t>> seed_matchings;
            for (size_t seed_index : cluster.seeds) {
                const Seed& seed = seeds[seed_index];
                seed_matchings.emplace_back(seed.node, (int64_t)minimizers[seed.source].offset - (int64_t)seed.offset);
                minimizer_extended_cluster_count.back()[seed.source]++;
            }
            if (g_stage_trace) g_stage_trace->reads[0].items.push_back(StageTrace::Item{cluster_num, cluster.fragment, seed_matchings});
            cluster_extensions.emplace_back(extend(g, scores, seed_matchings, sequence, P.max_extension_mismatches, 0.8, true));
            if (counters) counters->extend_calls++;
            kept_cluster_count++;
            return true;
        },
        [&](size_t) {}, [&](size_t) {});

    std::vector<int> cluster_extension_scores(cluster_extensions.size(), 0);
    for (size_t i = 0; i < cluster_extensions.size(); i++)
        cluster_extension_scores[i] = score_extension_group(sequence.size(), cluster_extensions[i], 4, scores.gap_open, scores.gap_extend);

    std::vector<Alignment> alignments;
    process_until_threshold_e<int>(cluster_extension_scores.size(),
        [&](size_t i) -> int { return cluster_extension_scores[i]; },
        [&](size_t a, size_t b) -> bool { return cluster_extension_scores[a] > cluster_extension_scores[b]; },
        [&](size_t) -> bool { return false; },
        P.extension_set_score_threshold, P.min_extension_sets, P.max_alignments, rng,
        [&](size_t extension_num, size_t, bool escaped) -> bool {
            if (cluster_extension_scores[extension_num] < P.extension_set_min_score && !escaped) return false;
            auto& extensions = cluster_extensions[extension_num];
            std::vector<Alignment> best_alignments(1);
            // GaplessExtender::full_length_extensions(extensions) with the default max_mismatches 4
            if (!extensions.empty() && extensions.front().full() && extensions.front().mismatch_positions.size() <= 4) {
                auto fill = [&](const GaplessExtension& e, Alignment& a) {
                    a.path = extension_to_path(g, e, sequence);
                    a.score = e.score;
                    a.identity = sequence.empty() ? 0.0 : (sequence.length() - e.mismatch_positions.size()) / (double)sequence.length();
                };
                fill(extensions.front(), best_alignments.front());
                for (auto it = extensions.begin() + 1; it != extensions.end() && it->full(); ++it) {
                    best_alignments.emplace_back();
                    fill(*it, best_alignments.back());
                }
                if (counters) counters->direct++;
            } else if (P.do_dp) {
                best_alignments.emplace_back();
                find_optimal_tail_alignments(g, scores, P, sequence, extensions, rng, best_alignments[0], best_alignments[1], counters);
            }
            for (auto it = best_alignments.begin(); it != best_alignments.end() && it->score != 0 &&
                                                    it->score >= best_alignments[0].score * 0.8; ++it) {
                alignments.emplace_back(std::move(*it));
            }
            for (size_t i = 0; i < minimizer_extended_cluster_count[extension_num].size(); i++)
                if (minimizer_extended_cluster_count[extension_num][i] > 0) minimizer_explored[i] = true;
            return true;
        },
        [&](size_t) {}, [&](size_t) {});

    if (alignments.empty()) alignments.emplace_back();

    std::vector<Alignment> mappings;
    std::vector<double> out_scores;
    process_until_threshold_e<double>(alignments.size(),
        [&](size_t i) -> double { return alignments[i].score; },
        [&](size_t a, size_t b) -> bool { return alignments[a].score > alignments[b].score; },
        [&](size_t) -> bool { return false; },
        0, 1, P.max_multimaps, rng,
        [&](size_t alignment_num, size_t, bool) -> bool {
            out_scores.emplace_back(alignments[alignment_num].score);
            mappings.emplace_back(std::move(alignments[alignment_num]));
            return true;
        },
        [&](size_t alignment_num) { out_scores.emplace_back(alignments[alignment_num].score); },
        [&](size_t) {});

    double mapq = mappings.front().path.empty() ? 0 : compute_max_mapping_quality(out_scores, log_base);
    std::vector<size_t> explored_minimizers;
    for (size_t i = 0; i < minimizers.size(); i++) if (minimizer_explored[i]) explored_minimizers.push_back(i);
    double escape_bonus = mapq < std::numeric_limits<int32_t>::max() ? 1.0 : 2.0;
    double mapq_explored_cap = escape_bonus * faster_cap(minimizers, explored_minimizers, sequence, quality);
    Alignment& out = mappings.front();
    out.mapq_uncapped = mapq;
    out.mapq_explored_cap = mapq_explored_cap;
    mapq = std::round(std::min(mapq_explored_cap, std::min(mapq, 60.0)));
    out.mapq = std::max(std::min(mapq, 60.0), 0.0);
    // mappings 1 .. max_multimaps - 1 are the secondaries, in score order (:1199-1206); only the primary carries a MAPQ
    Alignment result = std::move(out);
    for (size_t i = 1; i < mappings.size(); i++) { mappings[i].secondary = true; result.secondaries.push_back(std::move(mappings[i])); }
    return result;
}

// ---------------------------------------------------------------------------------------
// record packing (include/giraffe_b200.h gb_alignment / gb_mapping / edit words)
// ---------------------------------------------------------------------------------------
static uint32_t base_code(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 0; } }

int pack_alignment(const Alignment& a, uint32_t read_id, gb_alignment* rec, gb_mapping* mappings, uint32_t mapping_cap,
                   uint32_t* edits, uint32_t edit_cap, uint32_t mapping_base, uint32_t edit_base) {
    rec->read_id = read_id; rec->score = a.score;
    rec->mapq = (uint8_t)a.mapq;
    rec->flags = (a.path.empty() ? 0 : GB_ALN_MAPPED) | (a.rescued ? GB_ALN_RESCUED : 0) | (a.secondary ? GB_ALN_SECONDARY : 0);
    rec->mapping_off = mapping_base; rec->edit_off = edit_base;
    rec->mapq_uncapped = (float)a.mapq_uncapped; rec->mapq_explored_cap = (float)a.mapq_explored_cap;
    if (a.path.size() > mapping_cap) return -1;
    uint32_t ne = 0;
    for (size_t i = 0; i < a.path.size(); i++) {
        const Mapping& m = a.path[i];
        mappings[i].node = m.node; mappings[i].offset = (uint16_t)m.offset; mappings[i].n_edits = (uint16_t)m.edits.size();
        for (const Edit& e : m.edits) {
            if (ne >= edit_cap) return -1;
            uint32_t word;
            if (e.from_length == e.to_length && e.sequence.empty()) word = (e.from_length << 4) | GB_EDIT_MATCH;
            else if (e.from_length == e.to_length) word = (e.from_length << 4) | (base_code(e.sequence[0]) << 2) | GB_EDIT_SUB;
            else if (e.from_length == 0) word = (e.to_length << 4) | GB_EDIT_INS;
            else word = (e.from_length << 4) | GB_EDIT_DEL;
            edits[ne++] = word;
        }
    }
    rec->n_mappings = (uint16_t)a.path.size(); rec->n_edits = ne;
    return 0;
}

// Records of ranks 1 .. max_multimaps - 1 of read r (rank-major layout of gb_map_batch: record j * n_reads + r).
int pack_secondaries(const Alignment& primary, const gb_map_params& P, uint32_t n_reads, uint32_t r, uint32_t extra_flags,
                     gb_alignment* aln, gb_mapping* mappings, uint32_t* edits) {
    int rc = 0;
    for (uint32_t j = 1; j < P.max_multimaps; j++) {
        const size_t R = (size_t)j * n_reads + r;
        if (j - 1 < primary.secondaries.size()) {
            rc |= pack_alignment(primary.secondaries[j - 1], r, aln + R, mappings + R * P.mapping_cap_per_read, P.mapping_cap_per_read,
                                 edits + R * P.edit_cap_per_read, P.edit_cap_per_read, (uint32_t)(R * P.mapping_cap_per_read), (uint32_t)(R * P.edit_cap_per_read));
            aln[R].flags |= extra_flags;
        } else {
            memset(aln + R, 0, sizeof(gb_alignment));
            aln[R].read_id = r; aln[R].flags = GB_ALN_ABSENT;
            aln[R].mapping_off = (uint32_t)(R * P.mapping_cap_per_read); aln[R].edit_off = (uint32_t)(R * P.edit_cap_per_read);
        }
    }
    return rc;
}

} // namespace oracle

extern "C" void oracle_map_params_default(gb_map_params* p) {
    memset(p, 0, sizeof(*p));
    p->hit_cap = 10; p->hard_hit_cap = 500; p->minimizer_score_fraction = 0.9; p->minimizer_coverage_flank = 250;
    p->max_unique_min = 500; p->num_bp_per_min = 1000; p->distance_limit = 200; p->min_extensions = 2; p->max_extensions = 800;
    p->cluster_score_threshold = 50; p->pad_cluster_score_threshold = 20; p->cluster_coverage_threshold = 0.3;
    p->extension_set_score_threshold = 20; p->extension_score_threshold = 1; p->min_extension_sets = 2;
    p->extension_set_min_score = 20; p->max_alignments = 8; p->max_extension_mismatches = 4; p->max_multimaps = 1;
    p->max_dozeu_cells = (uint32_t)(1.5 * 1024 * 1024); p->do_dp = 1;
    p->fragment_mean = 0; p->fragment_stdev = 0; p->paired_distance_stdevs = 2.0; p->paired_rescue_score_limit = 0.9;
    p->rescue_subgraph_stdevs = 4.0; p->max_rescue_attempts = 15; p->max_fragment_length = 2000;
    p->rescue_seed_limit = 100; p->reserved0 = 0; p->rescue_likelihood_limit = 0.05;
    p->mapping_cap_per_read = 96; p->edit_cap_per_read = 160;
}

extern "C" int oracle_map_batch(const gb_flat_index* ix, const gb_scores* scores, const gb_map_params* p,
                                uint32_t n_reads, const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off,
                                gb_alignment* aln, gb_mapping* mappings, uint32_t* edits, uint8_t* status,
                                int n_threads, uint64_t* counters_out) {
    oracle::MapCounters total;
    int failed = 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        oracle::MapCounters local;
#pragma omp for schedule(dynamic, 256)
        for (int64_t r = 0; r < (int64_t)n_reads; r++) {
            std::string seq((const char*)reads + read_off[r], (size_t)(read_off[r + 1] - read_off[r]));
            std::string qual;
            if (quals) qual.assign((const char*)quals + read_off[r], seq.size());
            oracle::Alignment a = oracle::map_from_extensions(ix, *scores, *p, seq, qual, &local);
            int rc = oracle::pack_alignment(a, (uint32_t)r, aln + r, mappings + (size_t)r * p->mapping_cap_per_read,
                                            p->mapping_cap_per_read, edits + (size_t)r * p->edit_cap_per_read,
                                            p->edit_cap_per_read, (uint32_t)(r * p->mapping_cap_per_read),
                                            (uint32_t)(r * p->edit_cap_per_read));
            rc |= oracle::pack_secondaries(a, *p, n_reads, (uint32_t)r, 0, aln, mappings, edits);
            status[r] = rc == 0 ? GB_ITEM_OK : GB_ITEM_OUT_FULL;
            if (rc) {
#pragma omp atomic
                failed++;
            }
        }
#pragma omp critical
        total.add(local);
    }
    if (counters_out) total.store(counters_out);
    return failed ? -1 : 0;
}

// faster_cap on explicit minimizers (test entry for the reference's MAPQ-cap unit tests,
// unittest/minimizer_mapper.cpp:112-252): minimizer i has core [offset, offset + length), agglomeration
// [agg_start, agg_start + agg_len), forward strand, key = `length` Gs (hash = gbwtgraph key hash); all are explored.
extern "C" double oracle_faster_cap_all_g(uint32_t n, const uint32_t* offset, const uint32_t* length, const uint32_t* agg_start,
                                          const uint32_t* agg_len, uint32_t read_len, uint8_t quality) {
    std::vector<oracle::Minimizer> minimizers(n);
    std::vector<size_t> explored(n);
    for (uint32_t i = 0; i < n; i++) {
        oracle::Minimizer& m = minimizers[i];
        uint64_t key = 0;
        for (uint32_t b = 0; b < length[i]; b++) key = (key << 2) | 2u;          // G = 2
        m.key = key; m.hash = oracle::wang_hash_64(key); m.offset = offset[i]; m.is_reverse = false; m.length = (int32_t)length[i];
        m.agglomeration_start = agg_start[i]; m.agglomeration_length = agg_len[i]; m.hit_cnt = 1; m.score = 1;
        explored[i] = i;
    }
    return oracle::faster_cap(minimizers, explored, std::string(read_len, 'G'), std::string(read_len, (char)quality));
}

// cluster_seeds on explicit positions (test entry for the reference's clusterer unit tests,
// unittest/snarl_seed_clusterer.cpp): seed i = (oriented node, offset on that strand) of read read_of[i];
// read clusters = components of "unoriented minimum distance <= read_limit" among the seeds of one read,
// fragment clusters = components at fragment_limit over all seeds (snarl_seed_clusterer.hpp:15-50).
// Labels are the smallest seed index of the component.
extern "C" void oracle_cluster_positions(const gb_flat_index* ix, uint32_t n, const uint32_t* node, const uint32_t* offset, const uint32_t* read_of,
                                         uint32_t read_limit, uint32_t fragment_limit, uint32_t* read_label, uint32_t* fragment_label) {
    std::vector<oracle::Seed> seeds(n);
    for (uint32_t i = 0; i < n; i++) seeds[i] = oracle::Seed{node[i], offset[i], 0, ix->dist[node[i] >> 1]};
    oracle::UnionFind reads(n), frags(n);
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = 0; j < i; j++) {
            const int64_t d = oracle::unoriented_distance(ix, seeds[i], seeds[j]);
            if (d == oracle::INF_DIST) continue;
            if (read_of[i] == read_of[j] && d <= (int64_t)read_limit) reads.unite(i, j);
            if (fragment_limit && d <= (int64_t)fragment_limit) frags.unite(i, j);
        }
    for (uint32_t i = 0; i < n; i++) { read_label[i] = (uint32_t)reads.find(i); fragment_label[i] = (uint32_t)frags.find(i); }
}

// minimizer_regions of one sequence (test entry): up to `cap` minimizers as (key, forward offset, is_reverse, agglomeration start,
// agglomeration length); returns their number.
extern "C" uint32_t oracle_minimizer_regions(const uint8_t* seq, uint32_t len, uint32_t k, uint32_t w, uint32_t cap,
                                             uint64_t* key, uint32_t* fwd_offset, uint8_t* is_reverse, uint32_t* agg_start, uint32_t* agg_len) {
    const std::vector<oracle::Minimizer> ms = oracle::minimizer_regions(std::string((const char*)seq, len), k, w);
    uint32_t n = 0;
    for (const oracle::Minimizer& m : ms) {
        if (n >= cap) break;
        key[n] = m.key; fwd_offset[n] = (uint32_t)m.forward_offset(); is_reverse[n] = m.is_reverse ? 1 : 0;
        agg_start[n] = (uint32_t)m.agglomeration_start; agg_len[n] = (uint32_t)m.agglomeration_length; n++;
    }
    return (uint32_t)ms.size();
}
