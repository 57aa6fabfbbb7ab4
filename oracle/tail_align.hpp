// ORACLE — TEST INFRASTRUCTURE ONLY.  Tail alignment of non-full-length gapless extensions:
//   find_optimal_tail_alignments            minimizer_mapper.cpp:5369-5622
//   get_best_alignment_against_any_tree     minimizer_mapper.cpp:5626-5743
//   get_tail_forest / dfs_gbwt              minimizer_mapper.cpp:5745-6013
//   Aligner::align_pinned(xdrop)            aligner.cpp:628-686
//   DozeuInterface::align_pinned/do_poa/... dozeu_interface.cpp:210-572, :724-766
// and the X-drop DP itself (vgteam/dozeu @ d0e9ba6, ABSENT: see xdrop_pinned() below).
#pragma once
#include "mapper_common.hpp"

namespace oracle {

struct MapCounters {
    uint64_t reads = 0, minimizers = 0, seeds = 0, clusters = 0, extend_calls = 0, direct = 0;
    uint64_t tail_dps = 0, tail_cells = 0, tail_nodes = 0, tail_bases = 0, path_nodes = 0, edits = 0, rescues = 0;
    void add(const MapCounters& o) {
        reads += o.reads; minimizers += o.minimizers; seeds += o.seeds; clusters += o.clusters; extend_calls += o.extend_calls;
        direct += o.direct; tail_dps += o.tail_dps; tail_cells += o.tail_cells; tail_nodes += o.tail_nodes; tail_bases += o.tail_bases;
        path_nodes += o.path_nodes; edits += o.edits; rescues += o.rescues;
    }
    void store(uint64_t* out) const {
        out[0] = reads; out[1] = minimizers; out[2] = seeds; out[3] = clusters; out[4] = extend_calls; out[5] = direct;
        out[6] = tail_dps; out[7] = tail_cells; out[8] = tail_nodes; out[9] = tail_bases; out[10] = path_nodes; out[11] = edits; out[12] = rescues;
    }
};

// One tree of the tail forest: (parent index or -1, oriented node) in DFS visit order; the
// root's sequence is trimmed by root_trim bases on the left (TreeSubgraph, tree_subgraph.hpp:33-136).
struct TailTree {
    std::vector<std::pair<int64_t, uint32_t>> nodes;
    size_t root_trim = 0;
};

// Result of a pinned X-drop alignment against one tree, in tree space.
struct PinnedAlignment {
    int32_t score = 0;
    std::vector<Mapping> path;   // Mapping.node = tree index + 1 (TreeSubgraph id), offsets in trimmed coordinates
};

// A read-vs-DAG alignment problem (full_dp.cpp, xdrop_dag.cpp): oriented graph nodes in topological order
// and, per node, the indices of its predecessors.  Results carry Mapping.node = index into `node`.
struct DagProblem { std::vector<uint32_t> node; std::vector<std::vector<uint32_t>> pred; };
struct LocalAlignmentResult { int32_t score = 0; std::vector<Mapping> path; };
LocalAlignmentResult sw_local_dag(const Graph& g, const gb_scores& sc, const DagProblem& P, const std::string& q, uint64_t* cells);
LocalAlignmentResult align_xdrop_dag(const Graph& g, const gb_scores& sc, const DagProblem& P, const std::string& query,
                                     bool has_seed, uint32_t seed_u, uint32_t seed_o, uint32_t seed_q, uint32_t max_gap, uint64_t* cells);
size_t longest_detectable_gap(const gb_scores& s, size_t read_length, size_t read_pos);

// The X-drop DP contract (see tail_align.cpp for the full statement).
PinnedAlignment xdrop_pinned(const Graph& g, const gb_scores& scores, const TailTree& tree,
                             const std::string& sequence, uint32_t max_gap_length, uint64_t* cells = nullptr);

std::vector<TailTree> get_tail_forest(const Graph& g, const gb_scores& scores, const GaplessExtension& ext,
                                      size_t read_length, bool left_tails, size_t* longest_detectable_gap);

void find_optimal_tail_alignments(const Graph& g, const gb_scores& scores, const gb_map_params& P,
                                  const std::string& sequence, const std::vector<GaplessExtension>& extended_seeds,
                                  LazyRNG& rng, Alignment& best, Alignment& second_best, MapCounters* counters);

std::vector<Mapping> extension_to_path(const Graph& g, const GaplessExtension& e, const std::string& sequence);
double path_identity(const std::vector<Mapping>& path);

} // namespace oracle
