// Compile / link check of include/giraffe_b200.hpp against libgiraffe_b200.so, plus the parts that run without a GPU:
// FragmentLengthDistribution through the mirror class, and the error behaviour of a mapper without a device.
#include "giraffe_b200.hpp"
#include <cmath>
#include <cstdio>

using namespace giraffe_b200;

int main() {
    FragmentLengthDistribution distr(100, 25, 0.9);
    for (int i = 0; i < 99; i++) distr.register_fragment_length(380 + (i * 37) % 80);
    if (distr.is_finalized() || distr.curr_sample_size() != 99) { puts("FAIL: finalized too early"); return 1; }
    distr.register_fragment_length(400);
    if (!distr.is_finalized() || !(std::fabs(distr.mean() - 420.0) < 15.0) || !(distr.std_dev() > 5.0)) { printf("FAIL: estimate %f %f\n", distr.mean(), distr.std_dev()); return 1; }
    // no device: the mapper must throw, not fall back
    MinimizerMapper mapper(nullptr);
    mapper.max_rescue_attempts = 0;
    if (mapper.params.max_rescue_attempts != 0 || mapper.hit_cap != 10) { puts("FAIL: parameter aliases"); return 1; }
    mapper.force_fragment_length_distr(400.0, 50.0);
    if (!mapper.fragment_distr_is_finalized() || mapper.get_fragment_length_mean() != 400.0) { puts("FAIL: forced distribution"); return 1; }
    Alignment a; a.sequence = "ACGTACGTACGTACGTACGTACGTACGTACGTACGT"; a.name = "r1";
    bool threw = false;
    try { mapper.map(a); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { puts("FAIL: map() without a device did not throw"); return 1; }
    Alignment b = a; b.name = "r2";
    threw = false;
    try { mapper.map_paired(a, b); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { puts("FAIL: map_paired() without a device did not throw"); return 1; }
    // chaining seams: to_anchors is host side; the device entries throw without a device
    {
        const char* node = "AAAAAAAAAA"; const uint64_t node_off[2] = {0, 10}; const uint32_t path[1] = {2}; const uint64_t path_off[2] = {0, 1};
        gb_host_index* host = nullptr; gb_flat_index flat;
        if (gb_index_build(1, (const uint8_t*)node, node_off, 1, path, path_off, nullptr, 5, 3, &host) != GB_OK || gb_index_view(host, &flat) != GB_OK) { puts("FAIL: index"); return 1; }
        gb_scores sc{1, 4, 6, 1, 5};
        // unittest/minimizer_mapper.cpp:882-1048 with both outer minimizers on the read's reverse strand
        auto anchors = algorithms::to_anchors(flat, sc, {{2, 2}, {2, 9}, {2, 1}, {2, 3}}, {2, 9, 1, 3}, {1, 1, 0, 0}, {3, 3, 3, 2});
        const uint32_t want_start[4] = {0, 7, 1, 3}, want_len[4] = {3, 3, 3, 2};
        for (int i = 0; i < 4; i++) if (anchors[i].read_start != want_start[i] || anchors[i].length != want_len[i]) { puts("FAIL: to_anchors"); return 1; }
        if (anchors[0].start_hint_offset != 2 || anchors[0].end_hint_offset != 1 || anchors[2].start_hint_offset != 0) { puts("FAIL: hint offsets"); return 1; }
        threw = false;
        try { algorithms::transition_candidates(nullptr, {{2, 2}, {2, 9}}, 100); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { puts("FAIL: transition_candidates() without a device did not throw"); return 1; }
        threw = false;
        try { algorithms::find_best_chains(nullptr, anchors, {}); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { puts("FAIL: find_best_chains() without a device did not throw"); return 1; }
        if (algorithms::find_best_chains(nullptr, {}, {}).front().first != 0) { puts("FAIL: empty chain"); return 1; }
        gb_index_free(host);
    }
    puts("shim ok");
    return 0;
}
