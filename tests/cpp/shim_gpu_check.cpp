// GPU run of the C++ mirror: load a flat index file, map single-end reads and pairs through giraffe_b200::MinimizerMapper,
// print one line per read (name, score, MAPQ, first node, first offset, mappings) for the Python side to compare with
// the records it gets from the C ABI directly.   usage: shim_gpu_check graph.gbflat reads.txt   (reads.txt: name seq per line, pairs adjacent)
#include "giraffe_b200.hpp"
#include <cstdio>
#include <fstream>
#include <iostream>

using namespace giraffe_b200;

static void print(const char* mode, const Alignment& a) {
    const bool mapped = !a.path.mapping.empty();
    std::printf("%s %s %d %d %lld %llu %zu\n", mode, a.name.c_str(), a.score, a.mapping_quality,
                mapped ? (long long)(2 * a.path.mapping[0].position.node_id + a.path.mapping[0].position.is_reverse) : -1LL,
                mapped ? (unsigned long long)a.path.mapping[0].position.offset : 0ULL, a.path.mapping.size());
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    gb_host_index* host = nullptr;
    if (gb_index_load(argv[1], &host) != GB_OK) { std::puts("FAIL load"); return 1; }
    gb_flat_index flat;
    gb_index_view(host, &flat);
    gb_device* dev = nullptr;
    if (gb_device_create(&flat, 0, &dev) != GB_OK) { std::printf("FAIL device: %s\n", gb_last_error()); return 1; }
    std::vector<Alignment> reads;
    { std::ifstream in(argv[2]); std::string name, seq; while (in >> name >> seq) { Alignment a; a.name = name; a.sequence = seq; a.quality = std::string(seq.size(), (char)30); reads.push_back(a); } }
    MinimizerMapper mapper(dev);
    mapper.set_alignment_scores(1, 4, 6, 1, 5);
    std::vector<Alignment> se = reads;
    mapper.map_batch(se);
    for (const Alignment& a : se) print("SE", a);
    Alignment one = reads[0];
    std::vector<Alignment> single = mapper.map(one);
    print("ONE", single.at(0));
    mapper.force_fragment_length_distr(400.0, 50.0);
    std::vector<std::pair<Alignment, Alignment>> pairs;
    for (size_t i = 0; i + 1 < reads.size(); i += 2) pairs.push_back({reads[i], reads[i + 1]});
    std::vector<uint8_t> route;
    mapper.map_paired_batch(pairs, &route);
    for (size_t i = 0; i < pairs.size(); i++) { print("PE", pairs[i].first); print("PE", pairs[i].second); if (route[i] != GB_PAIR_PAIRED) { std::puts("FAIL route"); return 1; } }
    gb_device_destroy(dev);
    gb_index_free(host);
    return 0;
}
