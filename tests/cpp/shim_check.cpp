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
    puts("shim ok");
    return 0;
}
