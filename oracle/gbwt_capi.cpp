// ORACLE — TEST INFRASTRUCTURE ONLY.  C wrappers around oracle::Graph primitives.
#include "oracle.h"
#include "gbwt_view.hpp"

static void pack(const oracle::BidirectionalState& s, int64_t* o) {
    o[0] = s.forward.node; o[1] = s.forward.lo; o[2] = s.forward.hi;
    o[3] = s.backward.node; o[4] = s.backward.lo; o[5] = s.backward.hi;
}
static oracle::BidirectionalState unpack(const int64_t* o) {
    oracle::BidirectionalState s;
    s.forward.node = (uint32_t)o[0]; s.forward.lo = o[1]; s.forward.hi = o[2];
    s.backward.node = (uint32_t)o[3]; s.backward.lo = o[4]; s.backward.hi = o[5];
    return s;
}
extern "C" int oracle_bd_state(const gb_flat_index* ix, uint32_t node, int64_t* state6) {
    oracle::Graph g(ix);
    pack(g.get_bd_state(node), state6);
    return 0;
}
extern "C" int oracle_follow_paths(const gb_flat_index* ix, const int64_t* state6, int backward,
                                   int64_t* out_states, int max_out) {
    oracle::Graph g(ix);
    int n = 0;
    g.follow_paths(unpack(state6), backward != 0, [&](const oracle::BidirectionalState& next) {
        if (n < max_out) pack(next, out_states + 6 * n);
        n++;
        return true;
    });
    return n;
}
