// minimizer_common.h — (k,w)-minimizer definition shared by the host index builder and
// (as __host__ __device__ arithmetic) the seeding kernel.
//
// Restates the published algorithm of jltsiren/gbwtgraph @ e27bc43 (MinimizerIndex:
// Key64 2-bit packing, Thomas Wang 64-bit hash, canonical strand = smaller hash with the
// forward strand winning ties, window minimum over w consecutive k-mers, every k-mer whose
// hash equals the window minimum is reported).  That library is absent from
// /root/reference; vg's call sites are minimizer_mapper.cpp:3930 (minimizer_regions) and
// :3933 (find).  Reverse-strand minimizers report the offset of their LAST forward base
// (minimizer_mapper.hpp:583-592).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

#if defined(__CUDACC__)
#define GBMIN_HD __host__ __device__ __forceinline__
#else
#define GBMIN_HD inline
#endif

namespace gbmin {

GBMIN_HD uint64_t hash64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

// A=0 C=1 G=2 T=3, anything else 4 (invalid).
GBMIN_HD uint32_t base_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

struct Minimizer {
    uint64_t key;
    uint64_t hash;
    uint32_t offset;      // forward: first base; reverse: last base (forward coordinates)
    uint32_t is_reverse;
};

struct Region { uint32_t start, length; };

#if !defined(__CUDA_ARCH__)
// Host implementation: candidates per k-mer start, then window minima.
inline void minimizers(const uint8_t* seq, size_t len, uint32_t k, uint32_t w,
                       std::vector<Minimizer>& out, std::vector<Region>* regions) {
    const size_t window_bp = (size_t)k + w - 1;
    if (len < window_bp) return;
    const size_t nk = len - k + 1;
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    std::vector<uint64_t> h(nk), key(nk);
    std::vector<uint8_t> rev(nk), valid(nk, 0);
    uint64_t fk = 0, rk = 0; size_t valid_chars = 0;
    for (size_t i = 0; i < len; i++) {
        uint32_t c = base_code(seq[i]);
        if (c < 4) {
            fk = ((fk << 2) | c) & mask;
            rk = (rk >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
            valid_chars++;
        } else { fk = 0; rk = 0; valid_chars = 0; }
        if (i + 1 >= k) {
            size_t s = i + 1 - k;
            if (valid_chars >= k) {
                uint64_t hf = hash64(fk), hr = hash64(rk);
                valid[s] = 1;
                if (hr < hf) { h[s] = hr; key[s] = rk; rev[s] = 1; }
                else { h[s] = hf; key[s] = fk; rev[s] = 0; }
            }
        }
    }
    // nearest strictly smaller valid k-mer to the left / right
    const long NONE_L = -1; const long NONE_R = (long)nk;
    std::vector<long> left(nk, NONE_L), right(nk, NONE_R), st;
    for (size_t s = 0; s < nk; s++) {
        if (!valid[s]) continue;
        while (!st.empty() && h[st.back()] >= h[s]) st.pop_back();
        left[s] = st.empty() ? NONE_L : st.back();
        st.push_back((long)s);
    }
    st.clear();
    for (size_t s = nk; s-- > 0;) {
        if (!valid[s]) continue;
        while (!st.empty() && h[st.back()] >= h[s]) st.pop_back();
        right[s] = st.empty() ? NONE_R : st.back();
        st.push_back((long)s);
    }
    const long last_window = (long)(len - window_bp);
    for (size_t s = 0; s < nk; s++) {
        if (!valid[s]) continue;
        long lo = (long)s - (long)w + 1; if (lo < left[s] + 1) lo = left[s] + 1; if (lo < 0) lo = 0;
        long hi = (long)s; if (hi > right[s] - (long)w) hi = right[s] - (long)w; if (hi > last_window) hi = last_window;
        if (lo > hi) continue;   // never a window minimum
        Minimizer m;
        m.key = key[s]; m.hash = h[s]; m.is_reverse = rev[s];
        m.offset = rev[s] ? (uint32_t)(s + k - 1) : (uint32_t)s;
        out.push_back(m);
        if (regions) regions->push_back(Region{(uint32_t)lo, (uint32_t)(hi - lo + (long)window_bp)});
    }
}
#endif

} // namespace gbmin
