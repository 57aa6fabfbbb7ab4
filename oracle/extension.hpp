// ORACLE — TEST INFRASTRUCTURE ONLY.  vg::GaplessExtension (gbwt_extender.hpp:30-109).
#pragma once
#include "gbwt_view.hpp"
#include <algorithm>
#include <utility>
#include <vector>

namespace oracle {

struct GaplessExtension {
    std::vector<uint32_t> path;
    size_t offset = 0;
    BidirectionalState state;
    std::pair<size_t, size_t> read_interval{0, 0};
    std::vector<size_t> mismatch_positions;
    int32_t score = 0;
    bool left_full = false, right_full = false;
    bool left_maximal = false, right_maximal = false;
    uint32_t internal_score = 0;
    uint32_t old_score = 0;

    size_t length() const { return read_interval.second - read_interval.first; }
    bool empty() const { return length() == 0; }
    bool full() const { return left_full & right_full; }
    bool exact() const { return mismatch_positions.empty(); }

    // gbwt_extender.cpp:23-53
    bool contains(const Graph& g, uint32_t node, int64_t diag) const {
        size_t read_offset = read_interval.first, node_offset = offset;
        for (uint32_t h : path) {
            size_t len = std::min<size_t>(g.get_length(h) - node_offset, read_interval.second - read_offset);
            if (h == node && (int64_t)read_offset - (int64_t)node_offset == diag) return true;
            read_offset += len;
            node_offset = 0;
        }
        return false;
    }

    // starting_position / tail_position, gbwt_extender.cpp:55-87: (oriented node, offset) of the first aligned base and
    // of the position just past the last one (the offset may equal the node length)
    std::pair<uint32_t, size_t> starting_position(const Graph&) const { return {path.front(), offset}; }
    std::pair<uint32_t, size_t> tail_position(const Graph& g) const {
        size_t tail_off = offset + length();
        for (size_t i = 0; i + 1 < path.size(); i++) tail_off -= g.get_length(path[i]);
        return {path.back(), tail_off};
    }

    // gbwt_extender.cpp:89-117
    size_t overlap(const Graph& g, const GaplessExtension& another) const {
        size_t result = 0;
        size_t this_pos = read_interval.first, another_pos = another.read_interval.first;
        auto this_iter = path.begin(), another_iter = another.path.begin();
        size_t this_offset = offset, another_offset = another.offset;
        while (this_pos < read_interval.second && another_pos < another.read_interval.second) {
            if (this_pos == another_pos && *this_iter == *another_iter && this_offset == another_offset) {
                size_t len = std::min({(size_t)g.get_length(*this_iter) - this_offset,
                                       read_interval.second - this_pos,
                                       another.read_interval.second - another_pos});
                result += len; this_pos += len; another_pos += len;
                ++this_iter; ++another_iter; this_offset = 0; another_offset = 0;
            } else if (this_pos <= another_pos) {
                this_pos += g.get_length(*this_iter) - this_offset; ++this_iter; this_offset = 0;
            } else {
                another_pos += g.get_length(*another_iter) - another_offset; ++another_iter; another_offset = 0;
            }
        }
        return result;
    }

    bool operator<(const GaplessExtension& another) const { return score < another.score; }
    bool operator==(const GaplessExtension& another) const {
        return read_interval == another.read_interval && state == another.state && offset == another.offset;
    }
    bool operator!=(const GaplessExtension& another) const { return !(*this == another); }
};


} // namespace oracle
