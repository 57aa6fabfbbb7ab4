// emit.cpp — host-side emission of alignment records (no device work):
//   GAF lines   the text format `vg giraffe -o gaf` writes (alignment_to_gaf lives in libvgio @ d029989,
//               ABSENT from the reference tree; columns follow the published GAF specification and the
//               cs difference string of minimap2: ":n" matches, "*RQ" substitution, "+Q" insertion, "-R" deletion; the
//               conventions the reference's tests fix, unittest/alignment.cpp:398-470 and :793-820, are followed)
//   JSON lines  the protobuf JSON of vg.proto's Alignment as `vg view -aj` prints it (field names as used at
//               gbwt_extender.cpp:119-156, aligner.cpp:120-241, minimizer_mapper.cpp:1146-1216: sequence, path.mapping[]
//               {position{node_id, offset, is_reverse}, edit[]{from_length, to_length, sequence}, rank}, name, quality
//               (base64), mapping_quality, score, identity, fragment_next / fragment_prev, annotation)
// giraffe_main.cpp:2209-2226 hands alignments to an AlignmentEmitter; this is the stand-in until the shim converts
// records to vg::Alignment itself.  PARITY UNPINNED: no libvgio here to compare bytes with.
#include "giraffe_b200.h"

#include <cstdio>
#include <cstring>
#include <string>

namespace {

struct Out {
    char* buf; uint64_t cap; uint64_t n; bool overflow;
    void put(const char* s, size_t len) { if (n + len > cap) { overflow = true; return; } memcpy(buf + n, s, len); n += len; }
    void str(const char* s) { put(s, strlen(s)); }
    void str(const std::string& s) { put(s.data(), s.size()); }
    void ch(char c) { put(&c, 1); }
    void num(long long v) { char t[32]; const int k = snprintf(t, sizeof t, "%lld", v); put(t, (size_t)k); }
    void real(double v) { char t[40]; const int k = snprintf(t, sizeof t, "%.6g", v); put(t, (size_t)k); }
};


// base i of oriented node v (gb_flat_index stores the sequence of both orientations)
inline char node_base(const gb_flat_index* ix, uint32_t v, uint32_t i) {
    const gb_node_rec& r = ix->nodes[v];
    return (char)ix->seq[r.seq_off + i];
}

std::string read_name(const uint8_t* names, const uint64_t* name_off, uint32_t r) {
    if (names && name_off) return std::string((const char*)names + name_off[r], (size_t)(name_off[r + 1] - name_off[r]));
    return "read" + std::to_string(r);
}

void base64(Out& o, const uint8_t* p, size_t n) {
    static const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (size_t i = 0; i < n; i += 3) {
        const uint32_t a = p[i], b = i + 1 < n ? p[i + 1] : 0, c = i + 2 < n ? p[i + 2] : 0;
        o.ch(T[a >> 2]); o.ch(T[((a & 3) << 4) | (b >> 4)]);
        o.ch(i + 1 < n ? T[((b & 15) << 2) | (c >> 6)] : '=');
        o.ch(i + 2 < n ? T[c & 63] : '=');
    }
}

void json_string(Out& o, const std::string& s) {
    o.ch('"');
    for (char c : s) {
        if (c == '"' || c == '\\') { o.ch('\\'); o.ch(c); }
        else if ((unsigned char)c < 0x20) { char t[8]; snprintf(t, sizeof t, "\\u%04x", c); o.str(t); }
        else o.ch(c);
    }
    o.ch('"');
}

} // namespace

extern "C" int gb_emit_gaf(const gb_flat_index* ix, uint32_t n, const gb_alignment* aln, const gb_mapping* mappings, const uint32_t* edits,
                           const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off, const uint8_t* names, const uint64_t* name_off,
                           char* out, uint64_t out_cap, uint64_t* out_used) {
    if (!ix || !aln || !reads || !read_off || !out || !out_used) return GB_ERR_ARG;
    Out o{out, out_cap, 0, false};
    for (uint32_t x = 0; x < n; x++) {
        const gb_alignment& a = aln[x];
        const uint32_t r = a.read_id;
        const uint8_t* seq = reads + read_off[r];
        const uint32_t L = (uint32_t)(read_off[r + 1] - read_off[r]);
        o.str(read_name(names, name_off, r)); o.ch('\t'); o.num(L); o.ch('\t');
        // Conventions pinned by the reference's own GAF tests (unittest/alignment.cpp:398-470, :793-820): the query
        // interval is the whole read (soft clips stay in the difference string as insertions), cs bases are upper
        // case, match runs merge across edits and mappings, a mapping that only carries an insertion (a soft clip
        // parked on an unused node) is left out of the path, an unaligned read has an empty path and cs "+<read>".
        if (!(a.flags & GB_ALN_MAPPED) || a.n_mappings == 0) {
            o.num(0); o.ch('\t'); o.num(L); o.str("\t*\t*\t*\t*\t*\t*\t*\t255");
            o.str("\tcs:Z:+"); o.put((const char*)seq, L);
        } else {
            const gb_mapping* m = mappings + a.mapping_off;
            const uint32_t* e = edits + a.edit_off;
            std::string path, cs;
            uint64_t path_len = 0, matches = 0, block = 0, ref_used_total = 0, pstart = 0;
            uint32_t q = 0, ei = 0, run = 0;
            bool have_start = false;
            auto flush_run = [&]() { if (run) { cs += ':'; cs += std::to_string(run); run = 0; } };
            for (uint32_t i = 0; i < a.n_mappings; i++) {
                const uint32_t v = m[i].node;
                bool uses_node = false;
                for (uint32_t j = 0; j < m[i].n_edits; j++) uses_node |= (e[ei + j] & 3u) != GB_EDIT_INS;
                if (uses_node) {
                    path += (v & 1u) ? '<' : '>'; path += std::to_string(v >> 1);
                    path_len += ix->nodes[v].len;
                    if (!have_start) { pstart = m[i].offset; have_start = true; }
                }
                uint32_t off = m[i].offset;
                for (uint32_t j = 0; j < m[i].n_edits; j++, ei++) {
                    const uint32_t wd = e[ei], op = wd & 3u, len = op == GB_EDIT_SUB ? 1u : wd >> 4;
                    if (op == GB_EDIT_MATCH) { run += len; matches += len; block += len; q += len; off += len; ref_used_total += len; }
                    else if (op == GB_EDIT_SUB) { flush_run(); cs += '*'; cs += node_base(ix, v, off); cs += (char)seq[q]; block += 1; q += 1; off += 1; ref_used_total += 1; }
                    else if (op == GB_EDIT_INS) { flush_run(); cs += '+'; cs.append((const char*)seq + q, len); block += len; q += len; }
                    else { flush_run(); cs += '-'; for (uint32_t t = 0; t < len; t++) cs += node_base(ix, v, off + t); block += len; off += len; ref_used_total += len; }
                }
            }
            flush_run();
            o.num(0); o.ch('\t'); o.num(L); o.str("\t+\t"); o.str(path.empty() ? std::string("*") : path); o.ch('\t'); o.num((long long)path_len); o.ch('\t');
            o.num((long long)pstart); o.ch('\t'); o.num((long long)(pstart + ref_used_total)); o.ch('\t');
            o.num((long long)matches); o.ch('\t'); o.num((long long)block); o.ch('\t'); o.num(a.mapq);
            o.str("\tAS:i:"); o.num(a.score);
            if (quals) { o.str("\tbq:Z:"); for (uint32_t t = 0; t < L; t++) o.ch((char)(quals[read_off[r] + t] + 33)); }
            o.str("\tcs:Z:"); o.str(cs);
            o.str("\tdv:f:"); o.real(block ? 1.0 - (double)matches / (double)block : 0.0);
        }
        if (a.flags & GB_ALN_PAIRED) {
            // mates are interleaved: fragment_next on mate 1, fragment_prev on mate 2 (pair_all, minimizer_mapper.cpp:1280-1300)
            if ((r & 1u) == 0) { o.str("\tfn:Z:"); o.str(read_name(names, name_off, r + 1)); }
            else { o.str("\tfp:Z:"); o.str(read_name(names, name_off, r - 1)); }
        }
        o.ch('\n');
    }
    *out_used = o.n;
    return o.overflow ? GB_ERR_CAPACITY : GB_OK;
}

extern "C" int gb_emit_json(const gb_flat_index* ix, uint32_t n, const gb_alignment* aln, const gb_mapping* mappings, const uint32_t* edits,
                            const uint8_t* reads, const uint8_t* quals, const uint64_t* read_off, const uint8_t* names, const uint64_t* name_off,
                            char* out, uint64_t out_cap, uint64_t* out_used) {
    if (!ix || !aln || !reads || !read_off || !out || !out_used) return GB_ERR_ARG;
    Out o{out, out_cap, 0, false};
    for (uint32_t x = 0; x < n; x++) {
        const gb_alignment& a = aln[x];
        const uint32_t r = a.read_id;
        const uint8_t* seq = reads + read_off[r];
        const uint32_t L = (uint32_t)(read_off[r + 1] - read_off[r]);
        o.str("{\"sequence\": "); json_string(o, std::string((const char*)seq, L));
        uint64_t matches = 0;
        if ((a.flags & GB_ALN_MAPPED) && a.n_mappings) {
            const gb_mapping* m = mappings + a.mapping_off;
            const uint32_t* e = edits + a.edit_off;
            o.str(", \"path\": {\"mapping\": [");
            uint32_t q = 0, ei = 0;
            for (uint32_t i = 0; i < a.n_mappings; i++) {
                if (i) o.str(", ");
                o.str("{\"position\": {\"node_id\": \""); o.num(m[i].node >> 1); o.ch('"');
                if (m[i].offset) { o.str(", \"offset\": \""); o.num(m[i].offset); o.ch('"'); }
                if (m[i].node & 1u) o.str(", \"is_reverse\": true");
                o.str("}, \"edit\": [");
                for (uint32_t j = 0; j < m[i].n_edits; j++, ei++) {
                    const uint32_t wd = e[ei], op = wd & 3u, len = op == GB_EDIT_SUB ? 1u : wd >> 4;
                    if (j) o.str(", ");
                    if (op == GB_EDIT_MATCH) { o.str("{\"from_length\": "); o.num(len); o.str(", \"to_length\": "); o.num(len); o.ch('}'); matches += len; q += len; }
                    else if (op == GB_EDIT_SUB) { o.str("{\"from_length\": 1, \"to_length\": 1, \"sequence\": "); json_string(o, std::string(1, (char)seq[q])); o.ch('}'); q += 1; }
                    else if (op == GB_EDIT_INS) { o.str("{\"to_length\": "); o.num(len); o.str(", \"sequence\": "); json_string(o, std::string((const char*)seq + q, len)); o.ch('}'); q += len; }
                    else { o.str("{\"from_length\": "); o.num(len); o.ch('}'); }
                }
                o.str("], \"rank\": \""); o.num(i + 1); o.str("\"}");
            }
            o.str("]}");
        }
        o.str(", \"name\": "); json_string(o, read_name(names, name_off, r));
        if (quals) { o.str(", \"quality\": \""); base64(o, quals + read_off[r], L); o.ch('"'); }
        if (a.mapq) { o.str(", \"mapping_quality\": "); o.num(a.mapq); }
        if (a.score) { o.str(", \"score\": "); o.num(a.score); }
        if ((a.flags & GB_ALN_MAPPED) && L) { o.str(", \"identity\": "); o.real((double)matches / (double)L); }      // identity(path), alignment.cpp
        if (a.flags & GB_ALN_PAIRED) {
            if ((r & 1u) == 0) { o.str(", \"fragment_next\": {\"name\": "); json_string(o, read_name(names, name_off, r + 1)); o.ch('}'); }
            else { o.str(", \"fragment_prev\": {\"name\": "); json_string(o, read_name(names, name_off, r - 1)); o.ch('}'); }
        }
        o.str(", \"annotation\": {\"mapq_uncapped\": "); o.real(a.mapq_uncapped); o.str(", \"mapq_explored_cap\": ");
        if (a.mapq_explored_cap > 1e30f) o.str("\"Infinity\""); else o.real(a.mapq_explored_cap);
        if (a.flags & GB_ALN_RESCUED) o.str(", \"rescued\": true");
        o.str("}}\n");
    }
    *out_used = o.n;
    return o.overflow ? GB_ERR_CAPACITY : GB_OK;
}
