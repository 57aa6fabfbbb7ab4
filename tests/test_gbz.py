"""gb_index_from_gbz on the GBZ the reference ships as test data (test/primers/y.giraffe.gbz, copied to
tests/golden/gbz/): 66 nodes, 1012 bp, 3 haplotypes.  An independent reader of the published simple-sds / GBWT /
GBWTGraph layout (below, Python) must agree with the library's reader on every array of the flat index, the
distance payload must be consistent with the haplotypes, and reads drawn from the haplotypes must map back."""
import struct
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

GBZ = Path(__file__).parent / "golden" / "gbz" / "y.giraffe.gbz"


def read_gbz(path):
    """-> (node sequences by id order, forward haplotype paths as GBWT nodes, header facts)."""
    raw = path.read_bytes()
    W = struct.unpack("<%dQ" % (len(raw) // 8), raw)
    pos = [0]

    def u():
        pos[0] += 1
        return W[pos[0] - 1]

    def raw_vector():
        bits, n = u(), u()
        return bits, [u() for _ in range(n)]

    def skip_optional():
        n = u()
        pos[0] += n

    def int_vector():
        n, width = u(), u()
        bits, words = raw_vector()
        big = 0
        for k, x in enumerate(words):
            big |= x << (64 * k)
        return [(big >> (k * width)) & ((1 << width) - 1) for k in range(n)], width

    def sparse_values():
        universe, ones = u(), u()
        bits, words = raw_vector()
        skip_optional(); skip_optional(); skip_optional()
        low, width = int_vector()
        ones_at = [p for p in range(bits) if (words[p >> 6] >> (p & 63)) & 1]
        assert len(ones_at) == ones == len(low)
        return [((p - i) << width) | low[i] for i, p in enumerate(ones_at)], universe

    def vector_u8():
        n = u()
        b = raw[pos[0] * 8: pos[0] * 8 + n]
        pos[0] += (n + 7) // 8
        return b

    def string_array():
        starts, _ = sparse_values()
        alphabet = vector_u8()
        ranks, _ = int_vector()
        s = "".join(chr(alphabet[c]) for c in ranks)
        return [s[a:b] for a, b in zip(starts, starts[1:] + [len(s)])]

    h = u(); u()
    assert h & 0xFFFFFFFF == 0x205A4247 and h >> 32 == 1
    gbz_tags = string_array()
    g = u()
    assert g & 0xFFFFFFFF == 0x6B376B37 and g >> 32 == 5
    sequences, size, offset, sigma, flags = u(), u(), u(), u(), u()
    gbwt_tags = string_array()
    starts, universe = sparse_values()
    data = vector_u8()
    assert universe == len(data) and len(starts) == sigma - offset
    skip_optional(); skip_optional()
    gg = u()
    assert gg & 0xFFFFFFFF == 0x6B3764AF and gg >> 32 == 3
    n_nodes = u(); u()
    seqs = string_array()
    assert len(seqs) == n_nodes

    def bytecode(b, i):
        v = s = 0
        while True:
            c = b[i]; i += 1; v |= (c & 0x7F) << s; s += 7
            if not c & 0x80:
                return v, i

    records = []
    for k in range(len(starts)):
        b = data[starts[k]: starts[k + 1] if k + 1 < len(starts) else len(data)]
        edges, runs, i = [], [], 0
        if b:
            deg, i = bytecode(b, i); prev = 0
            for _ in range(deg):
                d, i = bytecode(b, i); o, i = bytecode(b, i); prev += d; edges.append((prev, o))
            cont = 256 // deg if 0 < deg < 255 else 0
            while i < len(b):
                if deg >= 255:
                    v, i = bytecode(b, i); ln, i = bytecode(b, i); ln += 1
                else:
                    c = b[i]; i += 1; v, ln = c % deg, c // deg + 1
                    if ln == cont:
                        extra, i = bytecode(b, i); ln += extra
                runs.append((v, ln))
        records.append((edges, runs))

    def lf(comp, p):
        edges, runs = records[comp]
        seen, at = [0] * len(edges), 0
        for v, ln in runs:
            if p < at + ln:
                return edges[v][0], edges[v][1] + seen[v] + (p - at)
            seen[v] += ln; at += ln
        raise IndexError

    paths, total = [], 0
    for s in range(sequences):
        node, p = lf(0, s); walk = []
        while node:
            walk.append(node); node, p = lf(node - offset, p)
        total += len(walk) + 1
        if s % 2 == 0:
            paths.append(walk)
    return seqs, paths, {"sequences": sequences, "size": size, "total": total, "tags": gbz_tags + gbwt_tags, "records": records, "offset": offset}


def test_independent_reader_agrees_with_the_file_header():
    seqs, paths, facts = read_gbz(GBZ)
    assert facts["sequences"] == 6 and facts["total"] == facts["size"] == 322          # every BWT position is walked exactly once
    assert len(seqs) == 66 and sum(map(len, seqs)) == 1012 and max(map(len, seqs)) <= 32
    assert "jltsiren/gbwtgraph" in facts["tags"] and "reference_samples" in facts["tags"]
    assert len(paths) == 3 and all(v % 2 == 0 for p in paths for v in p)


def test_library_reader_builds_the_same_index_as_the_independent_reader():
    seqs, paths, _ = read_gbz(GBZ)
    from_gbz = capi.HostIndex.from_gbz(GBZ)
    assert (from_gbz.view.n_nodes, from_gbz.view.n_paths, from_gbz.view.k, from_gbz.view.w) == (2 * 67, 3, 29, 11)
    rebuilt = capi.HostIndex(seqs, paths, from_gbz.array("dist").copy(), k=29, w=11)            # same payload, everything else independent
    for name in ("nodes", "seq", "gbwt", "table", "hits"):
        assert from_gbz.array(name).tobytes() == rebuilt.array(name).tobytes(), name
    from_gbz.close(); rebuilt.close()


def test_payload_is_consistent_with_the_haplotypes():
    """The derived chain model on the graph vg built: slots never decrease along a haplotype, the payload distance between
    two nodes of a haplotype is a minimum over all walks (never more than the haplotype's own), neighbours touch, and the
    cut nodes lie on every haplotype."""
    seqs, paths, _ = read_gbz(GBZ)
    index = capi.HostIndex.from_gbz(GBZ)
    dist, slots, table = index.array("dist"), index.array("slots"), index.array("site_dist")
    length = {i + 1: len(s) for i, s in enumerate(seqs)}
    on_all = set.intersection(*[set(v >> 1 for v in p) for p in paths])

    def payload_distance(u, v):
        pu, pv = dist[u], dist[v]
        if int(pu["slot"]) != int(pv["slot"]):
            return int(pv["x_in"]) - int(pu["x_out"])
        sr = slots[int(pu["slot"])]
        t = int(table[int(sr["table_off"]) + int(pu["allele"]) * int(sr["n"]) + int(pv["allele"])])
        assert t != 0xFFFF, "nodes of one haplotype reach each other"
        return t
    n_same_site = 0
    for p in paths:
        ids = [v >> 1 for v in p]
        assert [int(dist[i]["slot"]) for i in ids] == sorted(int(dist[i]["slot"]) for i in ids)          # slots never decrease along a haplotype
        for a in range(len(ids)):
            walked = 0
            for b in range(a + 1, len(ids)):
                d = payload_distance(ids[a], ids[b])
                assert 0 <= d <= walked                                                                 # a minimum over all walks
                if b == a + 1:
                    assert d == 0                                                                       # neighbours on a haplotype touch
                n_same_site += int(dist[ids[a]]["slot"]) == int(dist[ids[b]]["slot"])
                walked += length[ids[b]]
    cut = {i for i in range(1, len(seqs) + 1) if int(dist[i]["allele"]) == 0xFFFF}
    assert cut and cut <= on_all                                                                        # every walk passes a cut node, so every haplotype does
    assert all(int(slots[int(dist[i]["slot"])]["n"]) == 1 for i in cut)
    index.close()


def test_reads_from_the_haplotypes_map_back():
    seqs, paths, _ = read_gbz(GBZ)
    index = capi.HostIndex.from_gbz(GBZ)
    rng = np.random.default_rng(7)
    haps = ["".join(seqs[(v >> 1) - 1] for v in p) for p in paths]
    reads = []
    for _ in range(60):
        h = haps[int(rng.integers(0, len(haps)))]
        s = int(rng.integers(0, len(h) - 150))
        r = np.frombuffer(h[s:s + 150].encode(), dtype=np.uint8).copy()
        if rng.random() < 0.5:
            r = synth.revcomp_bytes(r[None, :])[0]
        reads.append(r)
    reads = np.stack(reads); quals = np.full(reads.shape, 30, dtype=np.uint8)
    res = H.oracle_map(index, reads, quals, threads=4)
    assert (res[0]["flags"] & 1).all() and (res[0]["score"] == 160).all()                              # exact, full length, both bonuses
    index.close()


def test_foreign_and_truncated_files_are_refused(tmp_path):
    raw = GBZ.read_bytes()
    (tmp_path / "short.gbz").write_bytes(raw[:1000])
    (tmp_path / "other.gbz").write_bytes(b"\\x00" * 8 + raw[8:])
    (tmp_path / "odd.gbz").write_bytes(raw[:-3])
    for name in ("short.gbz", "other.gbz", "odd.gbz", "missing.gbz"):
        with pytest.raises(capi.GbError):
            capi.HostIndex.from_gbz(tmp_path / name)


def test_minimizer_table_equals_the_reference_min_file():
    """test/primers/y.min is the gbwtgraph minimizer index vg built for the same GBZ (k = 31, w = 50, 62 unique keys,
    32-byte cells: key, encoded position, 16-byte payload; empty key 2^63 - 1).  gbwtgraph is absent from the reference
    tree, so this file is the ground truth for find_minimizers' definition of a minimizer (hash order, canonical strand,
    windows, position convention id << 11 | is_reverse << 10 | offset): the library's builder must select exactly the
    same k-mers at exactly the same positions."""
    raw = (GBZ.parent / "y.min").read_bytes()
    W = struct.unpack("<%dQ" % (len(raw) // 8), raw[: len(raw) // 8 * 8])
    assert W[0] & 0xFFFFFFFF == 0x31513151
    k, w, n_keys, capacity = W[1], W[2], W[3], W[9]
    assert (k, w, n_keys, capacity) == (31, 50, 62, 1024) and len(raw) == 8 * 10 + 32 * capacity + (len(raw) - 8 * 10 - 32 * capacity)
    cells = {}
    for c in range(capacity):
        key, pos = W[10 + 4 * c], W[11 + 4 * c]
        if key != 0x7FFFFFFFFFFFFFFF:
            cells[key] = pos
    assert len(cells) == n_keys
    index = capi.HostIndex.from_gbz(GBZ, k=k, w=w)
    table, hits = index.array("table"), index.array("hits")
    mine = {int(c["key"]): sorted(int(hits[int(c["hit_off"]) + i]["pos"]) for i in range(int(c["hit_cnt"])))
            for c in table if int(c["key"]) != 0xFFFFFFFFFFFFFFFF}
    assert set(mine) == set(cells)
    assert all(mine[key] == [pos] for key, pos in cells.items())
    index.close()


def test_distance_payload_equals_vgs_prefix_sums():
    """The 16-byte payload of y.min's cells is vg's zipcode (zip_code.cpp:1943-2010: byte count, the zipcode's varints,
    then decoder offsets).  For a node of the top-level chain it reads [1, component, chain component, connectivity]
    (zip_code.cpp:57-100) + [prefix sum + 1, length + 1, is_reversed, chain component] (:105-111): the prefix sum is the
    chain coordinate vg's distance index gives the node.  The payload derived from the GBZ's haplotypes must give every
    such node the same coordinate (x_in), i.e. the two distance models agree on this graph."""
    raw = (GBZ.parent / "y.min").read_bytes()
    W = struct.unpack("<%dQ" % (len(raw) // 8), raw[: len(raw) // 8 * 8])
    index = capi.HostIndex.from_gbz(GBZ, k=31, w=50)
    dist, nodes = index.array("dist"), index.array("nodes")

    def varints(bs):
        out, v, s = [], 0, 0
        for b in bs:
            v |= (b & 0x7F) << s; s += 7
            if not b & 0x80:
                out.append(v); v = s = 0
        return out

    checked = set()
    for c in range(1024):
        key, pos, p0, p1 = W[10 + 4 * c: 14 + 4 * c]
        if key == 0x7FFFFFFFFFFFFFFF:
            continue
        b = list(p0.to_bytes(8, "little") + p1.to_bytes(8, "little"))
        zc = varints(b[1: 1 + b[0]])
        if len(zc) != 8 or zc[0] != 1:
            continue                              # nodes inside bubbles carry longer codes (or oversized ones, byte count 0)
        nid = pos >> 11
        prefix_sum, length = zc[4] - 1, zc[5] - 1
        assert length == int(nodes[2 * nid]["len"])
        assert prefix_sum == int(dist[nid]["x_in"]) and prefix_sum + length == int(dist[nid]["x_out"])
        assert int(dist[nid]["allele"]) == 0xFFFF and zc[6] == 0
        checked.add(nid)
    assert len(checked) >= 20                     # most minimizers of this graph sit on backbone nodes

    # nodes inside sites: their zipcodes do not fit 15 bytes and live in y.zipcodes ("SPIZ", version, then per zipcode
    # a varint byte count, the zipcode, its decoder; zip_code.cpp:2111-2170), the cell's payload being (0, index).
    # A site is a snarl code after the root chain: [is_regular, offset in chain + 1, minimum length + 1, ...] (:128-147);
    # the offset is the coordinate right after the site's start node, i.e. x_in of its alleles.
    zraw = (GBZ.parent / "y.zipcodes").read_bytes()
    assert zraw[:4] == b"SPIZ"
    codes, i = [], 8

    def varint_at(j):
        v = sh = 0
        while True:
            c = zraw[j]; j += 1; v |= (c & 0x7F) << sh; sh += 7
            if not c & 0x80:
                return v, j

    while i < len(zraw):
        n, i = varint_at(i)
        codes.append(varints(zraw[i: i + n])); i += n
        dn, i = varint_at(i); i += dn                          # the decoder (is_chain, offset pairs): not needed here
    in_sites = 0
    for c in range(1024):
        key, pos, p0, p1 = W[10 + 4 * c: 14 + 4 * c]
        if key == 0x7FFFFFFFFFFFFFFF or (p0 & 0xFF) != 0:
            continue
        zc = codes[p1]; nid = pos >> 11
        assert zc[:4] == [1, 0, 0, 0]
        assert zc[5] - 1 == int(dist[nid]["x_in"]) and int(dist[nid]["allele"]) != 0xFFFF          # site offset = x_in of its alleles
        if zc[4] == 1:                                                                              # regular snarl: one site
            assert zc[6] - 1 == int(dist[nid]["x_out"]) - int(dist[nid]["x_in"])                    # minimum length across the site
            assert zc[11] - 1 == int(nodes[2 * nid]["len"])                                         # the child chain is the allele node
        else:                                                                                       # irregular snarl: adjacent sites, one slot each here
            assert zc[6] - 1 >= int(dist[nid]["x_out"]) - int(dist[nid]["x_in"]) and zc[15] - 1 == int(nodes[2 * nid]["len"])
        in_sites += 1
    assert in_sites == 2
    index.close()


def test_flat_gbwt_records_equal_the_files_gbwt_records():
    """The flat index keeps its own GBWT (built from the haplotype paths, both orientations).  Decoded record by record
    it must be the GBWT of the file: the same (successor, offset) edges and the same body — the sequence of successor
    ranks in BWT order — for every oriented node.  gbwt is absent from the reference tree; this file is its output."""
    seqs, paths, facts = read_gbz(GBZ)
    index = capi.HostIndex.from_gbz(GBZ)
    nodes, gbwt = index.array("nodes"), index.array("gbwt")
    compared = 0
    for comp in range(1, len(facts["records"])):
        v = comp + facts["offset"]
        edges, runs = facts["records"][comp]
        rec = gbwt[int(nodes[v]["rec_off"]):]
        n_edges, n_runs = int(rec[0]), int(rec[1])
        mine_edges = [(int(rec[2 + 2 * e]), int(rec[3 + 2 * e])) for e in range(n_edges)]
        mine_body = [w & 1023 for w in map(int, rec[2 + 2 * n_edges: 2 + 2 * n_edges + n_runs]) for _ in range(w >> 10)]
        body = [r for r, ln in runs for _ in range(ln)]
        assert mine_edges == edges and mine_body == body, v
        assert int(nodes[v]["size"]) == len(body)
        compared += 1
    assert compared == 132
    index.close()


def test_oracle_minimizer_regions_find_exactly_the_keys_of_the_reference_min_file():
    """The oracle's own (brute force over windows) definition of a minimizer, run over the three haplotype sequences of
    the GBZ with k = 31, w = 50, must select the k-mers of y.min and no others, and each at the haplotype position the
    file's graph position corresponds to."""
    import ctypes as C
    seqs, paths, _ = read_gbz(GBZ)
    raw = (GBZ.parent / "y.min").read_bytes()
    W = struct.unpack("<%dQ" % (len(raw) // 8), raw[: len(raw) // 8 * 8])
    file_cells = {W[10 + 4 * c]: W[11 + 4 * c] for c in range(1024) if W[10 + 4 * c] != 0x7FFFFFFFFFFFFFFF}
    lib = H.oracle_lib()
    lib.oracle_minimizer_regions.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
    lib.oracle_minimizer_regions.restype = C.c_uint32
    found = {}
    for p in paths:
        hap = "".join(seqs[(v >> 1) - 1] for v in p)
        starts = np.cumsum([0] + [len(seqs[(v >> 1) - 1]) for v in p])
        s = np.frombuffer(hap.encode(), dtype=np.uint8).copy()
        cap = 4096
        key = np.zeros(cap, dtype=np.uint64); fwd = np.zeros(cap, dtype=np.uint32); rev = np.zeros(cap, dtype=np.uint8)
        a0 = np.zeros(cap, dtype=np.uint32); a1 = np.zeros(cap, dtype=np.uint32)
        n = lib.oracle_minimizer_regions(capi.ptr(s), len(s), 31, 50, cap, capi.ptr(key), capi.ptr(fwd), capi.ptr(rev), capi.ptr(a0), capi.ptr(a1))
        assert 0 < n <= cap
        for i in range(n):
            # graph position of the minimizer: first base of the k-mer on its strand (forward: offset; reverse: last base, other strand)
            at = int(fwd[i]) + (30 if rev[i] else 0)
            j = int(np.searchsorted(starts, at, side="right") - 1)
            nid, off = p[j] >> 1, at - int(starts[j])
            if rev[i]:
                off = len(seqs[nid - 1]) - 1 - off
            found.setdefault(int(key[i]), set()).add((nid << 11) | (int(rev[i]) << 10) | off)
    assert set(found) == set(file_cells)
    assert all(found[k] == {pos} for k, pos in file_cells.items())


@pytest.mark.gpu
def test_cuda_path_equals_oracle_on_the_reference_gbz_graph():
    """Single-end and paired mapping (rescue on) on the graph vg built: CUDA path vs oracle."""
    seqs, paths, _ = read_gbz(GBZ)
    index = capi.HostIndex.from_gbz(GBZ)
    haps = ["".join(seqs[(v >> 1) - 1] for v in p) for p in paths]
    rng = np.random.default_rng(11)
    reads = []
    for _ in range(300):
        h = haps[int(rng.integers(0, len(haps)))]
        s = int(rng.integers(0, len(h) - 150))
        r = np.frombuffer(h[s:s + 150].encode(), dtype=np.uint8).copy()
        m = rng.random(150) < 0.02
        r[m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
        reads.append(synth.revcomp_bytes(r[None, :])[0] if rng.random() < 0.5 else r)
    reads = np.stack(reads); quals = np.full(reads.shape, 30, dtype=np.uint8)
    dev = capi.Device(index)
    assert not H.compare_alignments(H.gpu_map(dev, reads, quals), H.oracle_map(index, reads, quals, threads=8), len(reads))
    pairs = []
    for _ in range(150):
        h = haps[int(rng.integers(0, len(haps)))]
        frag = int(np.clip(rng.normal(400, 50), 160, len(h) - 1)); s = int(rng.integers(0, len(h) - frag))
        a = np.frombuffer(h[s:s + 150].encode(), dtype=np.uint8).copy()
        b = synth.revcomp_bytes(np.frombuffer(h[s + frag - 150:s + frag].encode(), dtype=np.uint8).copy()[None, :])[0]
        pairs += [a, b] if rng.random() < 0.5 else [b, a]
    pairs = np.stack(pairs); pq = np.full(pairs.shape, 30, dtype=np.uint8)
    p = H.paired_params(); p.max_rescue_attempts = 15
    assert not H.compare_alignments(H.gpu_map(dev, pairs, pq, p, paired=True), H.oracle_map_paired(index, pairs, pq, p, threads=8), len(pairs))
    dev.close(); index.close()


# ---- the .min / .zipcodes files giraffe loads beside the GBZ (giraffe_main.cpp:1825-1881) -----------------------------------

def _table_as_dict(index):
    table, hits = index.array("table"), index.array("hits")
    return {int(c["key"]): sorted((int(hits[int(c["hit_off"]) + i]["pos"]), bytes(hits[int(c["hit_off"]) + i]["payload"].tobytes()))
                                  for i in range(int(c["hit_cnt"])))
            for c in table if int(c["key"]) != 0xFFFFFFFFFFFFFFFF}


def test_index_from_the_reference_min_file_equals_the_rederived_one():
    """gb_index_from_gbz_min takes k, w, keys and positions from the reference's y.min (and checks y.zipcodes holds the
    oversized codes it points at); the result must be the index gb_index_from_gbz derives by scanning the haplotypes:
    same table, same hits, same distance payload, and the oracle maps reads identically on both."""
    d = GBZ.parent
    a = capi.HostIndex.from_gbz_min(GBZ, d / "y.min", d / "y.zipcodes")
    b = capi.HostIndex.from_gbz(GBZ, k=31, w=50)
    assert (a.k, a.w) == (31, 50)
    assert _table_as_dict(a) == _table_as_dict(b) and len(_table_as_dict(a)) == 62
    for name in ("nodes", "gbwt", "dist", "slots", "site_dist"):
        assert a.array(name).tobytes() == b.array(name).tobytes(), name
    seqs, paths, _ = read_gbz(GBZ)
    hap = "".join(seqs[(v >> 1) - 1] for v in paths[0])
    reads = np.stack([np.frombuffer(hap[s:s + 150].encode(), dtype=np.uint8) for s in range(0, len(hap) - 150, 37)])
    quals = np.full(reads.shape, 30, dtype=np.uint8)
    ra, rb = H.oracle_map(a, reads, quals, threads=2), H.oracle_map(b, reads, quals, threads=2)
    assert ra[0].tobytes() == rb[0].tobytes() and (ra[0]["flags"] & 1).all()
    a.close(); b.close()


def test_min_files_that_do_not_fit_are_refused(tmp_path):
    d = GBZ.parent
    raw = bytearray((d / "y.min").read_bytes())
    W = lambda i: struct.unpack_from("<Q", raw, 8 * i)[0]

    def variant(name, edit):
        b = bytearray(raw); edit(b); p = tmp_path / name; p.write_bytes(bytes(b)); return p

    first = next(c for c in range(1024) if W(10 + 4 * c) != 0x7FFFFFFFFFFFFFFF)
    cases = {
        "tag.min": lambda b: struct.pack_into("<I", b, 0, 0x12345678),
        "version.min": lambda b: struct.pack_into("<I", b, 4, 9),
        "syncmers.min": lambda b: struct.pack_into("<Q", b, 64, W(8) | 0x100),
        "key128.min": lambda b: struct.pack_into("<Q", b, 64, 128),
        "otherflag.min": lambda b: struct.pack_into("<Q", b, 64, W(8) | 0x400),
        "pointer.min": lambda b: struct.pack_into("<Q", b, 8 * (10 + 4 * first), W(10 + 4 * first) | (1 << 63)),
        "otherkmer.min": lambda b: struct.pack_into("<Q", b, 8 * (10 + 4 * first), W(10 + 4 * first) ^ (3 << 60)),   # first base of the key changed
        "offgraph.min": lambda b: struct.pack_into("<Q", b, 8 * (11 + 4 * first), (5000 << 11)),
        "pastnode.min": lambda b: struct.pack_into("<Q", b, 8 * (11 + 4 * first), (W(11 + 4 * first) & ~1023) | 900),
        "short.min": lambda b: b.__delitem__(slice(len(b) - 16, len(b))),
    }
    for name, edit in cases.items():
        with pytest.raises(capi.GbError):
            capi.HostIndex.from_gbz_min(GBZ, variant(name, edit), d / "y.zipcodes")
    # a .zipcodes file that lacks the oversized codes the table points at, or is not one
    z = (d / "y.zipcodes").read_bytes()
    (tmp_path / "empty.zipcodes").write_bytes(z[:8])
    (tmp_path / "foreign.zipcodes").write_bytes(b"NOPE" + z[4:])
    for name in ("empty.zipcodes", "foreign.zipcodes", "missing.zipcodes"):
        with pytest.raises(capi.GbError):
            capi.HostIndex.from_gbz_min(GBZ, d / "y.min", tmp_path / name)
    capi.HostIndex.from_gbz_min(GBZ, d / "y.min").close()                                     # the zipcode file is optional
    # a .min that says it holds keys with several occurrences: its table is not read, k and w are, the minimizers are re-derived
    multi = variant("multi.min", lambda b: struct.pack_into("<Q", b, 48, W(6) + 1))
    a = capi.HostIndex.from_gbz_min(GBZ, multi, d / "y.zipcodes"); b = capi.HostIndex.from_gbz(GBZ, k=31, w=50)
    assert (a.k, a.w) == (31, 50)
    for name in ("nodes", "gbwt", "dist", "table", "hits"):
        assert a.array(name).tobytes() == b.array(name).tobytes(), name
    a.close(); b.close()


def test_build_with_hits_equals_the_haplotype_scan():
    """gb_index_build_with_hits on a graph with repeated k-mers: handing the builder the (key, position) pairs it would
    find itself (several occurrences per key here, unlike y.min) gives the identical index."""
    g = synth.make_variant_graph(length=6000, n_snp=20, n_ins=3, n_del=3, n_haps=4, seed=9)
    ref = g.build_index()
    table, hits = ref.array("table"), ref.array("hits")
    keys, pos = [], []
    for c in table:
        if int(c["key"]) == 0xFFFFFFFFFFFFFFFF:
            continue
        for i in range(int(c["hit_cnt"])):
            keys.append(int(c["key"])); pos.append(int(hits[int(c["hit_off"]) + i]["pos"]))
    order = np.random.default_rng(3).permutation(len(keys))
    got = capi.HostIndex(g.node_seqs, g.paths, g.dist, k=ref.k, w=ref.w,
                         hits=(np.asarray(keys, dtype=np.uint64)[order], np.asarray(pos, dtype=np.uint64)[order]))
    assert _table_as_dict(got) == _table_as_dict(ref)
    for name in ("nodes", "gbwt", "dist", "table", "hits"):
        assert got.array(name).tobytes() == ref.array(name).tobytes(), name
    got.close(); ref.close()


@pytest.mark.parametrize("which", ["variants", "repeats", "nested", "branchy", "reference gbz"])
def test_window_enumeration_builds_the_same_index_as_the_haplotype_scan(which, monkeypatch):
    """The builder's two routes to the minimizer table — every haplotype end to end, or every haplotype-consistent
    (k + w - 1)-window once by following GBWT search states (what gbwtgraph's index_haplotypes does; chosen when there are
    more than 32 haplotypes) — must give byte-identical indexes."""
    def build():
        if which == "reference gbz":
            return capi.HostIndex.from_gbz(GBZ, k=31, w=50)
        g = {"variants": lambda: synth.make_variant_graph(length=40000, n_snp=120, n_ins=15, n_del=15, n_haps=8, seed=4),
             "repeats": lambda: synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6),
             "nested": lambda: synth.make_nested_graph(n_items=120, n_haps=10, seed=5),
             "branchy": lambda: synth.make_branchy_graph(n_layers=600, n_haps=16, seed=4)}[which]()
        return g.build_index() if which != "nested" else g.build_index(k=11, w=5)
    monkeypatch.setenv("GIRAFFE_B200_WINDOW_BUILDER", "0"); a = build()
    monkeypatch.setenv("GIRAFFE_B200_WINDOW_BUILDER", "1"); b = build()
    assert int(a.view.n_hits) == int(b.view.n_hits) and int(a.view.n_hits) > 50
    for name in ("nodes", "gbwt", "dist", "table", "hits"):
        assert a.array(name).tobytes() == b.array(name).tobytes(), name
    a.close(); b.close()


@pytest.mark.parametrize("which", ["variants", "repeats", "nested", "branchy", "reference gbz"])
def test_index_from_the_flat_gbwt_equals_the_index_from_the_paths(which):
    """gb_index_build_from_gbwt: node sequences + GBWT records, no haplotype paths — distance model from the edges of the
    forward records, minimizers by window enumeration, record blobs copied after validation.  Feeding it the flat GBWT of a
    path-built index must give that index back byte for byte (hand-made payloads are passed through, derived ones re-derived)."""
    if which == "reference gbz":
        seqs, paths, _ = read_gbz(GBZ)
        ref = capi.HostIndex(seqs, paths, k=31, w=50); dist = None; k, w = 31, 50
    else:
        g = {"variants": lambda: synth.make_variant_graph(length=40000, n_snp=120, n_ins=15, n_del=15, n_haps=8, seed=4),
             "repeats": lambda: synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6),
             "nested": lambda: synth.make_nested_graph(n_items=120, n_haps=10, seed=5),
             "branchy": lambda: synth.make_branchy_graph(n_layers=600, n_haps=16, seed=4)}[which]()
        k, w = (11, 5) if which == "nested" else (29, 11)
        ref = g.build_index(k=k, w=w); seqs, paths, dist = g.node_seqs, g.paths, g.dist
    nodes = ref.array("nodes")
    got = capi.HostIndex.from_gbwt(seqs, len(paths), ref.array("gbwt"), nodes["rec_off"], dist, k=k, w=w)
    assert int(got.view.n_paths) == len(paths)
    for name in ("nodes", "seq", "gbwt", "dist", "slots", "site_dist", "table", "hits"):
        assert got.array(name).tobytes() == ref.array(name).tobytes(), name
    got.close(); ref.close()


def test_damaged_gbwt_records_are_refused():
    g = synth.make_variant_graph(length=6000, n_snp=20, n_ins=3, n_del=3, n_haps=4, seed=9)
    ref = g.build_index()
    words, rec_off = ref.array("gbwt").copy(), ref.array("nodes")["rec_off"].copy()
    v = int(np.flatnonzero(rec_off)[5]); off = int(rec_off[v])
    n_edges = int(words[off])

    def attempt(mut):
        w2, r2 = words.copy(), rec_off.copy(); mut(w2, r2)
        with pytest.raises(capi.GbError):
            capi.HostIndex.from_gbwt(g.node_seqs, len(g.paths), w2, r2, g.dist)
    attempt(lambda w2, r2: r2.__setitem__(v, len(w2) + 5))                                        # record outside the blob
    attempt(lambda w2, r2: w2.__setitem__(off, 5000))                                             # absurd edge count
    attempt(lambda w2, r2: w2.__setitem__(off + 2, 2 * (len(g.node_seqs) + 9)))                   # successor outside the graph
    attempt(lambda w2, r2: w2.__setitem__(off + 3, 1 << 30))                                      # offset outside the successor's record
    attempt(lambda w2, r2: w2.__setitem__(off + 2 + 2 * n_edges, (1 << 10) | 900))                # run rank outside the edge list
    ref.close()


def test_unused_low_node_ids_are_allowed():
    """A graph whose node ids do not start at 1 (a GBWT with an alphabet offset: ids below the first stay unused): the same
    graph with every id shifted by 100 builds (from paths and from its flat GBWT), keeps the original ids in its outputs, and
    the oracle maps reads to the shifted nodes exactly as to the unshifted ones; a path or record naming an unused id is refused."""
    shift = 100
    g = synth.make_variant_graph(length=20000, n_snp=40, n_ins=5, n_del=5, n_haps=4, seed=12)
    ref = g.build_index()
    seqs = [""] * shift + list(g.node_seqs)
    paths = [[v + 2 * shift for v in p] for p in g.paths]
    dist = np.zeros(len(seqs) + 1, dtype=capi.dist_dt); dist["allele"] = 0xFFFF
    dist[shift + 1:] = np.asarray(g.dist)[1:]
    shifted = capi.HostIndex(seqs, paths, dist)
    assert int(shifted.view.n_nodes) == int(ref.view.n_nodes) + 2 * shift
    assert (shifted.array("nodes")["len"][: 2 * shift + 2] == 0).all()
    assert shifted.array("nodes")["len"][2 * shift + 2:].tobytes() == ref.array("nodes")["len"][2:].tobytes()
    again = capi.HostIndex.from_gbwt(seqs, len(paths), shifted.array("gbwt"), shifted.array("nodes")["rec_off"], dist)
    for name in ("nodes", "seq", "gbwt", "table", "hits"):
        assert again.array(name).tobytes() == shifted.array(name).tobytes(), name
    rs = synth.simulate_reads(g, 300, length=150, sub_rate=0.01, seed=4)
    a = H.oracle_map(ref, rs.reads, rs.quals, threads=4)
    b = H.oracle_map(shifted, rs.reads, rs.quals, threads=4)
    assert a[0]["score"].tobytes() == b[0]["score"].tobytes() and a[0]["mapq"].tobytes() == b[0]["mapq"].tobytes()
    for i in range(rs.n):
        pa, pb = H.decode_alignment(a[0][i], a[1], a[2])[2], H.decode_alignment(b[0][i], b[1], b[2])[2]
        assert [(n + 2 * shift, o, e) for n, o, e in pa] == pb
    with pytest.raises(capi.GbError):
        capi.HostIndex(seqs, [[2 * 5] + paths[0]], dist)                                     # a path through an unused id
    bad_off = shifted.array("nodes")["rec_off"].copy(); bad_off[2 * 7] = bad_off[2 * shift + 2]
    with pytest.raises(capi.GbError):
        capi.HostIndex.from_gbwt(seqs, len(paths), shifted.array("gbwt"), bad_off, dist)      # a record on an unused id
    for x in (ref, shifted, again):
        x.close()
