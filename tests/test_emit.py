"""gb_emit_gaf / gb_emit_json: host-side text emission of alignment records (the AlignmentEmitter stand-in,
giraffe_main.cpp:2209-2226).  libvgio is absent from the reference tree, so these tests check the text against
the records and the graph: replaying a GAF line's cs string along its path must reproduce the aligned part of the
read, and a JSON line must decode to the same path / edits as the binary record (vg.proto field names)."""
import base64
import json
import re

import numpy as np

import helpers as H
from vg_b200 import capi, synth

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def _oriented_seq(g, v):
    s = g.node_seqs[(v >> 1) - 1]
    return "".join(COMP[c] for c in reversed(s)) if v & 1 else s


def _records(paired):
    g = synth.make_variant_graph(length=60000, n_snp=100, n_ins=12, n_del=12, n_haps=4, seed=8)
    index = g.build_index()
    if paired:
        rs = synth.simulate_pairs(g, 150, sub_rate=0.02, seed=5)
    else:
        rs = synth.simulate_reads(g, 300, length=150, sub_rate=0.02, ins_rate=0.004, del_rate=0.004, seed=6)
    rng = np.random.default_rng(1)
    for i in rng.integers(0, rs.n, size=20):                   # soft clips: garbage at one end
        rs.reads[i, :12] = synth.BASES[rng.integers(0, 4, size=12)]
    rs.reads[7] = synth.BASES[rng.integers(0, 4, size=rs.length)]      # an unmapped read
    if paired:
        res = H.oracle_map_paired(index, rs.reads, rs.quals, H.paired_params(), threads=8)
    else:
        res = H.oracle_map(index, rs.reads, rs.quals, threads=8)
    return g, index, rs, res


def _check_gaf(g, rs, res, text, paired, names):
    lines = text.splitlines()
    assert len(lines) == rs.n
    n_clipped = n_indel = 0
    for i, line in enumerate(lines):
        f = line.split("\t")
        read = bytes(rs.reads[i]).decode()
        assert f[0] == names[i] and int(f[1]) == rs.length
        assert (int(f[2]), int(f[3])) == (0, rs.length)                  # the whole read, soft clips included
        score, mapq, path = H.decode_alignment(res[0][i], res[1], res[2])
        tags = {t[:2]: t[5:] for t in f[12:]}
        if paired:
            assert tags["fn" if i % 2 == 0 else "fp"] == names[i ^ 1]
        if not path:
            assert f[4:12] == ["*"] * 7 + ["255"] and tags["cs"] == "+" + read
            continue
        assert f[4] == "+" and int(f[11]) == mapq and int(tags["AS"]) == score
        steps = [(int(x[1:]) << 1) | (x[0] == "<") for x in re.findall(r"[<>]\d+", f[5])]
        assert steps == [m[0] for m in path if any(e[0] != "I" for e in m[2])]
        ref = "".join(_oriented_seq(g, v) for v in steps)
        assert int(f[6]) == len(ref)
        pos, matches, block = int(f[7]), 0, 0
        query = []
        tokens = [(m[0] or m[2] or m[4] or m[6], m[1] or m[3] or m[5] or m[7])
                  for m in re.findall(r"(:)(\d+)|(\*)([A-Z]{2})|(\+)([A-Z]+)|(-)([A-Z]+)", tags["cs"])]
        assert "".join(op + arg for op, arg in tokens) == tags["cs"]
        assert not any(a[0] == ":" and b[0] == ":" for a, b in zip(tokens, tokens[1:]))     # match runs are merged
        for op, arg in tokens:
            if op == ":":
                n = int(arg); query.append(ref[pos:pos + n]); pos += n; matches += n; block += n
            elif op == "*":
                assert ref[pos] == arg[0] and arg[0] != arg[1]
                query.append(arg[1]); pos += 1; block += 1
            elif op == "+":
                query.append(arg); block += len(arg); n_indel += 1
            else:
                assert ref[pos:pos + len(arg)] == arg
                pos += len(arg); block += len(arg); n_indel += 1
        assert "".join(query) == read, (i, line)
        assert pos == int(f[8]) and matches == int(f[9]) and block == int(f[10])
        assert tags["bq"] == "".join(chr(int(c) + 33) for c in rs.quals[i])
        assert abs(float(tags["dv"]) - (1 - matches / block)) < 1e-5
        n_clipped += tokens[0][0] == "+" or tokens[-1][0] == "+"
    return n_clipped, n_indel


def _hand_record(mappings):
    """mappings: [(oriented node, offset, [(op, length, base2)...])] -> (aln[1], maps, edits) as the library lays them out."""
    aln = np.zeros(1, dtype=capi.alignment_dt)
    maps = np.zeros(len(mappings), dtype=capi.mapping_dt)
    words = []
    for i, (node, off, eds) in enumerate(mappings):
        maps[i]["node"] = node; maps[i]["offset"] = off; maps[i]["n_edits"] = len(eds)
        for op, ln, base in eds:
            words.append((ln << 4) | (base << 2) | "MSID".index(op))
    aln[0]["read_id"] = 0; aln[0]["flags"] = capi.GB_ALN_MAPPED if mappings else 0
    aln[0]["n_mappings"] = len(mappings); aln[0]["n_edits"] = len(words); aln[0]["score"] = 7; aln[0]["mapq"] = 60
    return aln, maps if len(maps) else np.zeros(1, dtype=capi.mapping_dt), np.array(words + [0], dtype=np.uint32)


def test_gaf_reference_vector_unused_final_node_is_removed():
    """unittest/alignment.cpp:398-470: GATTACA -> CAT -> GATTA, read TACACTTAC = TACA on 1:3, C *AT T on 2:0, AC soft-clipped
    onto node 3: query 0..9, path >1>2 of length 10, start 3, end 10, cs ":5*AT:1+AC"."""
    index = capi.HostIndex(["GATTACA", "CAT", "GATTA"], [[2, 4, 6]], None, k=5, w=3)
    aln, maps, edits = _hand_record([(2, 3, [("M", 4, 0)]), (4, 0, [("M", 1, 0), ("S", 1, 3), ("M", 1, 0)]), (6, 0, [("I", 2, 0)])])
    rbuf, qbuf, read_off = H.pack_reads(np.frombuffer(b"TACACTTAC", dtype=np.uint8).reshape(1, -1).copy(), None)
    f = capi.emit_text("gaf", index.view, aln, maps, edits, rbuf, None, read_off, ["softclip-at-end"]).rstrip("\n").split("\t")
    assert f[:9] == ["softclip-at-end", "9", "0", "9", "+", ">1>2", "10", "3", "10"]
    assert dict(t.split(":", 2)[::2] for t in f[12:])["cs"] == ":5*AT:1+AC"


def test_gaf_reference_vector_unaligned_read():
    """unittest/alignment.cpp:793-820: no path, query 0..9, cs "+TACACTTAC"."""
    index = capi.HostIndex(["GATTACA", "CAT", "GATTA"], [[2, 4, 6]], None, k=5, w=3)
    aln, maps, edits = _hand_record([])
    rbuf, qbuf, read_off = H.pack_reads(np.frombuffer(b"TACACTTAC", dtype=np.uint8).reshape(1, -1).copy(), None)
    f = capi.emit_text("gaf", index.view, aln, maps, edits, rbuf, None, read_off, ["unaligned"]).rstrip("\n").split("\t")
    assert f[:4] == ["unaligned", "9", "0", "9"] and f[5] == "*"
    assert dict(t.split(":", 2)[::2] for t in f[12:])["cs"] == "+TACACTTAC"


def test_gaf_lines_replay_to_the_reads_single_end():
    g, index, rs, res = _records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    names = [f"frag{i}" for i in range(rs.n)]
    text = capi.emit_text("gaf", index.view, res[0], res[1], res[2], rbuf, qbuf, read_off, names)
    n_clipped, n_indel = _check_gaf(g, rs, res, text, False, names)
    assert n_clipped >= 5 and n_indel >= 20          # soft clips and gapped alignments were exercised


def test_gaf_lines_replay_to_the_reads_paired():
    g, index, rs, res = _records(paired=True)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    names = [f"pair{i // 2}/{i % 2 + 1}" for i in range(rs.n)]
    text = capi.emit_text("gaf", index.view, res[0], res[1], res[2], rbuf, qbuf, read_off, names)
    _check_gaf(g, rs, res, text, True, names)
    # default names, a subset of the records in another order
    sub = res[0][[5, 2, 9]]
    lines = capi.emit_text("gaf", index.view, sub, res[1], res[2], rbuf, qbuf, read_off).splitlines()
    assert [l.split("\t")[0] for l in lines] == ["read5", "read2", "read9"]


def test_json_lines_decode_to_the_records():
    g, index, rs, res = _records(paired=True)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    names = [f"pair{i // 2}/{i % 2 + 1}" for i in range(rs.n)]
    text = capi.emit_text("json", index.view, res[0], res[1], res[2], rbuf, qbuf, read_off, names)
    lines = text.splitlines()
    assert len(lines) == rs.n
    for i, line in enumerate(lines):
        d = json.loads(line)
        read = bytes(rs.reads[i]).decode()
        score, mapq, path = H.decode_alignment(res[0][i], res[1], res[2])
        assert d["sequence"] == read and d["name"] == names[i]
        assert base64.b64decode(d["quality"]) == bytes(rs.quals[i])
        assert d.get("score", 0) == score and d.get("mapping_quality", 0) == mapq
        assert d["fragment_next" if i % 2 == 0 else "fragment_prev"]["name"] == names[i ^ 1]
        if not path:
            assert "path" not in d
            continue
        q = 0
        assert len(d["path"]["mapping"]) == len(path)
        for rank, (mj, (node, offset, edits)) in enumerate(zip(d["path"]["mapping"], path), 1):
            pos = mj["position"]
            assert int(pos["node_id"]) == node >> 1 and int(pos.get("offset", "0")) == offset and pos.get("is_reverse", False) == bool(node & 1)
            assert mj["rank"] == str(rank) and len(mj["edit"]) == len(edits)
            for ej, (op, length, base) in zip(mj["edit"], edits):
                fl, tl = ej.get("from_length", 0), ej.get("to_length", 0)
                if op == "M":
                    assert (fl, tl) == (length, length) and "sequence" not in ej
                elif op == "S":
                    assert (fl, tl) == (1, 1) and ej["sequence"] == read[q]
                elif op == "I":
                    assert (fl, tl) == (0, length) and ej["sequence"] == read[q:q + length]
                else:
                    assert (fl, tl) == (length, 0)
                q += tl
        assert q == rs.length
        assert 0.0 < d["identity"] <= 1.0


def test_emit_reports_a_short_buffer():
    import ctypes as C
    g, index, rs, res = _records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    out = np.zeros(64, dtype=np.uint8); used = C.c_uint64()
    lib = capi.load_library()
    args = lambda aln, n_maps, n_edits, n_reads: (C.byref(index.view), len(aln), capi.ptr(aln), capi.ptr(res[1]), n_maps, capi.ptr(res[2]), n_edits, n_reads,
                                                  capi.ptr(rbuf), capi.ptr(qbuf), capi.ptr(read_off), None, None, capi.ptr(out), 64, C.byref(used))
    ten = np.ascontiguousarray(res[0][:10])
    assert lib.gb_emit_gaf(*args(ten, len(res[1]), len(res[2]), rs.n)) == capi.GB_ERR_CAPACITY


def test_emit_refuses_records_that_point_outside_their_pools():
    import ctypes as C
    g, index, rs, res = _records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    lib = capi.load_library()
    out = np.zeros(1 << 20, dtype=np.uint8); used = C.c_uint64()

    def emit(fn, aln, n_maps=len(res[1]), n_edits=len(res[2]), n_reads=rs.n):
        aln = np.ascontiguousarray(aln)
        return fn(C.byref(index.view), len(aln), capi.ptr(aln), capi.ptr(res[1]), n_maps, capi.ptr(res[2]), n_edits, n_reads, capi.ptr(rbuf), capi.ptr(qbuf),
                  capi.ptr(read_off), None, None, capi.ptr(out), len(out), C.byref(used))

    mapped = res[0][(res[0]["flags"] & 1) != 0][:20]
    for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
        assert emit(fn, mapped) == capi.GB_OK
        assert emit(fn, mapped, n_maps=int(mapped["mapping_off"].max())) == capi.GB_ERR_ARG                  # mapping pool shorter than a record needs
        assert emit(fn, mapped, n_edits=int(mapped["edit_off"].max())) == capi.GB_ERR_ARG
        assert emit(fn, mapped, n_reads=int(mapped["read_id"].max())) == capi.GB_ERR_ARG                     # read id outside the batch
        bad = mapped.copy(); bad[3]["n_edits"] += 1                                                          # edits no longer add up to the read
        assert emit(fn, bad) == capi.GB_ERR_ARG
        bad = mapped.copy(); bad[5]["mapping_off"] = bad[6]["mapping_off"]                                   # another record's mappings
        assert emit(fn, bad) == capi.GB_ERR_ARG
