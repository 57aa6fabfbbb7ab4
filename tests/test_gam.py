"""gb_emit_gam: vg.proto Alignment messages in vg::io framing.  libvgio is absent from the reference tree, so the wire
layout is pinned to GAM files vg itself wrote (reference test data copied to tests/golden/gam/ by
scripts/extract_gam_fixtures.py; perpendicular.gam is Giraffe output): the decoder below only knows the field numbers
gb_emit_gam writes, must decode those files consistently, and must decode the library's output back to the records."""
import struct
import zlib
from pathlib import Path

import numpy as np

import helpers as H
from vg_b200 import capi
import test_emit as TE

GAM_DIR = Path(__file__).parent / "golden" / "gam"


def _varint(b, i):
    v = s = 0
    while True:
        c = b[i]; i += 1; v |= (c & 0x7F) << s; s += 7
        if not c & 0x80:
            return v, i


def _fields(b):
    i, out = 0, []
    while i < len(b):
        key, i = _varint(b, i); f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v = struct.unpack("<d", b[i:i + 8])[0]; i += 8
        elif wt == 5:
            v = b[i:i + 4]; i += 4
        else:
            assert wt == 2, wt
            ln, i = _varint(b, i); v = b[i:i + ln]; i += ln
        out.append((f, v))
    return out


def _one(fs, f, default=None):
    vals = [v for k, v in fs if k == f]
    return vals[0] if vals else default


def decode_alignment(msg):
    fs = _fields(msg)
    aln = {"sequence": _one(fs, 1, b"").decode(), "name": _one(fs, 3, b"").decode(), "quality": _one(fs, 4), "mapq": _one(fs, 5, 0),
           "score": _one(fs, 6, 0), "identity": _one(fs, 16), "path": [], "annotation": {}}
    for k, key in ((11, "fragment_prev"), (12, "fragment_next")):
        if _one(fs, k) is not None:
            aln[key] = _one(_fields(_one(fs, k)), 3).decode()
    if _one(fs, 2) is not None:
        for _, mb in [kv for kv in _fields(_one(fs, 2)) if kv[0] == 2]:
            mf = _fields(mb); pos = _fields(_one(mf, 1, b""))
            edits = [(_one(ef, 1, 0), _one(ef, 2, 0), _one(ef, 3, b"").decode()) for ef in (_fields(e) for k, e in mf if k == 2)]
            aln["path"].append({"node_id": _one(pos, 1, 0), "offset": _one(pos, 2, 0), "is_reverse": bool(_one(pos, 4, 0)), "edits": edits, "rank": _one(mf, 5, 0)})
    if _one(fs, 100) is not None:
        for _, entry in [kv for kv in _fields(_one(fs, 100)) if kv[0] == 1]:
            ef = _fields(entry); val = _fields(_one(ef, 2, b""))
            aln["annotation"][_one(ef, 1).decode()] = val[0] if val else None          # (Value field number, payload)
    return aln


def read_stream(data):
    """vg::io groups: [count][len][msg]...; a leading message that is a bare type tag ("GAM") is skipped."""
    i, msgs, tagged = 0, [], 0
    while i < len(data):
        cnt, i = _varint(data, i)
        for k in range(cnt):
            ln, i = _varint(data, i); m = data[i:i + ln]; i += ln
            if k == 0 and m == b"GAM":
                tagged += 1
                continue
            msgs.append(m)
    return msgs, tagged


def _inflate(raw):
    data, rest = b"", raw
    while rest:
        d = zlib.decompressobj(31); data += d.decompress(rest); rest = d.unused_data
    return data


def test_decoder_reads_gam_written_by_vg():
    """The field numbers gb_emit_gam writes are the ones vg's own files use."""
    n_msgs = 0
    for name in ("x-s13241-n1-p500-v300.gam", "perpendicular.gam", "flat-s69-n1-l50-e0.05.gam"):
        msgs, _ = read_stream(_inflate((GAM_DIR / name).read_bytes()))
        for m in msgs:
            a = decode_alignment(m); n_msgs += 1
            assert a["sequence"] and set(a["sequence"]) <= set("ACGTN") and a["name"]
            assert a["path"]            # (ranks are free-form in vg's own files: absent in vg sim output, not renumbered by Giraffe)
            assert sum(e[1] for mp in a["path"] for e in mp["edits"]) == len(a["sequence"])          # to_lengths cover the read
            q = 0
            for mp in a["path"]:
                for frm, to, seq in mp["edits"]:
                    if seq:
                        assert len(seq) == to and seq == a["sequence"][q:q + to]                      # substitutions / insertions carry the read bases
                    q += to
            if a["quality"] is not None:
                assert len(a["quality"]) == len(a["sequence"]) and max(a["quality"]) <= 60           # raw phred, not ASCII
    assert n_msgs == 4
    pair, _ = read_stream(_inflate((GAM_DIR / "x-s13241-n1-p500-v300.gam").read_bytes()))
    first, second = decode_alignment(pair[0]), decode_alignment(pair[1])
    assert first["fragment_next"] == second["name"] and second["fragment_prev"] == first["name"]
    giraffe, tagged = read_stream(_inflate((GAM_DIR / "perpendicular.gam").read_bytes()))
    assert tagged == 1                                                                               # [2]["GAM"][message]
    g = decode_alignment(giraffe[0])
    assert g["annotation"]["mapq_explored_cap"][0] == 2 and g["annotation"]["proper_pair"][0] == 4   # number_value / bool_value
    assert 0.0 < g["identity"] <= 1.0 and g["score"] > 0


def test_library_gam_decodes_to_the_records():
    g, index, rs, res = TE._records(paired=True)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    names = [f"pair{i // 2}/{i % 2 + 1}" for i in range(rs.n)]
    data = capi.emit_text("gam", index.view, res[0], res[1], res[2], rbuf, qbuf, read_off, names)
    msgs, tagged = read_stream(data)
    assert len(msgs) == rs.n and tagged == 1
    for i, m in enumerate(msgs):
        a = decode_alignment(m)
        read = bytes(rs.reads[i]).decode()
        score, mapq, path = H.decode_alignment(res[0][i], res[1], res[2])
        assert a["sequence"] == read and a["name"] == names[i] and a["quality"] == bytes(rs.quals[i])
        assert a["score"] == score and a["mapq"] == mapq
        assert a["fragment_next" if i % 2 == 0 else "fragment_prev"] == names[i ^ 1]
        assert a["annotation"]["mapq_explored_cap"][0] == 2 and a["annotation"]["mapq_uncapped"][0] == 2
        assert len(a["path"]) == len(path)
        q = 0
        for rank, (mp, (node, offset, edits)) in enumerate(zip(a["path"], path), 1):
            assert (mp["node_id"], mp["offset"], mp["is_reverse"], mp["rank"]) == (node >> 1, offset, bool(node & 1), rank)
            assert len(mp["edits"]) == len(edits)
            for (frm, to, seq), (op, length, base) in zip(mp["edits"], edits):
                want = {"M": (length, length, ""), "S": (1, 1, read[q:q + 1]), "I": (0, length, read[q:q + length]), "D": (length, 0, "")}[op]
                assert (frm, to, seq) == want
                q += to
        if path:
            assert q == rs.length and 0.0 < a["identity"] <= 1.0
        else:
            assert a["identity"] is None


def test_gam_groups_hold_a_thousand_messages():
    g, index, rs, res = TE._records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    many = np.concatenate([res[0]] * 7)                      # 2100 records -> 3 groups
    data = capi.emit_text("gam", index.view, many, res[1], res[2], rbuf, qbuf, read_off)
    msgs, tagged = read_stream(data)
    assert len(msgs) == len(many) and tagged == 3
    assert decode_alignment(msgs[0])["name"] == "read0" and "fragment_next" not in decode_alignment(msgs[0])


def _bgzf_blocks(raw):
    """Split a BGZF stream into (header16, bsize, payload bytes) per block, checking htslib's layout."""
    i, out = 0, []
    while i < len(raw):
        head = raw[i:i + 16]
        assert head[:4] == b"\x1f\x8b\x08\x04" and head[10:12] == b"\x06\x00" and head[12:16] == b"BC\x02\x00"
        bsize = struct.unpack("<H", raw[i + 16:i + 18])[0] + 1
        block = raw[i:i + bsize]
        payload = zlib.decompress(block[18:-8], -15)
        crc, isize = struct.unpack("<II", block[-8:])
        assert crc == zlib.crc32(payload) and isize == len(payload)
        out.append((head, bsize, payload)); i += bsize
    assert i == len(raw)
    return out


def test_bgzf_container_is_the_one_vg_writes():
    """gb_bgzf_compress against the GAM files vg wrote: same 16 header bytes on every block, the same 28-byte EOF block,
    and any gzip reader (zlib, as vg::io's BlockedGzipInputStream / htslib do) inflates it back to the input."""
    vg_raw = (GAM_DIR / "perpendicular.gam").read_bytes()
    vg_blocks = _bgzf_blocks(vg_raw)
    assert vg_blocks[-1][2] == b"" and len(vg_blocks) >= 2
    rng = np.random.default_rng(5)
    for size, level in ((0, 6), (1, 6), (1000, 1), (0xff00, 6), (0xff00 + 1, 9), (300000, 6), (70000, 0)):
        data = bytes(rng.integers(0, 4, size=size, dtype=np.uint8) + 65) if size != 300000 else bytes(rng.integers(0, 256, size=size, dtype=np.uint8))
        raw = capi.bgzf_compress(data, level)
        blocks = _bgzf_blocks(raw)
        assert all(b[0] == vg_blocks[0][0] for b in blocks)                   # byte-identical block headers
        assert raw[-28:] == vg_raw[-28:]                                      # the EOF marker block
        assert all(len(b[2]) <= 0xff00 and b[1] <= 0x10000 for b in blocks)
        assert len(blocks) == (size + 0xff00 - 1) // 0xff00 + 1
        assert b"".join(b[2] for b in blocks) == data and _inflate(raw) == data


def test_library_gam_in_bgzf_reads_back_like_vgs_files():
    """The whole emission path a drop-in needs: records -> GAM messages -> BGZF, read back with the reader used on vg's files."""
    g, index, rs, res = TE._records(paired=True)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    plain = capi.emit_text("gam", index.view, res[0], res[1], res[2], rbuf, qbuf, read_off)
    msgs, tagged = read_stream(_inflate(capi.bgzf_compress(plain)))
    assert tagged == 1 and len(msgs) == rs.n and [decode_alignment(m)["score"] for m in msgs] == [int(a["score"]) for a in res[0]]


def test_gam_negative_score_is_a_sign_extended_varint():
    """proto3 int32: a negative score is written as the 10-byte two's-complement varint protobuf parsers expect."""
    g, index, rs, res = TE._records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    aln = res[0][:1].copy(); aln["score"] = -7
    msgs, _ = read_stream(capi.emit_text("gam", index.view, aln, res[1], res[2], rbuf, qbuf, read_off))
    raw = _one(_fields(msgs[0]), 6)
    assert raw == (1 << 64) - 7


def test_emitters_refuse_decreasing_name_offsets():
    g, index, rs, res = TE._records(paired=False)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    lib = capi.load_library()
    import ctypes as C
    names = np.frombuffer(b"abcdefgh" * rs.n, dtype=np.uint8).copy()
    noff = np.arange(rs.n + 1, dtype=np.uint64) * 8
    noff[3] = 1                                                  # offset 3 below offset 2
    out = np.zeros(1 << 20, dtype=np.uint8); used = C.c_uint64()
    aln = np.ascontiguousarray(res[0])
    for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
        rc = fn(C.byref(index.view), len(aln), capi.ptr(aln), capi.ptr(res[1]), len(res[1]), capi.ptr(res[2]), len(res[2]), rs.n, capi.ptr(rbuf), capi.ptr(qbuf),
                capi.ptr(read_off), capi.ptr(names), capi.ptr(noff), capi.ptr(out), len(out), C.byref(used))
        assert rc == capi.GB_ERR_ARG
