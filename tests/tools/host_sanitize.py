"""AddressSanitizer / UBSan pass over the library's host-only C++ (emit.cpp, index_builder.cpp, gbz_reader.cpp): builds the files
with g++ -fsanitize=address,undefined into gpurun_out/libhost_asan.so and drives index build / save / load (including
truncated and corrupted files) and the three emitters (including output buffers that are too small) through ctypes.
usage: python tests/tools/host_sanitize.py        (re-executes itself under LD_PRELOAD=libasan.so)"""
import ctypes as C, os, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
LIB = ROOT / "gpurun_out" / "libhost_asan.so"

if os.environ.get("HOST_SANITIZE_CHILD") != "1":
    LIB.parent.mkdir(exist_ok=True)
    src = [str(ROOT / "vg_b200" / "csrc" / f) for f in ("emit.cpp", "index_builder.cpp", "gbz_reader.cpp")]
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                    "-I", str(ROOT / "include"), "-I", str(ROOT / "vg_b200" / "csrc"), "-o", str(LIB)] + src, check=True)
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, HOST_SANITIZE_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    sys.exit(subprocess.run([sys.executable, __file__], env=env).returncode)

sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from vg_b200 import capi, synth
import helpers as H

lib = C.CDLL(str(LIB))
vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
lib.gb_index_build.argtypes = [u32, vp, vp, u32, vp, vp, vp, u32, u32, C.POINTER(vp)]
lib.gb_index_view.argtypes = [vp, C.POINTER(capi.FlatIndex)]
lib.gb_index_free.argtypes = [vp]
lib.gb_index_save.argtypes = [C.POINTER(capi.FlatIndex), C.c_char_p]
lib.gb_index_load.argtypes = [C.c_char_p, C.POINTER(vp)]
for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
    fn.argtypes = [C.POINTER(capi.FlatIndex), u32, vp, vp, u64, vp, u64, u32, vp, vp, vp, vp, vp, vp, u64, vp]

g = synth.make_variant_graph(length=30000, n_snp=48, n_ins=6, n_del=6, n_haps=4, seed=3)
node_off = np.zeros(len(g.node_seqs) + 1, dtype=np.uint64); node_off[1:] = np.cumsum([len(s) for s in g.node_seqs])
seq = np.frombuffer("".join(g.node_seqs).encode(), dtype=np.uint8).copy()
path_off = np.zeros(len(g.paths) + 1, dtype=np.uint64); path_off[1:] = np.cumsum([len(p) for p in g.paths])
flat = np.concatenate([np.asarray(p, dtype=np.uint32) for p in g.paths])
dist = np.ascontiguousarray(g.dist, dtype=capi.dist_dt)
h = vp()
assert lib.gb_index_build(len(g.node_seqs), capi.ptr(seq), capi.ptr(node_off), len(g.paths), capi.ptr(flat), capi.ptr(path_off), capi.ptr(dist), 29, 11, C.byref(h)) == 0
view = capi.FlatIndex(); assert lib.gb_index_view(h, C.byref(view)) == 0

with tempfile.TemporaryDirectory() as tmp:
    p = os.path.join(tmp, "x.gbflat").encode()
    assert lib.gb_index_save(C.byref(view), p) == 0
    h2 = vp(); assert lib.gb_index_load(p, C.byref(h2)) == 0
    lib.gb_index_free(h2)
    raw = open(p, "rb").read()
    rng = np.random.default_rng(1)
    refused = 0
    for t in range(200):                                    # truncations and random corruptions of header / offsets
        b = bytearray(raw)
        if t % 2 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, min(len(b), 64 + 16 * view.n_nodes)))
                b[pos] = int(rng.integers(0, 256))
        q = os.path.join(tmp, "bad.gbflat")
        open(q, "wb").write(bytes(b))
        hb = vp(); rc = lib.gb_index_load(q.encode(), C.byref(hb))
        if rc == 0:
            lib.gb_index_free(hb)
        else:
            refused += 1
    print("corrupted files refused:", refused, "of 200 (the rest still passed every check)")

# the GBZ reader on the reference's test GBZ: intact, truncated, and with random bytes overwritten (must refuse or build, never fault)
lib.gb_index_from_gbz.argtypes = [C.c_char_p, u32, u32, C.POINTER(vp)]
gbz = (ROOT / "tests" / "golden" / "gbz" / "y.giraffe.gbz").read_bytes()
with tempfile.TemporaryDirectory() as tmp:
    q = os.path.join(tmp, "y.gbz"); open(q, "wb").write(gbz)
    hg = vp(); assert lib.gb_index_from_gbz(q.encode(), 29, 11, C.byref(hg)) == 0; lib.gb_index_free(hg)
    rng = np.random.default_rng(3); built = 0
    for t in range(400):
        b = bytearray(gbz)
        if t % 4 == 0:
            b = b[: 8 * int(rng.integers(0, len(b) // 8))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        open(q, "wb").write(bytes(b))
        hg = vp()
        if lib.gb_index_from_gbz(q.encode(), 29, 11, C.byref(hg)) == 0:
            built += 1; lib.gb_index_free(hg)
    print("damaged GBZ files: refused", 400 - built, "built", built)

# the .min / .zipcodes readers beside it: intact, truncated, random bytes overwritten (must refuse or build, never fault)
lib.gb_index_from_gbz_min.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp)]
gdir = ROOT / "tests" / "golden" / "gbz"
minraw, zipraw = (gdir / "y.min").read_bytes(), (gdir / "y.zipcodes").read_bytes()
with tempfile.TemporaryDirectory() as tmp:
    qg = os.path.join(tmp, "y.gbz"); open(qg, "wb").write(gbz)
    qm = os.path.join(tmp, "y.min"); qz = os.path.join(tmp, "y.zipcodes")
    open(qm, "wb").write(minraw); open(qz, "wb").write(zipraw)
    hm = vp(); assert lib.gb_index_from_gbz_min(qg.encode(), qm.encode(), qz.encode(), C.byref(hm)) == 0; lib.gb_index_free(hm)
    rng = np.random.default_rng(7); built = 0
    for t in range(600):
        bm, bz = bytearray(minraw), bytearray(zipraw)
        which = t % 3
        target = bm if which < 2 else bz
        if t % 5 == 0:
            del target[int(rng.integers(0, len(target))):]
        else:
            for _ in range(int(rng.integers(1, 4))):
                hot = 96 if (which == 0) else len(target)                     # every third file: damage the header words
                target[int(rng.integers(0, hot))] = int(rng.integers(0, 256))
        open(qm, "wb").write(bytes(bm)); open(qz, "wb").write(bytes(bz))
        hm = vp()
        if lib.gb_index_from_gbz_min(qg.encode(), qm.encode(), qz.encode(), C.byref(hm)) == 0:
            built += 1; lib.gb_index_free(hm)
    print("damaged .min / .zipcodes files: refused", 600 - built, "built", built)

# gb_index_build_with_hits: hits that lie off the graph, past a node, on the wrong k-mer (must refuse, never fault)
lib.gb_index_build_with_hits.argtypes = [u32, vp, vp, u32, vp, vp, vp, u32, u32, u64, vp, vp, C.POINTER(vp)]
ref_index = g.build_index()
tb, hits = ref_index.array("table"), ref_index.array("hits")
keys, poss = [], []
for c in tb:
    if int(c["key"]) != 0xFFFFFFFFFFFFFFFF:
        for i in range(int(c["hit_cnt"])):
            keys.append(int(c["key"])); poss.append(int(hits[int(c["hit_off"]) + i]["pos"]))
keys = np.asarray(keys, dtype=np.uint64); poss = np.asarray(poss, dtype=np.uint64)
hh = vp(); assert lib.gb_index_build_with_hits(len(g.node_seqs), capi.ptr(seq), capi.ptr(node_off), len(g.paths), capi.ptr(flat), capi.ptr(path_off), capi.ptr(dist), 29, 11,
                                               len(keys), capi.ptr(keys), capi.ptr(poss), C.byref(hh)) == 0
lib.gb_index_free(hh)
rng = np.random.default_rng(8); refused = 0
for t in range(200):
    k2, p2 = keys.copy(), poss.copy()
    i = int(rng.integers(0, len(k2)))
    if t % 2 == 0:
        p2[i] = np.uint64(int(rng.integers(0, 1 << 40)))
    else:
        k2[i] = np.uint64(int(k2[i]) ^ (1 << int(rng.integers(0, 58))))
    hh = vp()
    rc = lib.gb_index_build_with_hits(len(g.node_seqs), capi.ptr(seq), capi.ptr(node_off), len(g.paths), capi.ptr(flat), capi.ptr(path_off), capi.ptr(dist), 29, 11,
                                      len(k2), capi.ptr(k2), capi.ptr(p2), C.byref(hh))
    if rc == 0:
        lib.gb_index_free(hh)
    else:
        refused += 1
print("damaged hit lists refused:", refused, "of 200")
ref_index.close()

# multi-mapping records (secondaries, absent ranks) through the emitters
index_r = synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6)
ix_r = index_r.build_index()
rs_r = synth.simulate_reads(index_r, 90, length=150, sub_rate=0.01, seed=36)
pm = H.default_map_params(); pm.max_multimaps = 3
res_r = H.oracle_map(ix_r, rs_r.reads, rs_r.quals, pm, threads=2)
rb_r, qb_r, ro_r = H.pack_reads(rs_r.reads, rs_r.quals)
aln_r = np.ascontiguousarray(res_r[0]); out_r = np.zeros(1 << 22, dtype=np.uint8)
for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
    used = u64()
    assert fn(C.byref(ix_r.view), len(aln_r), capi.ptr(aln_r), capi.ptr(res_r[1]), len(res_r[1]), capi.ptr(res_r[2]), len(res_r[2]), rs_r.n, capi.ptr(rb_r), capi.ptr(qb_r),
              capi.ptr(ro_r), None, None, capi.ptr(out_r), len(out_r), C.byref(used)) == 0
print("multi-mapping records emitted:", int(((aln_r["flags"] & 16) == 0).sum()), "of", len(aln_r), "ranks present")
ix_r.close()

# records from the oracle (tail alignments, soft clips, unmapped reads, pairs), through all three emitters
index = g.build_index()
rs = synth.simulate_pairs(g, 120, sub_rate=0.02, seed=5, indel_rate=0.002)
rng = np.random.default_rng(2)
for i in rng.integers(0, rs.n, size=12):
    rs.reads[i, :15] = synth.BASES[rng.integers(0, 4, size=15)]
rs.reads[9] = synth.BASES[rng.integers(0, 4, size=rs.length)]
res = H.oracle_map_paired(index, rs.reads, rs.quals, H.paired_params(), threads=4)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
aln = np.ascontiguousarray(res[0])
for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
    for quals in (qbuf, None):
        need = None
        for cap in (1 << 22, 0, 1, 17, 1000, 40000):
            out = np.zeros(max(cap, 1), dtype=np.uint8); used = u64()
            rc = fn(C.byref(view), len(aln), capi.ptr(aln), capi.ptr(res[1]), len(res[1]), capi.ptr(res[2]), len(res[2]), rs.n, capi.ptr(rbuf),
                    capi.ptr(quals) if quals is not None else None, capi.ptr(read_off), None, None, capi.ptr(out), cap, C.byref(used))
            if cap == 1 << 22:
                assert rc == 0; need = used.value
            else:
                assert (rc == 0) == (cap >= need), (cap, need, rc)
                assert used.value <= cap
# damaged records: random bytes overwritten in the 32-byte headers; every emitter must refuse or write, never fault
rng = np.random.default_rng(5); refused = 0
out = np.zeros(1 << 22, dtype=np.uint8)
for t in range(300):
    bad = aln.copy(); raw_view = bad.view(np.uint8).reshape(-1)
    for _ in range(int(rng.integers(1, 5))):
        raw_view[int(rng.integers(0, len(raw_view)))] = int(rng.integers(0, 256))
    for fn in (lib.gb_emit_gaf, lib.gb_emit_json, lib.gb_emit_gam):
        used = u64()
        rc = fn(C.byref(view), len(bad), capi.ptr(bad), capi.ptr(res[1]), len(res[1]), capi.ptr(res[2]), len(res[2]), rs.n, capi.ptr(rbuf), capi.ptr(qbuf),
                capi.ptr(read_off), None, None, capi.ptr(out), len(out), C.byref(used))
        refused += rc != 0
print("damaged record batches refused:", refused, "of 900 emitter calls")
lib.gb_index_free(h)
print("host sanitizer pass: no AddressSanitizer / UBSan report")
