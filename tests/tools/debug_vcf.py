import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H, vcf_graph as V
from vg_b200 import capi, synth
g, _ = V.kgp_100k()
index = g.build_index()
rs = synth.simulate_reads(g, 3000, length=150, sub_rate=0.01, seed=7)
want = H.oracle_map(index, rs.reads, rs.quals, threads=8)
for tiles in ("1", "0"):
    os.environ["GIRAFFE_B200_TILES"] = tiles
    dev = capi.Device(index)
    got = H.gpu_map(dev, rs.reads, rs.quals)
    bad = np.nonzero(got[3] != 0)[0]
    print("tiles", tiles, "status errors", len(bad), [(int(i), int(got[3][i])) for i in bad[:10]], flush=True)
    if len(bad):
        i = int(bad[0])
        sc, mq, path = H.decode_alignment(want[0][i], want[1], want[2])
        print("  oracle read", i, "score", sc, "mapq", mq, "mappings", len(path), "edits", sum(len(e) for _, _, e in path), "oracle status", int(want[3][i]))
        rbuf, qbuf, read_off = H.pack_reads(rs.reads[i:i+1], rs.quals[i:i+1])
        st = dev.seed_stage(rbuf, qbuf, read_off)
        print("  stage:", {k: int(st[0][0][k]) for k in ("min_cnt", "seed_cnt", "cluster_cnt", "item_cnt", "status")})
        one = H.gpu_map(dev, rs.reads[i:i+1], rs.quals[i:i+1])
        print("  alone status", int(one[3][0]))
    dev.close()
