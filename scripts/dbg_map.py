import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from vg_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
g = synth.make_variant_graph(length=200000, n_snp=320, n_ins=40, n_del=40, n_haps=8, seed=2)
rs = synth.simulate_reads(g, n, length=150, sub_rate=0.002, ins_rate=0.0002, del_rate=0.0002, seed=55)
index = g.build_index()
dev = capi.Device(index)
got = H.gpu_map(dev, rs.reads, rs.quals)
want = H.oracle_map(index, rs.reads, rs.quals, threads=8)
bad = H.compare_alignments(got, want, rs.n)
print("bad", len(bad))
for b in bad[:3]:
    print(b)
