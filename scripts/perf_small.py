import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from vg_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
g = synth.make_variant_graph(length=200000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=2)
index = g.build_index()
rs = synth.simulate_pairs(g, n // 2, sub_rate=0.002, seed=22)
dev = capi.Device(index)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    got = dev.map_arrays(rbuf, qbuf, read_off)
    print(f"kernels {dev.kernel_ms():.1f} ms -> {n / (dev.kernel_ms() / 1e3) / 1e6:.2f} M reads/s")
