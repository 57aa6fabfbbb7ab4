"""One device call on a secondary config for ncu.  usage: python scripts/profile_cfg.py {4|5} [n_reads]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
from vg_b200 import capi, synth
cfg = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
if cfg == "4":
    g = synth.make_branchy_graph(); rs = synth.simulate_reads(g, n, length=150, sub_rate=0.005, seed=44)
else:
    g = synth.make_variant_graph(); rs = synth.simulate_reads(g, n, length=250, sub_rate=0.03, ins_rate=0.01, del_rate=0.01, seed=55)
index = g.build_index(); dev = capi.Device(index)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
got = dev.map_arrays(rbuf, qbuf, read_off)
print("kernel ms", dev.kernel_ms(), dev.stage_times())
