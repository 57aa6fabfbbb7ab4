"""Run one secondary workload once (for ncu captures): python tests/tools/run_config.py config5|config4|config2se [n_reads]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
from vg_b200 import capi, synth
which = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
if which == "config4":
    g = synth.make_branchy_graph(); kw = dict(length=150, sub_rate=0.005, seed=44)
elif which == "config5":
    g = synth.make_variant_graph(); kw = dict(length=250, sub_rate=0.03, ins_rate=0.01, del_rate=0.01, seed=55)
else:
    g = synth.make_variant_graph(); kw = dict(length=150, sub_rate=0.002, seed=23)
index = g.build_index(); rs = synth.simulate_reads(g, n, **kw)
dev = capi.Device(index)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
import torch
for rep in range(2):
    if rep == 1:
        torch.cuda.profiler.start()           # ncu --profile-from-start off: only the second (warm) call is captured
    got = dev.map_arrays(rbuf, qbuf, read_off)
torch.cuda.profiler.stop()
print(which, n, dev.kernel_ms(), dev.kernel_times())
