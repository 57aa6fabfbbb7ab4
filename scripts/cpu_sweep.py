import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from vg_b200 import synth
import bench
g, index = bench.make_graph_and_index()
lib, flags = bench.oracle_library()
print("flags", flags, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
rs_reads, rs_quals = bench.simulate_pairs_numpy(g, 400000, 22)
p = H.paired_params(400.0, 50.0)
for th in (1, 8, 16, 32, 64, 96, 128):
    n = 40000 if th == 1 else 800000
    H.oracle_map_paired(index, rs_reads[:n], rs_quals[:n], p, threads=th)
    t = time.time(); H.oracle_map_paired(index, rs_reads[:n], rs_quals[:n], p, threads=th); dt = time.time() - t
    print(f"threads {th}: {n/dt/1e3:.1f} k reads/s  ({n/dt/th/1e3:.2f} k/thread)")
