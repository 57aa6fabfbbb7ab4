import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
import bench
g, index = bench.make_graph_and_index()
lib, flags = bench.oracle_library()
rs_reads, rs_quals = bench.simulate_pairs_numpy(g, 400000, 22)
p = H.paired_params(400.0, 50.0)
for th in (16, 64, 128):
    n = 800000
    H.oracle_map_paired(index, rs_reads[:n], rs_quals[:n], p, threads=th)
    t = time.time(); H.oracle_map_paired(index, rs_reads[:n], rs_quals[:n], p, threads=th); dt = time.time() - t
    print(f"threads {th}: {n/dt/1e3:.1f} k reads/s", flush=True)
