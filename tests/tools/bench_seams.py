"""Throughput of the DP stage seams on rescue-sized problems: gb_sw_batch (full local DP over a DAG)
and the CPU oracle on the same problems.  Reports cell updates per second (CUPS).
usage: python tests/tools/bench_seams.py [n_problems]"""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from vg_b200 import capi, synth
import test_sw_golden as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
g = synth.make_variant_graph(length=200000, n_snp=320, n_ins=40, n_del=40, n_haps=8, seed=2)
index = g.build_index()
rng = np.random.default_rng(5)
# problems: ~600 bp windows of haplotype 0's node walk (a chain; bubbles come in with the other haplotypes' nodes skipped)
walk = [int(x) for x in g.paths[0]]
lens = np.array([len(g.node_seqs[v // 2 - 1]) for v in walk])
cum = np.concatenate([[0], np.cumsum(lens)])
problems = []
hs = g.hap_seq[0]
for _ in range(n):
    a = int(rng.integers(0, len(walk) - 40))
    b = a
    while b < len(walk) and cum[b] - cum[a] < 600:
        b += 1
    nodes = walk[a:b]
    preds = [[i - 1] if i else [] for i in range(len(nodes))]
    start = int(cum[a] + rng.integers(0, max(1, cum[b] - cum[a] - 150)))
    q = hs[start:start + 150].copy()
    mut = rng.random(150) < 0.05
    q[mut] = synth.BASES[rng.integers(0, 4, size=int(mut.sum()))]
    problems.append((nodes, preds, bytes(q)))
cells = sum(sum(len(g.node_seqs[v // 2 - 1]) for v in p[0]) * len(p[2]) for p in problems)
dev = capi.Device(index)
best = None
for rep in range(3):
    got = dev.sw_batch(problems, map_cap=64, edit_cap=256)
    ms = dev.kernel_ms(); best = ms if best is None else min(best, ms)
t = time.time()
sub = problems[: min(n, 300)]
sc = capi.Scores(1, 4, 6, 1, 5)
want = [T.oracle_sw(index, {"node": p[0], "pred": p[1]}, p[2].decode(), sc) for p in sub]
dt = time.time() - t
sub_cells = sum(sum(len(g.node_seqs[v // 2 - 1]) for v in p[0]) * len(p[2]) for p in sub)
bad = sum(1 for i in range(len(sub)) if got[i] != (want[i][0], want[i][1]))
out = {"problems": n, "cells": int(cells), "gpu_kernel_ms": best, "gpu_gcups": cells / (best / 1e3) / 1e9,
       "cpu_1thread_gcups_incl_ctypes": sub_cells / dt / 1e9, "parity_mismatches": bad, "checked": len(sub)}
print(json.dumps(out))
json.dump(out, open("gpurun_out/bench_seams.json", "w"), indent=1)

# ---- gb_wfa_batch: suffix / connect problems cut from haplotype 0 -------------------------------------------
import test_wfa_golden as TW
wprobs, wcases = [], []
for _ in range(20000):
    a = int(rng.integers(50, len(hs) - 300)); ln = int(rng.integers(20, 100))
    seq = hs[a + 1:a + 1 + ln].copy()
    mut = rng.random(ln) < 0.02
    seq[mut] = synth.BASES[rng.integers(0, 4, size=int(mut.sum()))]
    frm = (2 * int(g.hap_node[0][a]), int(g.hap_off[0][a])); b_ = a + 1 + ln
    to = (2 * int(g.hap_node[0][b_]), int(g.hap_off[0][b_]))
    mode = int(rng.integers(0, 2))
    wprobs.append((mode, bytes(seq), frm, to if mode == 0 else None))
    wcases.append({"call": ["connect", "suffix"][mode], "sequence": bytes(seq).decode(), "error_model": None,
                   "from": [frm[0] >> 1, False, frm[1]], "to": None if mode == 1 else [to[0] >> 1, False, to[1]]})
wbest = None
for rep in range(3):
    wgot = dev.wfa_batch(wprobs)
    ms = dev.kernel_ms(); wbest = ms if wbest is None else min(wbest, ms)
t = time.time()
wwant = [TW.oracle_wfa(index, c) for c in wcases[:500]]
wdt = time.time() - t
wbad = sum(1 for i in range(500) if wgot[i] != wwant[i])
out["wfa"] = {"problems": len(wprobs), "gpu_kernel_ms": wbest, "gpu_problems_per_s": len(wprobs) / (wbest / 1e3),
              "cpu_1thread_problems_per_s_incl_ctypes": 500 / wdt, "parity_mismatches": wbad, "checked": 500}
print(json.dumps(out["wfa"]))
json.dump(out, open("gpurun_out/bench_seams.json", "w"), indent=1)
