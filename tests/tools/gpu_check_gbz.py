"""GPU parity on a graph vg built: the reference's test GBZ (tests/golden/gbz/y.giraffe.gbz) read by gb_index_from_gbz,
single-end and paired reads drawn from its three haplotypes (with errors), CUDA path vs oracle.  Not part of the pytest
suite yet (written when no GPU time was left in round 1): run it first in round 2, then turn it into a test.
usage: python tests/tools/gpu_check_gbz.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers as H
import test_gbz as TG
from vg_b200 import capi, synth

seqs, paths, _ = TG.read_gbz(TG.GBZ)
index = capi.HostIndex.from_gbz(TG.GBZ)
haps = ["".join(seqs[(v >> 1) - 1] for v in p) for p in paths]
rng = np.random.default_rng(11)

def draw(n, length=150):
    out = []
    for _ in range(n):
        h = haps[int(rng.integers(0, len(haps)))]
        s = int(rng.integers(0, len(h) - length))
        r = np.frombuffer(h[s:s + length].encode(), dtype=np.uint8).copy()
        m = rng.random(length) < 0.02
        r[m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
        if rng.random() < 0.5:
            r = synth.revcomp_bytes(r[None, :])[0]
        out.append(r)
    return np.stack(out)

dev = capi.Device(index)
reads = draw(400); quals = np.full(reads.shape, 30, dtype=np.uint8)
bad = H.compare_alignments(H.gpu_map(dev, reads, quals), H.oracle_map(index, reads, quals, threads=8), len(reads))
print("single-end:", len(reads), "reads,", len(bad), "differ")
pairs = []
for _ in range(200):
    h = haps[int(rng.integers(0, len(haps)))]
    frag = int(np.clip(rng.normal(400, 50), 160, len(h) - 1)); s = int(rng.integers(0, len(h) - frag))
    a = np.frombuffer(h[s:s + 150].encode(), dtype=np.uint8).copy()
    b = synth.revcomp_bytes(np.frombuffer(h[s + frag - 150:s + frag].encode(), dtype=np.uint8).copy()[None, :])[0]
    pairs += [a, b] if rng.random() < 0.5 else [b, a]
pairs = np.stack(pairs); pq = np.full(pairs.shape, 30, dtype=np.uint8)
p = H.paired_params(); p.max_rescue_attempts = 15
badp = H.compare_alignments(H.gpu_map(dev, pairs, pq, p, paired=True), H.oracle_map_paired(index, pairs, pq, p, threads=8), len(pairs))
print("paired:", len(pairs), "reads,", len(badp), "differ")
dev.close()
sys.exit(1 if bad or badp else 0)
