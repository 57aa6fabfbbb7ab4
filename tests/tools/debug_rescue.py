"""Mate-rescue parity triage: status histogram and the first differing pairs (GPU vs oracle)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers as H
import test_map_paired_parity as T
from vg_b200 import capi

attempts = int(sys.argv[1]) if len(sys.argv) > 1 else 15
g, rs, wrecked = T._wrecked_pairs()
index = g.build_index()
dev = capi.Device(index)
p = H.paired_params(); p.max_rescue_attempts = attempts
got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
want = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=8)
print("gpu status histogram", np.bincount(got[3], minlength=8)[:8], "oracle", np.bincount(want[3], minlength=8)[:8])
ok = [i for i in range(rs.n) if got[3][i] == 0 and want[3][i] == 0]
bad = H.compare_alignments(got, want, rs.n, indices=ok)
print("compared", len(ok), "differ", len(bad))
gr = (got[0]["flags"] & capi.GB_ALN_RESCUED) != 0; wr = (want[0]["flags"] & capi.GB_ALN_RESCUED) != 0
print("rescued gpu", int(gr.sum()), "oracle", int(wr.sum()), "flag mismatch", int((gr != wr).sum()))
for i, gd, wd in bad[:12]:
    print("read", i, "rescued gpu/oracle", bool(gr[i]), bool(wr[i]), "uncapped", got[0][i]["mapq_uncapped"], want[0][i]["mapq_uncapped"],
          "cap", got[0][i]["mapq_explored_cap"], want[0][i]["mapq_explored_cap"])
    print("  got ", gd)
    print("  want", wd)
