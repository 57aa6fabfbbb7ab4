"""Debug driver: the pinned X-drop seam (tile kernel) against the oracle on the golden vectors and random trees."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_xdrop_golden as T
import numpy as np
import helpers as H
from vg_b200 import capi
bad = 0
for case in T.GOLD["cases"][: int(sys.argv[1]) if len(sys.argv) > 1 else 3]:
    index = T.case_index(case)
    sc = capi.Scores(*case["scores"])
    dev = capi.Device(index, scores=sc)
    nodes = [2 * (i + 1) for i in range(len(case["nodes"]))]
    want = T.oracle_xdrop(index, case["parents"], nodes, 0, case["read"], sc, max(case["max_gap"], 1))
    print("CASE", case["name"], "scores", case["scores"], "read", case["read"], "nodes", case["nodes"], "parents", case["parents"], "gap", case["max_gap"], flush=True)
    try:
        got = dev.xdrop_pinned_batch([(case["parents"], nodes, 0, case["read"].encode(), max(case["max_gap"], 1))])[0]
        print("  got ", got[0], got[1]); print("  want", want[0], want[1], flush=True)
        bad += got[0] != want[0] or got[1] != want[1]
    except AssertionError as e:
        print("  FAIL", e, "want", want, flush=True); bad += 1
    dev.close()
print("bad", bad)
