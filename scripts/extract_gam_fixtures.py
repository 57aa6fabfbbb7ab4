"""Copy three small GAM files written by vg itself out of the reference's test data into tests/golden/gam/.
They are the only place in the reference tree where vg.proto's wire layout (libvgio is absent) can be read off:
field numbers and types of Alignment / Path / Mapping / Position / Edit and of the annotation Struct.
  test/small/x-s13241-n1-p500-v300.gam   one simulated pair (fragment_prev / fragment_next, refpos)
  test/surject/perpendicular.gam          a mapped read with quality, MAPQ, sample, read group, annotations
  test/tiny/flat-s69-n1-l50-e0.05.gam     a simulated read with errors (edits carrying sequence)
Development container only; tests read the copies.   usage: python scripts/extract_gam_fixtures.py"""
import shutil
from pathlib import Path

REF = Path("/root/reference/test")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "gam"
OUT.mkdir(parents=True, exist_ok=True)
for rel in ("small/x-s13241-n1-p500-v300.gam", "surject/perpendicular.gam", "tiny/flat-s69-n1-l50-e0.05.gam"):
    shutil.copyfile(REF / rel, OUT / Path(rel).name)
    print(rel, "->", OUT / Path(rel).name)
