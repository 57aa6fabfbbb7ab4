"""Top source lines by executed instructions from an ncu report (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; data = []; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples")
        try:
            data.append((int(r[ie]), int(r[ss]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
tots = sum(d[1] for d in data) or 1
print(f"kernel ~{kern}: {tot} warp instructions attributed, {tots} samples")
for n, s, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% samp  {f}:{ln}  {src}")

# optional 4th argument: "file:lo-hi=name,..." ranges to aggregate (phase breakdown)
if len(sys.argv) > 4:
    print("phase breakdown:")
    for spec in sys.argv[4].split(","):
        rng, name = spec.split("=")
        f, lh = rng.split(":"); lo, hi = map(int, lh.split("-"))
        n = sum(d[0] for d in data if d[2] == f and lo <= d[3] <= hi)
        s = sum(d[1] for d in data if d[2] == f and lo <= d[3] <= hi)
        print(f"{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% samp  {name} ({rng})")
