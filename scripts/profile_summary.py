"""Summarise an ncu launch list + a `--set full` report into profiles/<tag>.{json,md} (round 2 layout).
usage: python scripts/profile_summary.py TAG launches.csv report.ncu-rep --reads N [--commit SHA]
  launches.csv   ncu --metrics gpu__time_duration.sum --clock-control none --csv  (every launch of the bench command)
  report         ncu --set full --clock-control none --import-source on          (one launch of each kernel of interest)
Kernels are keyed by the names gb_kernel_times() reports, so bench.py can look up roofline.traffic for whatever kernel is
dominant in the run it measures.  Nothing in here is a bench value (ncu serialises and cold-starts every launch)."""
import csv, io, json, re, subprocess, sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, launches = sys.argv[1], sys.argv[2]
rep = next((a for a in sys.argv[3:] if a.endswith(".ncu-rep") or a.endswith("_raw.csv")), None)      # a report, or its `--page raw --csv` export
reads = int(sys.argv[sys.argv.index("--reads") + 1]) if "--reads" in sys.argv else None
commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()


def short(name):
    if "DeviceScan" in name: return "cub::DeviceScan"
    if "at::" in name or "at_cuda_detail" in name or "elementwise" in name: return "torch (input synthesis)"
    m = re.search(r"(?:gb::)?(\w+)(?:<([^>]*)>)?\(", name)
    if not m: return name[:40]
    base, targ = m.group(1), m.group(2)
    if base == "xdrop_tile_kernel": return f"xdrop_tile_kernel<{targ.strip().split(')')[-1].strip()}>"
    if base == "align_kernel_pe": return "align_kernel_pe<rescue>" if targ and targ.strip().endswith("1") else "align_kernel_pe"
    if base in ("tail_plan_kernel", "tail_decide_kernel", "seed_kernel_pe"): return base
    return base


rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 10 and r[0].isdigit()]
per = OrderedDict()
for r in rows:
    k = short(r[4]); ns = float(r[-1])
    e = per.setdefault(k, {"launches": 0, "ns": 0.0}); e["launches"] += 1; e["ns"] += ns
ours = {k: v for k, v in per.items() if not k.startswith("torch")}
tot = sum(v["ns"] for v in ours.values()) or 1.0
summary = {"tag": tag, "commit": commit, "reads_per_launch": reads, "source": {"launches": launches, "report": rep}, "kernels": {}}
for k, v in sorted(ours.items(), key=lambda kv: -kv[1]["ns"]):
    summary["kernels"][k] = {"launches": v["launches"], "total_ms": v["ns"] / 1e6, "share_of_step": v["ns"] / tot}

KEYS = {"gpu__time_duration.sum": "duration_ms", "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
        "launch__shared_mem_per_block_dynamic": "dyn_smem_kb", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
        "smsp__inst_executed.sum": "warp_instructions",
        "sm__inst_executed_pipe_alu.sum": "alu_pipe_warp_instructions", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_active_pct_of_peak",
        "sm__inst_executed_pipe_lsu.sum": "lsu_pipe_warp_instructions",
        "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
if rep:
    out = open(rep, errors="replace").read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    h, units = rr[0], rr[1]
    for r in rr[2:]:
        k = short(r[h.index("Kernel Name")])
        e = summary["kernels"].setdefault(k, {})
        f = {}
        for m, nm in KEYS.items():
            if m in h:
                i = h.index(m); v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None
                if nm in ("dram_read", "dram_write") and v is not None: v *= UNIT.get(units[i], 1.0)
                f[nm] = v
        # keep the longest launch of each kernel (a kernel may run several times per chunk: the retry pass, the two tile waves)
        if "ncu_full" in e and (e["ncu_full"].get("duration_ms") or 0) >= (f.get("duration_ms") or 0): continue
        f["dram_bytes_per_launch"] = (f.get("dram_read") or 0) + (f.get("dram_write") or 0)
        e["ncu_full"] = f
        e["dram_bytes_per_launch"] = f["dram_bytes_per_launch"]
(ROOT / "profiles" / f"{tag}.json").write_text(json.dumps(summary, indent=1))
md = [f"# ncu summary {tag}", "", f"commit {commit}; reads per launch: {reads}", "", "| kernel | launches | total ms (ncu, serialised) | share |", "|---|---|---|---|"]
for k, v in summary["kernels"].items():
    if "total_ms" in v: md.append(f"| {k} | {v['launches']} | {v['total_ms']:.3f} | {100 * v['share_of_step']:.1f} % |")
for k, v in summary["kernels"].items():
    if "ncu_full" in v:
        md += ["", f"## {k} (`--set full`, one launch)", ""] + [f"- {a}: {b}" for a, b in v["ncu_full"].items()]
(ROOT / "profiles" / f"{tag}.md").write_text("\n".join(md) + "\n")
print("\n".join(md))
