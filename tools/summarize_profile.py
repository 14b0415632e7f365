"""Turn an ncu report (--set full, one build) + a launch list into the tracked summaries
under profiles/: a per-kernel markdown table and ncu_traffic.json (DRAM bytes per build)."""
import csv, io, json, subprocess, sys
rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return float("nan")
cols = [("us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("Minst", "smsp__inst_executed.sum"),
        ("regs", "launch__registers_per_thread"), ("smem_KB", "launch__shared_mem_per_block_dynamic")]
lines = ["| kernel | " + " | ".join(c for c, _ in cols) + " |", "|---|" + "---|" * len(cols)]
tot_rd = tot_wr = tot_us = 0.0
per_kernel = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = {c: g(r, k) for c, k in cols}
    vals["Minst"] /= 1e6
    if vals["smem_KB"] == vals["smem_KB"]: vals["smem_KB"] /= (1.0 if r[idx["launch__shared_mem_per_block_dynamic"]] == "" else 1.0)
    unit = rows[1][idx["dram__bytes_read.sum"]]
    scale = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}.get(unit, 1.0)
    vals["dram_rd_MB"] *= scale; vals["dram_wr_MB"] *= scale
    tot_rd += vals["dram_rd_MB"]; tot_wr += vals["dram_wr_MB"]; tot_us += vals["us"]
    per_kernel.append({"kernel": name, **{k: (None if v != v else round(v, 3)) for k, v in vals.items()}})
    lines.append("| " + name + " | " + " | ".join(f"{vals[c]:.1f}" for c, _ in cols) + " |")
lines.append(f"| **sum** | {tot_us:.1f} | {tot_rd:.1f} | {tot_wr:.1f} | | | | | | |")
lrows = [r for r in csv.reader(open(launches)) if len(r) > 5 and r[0].isdigit()]
half = lrows[-14:]  # the last build of the process (14 launches per build)
ll = ["| kernel | grid | block | us |", "|---|---|---|---|"] + [f"| {r[4].split('(')[0].replace('void ', '')} | {r[8]} | {r[7]} | {float(r[-1].replace(',', '')) / 1000:.1f} |" for r in half]
tot_l = sum(float(r[-1].replace(",", "")) for r in half) / 1000
open(f"profiles/ncu_summary_{tag}.md", "w").write(
    f"# ncu summary {tag}: one map build, cfg2 (10 M points, 0.2 m cells), B200, second build of the process\n\n"
    "## `ncu --set full --clock-control none` (per kernel; replayed, cold cache)\n\n" + "\n".join(lines) +
    "\n\n## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), same command\n\n" + "\n".join(ll) +
    f"\n\nsum of kernel times: {tot_l:.1f} us\n")
json.dump({"tag": tag, "workload": "cfg2 10M points", "build_dram_bytes": (tot_rd + tot_wr) * 1e6, "build_dram_read_bytes": tot_rd * 1e6,
           "build_dram_write_bytes": tot_wr * 1e6, "kernels": per_kernel}, open("profiles/ncu_traffic.json", "w"), indent=1)
print(open(f"profiles/ncu_summary_{tag}.md").read())
