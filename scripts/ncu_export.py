"""Turn gpurun_out/*.ncu-rep + launch lists into the tracked summaries under profiles/<round>/.
usage: python scripts/ncu_export.py r01"""
import csv, io, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles", rnd)
os.makedirs(out, exist_ok=True)
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
lines = ["# ncu summaries (%s)\n" % rnd,
         "Per-kernel metrics from `ncu --set full --clock-control none` captures (one launch each, ~40 replays, cold caches);",
         "launch lists from `ncu --metrics gpu__time_duration.sum` over the plain-launch bench step.  Raw exports: the `*_raw_metrics.csv` beside this file.\n"]
for rep in sorted(f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.endswith(".ncu-rep")):
    path = os.path.join(ROOT, "gpurun_out", rep)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    name = rep[:-8]
    with open(os.path.join(out, name + "_raw_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        for j, h in enumerate(hdr):
            w.writerow([h, units[j]] + [r[j] if j < len(r) else "" for r in rows[2:]])
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, units))
    lines.append("## %s — `%s`\n" % (name, d.get("Kernel Name", "?")[:100]))
    lines.append("| metric | value | unit |\n|---|---|---|")
    for k in KEEP:
        if k in d and d[k] != "":
            lines.append("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
    st = sorted(((float(d[k]), k.split("stalled_")[1].split("_per")[0]) for k in hdr
                 if "issue_stalled" in k and "per_issue_active" in k and d.get(k) not in (None, "")), reverse=True)[:5]
    lines.append("\nTop warp stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st) + "\n")
for ll in ("launches_c2.csv", "launches_c3.csv", "launches_c4.csv", "launches_c5.csv"):
    p = os.path.join(ROOT, "gpurun_out", ll)
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(open(p)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ik].split("(")[0].split("::")[-1][:48], float(r[iv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > iv]
    with open(os.path.join(out, ll.replace(".csv", "_step.csv")), "w") as f:
        f.write("kernel,duration_ns\n")
        for k, v in seq:
            f.write("%s,%.0f\n" % (k, v))
    # the last full step = last 6..8 launches starting at a forward kernel
    last = [i for i, (k, v) in enumerate(seq) if "forward" in k or "impala_direct" in k][-1]
    step = seq[last:]
    tot = sum(v for k, v in step)
    lines.append("## %s — one step, per-launch device time (cold cache, serialised: compare SHARES)\n" % ll)
    lines.append("| kernel | ns | share |\n|---|---|---|")
    for k, v in step:
        lines.append("| %s | %.0f | %.1f %% |" % (k, v, 100 * v / tot))
    lines.append("| total | %.0f | |\n" % tot)
open(os.path.join(out, "SUMMARY.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
