"""Builds the tracked summaries under profiles/ from the ncu captures gpurun brought back in gpurun_out/ (scratch).
usage: make_profiles.py <tag>   (expects gpurun_out/prof_{trace,shade}_<tag>_raw.csv, src_{trace,shade}_<tag>.csv,
launches_bench_<tag>.csv, bench_<tag>.json, elf_<tag>/b2pt_kernels.sm_100a.cubin)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__sass_branch_targets.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def raw(name):
    rows = list(csv.reader(open(os.path.join(G, "prof_%s_%s_raw.csv" % (name, tag)))))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


traffic = {}
with open(os.path.join(P, "%s_ncu_full.csv" % tag), "w") as f:
    wcsv = csv.writer(f)
    wcsv.writerow(["capture", "metric", "unit", "value"])
    for name in ("trace", "shade"):
        v, u = raw(name)
        for k in KEEP:
            if k in v:
                wcsv.writerow([name, k, u[k], v[k]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        traffic[name] = sum(float(v[k]) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[name + "_ms"] = float(v["gpu__time_duration.sum"]) * (1e-3 if u["gpu__time_duration.sum"] == "us" else 1.0)

bench = json.loads(open(os.path.join(G, "bench_%s.json" % tag)).read().strip().splitlines()[-1])
rays = [p for p in bench["roofline"]["stage_profile_trace_ms_shade_ms_rays"]][1][2]
json.dump({
    "kernel": "bounce 1 of a 128-sample batch at 1024x1024 (%d rays in, the launch pair bench.py times for `achieved`): "
              "k_trace<queue> launch + k_shade launch" % rays,
    "dram_bytes_per_launch": traffic["trace"] + traffic["shade"],
    "dram_bytes_k_trace": traffic["trace"], "dram_bytes_k_shade": traffic["shade"],
    "ncu_ms_k_trace": traffic["trace_ms"], "ncu_ms_k_shade": traffic["shade_ms"],
    "algorithmic_bytes_per_launch": 88 * rays,
    "source": "profiles/%s_ncu_full.csv (ncu --set full --clock-control none, scripts/profile_target.py)" % tag,
    "note": "traffic above the algorithmic 88 B/segment is the sorted hit bins: k_trace writes one 52-byte ray+hit record per "
            "surviving hit, k_shade reads it back (DESIGN.md 4)",
}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

# per-source-line breakdowns
cubin = os.path.join(G, "elf_%s" % tag, "b2pt_kernels.sm_100a.cubin")
for name, sub in (("trace", "k_traceILb0E12B2SmallSceneLb0"), ("shade", "k_shadeI12B2SmallSceneLb0ELb0")):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_agg.py"), cubin, sub,
                          os.path.join(G, "src_%s_%s.csv" % (name, tag)), "40"], capture_output=True, text=True).stdout
    open(os.path.join(P, "%s_%s_source_breakdown.txt" % (tag, name)), "w").write(out)

# launch list of the bench command, one row per launch
rows = [r for r in csv.reader(open(os.path.join(G, "launches_bench_%s.csv" % tag))) if len(r) > 5]
hdr = rows[0]
iK, iV, iM, iI = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
byid = collections.OrderedDict()
for r in rows[1:]:
    try:
        byid.setdefault(r[iI], {"k": r[iK].split("(")[0].replace("void ", "")})[r[iM]] = float(r[iV].replace(",", ""))
    except ValueError:
        pass
with open(os.path.join(P, "%s_launches_bench.csv" % tag), "w") as f:
    wcsv = csv.writer(f)
    wcsv.writerow(["id", "kernel", "gpu__time_duration.sum [ns]", "smsp__inst_executed.sum",
                   "thread_inst_per_inst (active lanes)", "issue_active pct"])
    for i, d in byid.items():
        wcsv.writerow([i, d["k"], d.get("gpu__time_duration.sum"), d.get("smsp__inst_executed.sum"),
                       d.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                       d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")])
json.dump(bench, open(os.path.join(P, "bench_%s.json" % tag), "w"), indent=1)
print("wrote profiles for", tag, traffic)
