"""Condenses the raw round-2 GPU outputs (scratch: gpurun_out/r2/) into the tracked summaries under profiles/.
Nothing here was measured under a profiler unless it is an ncu metric.  usage: python scripts/make_profiles_r02.py"""
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r2")
DST = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__branch_targets_threads_uniform.pct",
]


def raw_csv(rep):
    out = os.path.join(SRC, rep + "_raw.csv")
    if not os.path.exists(out) and os.path.exists(os.path.join(SRC, rep + ".ncu-rep")):
        with open(out, "w") as f:
            subprocess.run(["ncu", "-i", os.path.join(SRC, rep + ".ncu-rep"), "--page", "raw", "--csv"], stdout=f,
                           stderr=subprocess.DEVNULL)
    return out if os.path.exists(out) else None


def condense(dst_name, captures):
    rows_out = [["capture", "kernel", "metric", "unit", "value"]]
    for label, rep in captures:
        path = raw_csv(rep)
        if not path:
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            kern = r[idx["Kernel Name"]]
            for m in METRICS:
                if m in idx:
                    rows_out.append([label, kern, m, units[idx[m]], r[idx[m]]])
    with open(os.path.join(DST, dst_name), "w", newline="") as f:
        csv.writer(f).writerows(rows_out)
    print("wrote", dst_name, len(rows_out) - 1, "rows")


condense("r02_cornell_ncu.csv", [
    ("two-kernel pipeline (default), final build, 32-sample batch at 1024^2: k_trace<PRIMARY> + k_shade of bounce 0, then k_trace<queue> + k_shade of bounce 1", "prof_final2"),
    ("one-kernel pipeline (B2PT_FLAG_ONE_KERNEL_BOUNCE), the same bounce: k_bounce", "prof_fused"),
])
condense("r02_bvh_ncu.csv", [
    ("1M spheres, 1920x1080, 16 spp, bounce 1: binary tree, queue order (round-1 configuration)", "prof_sph_bin"),
    ("the same launch: 8-wide compressed tree (B2PT_FLAG_WIDE_BVH), queue order", "prof_sph_wide"),
    ("the same launch: binary tree, rays sorted spatially (default)", "prof_sph_bin_sort"),
])
for src, dst in (("launches_sph_bin_sort.csv", "r02_launches_spheres.csv"), ("launches_bench.csv", "r02_launches_bench.csv"),
                 ("bench_n1.json", "bench_r02.json"), ("bench_ref.json", "bench_ref_r02.json"),
                 ("configs.jsonl", "configs_r02.jsonl"), ("bench_n2.json", "bench_n2_r02.json"),
                 ("bench_n4.json", "bench_n4_r02.json"), ("bench_n8.json", "bench_n8_r02.json"),
                 ("bench_n8_4096.json", "bench_n8_4096_r02.json"), ("hist.log", "r02_phase2_hist.jsonl"),
                 ("bvh_hist.log", "r02_bvh_traversal_counts.jsonl"), ("multi_rank_check.json", "r02_multi_rank_check.jsonl")):
    if os.path.exists(os.path.join(SRC, src)):
        shutil.copy(os.path.join(SRC, src), os.path.join(DST, dst))
        print("copied", dst)
