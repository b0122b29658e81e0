#!/usr/bin/env python
"""Summarise an ncu report into profiles/: key raw metrics (text) and the per-opcode instruction mix.

usage: python scripts/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep profiles/<name> [frames_per_launch]
Writes <name>_metrics.txt (selected `--page raw` metrics per captured launch) and <name>_sass_mix.txt
(executed warp instructions and shared-memory wavefronts per opcode from `--page source`).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst + "_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none; source report {rep}\n")
        for d in data:
            rec = dict(zip(hdr, d))
            f.write(f"\n== {rec.get('Kernel Name')}  (launch id {rec.get('ID')})\n")
            for k in KEYS:
                if k in rec:
                    f.write(f"{k:85s} {rec[k]:>16s} {units[hdr.index(k)]}\n")
    rows = page(rep, "source")
    # several kernels may follow each other; take the first table
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
    ops, wf, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[start + 1:]:
        if len(r) <= iW or not r[iE].isdigit():
            break
        s = r[iS].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        full = s.split()[0] if s else "?"
        op = full.split(".")[0]
        key = full if op in ("LDS", "STS", "LDG", "STG", "SHFL") else op
        n = int(r[iE])
        ops[key] += n
        wf[key] += int(r[iW] or 0)
        tot += n
    with open(dst + "_sass_mix.txt", "w") as f:
        unit = f" (per frame, {frames:.0f} frames/launch)" if frames else ""
        f.write(f"# executed warp instructions by opcode{unit}; source report {rep}\n")
        sc = 1.0 / frames if frames else 1.0
        f.write(f"{'TOTAL':24s} {tot * sc:14.1f}  smem wavefronts {sum(wf.values()) * sc:12.1f}\n")
        for k, v in ops.most_common(45):
            f.write(f"{k:24s} {v * sc:14.1f}  smem wavefronts {wf[k] * sc:12.1f}\n")


if __name__ == "__main__":
    main()
