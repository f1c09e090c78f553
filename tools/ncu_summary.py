#!/usr/bin/env python3
"""Summarises an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`) into profiles/*.json.
usage: ncu_summary.py <report.ncu-rep> <out.json> [frames]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return None


def unit(k):
    return rows[1][ix[k]] if k in ix else ""


def to_bytes(v, k):
    u = unit(k).lower()
    return None if v is None else v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, k):
    u = unit(k).lower()
    return None if v is None else v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


keep = {
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warp_instructions": "smsp__inst_executed.sum",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "smem_dynamic_KB": "launch__shared_mem_per_block_dynamic",
    "dram_pct_ncu": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lsu_pipe_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "global_load_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "global_store_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
}
launches = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = {"kernel": r[ix["Kernel Name"]][:90]}
    d["time_us"] = to_us(num(r, "gpu__time_duration.sum"), "gpu__time_duration.sum")
    rd, wr = to_bytes(num(r, "dram__bytes_read.sum"), "dram__bytes_read.sum"), to_bytes(num(r, "dram__bytes_write.sum"), "dram__bytes_write.sum")
    d["dram_read"], d["dram_write"], d["traffic"] = rd, wr, (rd or 0) + (wr or 0)
    for k, m in keep.items():
        if m in ix:
            d[k] = num(r, m)
    launches.append(d)
json.dump({"command": "ncu --set full --clock-control none (see tools/gpu_profile.sh); times are cold-cache and serialised",
           "frames": frames, "launches": launches}, open(out, "w"), indent=1)
for d in launches:
    print(f'{d["kernel"][:60]:60s} {d["time_us"]:9.1f} us  traffic {d["traffic"] / 1e6:9.1f} MB  issue {d.get("issue_active_pct")}')
