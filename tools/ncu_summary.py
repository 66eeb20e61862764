#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics per kernel, opcode mix, hottest SASS lines.
usage: tools/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_sector_op_atom_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in raw[2:]:
    name = r[idx["Kernel Name"]]
    print("=" * 100)
    print(name[:160])
    for k in KEYS:
        if k in idx:
            print(f"  {k:75s} {r[idx[k]]:>18s} {units[idx[k]]}")
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    vals = []
    for h in stalls:
        try:
            vals.append((float(r[idx[h]] or 0), h))
        except ValueError:
            pass
    tot = sum(v for v, _ in vals) or 1.0
    print("  warp stall reasons (share of pc samples):")
    for v, h in sorted(vals, reverse=True)[:8]:
        print(f"    {100 * v / tot:6.2f}%  {h.replace('smsp__pcsamp_warps_issue_stalled_', '')}")

args = ["--page", "source", "--csv"]
if kre:
    args += ["--kernel-name", "regex:" + kre]
src = run(args)
blocks = src.split('"Kernel Name",')
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
    kname = rows[0][1]
    h = rows[1]
    ix = {x: i for i, x in enumerate(h)}
    data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].isdigit()]
    ti = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    ts = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    print("=" * 100)
    print("SOURCE", kname[:140], " warp-instructions:", ti, " samples:", ts)
    ci, cs = Counter(), Counter()
    for r in data:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ci[op] += int(r[ix["Instructions Executed"]])
        cs[op] += int(r[ix["# Samples"]] or 0)
    print("  opcode mix (% of executed warp instructions, % of stall samples):")
    for op, v in ci.most_common(18):
        print(f"    {op:16s} {100 * v / ti:6.2f}%  {100 * cs[op] / ts:6.2f}%")
    print("  hottest SASS lines by samples:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:18]:
        print(f"    {100 * int(r[ix['# Samples']] or 0) / ts:6.2f}%  x{r[ix['Instructions Executed']]:>12s}  {r[ix['Source']][:90]}")
