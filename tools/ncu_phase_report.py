#!/usr/bin/env python
"""Summarise an ncu report of the fused kernel: headline metrics + instruction / stall-sample share per
barrier-delimited phase (SASS source page).  usage: ncu_phase_report.py <report.ncu-rep> <n_clips> [kernel-name regex]"""
import csv
import io
import subprocess
import sys

rep, n_clips = sys.argv[1], int(sys.argv[2])
sel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, *sel, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in keys:
    if k in d:
        print(f"{k} = {d[k]}")
inst = float(d['smsp__inst_executed.sum'])
print(f"per clip: {inst / n_clips:.0f} warp instructions; dram read {float(d['dram__bytes_read.sum']) * 1e6 / n_clips:.0f} B "
      f"(unit {rows[1][rows[0].index('dram__bytes_read.sum')]})")
src = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# one section per kernel: a "Kernel Name" row, the header row, the instruction rows; take the first section that matches
import re
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if len(sys.argv) <= 3 or re.search(sys.argv[3], rows[i][1])), starts[0])
end = next((i for i in starts if i > pick), len(rows))
hdr = rows[pick + 1]
data = [r for r in rows[pick + 2:end] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
seg, cur = [], None


def new():
    return {'samp': 0, 'inst': 0, 'st': {k: 0 for k in stalls}, 'ops': {}}


cur = new()
for r in data:
    s = r[ix['Source']].strip()
    n, sm = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
    cur['samp'] += sm
    cur['inst'] += n
    for k in stalls:
        cur['st'][k] += int(r[ix[k]])
    op = (s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + n
    if 'BAR.SYNC' in s or 'SYNCS.PHASECHK' in s:
        seg.append(cur)
        cur = new()
seg.append(cur)
ti, ts = sum(s['inst'] for s in seg), sum(s['samp'] for s in seg)
print("\nphase (barrier-delimited)  inst/clip  inst%  samples%  top opcodes | top stall reasons")
for k, s in enumerate(seg):
    if s['inst'] < ti * 0.002 and s['samp'] < ts * 0.01:
        continue
    ops = " ".join(f"{o}:{100 * c / max(1, s['inst']):.0f}" for o, c in sorted(s['ops'].items(), key=lambda x: -x[1])[:6])
    st = " ".join(f"{o[6:]}:{100 * c / max(1, s['samp']):.0f}" for o, c in sorted(s['st'].items(), key=lambda x: -x[1])[:4])
    print(f"seg{k:2d} {s['inst'] / n_clips:9.0f} {100 * s['inst'] / ti:6.1f} {100 * s['samp'] / ts:8.1f}   {ops} | {st}")
