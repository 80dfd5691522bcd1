#!/usr/bin/env python
"""Summarise an ncu report (raw page + source page) into the few numbers the roofline discussion uses.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'smsp__cycles_active.avg',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for row in r[2:]:
    print("kernel:", row[h.index('Kernel Name')][:70])
    for w in want:
        if w in h:
            i = h.index(w)
            print(f"  {w:75s} {r[1][i]:12s} {row[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, x in enumerate(rows) if x and x[0] == "Address"][0]
h2, body = rows[hi], rows[hi + 1:]
ix = {c: i for i, c in enumerate(h2)}
stalls = [c for c in h2 if c.startswith('stall_') and 'Not Issued' not in c]
tot = {s: 0 for s in stalls}
for row in body:
    for s_ in stalls:
        try:
            tot[s_] += int(row[ix[s_]])
        except (ValueError, IndexError):
            pass
T = sum(tot.values()) or 1
print("stall samples (all):", ", ".join(f"{k[6:]} {100 * v / T:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:10]))
def num(row, c):
    try:
        return int(row[ix[c]])
    except (ValueError, IndexError):
        return 0
print("total warp instructions:", sum(num(x, 'Instructions Executed') for x in body), " samples:", sum(num(x, '# Samples') for x in body))
print(f"top {top} SASS lines by samples:  samples  executed  instruction")
for x in sorted(body, key=lambda x: -num(x, '# Samples'))[:top]:
    print(f"  {num(x, '# Samples'):8d} {num(x, 'Instructions Executed'):10d}  {x[ix['Source']][:100]}")
