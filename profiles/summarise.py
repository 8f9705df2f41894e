#!/usr/bin/env python3
"""Condense one `tools/ncu_capture.sh` capture (ncu --set full of ONE launch, exported as CSV on the GPU box) into the
text summary that is committed under profiles/:  python profiles/summarise.py <prof_X_raw.csv> <prof_X_source.csv> > out.txt
  * the launch's headline metrics (duration, DRAM bytes, issue-slot / pipe utilisation, occupancy limits, stall reasons)
  * executed warp-instructions per SASS opcode, and grouped by how often a SASS line ran (= per granule / per tile /
    per row work), with the stall samples of each group
  * the hottest SASS lines by stall samples"""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = dict(zip(hdr, zip(units, vals)))
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print("== metrics (ncu --set full --clock-control none, one launch)")
for w in want:
    if w in M:
        print(f"{w:78s} {M[w][1]} {M[w][0]}")
print("== warp stall reasons per issued instruction (>= 0.15)")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            v = float(M[h][1].replace(",", ""))
        except ValueError:
            continue
        if v >= 0.15:
            print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.2f}")
rows = list(csv.reader(open(src)))
h2 = None; data = []
for r in rows:
    if r and r[0] == "Address":
        if h2 is not None:
            break
        h2 = r
        continue
    if h2 and len(r) == len(h2):
        data.append(r)
iS, iE, iN = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iN]) for r in data)
print(f"== SASS: {tot} warp-instructions executed, {totS} stall samples, {len(data)} lines")
op = collections.Counter(); ops = collections.Counter()
for r in data:
    t = r[iS].split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o] += int(r[iE]); ops[o] += int(r[iN])
for k, v in op.most_common(18):
    print(f"{k:30s} {v:12d} {100 * v / tot:5.1f}%   samples {100 * ops[k] / max(totS, 1):5.1f}%")
print("== SASS lines grouped by execution count (a group = code that runs once per granule / tile / row ...)")
band = collections.Counter(); lines = collections.Counter(); smp = collections.Counter()
for r in data:
    e = int(r[iE]); band[e] += e; lines[e] += 1; smp[e] += int(r[iN])
for e, v in sorted(band.items(), key=lambda x: -x[1])[:10]:
    if e:
        print(f"{e:10d} x {lines[e]:4d} lines = {100 * v / tot:5.1f}% of instructions, {100 * smp[e] / max(totS, 1):5.1f}% of samples")
print("== hottest lines by stall samples")
for r in sorted(data, key=lambda r: -int(r[iN]))[:10]:
    print(f"{int(r[iN]):7d} {int(r[iE]):10d}  {r[iS].strip()}")
