#!/usr/bin/env python
"""Summarises an `ncu --page source --csv` export of one kernel: warp-sample share and executed-instruction mix per
barrier-delimited segment, plus the top stalled instructions.  usage: ncu_phases.py src.csv [raw.csv]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr, data = rows[hi], rows[hi + 1:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[isamp] or 0) for r in data)
totex = sum(int(r[iex] or 0) for r in data)
print("instructions", len(data), "samples", tot, "warp-instructions executed", totex)
FP = ('DFMA', 'DMUL', 'DADD', 'DSETP')
seg_start, acc, accex, fp = 0, 0, 0, 0
mix = collections.Counter()
for k, r in enumerate(data):
    s, ex = int(r[isamp] or 0), int(r[iex] or 0)
    src = r[isrc].strip()
    op = re.sub(r'^@!?U?P\d+\s+', '', src).split()[0].split('.')[0]
    acc += s; accex += ex; mix[op] += ex
    if op in FP: fp += ex
    if any(m in src for m in ('BAR.SYNC', 'BAR.RED', 'SYNCS.PHASECHK', 'UTCBAR', 'ATOMS', 'EXIT')) or k == len(data) - 1:
        if acc > tot * 0.004:
            print(f"{seg_start:5d}-{k:5d} samples {100*acc/tot:5.1f}%  instrs {100*accex/totex:5.1f}%  fp64 {100*fp/max(accex,1):4.0f}%  "
                  f"exec/instr {accex//max(k-seg_start+1,1):>10d}  ends: {src[:44]}")
        seg_start, acc, accex, fp = k + 1, 0, 0, 0
print("mix:", ", ".join(f"{m} {100*c/totex:.1f}%" for m, c in mix.most_common(14)))
top = sorted(((int(r[isamp] or 0), k, r[isrc].strip()[:60]) for k, r in enumerate(data)), reverse=True)[:12]
for s, k, src in top: print(f"  {k:5d} {100*s/tot:5.2f}% {src}")
if len(sys.argv) > 2:
    rr = list(csv.reader(open(sys.argv[2])))
    h, v = rr[0], rr[2]
    for name in ('gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
                 'sm__inst_executed.sum.pct_of_peak_sustained_elapsed', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
                 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum'):
        if name in h: print(name, v[h.index(name)], rr[1][h.index(name)])
    st = sorted(((float(v[i].replace(',', '')), x) for i, x in enumerate(h) if x.endswith('_per_issue_active.ratio') and 'stalled' in x), reverse=True)[:8]
    print("stalls per issue:", ", ".join(f"{x.split('stalled_')[1].split('_per')[0]} {a:.2f}" for a, x in st))
