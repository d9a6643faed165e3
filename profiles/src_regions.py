#!/usr/bin/env python3
"""Executed warp-instructions of a kernel aggregated over source regions given as marker substrings (first matching line at or after the
kernel's definition starts the region).  usage: src_regions.py <file.ncu-rep> <kernel substring> <source file> marker1 marker2 ..."""
import csv, io, subprocess, sys
rep, kname, srcfile = sys.argv[1:4]
markers = sys.argv[4:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
hdr = func = fpath = None; agg = {}; stall = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or func is None or kname not in func or not r[0]: continue
    ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
    try:
        k = (fpath, int(r[0])); agg[k] = agg.get(k, 0) + int(r[ie]); stall[k] = stall.get(k, 0) + int(r[ss] or 0)
    except ValueError: pass
src = open(srcfile).read().split('\n')
base = srcfile.split('/')[-1]
start = next(i + 1 for i, l in enumerate(src) if kname in l and '__global__' in ''.join(src[max(0, i - 2):i + 1]))
marks = [('(kernel head)', start)]
for m in markers:
    ln = next((i + 1 for i, l in enumerate(src) if i + 1 > marks[-1][1] and m in l), None)
    if ln: marks.append((m[:50], ln))
tot = sum(agg.values()); ts = sum(stall.values())
reg = {}; rs = {}
for (f, l), v in agg.items():
    name = 'other files (inlined helpers)'
    if f == base:
        name = '(before kernel)'
        for k in range(len(marks)):
            if l >= marks[k][1]: name = marks[k][0]
    reg[name] = reg.get(name, 0) + v; rs[name] = rs.get(name, 0) + stall.get((f, l), 0)
print(f'{kname}: {tot} warp-instructions, {ts} stall samples (all launches in the report)')
for k, v in sorted(reg.items(), key=lambda x: -x[1]): print(f'{100*v/tot:6.2f}% inst {100*rs[k]/max(1,ts):6.2f}% stall  {k}')
