#!/usr/bin/env python3
"""Per-source-line executed warp-instructions and stall samples from an .ncu-rep (needs -lineinfo and --import-source on).
usage: src_hot.py <file.ncu-rep> <kernel-name substring> [top]"""
import csv, io, subprocess, sys
def main(rep, kname, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    fpath = func = None; hdr = None; agg = {}; seen_funcs = set()
    for r in csv.reader(io.StringIO(out)):
        if not r: continue
        if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name': func = r[1]; continue
        if r[0] == 'Line No': hdr = r; continue
        if hdr is None or func is None or kname not in func or not r[0]: continue
        ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
        try: a = agg.setdefault((func[:60], fpath, int(r[0]), r[1].strip()[:110]), [0, 0]); a[0] += int(r[ie]); a[1] += int(r[ss])
        except ValueError: pass
    ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    print(f'{kname}: warp-instructions {ti}, stall samples {ts} (all launches in the report)')
    for (f, p, ln, src), (i, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f'{100*i/max(1,ti):6.2f}% inst {100*s/max(1,ts):6.2f}% stall  {p}:{ln}  {src}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
