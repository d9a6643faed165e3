#!/usr/bin/env python3
"""Join an ncu SASS-level source page (csv) with `nvdisasm -g` line info: executed warp-instructions and stall samples per CUDA source line.
usage: sass_by_line.py <ncu_source_page.csv> <nvdisasm -g -c output> <kernel mangled-name substring> [top]"""
import csv, re, sys
def main(csvp, sassp, kname, top=30):
    rows = list(csv.reader(open(csvp)))
    hdr = rows[1]
    ie, ss, src = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
    data = []
    for r in rows[2:]:
        try: data.append((int(r[ie]), int(r[ss]), r[src].strip()))
        except Exception: pass
    # nvdisasm: collect (line, opcode) for the kernel's section
    seq, cur, on = [], None, False
    for ln in open(sassp):
        if ln.startswith('.text.') or ln.startswith('\t.section\t.text.') or ln.lstrip().startswith('.section'):
            on = kname in ln
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', ln)
        if m: seq.append((cur, m.group(1).strip()))
    n = len(seq)
    # the ncu page may hold several launches of the kernel back to back
    reps = max(1, len(data) // n) if n else 0
    if n == 0 or len(data) % n: print('warning: instruction count mismatch', len(data), n)
    agg = {}
    ti = ts = 0
    for k in range(min(n, len(data))):
        key = seq[k][0]
        a = agg.setdefault(key, [0, 0])
        a[0] += data[k][0]; a[1] += data[k][1]; ti += data[k][0]; ts += data[k][1]
    print(f'kernel instructions {ti}, stall samples {ts}, sass {n}, launches in page {reps}')
    for key, (i, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f'{100*i/ti:6.2f}% inst {100*s/max(1,ts):6.2f}% stall  {key}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 30)
