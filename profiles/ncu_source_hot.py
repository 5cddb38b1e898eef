#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from `ncu --page source --csv` (needs -lineinfo).
usage: python profiles/ncu_source_hot.py rep.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    fpath = ''
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            fpath = r[1]
        if len(r) > 8 and r[0] == 'Line No':
            hdr = r
            continue
        if hdr and len(r) >= 9 and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                inst = int(d.get('Instructions Executed') or 0)
                samp = int(d.get('# Samples') or 0)
            except ValueError:
                continue
            lines.append((inst, samp, fpath.split('/')[-1], int(r[0]), r[1].strip()[:110]))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print('total warp-instructions %d, samples %d' % (tot_i, tot_s))
    for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[1])[:top]:
        print('%5.1f%% samp %5.1f%% inst  %s:%d  %s' % (100.0 * samp / tot_s, 100.0 * inst / tot_i, f, ln, src))


if __name__ == '__main__':
    main()
