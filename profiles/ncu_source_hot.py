#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo).
usage: python profiles/ncu_source_hot.py rep.ncu-rep [kernel-regex] [top_n] [inst|samp]"""
import csv
import io
import subprocess
import sys


def load(rep, kernel=None):
    cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']
    if kernel:
        cmd += ['-k', 'regex:' + kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, lines, fpath = None, [], ''
    for r in rows:
        if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
            fpath = r[1].split('/')[-1]
        if len(r) > 8 and r[0] == 'Line No':
            hdr = r
            continue
        if hdr and len(r) >= 9 and r[0].isdigit():
            d = dict(zip(hdr, r))

            def num(k):
                try:
                    return int(d.get(k) or 0)
                except ValueError:
                    return 0
            lines.append((fpath, int(r[0]), num('Instructions Executed'), num('# Samples'), r[1].strip()[:105]))
    return lines


def main():
    rep = sys.argv[1]
    kernel = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    key = 2 if (len(sys.argv) > 4 and sys.argv[4] == 'inst') else 3
    lines = load(rep, kernel)
    ti = sum(l[2] for l in lines) or 1
    ts = sum(l[3] for l in lines) or 1
    print('total warp-instructions %d, samples %d' % (ti, ts))
    for f, ln, inst, samp, src in sorted(lines, key=lambda x: -x[key])[:top]:
        print('%5.1f%% samp %5.1f%% inst  %s:%d  %s' % (100.0 * samp / ts, 100.0 * inst / ti, f, ln, src))


if __name__ == '__main__':
    main()
