#!/usr/bin/env python
"""Shared-memory wavefronts per source line (excessive = bank conflicts) from an .ncu-rep captured with --import-source on.
usage: python profiles/ncu_shared_conflicts.py rep.ncu-rep kernel-regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, lines, fname = None, [], ''
for r in rows:
    if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]
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
        lines.append((fname, int(r[0]), num('L1 Wavefronts Shared'), num('L1 Wavefronts Shared Excessive'), num('Instructions Executed'),
                      r[1].strip()[:100]))
tw = sum(l[2] for l in lines) or 1
te = sum(l[3] for l in lines)
print('shared wavefronts %d, excessive %d (%.1f%%)' % (tw, te, 100.0 * te / tw))
for f, ln, w, e, inst, src in sorted(lines, key=lambda x: -x[3])[:top]:
    print('%5.1f%% of excessive  wavefronts %9d excessive %9d  %s:%d  %s' % (100.0 * e / max(te, 1), w, e, f, ln, src))
