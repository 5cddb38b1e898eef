#!/usr/bin/env python
"""Instruction count of merge_fast_kernel per phase (source markers "---- X." in kernels.cuh) from an .ncu-rep with source.
usage: python profiles/ncu_phase_breakdown.py rep.ncu-rep n_particles"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_source_hot as H

rep, npart = sys.argv[1], float(sys.argv[2])
lines = H.load(rep, 'merge_fast')
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'cuda-phdslam_b200', 'csrc', 'kernels.cuh')).read().split('\n')
k0 = [i + 1 for i, l in enumerate(src) if 'merge_fast_kernel(MrgArgs a)' in l][0]
k1 = [i + 1 for i, l in enumerate(src) if 'particle weights: w += dw' in l][0]
marks = [(l.strip()[3:60], i + 1) for i, l in enumerate(src) if k0 <= i + 1 < k1 and l.strip().startswith('/* ---- ')]
marks.append(('end', k1))
r0 = [i + 1 for i, l in enumerate(src) if 'block_radix_pass(const unsigned* keys' in l][0]
ti = sum(l[2] for l in lines)
ts = sum(l[3] for l in lines)
b = {}
for f, ln, inst, samp, _ in lines:
    name = 'helpers / other files'
    if f == 'kernels.cuh':
        if r0 <= ln < k0:
            name = 'block_radix_pass'
        for (nm, l0), (_, l1) in zip(marks, marks[1:]):
            if l0 <= ln < l1:
                name = nm
    x = b.setdefault(name, [0, 0])
    x[0] += inst
    x[1] += samp
for k, v in sorted(b.items(), key=lambda x: -x[1][0]):
    print('%-60s %5.1f%% inst %5.1f%% samples %8.0f warp-instr/particle' % (k, 100.0 * v[0] / ti, 100.0 * v[1] / ts, v[0] / npart))
print('total warp-instructions per particle %.0f' % (ti / npart))
