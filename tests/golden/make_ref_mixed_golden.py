#!/usr/bin/env python
"""Generates tests/golden/ref_mixed_golden.npz from the REFERENCE's own mixed-feature-model device code
(predictMapKernelMixed, computeBirth / computePreUpdate on Gaussian4D, phdUpdateKernelMixed, computeMahalDist(Gaussian4D),
phdUpdateMergeKernel<Gaussian4D>: src/phdfilter.cu:244-521,910-963,2323-2635,2707-2898), run through the CUDA emulator on
the cases of tests/mixed_cases.py.

Run in a container that has /root/reference:   bash oracle/ref_build.sh && python tests/golden/make_ref_mixed_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mixed_cases as MC  # noqa: E402


def main():
    out = {}
    for name in MC.MIXED_CASES:
        r = MC.reference_case(name)
        assert np.isfinite(r["d_terms"]["weight"]).all() and np.isfinite(r["s_terms"]["weight"]).all(), name
        for k, v in r.items():
            out["%s/%s" % (name, k)] = v
        print("%-14s static %3d -> %3d terms -> %3d merged   dynamic %3d -> %3d terms -> %3d merged   dlogw %s" % (
            name, len(r["smap"]), len(r["s_terms"]), len(r["s_merged"]), len(r["dmap"]), len(r["d_terms"]), len(r["d_merged"]),
            r["dlogw"]))
    path = os.path.join(ROOT, "tests", "golden", "ref_mixed_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
