#!/usr/bin/env python
"""Generates tests/golden/ref_cphd_golden.npz from the REFERENCE's own CPHD kernels (see tests/ref_cases.py CPHD_CASES).

Run in a container that has /root/reference:   bash oracle/ref_build.sh && python tests/golden/make_ref_cphd_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_cases as RC  # noqa: E402


def main():
    out = {}
    for name in RC.CPHD_CASES:
        r = RC.reference_cphd(name)
        assert np.isfinite(r["ip0"]).all() and np.isfinite(r["detect"]["weight"]).all(), name
        for k, v in r.items():
            out["%s/%s" % (name, k)] = v
        print("%-16s M %2d  log<Psi0,p> %s" % (name, len(r["Z"]), r["ip0"]))
    path = os.path.join(ROOT, "tests", "golden", "ref_cphd_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
