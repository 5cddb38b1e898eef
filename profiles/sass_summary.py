#!/usr/bin/env python
"""Instruction mix of the shipped sm_100a kernels (cuobjdump -sass of libphdslam.so), and an excerpt of the inner
(component pair, measurement) loop of the dense update kernel.  usage: python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cuda-phdslam_b200", "libphdslam.so")
head = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("# cuobjdump -lelf:", head.strip().replace("\n", " | "))
arch = re.findall(r"arch = (sm_\w+)", sass)
print("# architectures of the embedded cubins:", sorted(set(arch)))
funcs, cur = collections.OrderedDict(), None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(2).strip())
want = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED",
        "SHFL", "REDUX", "MATCH", "VOTE", "BAR", "WARPSYNC", "UTMALDG", "UTMASTG", "HMMA", "UTCMMA"]
print("\n%-62s %6s  " % ("kernel", "instr") + " ".join("%6s" % w for w in want))
for name, ins in funcs.items():
    ops = collections.Counter()
    for i in ins:
        t = i.split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.split(".")[0]] += 1
    print("%-62s %6d  " % (name[:62], len(ins)) + " ".join("%6d" % ops.get(w, 0) for w in want))
print("\nNo tensor-core (HMMA / UTCMMA) and no TMA (UTMALDG / UTMASTG) instructions: the work is 2x2 EKF algebra per term (FFMA2/FMUL2/FADD2 = Blackwell's"
      "\npacked fp32x2) and the dense output leaves as coalesced 64-bit streaming stores, 1792 contiguous bytes per warp iteration (DESIGN.md section 4).")
name = [n for n in funcs if n.startswith("void update_kernel<true, false>")][0]
ins = funcs[name]
# the inner loop of pass 2: the longest run of instructions between two backward branches that contains FFMA2 and STG
idx = [k for k, i in enumerate(ins) if "STG.E.EF.64" in i or ("STG" in i and ".64" in i)]
if idx:
    lo, hi = max(idx[0] - 45, 0), min(idx[0] + 25, len(ins))
    print("\n# %s: around the first 64-bit streaming store of pass 2 (instructions %d-%d of %d)" % (name, lo, hi, len(ins)))
    for i in ins[lo:hi]:
        print("    " + i)
