#!/usr/bin/env python
"""Static SASS opcode histogram of the hot march kernels of libdiffrender.so (cuobjdump -sass), for profiles/.
usage: sass_hist.py [lib.so] [kernel-substring ...]   default: the C3 kernels fwd_kernel<float,2,0,0,1,1> and bwd_kernel<float,2,0,1,1,1>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "differender_b200", "libdiffrender.so")
want = sys.argv[2:] or ["fwd_kernelIfLi2ELb0ELi0ELb1ELb1E", "bwd_kernelIfLi2ELi0ELb1ELb1ELb1E"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = next((w for w in want if w in m.group(1)), None)
        if cur:
            hist.setdefault(cur, collections.Counter())
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            hist[cur][m.group(1)] += 1
for k, c in hist.items():
    tot = sum(c.values())
    name = subprocess.run(["c++filt", k if k.startswith("_Z") else "_ZN2dr" + k], capture_output=True, text=True).stdout.strip()
    print(f"== {k}: {tot} SASS instructions (static)")
    groups = collections.Counter()
    for op, n in c.items():
        groups[op.split(".")[0]] += n
    for op, n in groups.most_common(40):
        print(f"   {op:14s} {n:6d}  {100 * n / tot:5.1f}%")
    packed = sum(n for op, n in c.items() if op.split(".")[0] in ("FFMA2", "FMUL2", "FADD2"))
    red = sum(n for op, n in c.items() if op.startswith("RED"))
    spill = sum(n for op, n in c.items() if op.split(".")[0] in ("STL", "LDL"))
    print(f"   packed fp32 (FFMA2/FMUL2/FADD2): {packed}; RED.*: {red} ({[op for op in c if op.startswith('RED')]}); local-memory (spill) instructions: {spill}")
