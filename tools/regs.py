#!/usr/bin/env python
"""Registers / spills of every march-kernel instantiation: python tools/regs.py [-DNAME[=V] ...] [--out path.so]
Builds the library (or a tuning variant into --out) with -Xptxas -v and prints one line per kernel."""
import contextlib, io, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differender_b200.build import build_library
defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    path = build_library(force=True, verbose=True, defines=defs, out=os.path.abspath(out) if out else None)
log = buf.getvalue()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info    : Function properties for \S+\n\s+(\d+) bytes stack frame, "
                  r"(\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info    : Used (\d+) registers", log)
names = subprocess.run(['c++filt'] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
only = os.environ.get("REGS_ONLY", "")
for n, e in zip(names, ents):
    m = re.search(r'(fwd_kernel|bwd_kernel)<([^>]*)>', n)
    if m and only in m.group(0):
        print(f"{m.group(1)}<{m.group(2)}>  regs {e[4]}  spill st/ld {e[2]}/{e[3]}")
print(path)
