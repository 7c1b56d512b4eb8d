#!/usr/bin/env python
"""Warp-instructions executed per CUDA source line of one kernel, from `ncu -i rep --page source --csv --print-source sass,cuda`.
usage: ncu_lines.py rep kernel-regex [thread-samples-in-the-capture] [top-N]
With -lineinfo everything inlined into the kernel is attributed to its innermost source line (dr_math.cuh)."""
import csv, subprocess, sys, os, collections
rep, kern = sys.argv[1], sys.argv[2]
per = float(sys.argv[3]) if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None
ei = ti = None
cur = None
# A SASS instruction is listed under every CUDA line of its inline stack (and once per captured launch): keep ONE line per
# address -- the innermost, i.e. the one in dr_math.cuh / a CUDA header if there is one, else the first listed.
per_addr = collections.OrderedDict()      # address -> {(file, line, src): [warp, thread]}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        ei, ti = r.index('Instructions Executed'), r.index('Thread Instructions Executed'); continue
    if r[0] and r[0].isdigit():
        cur = (fname, int(r[0]), r[1].strip()); continue
    if not r[0] and cur is not None and ei is not None and len(r) > ti and r[2].startswith('0x') and r[ei].replace(',', '').isdigit():
        d = per_addr.setdefault(r[2], collections.OrderedDict()).setdefault(cur, [0, 0])
        d[0] += int(r[ei].replace(',', '')); d[1] += int(r[ti].replace(',', ''))
merged = collections.OrderedDict()
for addr, d in per_addr.items():
    keys = list(d)
    inner = [k for k in keys if k[0] != 'diffrender.cu']
    k = inner[-1] if inner else keys[0]
    m = merged.setdefault(k[:2], [k[0], k[1], k[2], 0, 0])
    m[3] += d[k][0]; m[4] += d[k][1]
lines = [tuple(v) for v in merged.values()]
tot = sum(l[3] for l in lines)
print(f"kernel {kern}: {tot:.4g} warp-instructions attributed to {len(lines)} source lines")
byfile = collections.defaultdict(int)
for l in lines:
    byfile[l[0]] += l[3]
for f, w in sorted(byfile.items(), key=lambda x: -x[1]):
    print(f"  {f:28s} {100 * w / tot:5.1f}%")
print("in source order (lines >= 0.3 %):")
for f, ln, src, w, t in sorted(lines, key=lambda l: (l[0], l[1])):
    if w / tot < float(os.environ.get("NCU_LINES_MIN", "0.003")):
        continue
    extra = f"  per-32-samples {32 * w / per:6.1f}" if per else ""
    print(f"  {f}:{ln:<4d} {100 * w / tot:5.1f}%  lanes {t / max(w, 1):4.1f}{extra}  | {src[:110]}")
