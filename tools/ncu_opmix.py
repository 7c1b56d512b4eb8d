#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from `ncu -i rep --page source --csv` (SASS view): warp-level executed counts per opcode."""
import csv, sys, collections, subprocess
rep, kern = sys.argv[1], sys.argv[2]
per = float(sys.argv[3]) if len(sys.argv) > 3 else None   # thread-samples in the launch, to print per-sample numbers
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]; si = h.index('Source'); ei = h.index('Instructions Executed'); ti = h.index('Thread Instructions Executed')
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_w = tot_t = 0
for r in rows[hi + 1:]:
    if len(r) <= ti or not r[ei] or not r[ei].replace(',', '').isdigit():
        continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0].rstrip(';')
    w, t = int(r[ei].replace(',', '')), int(r[ti].replace(',', ''))
    agg[op][0] += w; agg[op][1] += t; agg[op][2] += 1
    tot_w += w; tot_t += t
print(f"kernel {kern}: warp-instr {tot_w:.4g}, thread-instr {tot_t:.4g}, avg active lanes {tot_t / tot_w:.2f}")
if per:
    print(f"per thread-sample: {tot_t / per:.1f} thread-instr; per 32 samples: {32 * tot_w / per:.1f} warp-instr")
for op, (w, t, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:32]:
    line = f"  {op:12s} static {n:4d}  warp-exec {w:12d} ({100 * w / tot_w:5.1f}%)  lanes {t / max(w, 1):5.1f}"
    if per:
        line += f"  per-32-samples {32 * w / per:6.1f}"
    print(line)
