#!/bin/bash
# tools/repeat_steps.sh N "<bench args>": N separate processes of the in-tree library, device time of every step (spread / one-off hiccups)
n="$1"; args="$2"
for r in $(seq 1 $n); do
  python bench.py --no-e2e --no-cpu-baseline --no-others $args > gpurun_out/rep_$r.json 2> gpurun_out/rep_$r.err
  python - "$r" <<'PY'
import json, sys
r = sys.argv[1]
l = json.loads([x for x in open(f"gpurun_out/rep_{r}.json") if x.startswith("{")][-1])
print(f"run {r}: value {l['value']:.2f}  steps ms {l.get('ms_each_step_rank0')}  clocks {l['clocks'].get('sm_mhz')} samples {l['clocks'].get('samples')}")
PY
done
