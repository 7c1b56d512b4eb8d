#!/bin/bash
# tools/ab_repeat.sh "<bench args>" N tag1=lib1.so tag2=lib2.so ...: N interleaved rounds of separate processes per library (run-to-run spread)
args="$1"; n="$2"; shift; shift
for r in $(seq 1 $n); do
  for kv in "$@"; do
    tag="${kv%%=*}"; lib="${kv#*=}"
    if [ "$lib" = "-" ]; then unset DIFFRENDER_LIB; else export DIFFRENDER_LIB="$PWD/$lib"; fi
    python bench.py --no-e2e --no-cpu-baseline --no-others $args > gpurun_out/abr_${tag}_$r.json 2> gpurun_out/abr_${tag}_$r.err
    python tools/bench_line.py "$tag#$r" gpurun_out/abr_${tag}_$r.json
  done
done
