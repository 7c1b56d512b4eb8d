#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh "<bench args>" tag1=path1.so tag2=path2.so ...   (path "-" = the in-tree build)
args="$1"; shift
for kv in "$@"; do
  tag="${kv%%=*}"; lib="${kv#*=}"
  if [ "$lib" = "-" ]; then unset DIFFRENDER_LIB; else export DIFFRENDER_LIB="$PWD/$lib"; fi
  python bench.py --no-e2e --no-cpu-baseline $args > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python tools/bench_line.py "$tag" gpurun_out/ab_$tag.json
done
