#!/bin/bash
# full ncu capture of the kernels matching $2.. (regex list) from a short bench run
tag=${1:-prof}; shift
out=gpurun_out/$tag; mkdir -p $out
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/$k python bench.py --steps 2 --warmup 3 > /dev/null 2> $out/$k.err
  echo "$k rc=$? $(ls -la $out/$k.ncu-rep 2>/dev/null | awk '{print $5}')"
done
