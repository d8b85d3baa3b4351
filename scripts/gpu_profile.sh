#!/bin/bash
# ncu pass over bench.py: launch list (per-launch device time) + full captures of the main kernels.
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1
B="python bench.py --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv $B > $out/bench_under_ncu.json 2> $out/launches.err
echo "launch list rc=$? lines=$(wc -l < $out/launches.csv)"
for k in k_mlp_tc_train k_dw1_tc k_dw0_tc k_mask_colsum k_gather_fwd k_gather_bwd k_traverse_coop k_sample k_scan; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/$k $B > /dev/null 2> $out/$k.err
  echo "$k rc=$? $(ls -la $out/$k.ncu-rep 2>/dev/null | awk '{print $5}')"
done
