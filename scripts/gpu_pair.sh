#!/bin/bash
# NL_TC_PAIR=1 (CTA pairs sharing weight stages by multicast): correctness under a tight timeout first, then the A/B
tag=${1:-pair}; out=gpurun_out/$tag; mkdir -p $out
export NL_TC_PAIR=1
timeout 180 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider --timeout 120 -x -k "tensor_core or single_iteration or pipelined" > $out/pytest_tc.log 2>&1; rc=$?; echo "tc tests rc=$rc"; tail -6 $out/pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest_all.log 2>&1; echo "pipeline+configs rc=$?"; tail -4 $out/pytest_all.log
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1 NL_BENCH_SKIP_REFGPU=1 NL_BENCH_SKIP_CONFIGS=1
for mode in 1 0 1 0; do
  NL_TC_PAIR=$mode timeout 300 python bench.py --steps 40 --warmup 5 > $out/bench_pair$mode.json 2> $out/bench_pair$mode.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_pair$mode.json")); print("PAIR=$mode ms/step", round(d["ms_per_step"],4), "median", round(d["steady_state"]["ms_median"],4), "frozen", round(d["frozen_decoder"]["ms_per_step"],4), "mlp", round(d["stage_ms"]["mlp_fwd_bwd"],4), "frozen mlp", round(d["frozen_decoder"]["mlp_fwd_bwd_ms"],4), "loss", d["config"]["loss"])
except Exception as e: print("PAIR=$mode failed", e)
PY
done
