#!/bin/bash
tag=${1:-scan}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 300 -x > $out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 $out/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1 NL_BENCH_SKIP_REFGPU=1 NL_BENCH_SKIP_CONFIGS=1
for mode in chained single chained single; do
  NL_SCAN=$mode timeout 300 python bench.py --steps 40 --warmup 5 > $out/bench_$mode.json 2> $out/bench_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$mode.json")); print("$mode ms/step", round(d["ms_per_step"],4), "median", round(d["steady_state"]["ms_median"],4), "frozen", round(d["frozen_decoder"]["ms_per_step"],4), "front", round(d["stage_ms"]["traverse_sample"],4))
except Exception as e: print("$mode failed", e)
PY
done
