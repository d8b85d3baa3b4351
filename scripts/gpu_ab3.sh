#!/bin/bash
tag=${1:-ab3}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 300 -x > $out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 $out/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1 NL_BENCH_SKIP_REFGPU=1 NL_BENCH_SKIP_CONFIGS=1
for sp in 0 100 108 112 116 120 126 0 112; do
  NL_DW_SPLIT=$sp timeout 300 python bench.py --steps 40 --warmup 5 > $out/bench_$sp.json 2> $out/bench_$sp.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$sp.json")); print("split=$sp ms/step", round(d["ms_per_step"],4), "median", round(d["steady_state"]["ms_median"],4), "mlp(serial)", round(d["stage_ms"]["mlp_fwd_bwd"],4), "loss", round(d["config"]["loss"],6))
except Exception as e: print("split=$sp failed", e)
PY
done
