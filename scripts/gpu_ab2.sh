#!/bin/bash
tag=${1:-ab2}; out=gpurun_out/$tag; mkdir -p $out
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1 NL_BENCH_SKIP_REFGPU=1 NL_BENCH_SKIP_CONFIGS=1
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 > $out/$name.json 2> $out/$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/$name.json")); print("$name", "ms/step", round(d["ms_per_step"],4), "median", round(d["steady_state"]["ms_median"],4), "frozen", round(d["frozen_decoder"]["ms_per_step"],4), "stages", {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e: print("$name failed", e)
PY
}
run base A=1
run lanes8_mid NL_TRAVERSE_LANES=8 NL_DW_RINGS=mid
run lanes8_deep NL_TRAVERSE_LANES=8
run lanes4_mid NL_DW_RINGS=mid
run sort_deep NL_PACKED_OCTREE=0
run lanes8_shallow NL_TRAVERSE_LANES=8 NL_DW_RINGS=shallow
run base2 A=1
