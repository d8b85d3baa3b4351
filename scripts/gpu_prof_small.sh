#!/bin/bash
out=gpurun_out/small; mkdir -p $out
IT=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/track.csv python scripts/profile_tracking.py > /dev/null 2>&1
IT=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/map.csv python scripts/profile_mapping.py > /dev/null 2>&1
IT=3 FREEZE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/map_frozen.csv python scripts/profile_mapping.py > /dev/null 2>&1
for l in 4 8; do
NL_TRAVERSE_LANES=$l NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_REFGPU=1 timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_l$l.json 2> $out/bench_l$l.err
python - <<PY
import json
d=json.load(open("$out/bench_l$l.json"))
print("lanes $l: value", d["value"], "ms", d["ms_per_step"], "tracking", d["tracking"]["ms_per_scan_cuda_graph"], d["tracking"]["ms_per_scan_eager"])
for k,v in d["configs"].items():
    if k!="note": print(" ", k, {kk: round(vv,2) for kk,vv in v.items() if kk.endswith("_ms") and not isinstance(vv, list)})
PY
done
