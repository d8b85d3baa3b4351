#!/bin/bash
# A/B visit: GPU suite with the default path, then with NL_TC_TS=1, then short benches of both (CPU baseline skipped).
tag=${1:-ab}; out=gpurun_out/$tag; mkdir -p $out
export NL_BENCH_SKIP_CPU=1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest_default.log 2>&1; echo "pytest default rc=$?"; tail -3 $out/pytest_default.log
NL_TC_TS=0 timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -x > $out/pytest_ts.log 2>&1; echo "pytest TS=0 rc=$?"; tail -15 $out/pytest_ts.log
for ts in 0 1; do
  NL_TC_TS=$ts timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench_ts$ts.json 2> $out/bench_ts$ts.err; echo "bench ts=$ts rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_ts$ts.json")); print("ts=$ts", round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],4), "ms", d["stage_ms"], "frozen", round(d["frozen_decoder"]["ms_per_step"],4), d["frozen_decoder"]["mlp_fwd_bwd_ms"], "track", d["tracking"].get("ms_per_scan"))
except Exception as e: print("ts=$ts bench parse failed", e)
PY
done
NL_TC_TS=1 NL_TC_TIMELINE=1 timeout 300 python bench.py --steps 4 --warmup 3 > /dev/null 2> $out/timeline_ts1.txt; grep -c TL $out/timeline_ts1.txt
