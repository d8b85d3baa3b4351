#!/bin/bash
# quick visit: full GPU suite, then the bench line without the CPU / reference legs, then the SLAM demo
out=gpurun_out/quick; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x > $out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 $out/pytest.log
if [ $rc -ne 0 ]; then grep -n "^E " $out/pytest.log | head -20; fi
NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_REFGPU=1 timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
print("tracking", {k:round(v,2) for k,v in d["tracking"].items() if isinstance(v,float)})
for k,v in d["configs"].items():
    if k!="note": print(k, {kk: round(vv,2) for kk,vv in v.items() if kk.endswith("_ms") and not isinstance(vv, list)})
PY
timeout 300 python scripts/demo_slam.py 2>/dev/null | tail -1 | cut -c1-420
