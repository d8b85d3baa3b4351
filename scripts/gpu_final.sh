#!/bin/bash
tag=${1:-fin}; out=gpurun_out/$tag; mkdir -p $out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench.json"))
for k in ("value","ms_per_step","gpu_launches"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "clocks", d["clocks"])
r=json.load(open("$out/bench_ref.json")); print("ref", r["value"], r.get("reference_kind"), "ratio e2e", d["e2e"]["value"]/r["value"])
PY
