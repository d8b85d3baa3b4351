#!/bin/bash
tag=${1:-sel}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_dropin_reference.py -m gpu -q -p no:cacheprovider --timeout 300 -x > $out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 $out/pytest.log
if [ $rc -ne 0 ]; then grep -n "^E " $out/pytest.log | head -20; exit 0; fi
NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_REFGPU=1 timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"])
print("tracking", {k:v for k,v in d["tracking"].items() if k!="note"})
for k,v in d["configs"].items():
    if k!="note": print(k, {kk: round(vv,2) for kk,vv in v.items() if kk.endswith("_ms") and not isinstance(vv, list)})
PY
python - <<PY
import sys, torch, importlib
sys.path.insert(0, ".")
nl = importlib.import_module("nerf-loam_b200")
from importlib import import_module
rh = import_module("nerf-loam_b200.render_helpers")
for F in (1, 5):
    cap = 100000
    d = torch.randn(F, cap, 3, device="cuda"); g = torch.rand(F, cap, device="cuda"); c = torch.rand(F, cap, device="cuda")
    seed = torch.tensor([12345], dtype=torch.int32, device="cuda")
    out = rh.select_rays_device(d, g, c, 2048, seed=seed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): rh.select_rays_device(d, g, c, 2048, seed=seed, out=out)
    e1.record(); torch.cuda.synchronize()
    print("select_rays F=%d n=%d: %.1f us/launch" % (F, cap, e0.elapsed_time(e1) * 5))
PY
timeout 300 python scripts/demo_slam.py 2>/dev/null | tail -1 | cut -c1-700
for env in "A=1" "A=2" "A=3" "NL_MAP_GRAPH=0" "NL_MAP_GRAPH=0"; do
  echo -n "demo 6 scans $env: "; env $env timeout 300 python scripts/demo_slam.py --scans 6 --init-calls 8 2>&1 | tail -1 | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s); print(d['per_scan_translation_error_m'])
except Exception: print('ERR', s[-200:])"
done
