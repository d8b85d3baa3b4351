#!/bin/bash
# Round-2 evidence visit: full GPU suite, smoke under compute-sanitizer, ncu launch list + full captures, final bench lines.
tag=${1:-r2p}
out=gpurun_out/$tag
mkdir -p $out
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1 NL_BENCH_SKIP_REFGPU=1 NL_BENCH_SKIP_CONFIGS=1
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > $out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY" $out/sanitizer.log | head -3
B="python bench.py --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv $B > $out/bench_under_ncu.json 2> $out/launches.err
echo "launch list rc=$? lines=$(wc -l < $out/launches.csv)"
for k in k_mlp_tc_train k_dw1_tc k_dw0_tc k_gather_fwd k_gather_bwd k_traverse_coop; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $out/$k $B > /dev/null 2> $out/$k.err
  echo "$k rc=$? $(ls -la $out/$k.ncu-rep 2>/dev/null | awk '{print $5}')"
done
# real-size iterations (2048 rays tracking, 5 x 2048 rays mapping): launch lists + the ray-selection kernel
IT=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_tracking.csv python scripts/profile_tracking.py > /dev/null 2>&1
IT=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_mapping.csv python scripts/profile_mapping.py > /dev/null 2>&1
IT=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_select_rays -s 4 -c 1 -f -o $out/k_select_rays python scripts/profile_tracking.py > /dev/null 2> $out/k_select_rays.err
echo "k_select_rays rc=$? $(ls -la $out/k_select_rays.ncu-rep 2>/dev/null | awk '{print $5}')"
unset NL_BENCH_SKIP_CPU NL_BENCH_SKIP_TRACKING NL_BENCH_SKIP_REFGPU NL_BENCH_SKIP_CONFIGS
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -3 $out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench.json"))
for k in ("value","ms_per_step","steady_state","real_size","tracking","configs"): print(k, json.dumps(d.get(k))[:1500])
print(json.dumps(d.get("reference_gpu",{}).get("speedup")))
r=json.load(open("$out/bench_ref.json")); print("ref", r["value"], r["ms_per_step"], r.get("reference_kind"))
PY
