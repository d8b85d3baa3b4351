#!/bin/bash
tag=${1:-r2c}
out=gpurun_out/$tag
mkdir -p $out
echo "== debug track"; timeout 600 python scripts/debug_track.py 2>&1 | grep -v Warning > $out/debug_track.log; echo "rc=$?"; tail -22 $out/debug_track.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -s > $out/pytest.log 2>&1; echo "rc=$?"; grep -a "gradient parity\|embedding update" $out/pytest.log; tail -30 $out/pytest.log
echo "== bench"; NL_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$out/bench.json'))
for k in ('value','ms_per_step','steady_state','real_size','tracking','stage_ms'): print(k, json.dumps(d.get(k)))
print(json.dumps(d.get('reference_gpu',{}).get('speedup')))"; tail -5 $out/bench.err
