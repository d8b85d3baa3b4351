#!/bin/bash
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
echo "== debug track"; timeout 600 python scripts/debug_track.py > $out/debug_track.log 2>&1; echo "rc=$?"; tail -60 $out/debug_track.log
echo "== dropin/ref tests"; timeout 900 python -m pytest tests/test_gpu_dropin_reference.py tests/test_gpu_ops.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 600 -s > $out/pytest_a.log 2>&1; echo "rc=$?"; tail -30 $out/pytest_a.log
echo "== grad parity test"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider --timeout 600 -s -k "single_iteration" > $out/pytest_b.log 2>&1; echo "rc=$?"; grep -a "gradient parity" $out/pytest_b.log; tail -5 $out/pytest_b.log
echo "== bench"; NL_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$out/bench.json'))
for k in ('value','ms_per_step','steady_state','real_size','reference_gpu','tracking'): print(k, json.dumps(d.get(k)))"; tail -5 $out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json; tail -5 $out/bench_ref.err
