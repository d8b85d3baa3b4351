#!/bin/bash
# Round-2 visit A: new drop-in-with-reference-objects tests first (fast signal), then the whole GPU suite, then bench (ours + reference arm).
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
free -g > $out/mem.txt 2>&1; nproc >> $out/mem.txt
echo "== dropin/ref tests"; timeout 900 python -m pytest tests/test_gpu_dropin_reference.py -m gpu -q -p no:cacheprovider --timeout 600 -x -s > $out/pytest_dropin.log 2>&1; echo "rc=$?"; tail -40 $out/pytest_dropin.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 $out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -5 $out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps ${STEPS:-20} --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json; tail -5 $out/bench_ref.err
