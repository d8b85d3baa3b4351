#!/bin/bash
# One GPU-box visit: sanitizer smoke, GPU test-suite, short bench.  Outputs under gpurun_out/$1/.
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $out/smoke.log
if [ "${SANITIZE:-1}" = "1" ]; then
  echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > $out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" $out/sanitizer.log | head -20
fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -60 $out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -5 $out/bench.err
