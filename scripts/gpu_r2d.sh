#!/bin/bash
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
echo "== debug ba"; timeout 900 python scripts/debug_ba.py 2>&1 | grep -v Warning | grep -v "^  " > $out/debug_ba.log; echo "rc=$?"; tail -14 $out/debug_ba.log
echo "== new tests"; timeout 1500 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 600 -x > $out/pytest_new.log 2>&1; echo "rc=$?"; tail -25 $out/pytest_new.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -12 $out/pytest.log
