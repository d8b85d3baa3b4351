#!/bin/bash
tag=${1:-r2e}
out=gpurun_out/$tag
mkdir -p $out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -s > $out/pytest.log 2>&1; echo "rc=$?"; grep -a "gradient parity\|embedding update\|track_frame drop-in" $out/pytest.log; tail -15 $out/pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log
