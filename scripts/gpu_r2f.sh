#!/bin/bash
tag=${1:-r2f}; out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_peer.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -15 $out/pytest.log
