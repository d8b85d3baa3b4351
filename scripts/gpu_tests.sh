#!/bin/bash
# GPU test-suite only (optionally a -k filter), each run under a hard timeout so a hung kernel cannot wedge the box.
tag=${1:-t}; shift
out=gpurun_out/$tag; mkdir -p $out
timeout ${TMO:-900} python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 "$@" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -${TAIL:-40} $out/pytest.log
