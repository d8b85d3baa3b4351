#!/bin/bash
# A/B on N GPUs: fused peer-memory exchange (default) vs plain NCCL (NL_PEER=0)
N=${1:-2}; tag=${2:-ab$N}
out=gpurun_out/$tag; mkdir -p $out
export NCCL_DEBUG=WARN NL_BENCH_SKIP_TRACKING=1
for mode in 1 0; do
  NL_PEER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > $out/bench_peer$mode.json 2> $out/bench_peer$mode.err; echo "NL_PEER=$mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_peer$mode.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "median", d["steady_state"]["ms_median"], "max", d["steady_state"]["ms_max"], d["config"].get("multi_gpu_exchange"))
    s=d.get("strong_scaling",{}); print("strong", s.get("value"), s.get("ms_per_step"), s.get("grad_equiv_max_rel"))
except Exception as e: print("parse error", e)
PY
  tail -3 $out/bench_peer$mode.err | cut -c1-300
done
