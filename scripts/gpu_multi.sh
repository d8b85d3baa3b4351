#!/bin/bash
# multi-GPU visit: bench.py under torchrun on N GPUs (weak scaling line + strong-scaling block + gradient equivalence)
N=${1:-2}; tag=${2:-mg$N}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/gpus.txt 2>&1
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench.json"))
for k in ("value","ms_per_step","steady_state","strong_scaling","e2e","clocks"): print(k, json.dumps(d.get(k)))
PY
tail -5 $out/bench.err
