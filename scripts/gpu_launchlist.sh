#!/bin/bash
tag=${1:-ll}; out=gpurun_out/$tag; mkdir -p $out
export NL_BENCH_SKIP_CPU=1 NL_BENCH_SKIP_TRACKING=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 > $out/bench_under_ncu.json 2> $out/launches.err
python - <<PY
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open("$out/launches.csv")) if len(r)>10]
h=rows[0]; ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
tot=defaultdict(float);cnt=defaultdict(int)
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    v*={"ns":1e-3,"us":1.0,"ms":1e3}.get(r[ui],1.0)
    n=r[ki].split("(")[0].replace("void ","").replace("<unnamed>::","")[:60]
    tot[n]+=v;cnt[n]+=1
for k in sorted(tot,key=lambda k:-tot[k])[:14]: print(f"{k:60s} n={cnt[k]:4d} avg={tot[k]/cnt[k]:9.1f} us")
PY
