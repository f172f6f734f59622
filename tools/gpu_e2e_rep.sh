#!/bin/bash
# repeat the default bench a few times and print the slowest steps of both loops (host-stall hunting)
mkdir -p gpurun_out
for r in 1 2 3 4; do
  for v in "PGNN_DUMMY=1" "PGNN_BENCH_NO_CLOCKS=1"; do
    env $v timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/rep.json 2> gpurun_out/rep.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/rep.json").read().strip().splitlines()[-1]); dt = d["detail"]
    print("$v run $r: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "slowest", dt["slowest_steps"][:2], "e2e", dt["slowest_e2e_steps"][:2])
except Exception as e:
    print("FAILED", e, open("gpurun_out/rep.err").read()[-500:])
PY
  done
done
