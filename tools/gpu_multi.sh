#!/bin/bash
# usage (on the GPU box): tools/gpu_multi.sh N   -- 2-GPU correctness test + bench lines at N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_dist2.py -m gpu -q > gpurun_out/pytest_dist2.log 2>&1; tail -3 gpurun_out/pytest_dist2.log
for c in masking bio_supervised gcn contextpred; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --config $c --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${c}_n$N.json 2> gpurun_out/bench_${c}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${c}_n$N.json").read().strip().splitlines()[-1])
    print("$c N=$N", round(d["value"]), "graphs/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), d["detail"]["grad_allreduce"], d["detail"]["per_step_ms_e2e"])
except Exception as e:
    print("$c N=$N FAILED", e, open("gpurun_out/bench_${c}_n$N.err").read()[-800:])
PY
done
