#!/bin/bash
# usage (on the GPU box): tools/gpu_multi_ab.sh N [config]  -- 2-GPU correctness test of every transport, then the bench at N GPUs per transport
N=${1:-2}
C=${2:-masking}
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_dist2.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/pytest_dist2.log; tail -8 gpurun_out/pytest_dist2.log
for mode in fused nvls p2p; do
  PGNN_ALLREDUCE=$mode timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --config $C --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${C}_n${N}_$mode.json 2> gpurun_out/bench_${C}_n${N}_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${C}_n${N}_$mode.json").read().strip().splitlines()[-1])
    dt = d["detail"]
    print("$C N=$N $mode:", round(d["value"]), "graphs/s", round(d["ms_per_step"], 4), "ms (median %.4f)" % dt["per_step_ms"]["median"], " e2e", round(d["e2e"]["value"]),
          round(d["e2e"]["ms_per_step"], 4), "(median %.4f)" % dt["per_step_ms_e2e"]["median"], dt["grad_allreduce"], dt.get("slowest_steps"), dt.get("slowest_e2e_steps"))
except Exception as e:
    print("$C N=$N $mode FAILED", e, open("gpurun_out/bench_${C}_n${N}_$mode.err").read()[-1500:])
PY
done
