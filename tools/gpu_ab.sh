#!/bin/bash
# A/B of one environment switch on the headline config (run from the repo root on the GPU box):
#   tools/gpu_ab.sh "PGNN_TMA_STORE=0" ["OTHER=1" ...]   -> default build first, then each variant
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$tag.json").read().strip().splitlines()[-1])
    dt = d["detail"]
    print("$tag", round(d["value"]), "graphs/s mean", round(d["ms_per_step"], 4), "ms median", round(dt["per_step_ms"]["median"], 4),
          "e2e", round(d["e2e"]["value"]), "slowest", dt.get("slowest_steps"), dt.get("slowest_e2e_steps"))
    r = d["roofline"]
    print("   gemm family us/step", round(r.get("us_per_step", 0), 1), {k: v for k, v in list(r.get("step_kernels_us", {}).items())[:12]})
except Exception as e:
    print("$tag FAILED", e, open("gpurun_out/ab_$tag.err").read()[-800:])
PY
}
run default PGNN_DUMMY=1
i=0
for v in "$@"; do i=$((i+1)); run v$i $v; done
