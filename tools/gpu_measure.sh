#!/bin/bash
# Everything one gpurun call should bring back after the kernels changed (run from the repo root on the GPU box):
#   tools/gpu_measure.sh [quick]
mkdir -p gpurun_out
rm -rf gpurun_out/parity
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1
tail -12 gpurun_out/pytest_full.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# A/B of the round's scheduling changes on the headline config
for v in default nostream nostream_sepfold; do
  case $v in
    default) envs="";;
    nostream) envs="PGNN_WGRAD_STREAM=0";;
    nostream_sepfold) envs="PGNN_WGRAD_STREAM=0 PGNN_SPLITK_FOLD=separate";;
  esac
  env $envs timeout -s KILL 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_masking_$v.json 2> gpurun_out/bench_masking_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_masking_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["value"]), "graphs/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]), d["detail"]["per_step_ms"]["median"])
except Exception as e:
    print("$v FAILED", e, open("gpurun_out/bench_masking_$v.err").read()[-600:])
PY
done
[ "$1" = "quick" ] && exit 0
tools/run_all_configs.sh
# ncu: launch list of two masking steps, then a full capture of the GEMM family + gather of one step
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_masking.csv python tools/profile_step.py --config masking --steps 2 > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_gemm_3xtf32_tma|k_aggregate_fwd|k_aggregate_bwd|k_bn_bwd' -c 40 -f -o gpurun_out/prof_masking \
  python tools/profile_step.py --config masking --steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_masking.csv 2>&1 | tail -3
