#!/bin/bash
# One gpurun call that brings back everything the round's summaries are made from (run from the repo root on the GPU box):
#   tools/gpu_final.sh
mkdir -p gpurun_out
rm -rf gpurun_out/parity
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1
tail -6 gpurun_out/pytest_full.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
# bench lines of all six BASELINE configs, CPU reference arm included (bench.py bounds it)
for c in masking contextpred bio_supervised gcn gat graphsage; do
  timeout -s KILL 500 python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["value"]), "graphs/s", round(d["ms_per_step"], 4), "ms (median %.4f)" % d["detail"]["per_step_ms"]["median"], " e2e", round(d["e2e"]["value"]),
          " cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"), d.get("cpu_baseline", {}).get("kind"))
except Exception as e:
    print("$c FAILED", e, open("gpurun_out/bench_$c.err").read()[-600:])
PY
done
timeout -s KILL 300 python bench.py --config masking --precision fp32 --steps 30 --no-cpu-baseline > gpurun_out/bench_masking_fp32.json 2> gpurun_out/bench_masking_fp32.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_masking_fp32.json').read().strip().splitlines()[-1]); print('masking --precision fp32', round(d['value']), round(d['ms_per_step'],4))" || tail -3 gpurun_out/bench_masking_fp32.err
PGNN_GAT_OCC=1 timeout -s KILL 300 python bench.py --config gat --steps 30 --no-cpu-baseline > gpurun_out/bench_gat_occ4.json 2> gpurun_out/bench_gat_occ4.err
python -c "
import json
for f in ('gpurun_out/bench_gat.json', 'gpurun_out/bench_gat_occ4.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['step_kernels_us'].items() if 'gat' in k})"
timeout -s KILL 120 python tools/host_profile.py --config gcn --steps 40 > gpurun_out/host_profile_gcn.txt 2>&1; head -16 gpurun_out/host_profile_gcn.txt
# ncu launch lists (shares), then one full capture of the masking step's GEMM family + gathers + BatchNorm-backward sweeps
for c in masking bio_supervised gat; do
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$c.csv python tools/profile_step.py --config $c --steps 2 > gpurun_out/ncu_launches_$c.log 2>&1
  echo "launch list $c: $(wc -l < gpurun_out/launches_$c.csv) rows"
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_gemm_3xtf32_tma|k_aggregate_fwd|k_aggregate_bwd|k_bn_bwd' -c 40 -f -o gpurun_out/prof_masking \
  python tools/profile_step.py --config masking --steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>&1 | tail -2
