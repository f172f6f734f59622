#!/bin/bash
# One bench line per BASELINE config on one GPU -> gpurun_out/bench_<config>.json (stderr -> .err)
mkdir -p gpurun_out
for c in masking contextpred bio_supervised gcn gat graphsage; do
  timeout -s KILL 400 python bench.py --config $c "$@" > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  echo "$c rc=$? $(head -c 300 gpurun_out/bench_$c.json)"
done
