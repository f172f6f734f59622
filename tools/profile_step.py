"""A few training steps of one bench config between cudaProfilerStart/Stop, for ncu (`--profile-from-start off`):

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_masking.csv \
        python tools/profile_step.py --config masking --steps 2
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_gemm_3xtf32_tma|k_aggregate_fwd' -c 24 \
        -o gpurun_out/prof_masking python tools/profile_step.py --config masking --steps 1
"""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="masking")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
ts = importlib.import_module("pretrain-gnns_b200.train_steps")
dev = torch.device("cuda:0")
torch.manual_seed(0)
step = ts.CONFIGS[a.config](dev)
batches = [{k: v.to(dev) for k, v in b.items()} for b in step.make_batches(0, 2)]
for i in range(a.warmup):
    step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(a.steps):
    step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled %d step(s) of %s" % (a.steps, a.config))
