"""Where the HOST time of one training step goes (no device synchronisation inside the loop): wall per enqueued step, the share spent
inside the library's C entry points (ctypes), and the top Python frames (cProfile).  Run on the GPU box:
    python tools/host_profile.py [--config masking] [--steps 40]"""
import argparse, cProfile, importlib, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="masking")
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
ts = importlib.import_module("pretrain-gnns_b200.train_steps")
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dev = torch.device("cuda:0")
step = ts.CONFIGS[a.config](dev)
host = step.make_batches(0, 8)
res = [{k: v.to(dev) for k, v in b.items()} for b in host]
for i in range(10):
    step(res[i % 8]).item()
torch.cuda.synchronize()

# time inside every C entry point
calls = {}
lib = cabi.lib
class Timed:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *args):
        t = time.perf_counter(); r = self.fn(*args); d = time.perf_counter() - t
        c = calls.setdefault(self.name, [0, 0.0]); c[0] += 1; c[1] += d
        return r
wrapped = []
dll = lib.load()
for name in cabi.parse_header().keys():
    setattr(lib, name, Timed(name, getattr(dll, name))); wrapped.append(name)   # instance attribute shadows _Lib.__getattr__
import gc
gc.collect(); gc.disable()
t0 = time.perf_counter()
for i in range(a.steps):
    step(res[i % 8])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue per step: %.1f us (GPU drained %.1f ms after the last enqueue; %d steps)" % ((t1 - t0) / a.steps * 1e6, (t2 - t1) * 1e3, a.steps))
tot = sum(v[1] for v in calls.values())
print("inside C entry points: %.1f us per step" % (tot / a.steps * 1e6))
for k, v in sorted(calls.items(), key=lambda kv: -kv[1][1])[:12]:
    print("  %-40s %5.1f calls/step %8.1f us/step" % (k, v[0] / a.steps, v[1] / a.steps * 1e6))
for name in wrapped:
    delattr(lib, name)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(a.steps):
    step(res[i % 8])
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])
