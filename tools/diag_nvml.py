"""How long do the NVML calls of bench.py's ClockSampler take while the GPU runs training steps, and do they stall the launching
thread?  Prints per-call latency statistics and the slowest host iteration of the step loop with / without the poller."""
import importlib, os, sys, threading, time
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import pynvml as nv
ts = importlib.import_module("pretrain-gnns_b200.train_steps")
dev = torch.device("cuda:0")
step = ts.CONFIGS["masking"](dev)
res = [{k: v.to(dev) for k, v in b.items()} for b in step.make_batches(0, 8)]
for i in range(10):
    step(res[i % 8]).item()
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
calls = {"clock": lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), "maxclock": lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
         "reasons": lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)}
lat = {k: [] for k in calls}
stop = False


def poll(which, period):
    while not stop:
        for k in which:
            t = time.perf_counter(); calls[k](); lat[k].append(1e3 * (time.perf_counter() - t))
        time.sleep(period)


def loop(n=150):
    worst = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        t = time.perf_counter()
        step(res[i % 8])
        worst = max(worst, 1e3 * (time.perf_counter() - t))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n, worst


print("no poller: mean %.3f ms / step, slowest host iteration %.2f ms" % loop())
for which in (["clock"], ["maxclock"], ["reasons"], ["clock", "maxclock", "reasons"]):
    for k in lat: lat[k].clear()
    stop = False
    th = threading.Thread(target=poll, args=(which, 0.02), daemon=True); th.start()
    m, w = loop()
    stop = True; th.join()
    print("poll %s every 20 ms: mean %.3f ms / step, slowest host iteration %.2f ms; call latency ms: %s"
          % (which, m, w, {k: (round(min(v), 3), round(sorted(v)[len(v) // 2], 3), round(max(v), 3)) for k, v in lat.items() if v}), flush=True)
