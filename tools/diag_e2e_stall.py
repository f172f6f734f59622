"""Find the host call behind the sporadic 10-40 ms stall at step 1 of bench.py's end-to-end loop: the same loop with a timestamp
after every statement, repeated; prints every iteration that took more than 3 ms with its per-statement breakdown."""
import gc, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ts = importlib.import_module("pretrain-gnns_b200.train_steps")
pdata = importlib.import_module("pretrain-gnns_b200.data")
dev = torch.device("cuda:0")
step = ts.CONFIGS["masking"](dev)
host = step.make_batches(0, 8)
resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stager = pdata.BatchStager(dev)
packed = [stager.pack(b) for b in host]
LAG = 2
loss_host = torch.zeros(LAG + 1, dtype=torch.float64).pin_memory()
loss_ready = [torch.cuda.Event() for _ in range(LAG + 1)]
STEPS = 12
names = ["flush", "ev0", "take", "train_step", "submit", "loss_copy", "loss_rec", "consume", "ev1"]


def run(rep, with_resident_first):
    if with_resident_first:
        for i in range(20):
            flush.zero_()
            step(resident[i % 8])
        torch.cuda.synchronize()
    for i in range(5):
        step(stager.take(stager.submit(packed[i % 8]))).item()
    torch.cuda.synchronize()
    gc.collect(); gc.disable()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    ticket, rows = None, []
    for i in range(STEPS):
        t = [time.perf_counter()]
        flush.zero_(); t.append(time.perf_counter())
        ev[i][0].record(); t.append(time.perf_counter())
        if ticket is None:
            ticket = stager.submit(packed[i % 8])
        b = stager.take(ticket); t.append(time.perf_counter())
        loss = step(b); t.append(time.perf_counter())
        ticket = stager.submit(packed[(i + 1) % 8]) if i + 1 < STEPS else None; t.append(time.perf_counter())
        slot = i % (LAG + 1)
        loss_host[slot].copy_(loss.detach(), non_blocking=True); t.append(time.perf_counter())
        loss_ready[slot].record(); t.append(time.perf_counter())
        if i >= LAG:
            j = (i - LAG) % (LAG + 1)
            loss_ready[j].synchronize()
            float(loss_host[j])
        t.append(time.perf_counter())
        ev[i][1].record(); t.append(time.perf_counter())
        rows.append([1e3 * (b_ - a_) for a_, b_ in zip(t[:-1], t[1:])])
    torch.cuda.synchronize()
    gc.enable()
    for i, r in enumerate(rows):
        tot = sum(r)
        if tot > 3.0:
            print("rep %d step %d: host %.2f ms, device %.2f ms :: %s" % (rep, i, tot, ev[i][0].elapsed_time(ev[i][1]),
                  ", ".join("%s %.2f" % (n, v) for n, v in zip(names, r) if v > 0.2)), flush=True)


for rep in range(8):
    run(rep, rep % 2 == 0)
print("done")
