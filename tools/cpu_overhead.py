"""CPU issue time of one training step, by segment (no device sync inside the step)."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
syn = importlib.import_module("pretrain-gnns_b200.synthetic")
chem = importlib.import_module("pretrain-gnns_b200.chem.model")
ops = importlib.import_module("pretrain-gnns_b200.ops")
dev = torch.device("cuda:0")
model = chem.GNN(5, 300).to(dev).train()
head = torch.nn.Linear(300, 119).to(dev)
params = list(model.parameters()) + list(head.parameters())
host = bench.make_batches(syn, 0, 4)
pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
seg = {k: 0.0 for k in ("h2d", "zero", "fwd", "head", "loss", "bwd", "item")}
def step(i, rec):
    t = [time.perf_counter()]
    b = {k: v.to(dev, non_blocking=True) for k, v in pinned[i % 4].items()}; t.append(time.perf_counter())
    for p in params: p.grad = None
    t.append(time.perf_counter())
    rep = model(b["x"], b["edge_index"], b["edge_attr"]); t.append(time.perf_counter())
    logits = ops.linear(ops.row_gather(rep, b["masked_atom_indices"]), head.weight, head.bias); t.append(time.perf_counter())
    loss = torch.nn.functional.cross_entropy(logits.double(), b["labels"]); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    v = loss.item(); t.append(time.perf_counter())
    if rec:
        for k, a, c in zip(seg, t[:-1], t[1:]): seg[k] += c - a
for i in range(10): step(i, False)
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for i in range(N): step(i, True)
tot = time.perf_counter() - t0
print("wall per step %.1f us" % (tot / N * 1e6))
for k, v in seg.items(): print("  %-5s %.1f us" % (k, v / N * 1e6))
