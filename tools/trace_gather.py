"""Phase stamps of CTA (0,0) of the LAST fused gather + GEMM1 launch of one encoder forward (k_gin_gather_gemm writes trace slots
10-15; the plain GEMMs use 0-9): entry, PDL wait passed, gather phase done, first MMA, accumulators complete, exit."""
import ctypes, importlib, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ts = importlib.import_module("pretrain-gnns_b200.train_steps")
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dll = cabi.lib.load()
dev = torch.device("cuda:0")
step = ts.MaskingStep(dev, "gin", batch_size=256)
b = {k: v.to(dev) for k, v in step.make_batches(0, 1)[0].items()}
for _ in range(3):
    step(b)
torch.cuda.synchronize()
fn = dll.pgnn_debug_gather_trace
fn.argtypes = [ctypes.c_void_p]
for rep in range(3):
    with torch.no_grad():
        step.model.train()
        step.model(b["x"], b["edge_index"], b["edge_attr"])
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    fn(buf)
    t = list(buf)
    names = {1: "alloc + PDL wait passed", 2: "bias/BN fold/rowptr/S/T staged", 3: "nbr ids staged", 4: "gather loop done", 5: "threadfence done",
             7: "counter bumped", 6: "team complete (TMA thread)", 8: "first MMA", 9: "accumulators complete", 10: "exit"}
    print("fused gather+GEMM1, CTA (0,0), forward %d:" % rep, ", ".join("%s +%.2f" % (names[i], (t[i] - t[0]) / 1e3) for i in (1, 2, 3, 4, 5, 7, 6, 8, 9, 10)))
