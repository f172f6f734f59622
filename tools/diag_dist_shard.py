"""1-GPU replica of tests/test_gpu_dist2.py's shard computation: per-tensor error of the GPU gradients of one 8-graph shard against the
fp64 oracle (3-layer GIN, Linear(300,7) head, mean-square loss)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import gnn_oracle as O
syn = importlib.import_module("pretrain-gnns_b200.synthetic")
chem = importlib.import_module("pretrain-gnns_b200.chem.model")
dev = torch.device("cuda:0")
torch.manual_seed(0)
gnn = chem.GNN(3, 300).to(dev).train()
head = torch.nn.Linear(300, 7).to(dev)
state = {k: v.detach().cpu().clone() for k, v in gnn.state_dict().items()}
hs = {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}
for fused in (True, False):
    gnn.fused = fused
    for rank in (0, 1):
        b = syn.zinc_batch(8, rank)
        for p in list(gnn.parameters()) + list(head.parameters()):
            p.grad = None
        head(gnn(*(b[k].to(dev) for k in ("x", "edge_index", "edge_attr")))).square().mean().backward()
        L = O.leaf_params(state, torch.float64)
        W, bias = hs["weight"].double().requires_grad_(True), hs["bias"].double().requires_grad_(True)
        rep = O.chem_gnn(L, b["x"], b["edge_index"], b["edge_attr"], 3, "gin", True)
        torch.nn.functional.linear(rep, W, bias).square().mean().backward()
        want = {k: L[k].grad for k in state if k in L and L[k].requires_grad}
        got = dict(gnn.named_parameters())
        gmax = max(float(v.abs().max()) for v in want.values())
        worst = []
        for k, w in want.items():
            scale = max(float(w.abs().max()), 1e-3 * gmax)
            worst.append((float((got[k].grad.cpu().double() - w).abs().max()) / scale, k))
        worst.sort(reverse=True)
        print("fused=%s shard %d: worst %s" % (fused, rank, [(round(e, 7), k) for e, k in worst[:4]]), flush=True)
