"""Generate golden vectors by executing the REFERENCE's own, unmodified model.py files.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The absent torch_geometric 1.0.3 / torch_scatter 1.1.2 names are served by
`oracle/pyg103_standin` (see its README for the semantics it encodes).  Inputs and parameters are NOT
stored: they are regenerated from seeds by `pretrain-gnns_b200/synthetic.py` and
`oracle.gnn_oracle.make_params` (numpy PCG64 / torch CPU generator: machine independent); an input
checksum is stored so generator drift is detected.  Stored per (domain, gnn_type):

    out_eval, out_train            reference GNN outputs [N, 300] fp32
    loss                           L = sum(out_train * R), R seeded
    g:<param>                      full gradient for small tensors (< 4096 elements)
    gs0:/gs1:/gp:<param>           column sums / row sums / seeded-probe projection for large ones
    *64:<param>                    the same summaries from the reference run in float64 (`model.double()`):
                                   train-mode BatchNorm makes these gradients ill-conditioned in fp32, so
                                   the parity bound for gradients is stated relative to the reference's OWN
                                   fp32-vs-fp64 discrepancy (tests/golden_util.py)
    rs:<key>                       BatchNorm running stats after the train-mode forward
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg103_standin"))
syn = importlib.import_module("pretrain-gnns_b200.synthetic")
from oracle import gnn_oracle as O  # noqa: E402

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"chem": dict(graphs=4, data_seed=11, param_seed=5), "bio": dict(graphs=2, data_seed=12, param_seed=6)}


def load_reference_model(domain):
    from oracle import reference_runner
    assert reference_runner.root() == REF, "goldens are generated from /root/reference itself, not from a staged copy"
    return reference_runner.load(domain)


def golden_batch(domain):
    c = CASES[domain]
    if domain == "chem":
        return syn.zinc_batch(c["graphs"], c["data_seed"])
    return syn.ppi_batch(c["graphs"], c["data_seed"], n_lo=40, n_hi=60, num_tasks=16)


def probe(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def grad_summaries(name, grad, out, tag=""):
    if grad.numel() < 4096:
        out["g%s:" % tag + name] = grad.numpy()
    else:
        g2 = grad.reshape(grad.shape[0], -1)
        out["gs0%s:" % tag + name] = g2.sum(0).numpy()
        out["gs1%s:" % tag + name] = g2.sum(1).numpy()
        out["gp%s:" % tag + name] = (grad * probe(grad.shape, 77).to(grad.dtype)).sum().numpy()


def input_checksum(b):
    h = 0
    for k in ("x", "edge_index", "edge_attr"):
        h = (h * 1000003 + int(b[k].to(torch.float64).sum().item() * 8 + b[k].numel())) % (2 ** 61 - 1)
    return np.int64(h)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    for domain in ("chem", "bio"):
        ref = load_reference_model(domain)
        b = golden_batch(domain)
        for t in ("gin", "gcn", "graphsage", "gat"):
            P = O.make_params(domain, t, 5, 300, seed=CASES[domain]["param_seed"])
            model = ref.GNN(5, 300, JK="last", drop_ratio=0, gnn_type=t)
            assert str(model.load_state_dict(P)) == "<All keys matched successfully>"
            out = {"input_checksum": input_checksum(b)}
            model.eval()
            with torch.no_grad():
                out["out_eval"] = model(b["x"], b["edge_index"], b["edge_attr"]).numpy()
            model.train()
            y = model(b["x"], b["edge_index"], b["edge_attr"])
            loss = (y * probe(y.shape, 99)).sum()
            loss.backward()
            out["out_train"] = y.detach().numpy()
            out["loss"] = loss.detach().numpy()
            for k, p in model.named_parameters():
                grad_summaries(k, p.grad, out)
            # the same step in float64 (the conditioning yardstick)
            m64 = ref.GNN(5, 300, JK="last", drop_ratio=0, gnn_type=t)
            m64.load_state_dict(P)
            m64.double().train()
            ea64 = b["edge_attr"].double() if b["edge_attr"].is_floating_point() else b["edge_attr"]
            x64 = b["x"].double() if b["x"].is_floating_point() else b["x"]
            y64 = m64(x64, b["edge_index"], ea64)
            (y64 * probe(y64.shape, 99).double()).sum().backward()
            out["out_train64"] = y64.detach().float().numpy()
            for k, p in m64.named_parameters():
                grad_summaries(k, p.grad, out, tag="64")
            for k, v in model.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
                    out["rs:" + k] = v.numpy()
            path = os.path.join(HERE, f"{domain}_{t}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path) // 1024, "KiB", "N =", y.shape[0])

    # heads (chem/pretrain_contextpred.py:54-67, chem/model.py:369) on reference-produced node reps
    ref = load_reference_model("chem")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg103_standin"))
    from torch_geometric.nn import global_mean_pool
    b = golden_batch("chem")
    rep = probe((b["x"].shape[0], 300), 5)
    pooled = global_mean_pool(rep, b["batch"])
    np.savez_compressed(os.path.join(HERE, "heads.npz"), pooled=pooled.numpy())
    print("heads.npz")


if __name__ == "__main__":
    main()
