"""Shared helpers: regenerate the golden inputs from seeds and compare against the stored vectors."""
import importlib
import os

import numpy as np
import torch

from oracle import gnn_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
syn = importlib.import_module("pretrain-gnns_b200.synthetic")
CASES = {"chem": dict(graphs=4, data_seed=11, param_seed=5), "bio": dict(graphs=2, data_seed=12, param_seed=6)}
TYPES = ("gin", "gcn", "graphsage", "gat")


def golden_batch(domain):
    c = CASES[domain]
    if domain == "chem":
        return syn.zinc_batch(c["graphs"], c["data_seed"])
    return syn.ppi_batch(c["graphs"], c["data_seed"], n_lo=40, n_hi=60, num_tasks=16)


def golden_params(domain, t):
    return O.make_params(domain, t, 5, 300, seed=CASES[domain]["param_seed"])


def load(domain, t):
    return dict(np.load(os.path.join(HERE, "golden", f"{domain}_{t}.npz")))


def probe(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def input_checksum(b):
    h = 0
    for k in ("x", "edge_index", "edge_attr"):
        h = (h * 1000003 + int(b[k].to(torch.float64).sum().item() * 8 + b[k].numel())) % (2 ** 61 - 1)
    return np.int64(h)


def close(a, b, atol=1e-4, rtol=1e-4):
    """north_star tolerance: 1e-4 fp32, stated abs + rel."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    worst = (err - bound).max().item() if err.numel() else -1.0
    return worst <= 0, float(err.max().item() if err.numel() else 0.0)


def close_scaled(a, b, tol=1e-4, floor=1.0):
    """Gradient check: max error relative to the reference tensor's largest magnitude (sums over
    hundreds of rows cancel, so an element-wise relative bound is meaningless for them)."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    scale = max(float(b.abs().max().item()) if b.numel() else 0.0, floor)
    e = float((a - b).abs().max().item()) / scale if b.numel() else 0.0
    return e <= tol, e


def check_against_golden(G, out_eval, out_train, grads, stats, atol=1e-4, rtol=1e-4, gtol=2e-4):
    """Compare a full set of results with one golden file. `grads`: name -> tensor; `stats`: name -> tensor."""
    bad = []
    # floor for the error scale: the model's typical gradient magnitude (a bias feeding a train-mode
    # BatchNorm has an exactly-zero gradient; its computed value is rounding noise of that scale)
    gmax = sorted(float(np.abs(v).max()) for k, v in G.items() if k.startswith("g:") and v.size)
    floor = max(gmax[len(gmax) // 2], 1.0) if gmax else 1.0
    for name, mine in (("out_eval", out_eval), ("out_train", out_train)):
        if mine is None:
            continue
        ok, e = close(mine, G[name], atol, rtol)
        if not ok:
            bad.append((name, e))
    for k, ref in G.items():
        kind, _, name = k.partition(":")
        if kind == "g" and grads is not None:
            ok, e = close_scaled(grads[name], ref, gtol, floor)
        elif kind == "gs0" and grads is not None:
            g2 = torch.as_tensor(np.asarray(grads[name])).reshape(grads[name].shape[0], -1)
            ok, e = close_scaled(g2.sum(0), ref, gtol, floor)
        elif kind == "gs1" and grads is not None:
            g2 = torch.as_tensor(np.asarray(grads[name])).reshape(grads[name].shape[0], -1)
            ok, e = close_scaled(g2.sum(1), ref, gtol, floor)
        elif kind == "gp" and grads is not None:
            g = torch.as_tensor(np.asarray(grads[name]))
            ok, e = close_scaled((g * probe(g.shape, 77)).sum(), ref, gtol * 5, floor)
        elif kind == "rs" and stats is not None:
            ok, e = close(stats[name], ref, atol, rtol)
        else:
            continue
        if not ok:
            bad.append((k, e))
    return bad
