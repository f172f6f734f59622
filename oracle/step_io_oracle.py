"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the two steps either side of the hot path.

collate_chem follows BatchMasking.from_data_list (/root/reference/chem/batch.py:17-52): per graph i append
`full((n_i,), i)` to `batch`, add the running node count to `edge_index`, then concatenate (edge_index along the last
dimension, everything else along dim 0).  adam_step restates torch.optim.Adam's update; `legacy_eps` is torch 1.0.1's
form (requirements.txt:2).  Parity status: collate is integer work checked bit-exactly against this restatement, which
is itself checked in tests against a literal per-graph torch.cat loop; the modern Adam form is pinned against
torch.optim.Adam of this image, the legacy form is unpinned (torch 1.0.1 is not installable here).
"""
import numpy as np


def collate_chem(graphs, ids):
    """graphs: list of (x [n,2], edge_index [2,e] local, edge_attr [e,2]) integer arrays; ids: order of the batch."""
    xs, eis, eas, bs, node_off, edge_off = [], [], [], [], [0], [0]
    cumsum_node = cumsum_edge = 0
    for i, g in enumerate(ids):
        x, ei, ea = graphs[g]
        n = x.shape[0]
        bs.append(np.full((n,), i, dtype=np.int64))
        xs.append(np.asarray(x, dtype=np.int64))
        eis.append(np.asarray(ei, dtype=np.int64) + cumsum_node)
        eas.append(np.asarray(ea, dtype=np.int64))
        cumsum_node += n
        cumsum_edge += ei.shape[1]
        node_off.append(cumsum_node)
        edge_off.append(cumsum_edge)
    z = lambda *s: np.zeros(s, dtype=np.int64)
    return dict(x=np.concatenate(xs, 0) if xs else z(0, 2), edge_index=np.concatenate(eis, 1) if eis else z(2, 0),
                edge_attr=np.concatenate(eas, 0) if eas else z(0, 2), batch=np.concatenate(bs) if bs else z(0),
                node_off=np.array(node_off, dtype=np.int64), edge_off=np.array(edge_off, dtype=np.int64))


def adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0, legacy_eps=False):
    """One Adam update in fp64 on numpy arrays; returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, dtype=np.float64) for a in (p, g, m, v))
    g = g * grad_scale + weight_decay * p
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    if legacy_eps:
        p = p - lr * np.sqrt(bc2) / bc1 * m / (np.sqrt(v) + eps)
    else:
        p = p - lr / bc1 * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v
