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


# ---------------------------------------------------------------------------------------------------------------------
# transforms and list collation (SURVEY.md 8(f) f4 / f1 remainder): restatements of chem/util.py:189-241 (MaskAtom,
# mask_edge=False), chem/batch.py:141-210 (BatchSubstructContext) and bio/batch.py:17-50.
# ---------------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(seed, idx):
    """The draw pgnn_mask_atoms defines (csrc/transforms.cu): 64-bit key of batch position `idx` under `seed`."""
    z = (int(seed) + (int(idx) + 1) * 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def mask_atom_choice(node_off, rate, seed):
    """Per graph: the int(n * rate + 1) nodes with the smallest keys (ties by index), ascending; graph-LOCAL indices."""
    out = []
    for g in range(len(node_off) - 1):
        n0, n = int(node_off[g]), int(node_off[g + 1] - node_off[g])
        k = min(int(n * rate + 1), n) if n > 0 else 0            # chem/util.py:229
        keys = sorted(range(n), key=lambda i: (splitmix64(seed & _M64, n0 + i), i))
        out.append(sorted(keys[:k]))
    return out


def mask_atoms(x, node_off, rate, seed, mask_token=119):
    """MaskAtom.__call__(data, masked_atom_indices=choice) applied graph by graph and collated (chem/util.py:231-241,
    chem/batch.py:41-42).  -> (x_masked, masked_atom_indices, mask_node_label, mask_off)"""
    x = np.array(x, dtype=np.int64, copy=True)
    idx, off = [], [0]
    for g, local in enumerate(mask_atom_choice(node_off, rate, seed)):
        idx += [int(node_off[g]) + i for i in local]
        off.append(len(idx))
    idx = np.array(idx, dtype=np.int64)
    labels = x[idx].copy() if len(idx) else np.zeros((0, 2), np.int64)
    if len(idx):
        x[idx] = np.array([mask_token, 0])
    return x, idx, labels, np.array(off, dtype=np.int64)


def collate_lists(list_ptr, values, ids, add=None):
    """Ragged per-graph index lists of a batch, each entry offset by `add[i]` -> (out, seg, sizes, list_off)."""
    out, seg, sizes, off = [], [], [], [0]
    for i, g in enumerate(ids):
        v = np.asarray(values[list_ptr[g]:list_ptr[g + 1]], dtype=np.int64) + (0 if add is None else int(add[i]))
        out.append(v)
        seg.append(np.full(len(v), i, dtype=np.int64))
        sizes.append(len(v))
        off.append(off[-1] + len(v))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return cat(out), cat(seg), np.array(sizes, dtype=np.int64), np.array(off, dtype=np.int64)


def collate_bio(graphs, ids):
    """graphs: list of (num_nodes, edge_index [2,e] local int, edge_attr [e,9] 0/1); follows bio/batch.py:17-50."""
    xs, eis, eas, bs, node_off, edge_off = [], [], [], [], [0], [0]
    for i, g in enumerate(ids):
        n, ei, ea = graphs[g]
        xs.append(np.ones((n, 1), dtype=np.float32))                      # bio/loader.py:47
        eis.append(np.asarray(ei, dtype=np.int64) + node_off[-1])
        eas.append(np.asarray(ea, dtype=np.float32))
        bs.append(np.full((n,), i, dtype=np.int64))
        node_off.append(node_off[-1] + n)
        edge_off.append(edge_off[-1] + ei.shape[1])
    return dict(x=np.concatenate(xs, 0) if xs else np.zeros((0, 1), np.float32),
                edge_index=np.concatenate(eis, 1) if eis else np.zeros((2, 0), np.int64),
                edge_attr=np.concatenate(eas, 0) if eas else np.zeros((0, 9), np.float32),
                batch=np.concatenate(bs) if bs else np.zeros(0, np.int64),
                node_off=np.array(node_off, dtype=np.int64), edge_off=np.array(edge_off, dtype=np.int64))
