"""Seeded synthetic batches shaped like the reference's inputs.

No dataset or rdkit is available, so every test / bench input comes from here.
The layouts follow the reference's collation code exactly:

* chem molecules: ``x [N,2] int64`` (atom type, chirality), ``edge_index [2,E] int64`` with both
  directions of a bond adjacent and carrying the same ``edge_attr [E,2] int64``
  (/root/reference/chem/loader.py:53-100), node ids offset per graph and a sorted ``batch`` vector
  (/root/reference/chem/batch.py:17-52).
* MaskAtom: per graph ``int(n*rate + 1)`` distinct atoms, labels saved, ``x`` row overwritten with
  ``[119, 0]`` (/root/reference/chem/util.py:229-241); optional bond masking with type 5
  (/root/reference/chem/util.py:243-272).
* substructure/context pairs (/root/reference/chem/batch.py:141-210).
* bio PPI ego graphs: ``x [N,1] float32`` ones, ``edge_attr [E,9] float32`` 0/1 with cols 7,8 zero
  (/root/reference/bio/loader.py:47-75), ``center_node_idx`` offset per graph
  (/root/reference/bio/batch.py:17-50).

numpy's PCG64 is used (not torch's generator) so the same seed gives the same tensors on every
machine and torch build.
"""
from __future__ import annotations

import numpy as np
import torch

NUM_ATOM_TYPE_MASK = 119   # chem/pretrain_masking.py:122
NUM_BOND_TYPE_MASK = 5     # chem/pretrain_masking.py:122 (num_edge_type=5)


def _t(a, dtype=torch.int64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _molecule(rng, n, extra, tree_only=False):
    """One molecule's undirected bond list as (u, v) with u != v. Spanning tree + `extra` chords."""
    parent = (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)  # node k attaches to a node < k
    u = np.arange(1, n, dtype=np.int64)
    v = parent
    if not tree_only and extra > 0:
        a = rng.integers(0, n, size=extra)
        b = (a + 1 + rng.integers(0, n - 1, size=extra)) % n  # never a self pair
        u = np.concatenate([u, a])
        v = np.concatenate([v, b])
    return u, v


def _directed(u, v, attr):
    """Both directions adjacent, identical attributes (chem/loader.py:83-86)."""
    m = len(u)
    ei = np.empty((2, 2 * m), dtype=np.int64)
    ei[0, 0::2], ei[1, 0::2] = u, v
    ei[0, 1::2], ei[1, 1::2] = v, u
    ea = np.repeat(attr, 2, axis=0)
    return ei, ea


def zinc_batch(num_graphs: int, seed: int, n_lo: int = 18, n_hi: int = 28, extra: int = 3,
               tree_only: bool = False):
    """ZINC-shaped batch. Returns dict of CPU tensors: x, edge_index, edge_attr, batch, ptr."""
    rng = np.random.default_rng(seed)
    xs, eis, eas, bs = [], [], [], []
    ptr = [0]
    off = 0
    for g in range(num_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        u, v = _molecule(rng, n, extra, tree_only)
        attr = np.stack([rng.integers(0, 4, size=len(u)), rng.integers(0, 3, size=len(u))], axis=1)
        ei, ea = _directed(u, v, attr)
        xs.append(np.stack([rng.integers(0, 119, size=n), rng.integers(0, 3, size=n)], axis=1))
        eis.append(ei + off)
        eas.append(ea)
        bs.append(np.full(n, g, dtype=np.int64))
        off += n
        ptr.append(off)
    if num_graphs == 0:
        return dict(x=torch.zeros(0, 2, dtype=torch.int64), edge_index=torch.zeros(2, 0, dtype=torch.int64),
                    edge_attr=torch.zeros(0, 2, dtype=torch.int64), batch=torch.zeros(0, dtype=torch.int64),
                    ptr=torch.zeros(1, dtype=torch.int64), num_graphs=0)
    return dict(x=_t(np.concatenate(xs)), edge_index=_t(np.concatenate(eis, axis=1)),
                edge_attr=_t(np.concatenate(eas)), batch=_t(np.concatenate(bs)), ptr=_t(np.array(ptr)),
                num_graphs=num_graphs)


def mask_atoms(batch: dict, seed: int, rate: float = 0.15, mask_edge: bool = False) -> dict:
    """Apply MaskAtom semantics per graph, in place on a copy; adds the BatchMasking keys."""
    rng = np.random.default_rng(seed + 7919)
    out = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    ptr = batch["ptr"].numpy()
    idx = []
    for g in range(len(ptr) - 1):
        n = int(ptr[g + 1] - ptr[g])
        k = int(n * rate + 1)
        idx.append(ptr[g] + rng.choice(n, size=k, replace=False))
    idx = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
    out["masked_atom_indices"] = _t(idx)
    out["mask_node_label"] = out["x"][out["masked_atom_indices"]].clone()
    out["x"][out["masked_atom_indices"]] = torch.tensor([NUM_ATOM_TYPE_MASK, 0])
    if mask_edge:
        ei = out["edge_index"].numpy()
        hit = np.isin(ei[0], idx) | np.isin(ei[1], idx)
        conn = np.nonzero(hit)[0]
        conn_pairs = conn[conn % 2 == 0]  # one id per undirected bond (first of the adjacent pair)
        out["connected_edge_indices"] = _t(conn_pairs)
        out["mask_edge_label"] = out["edge_attr"][out["connected_edge_indices"]].clone()
        both = np.concatenate([conn_pairs, conn_pairs + 1])
        out["edge_attr"][_t(both)] = torch.tensor([NUM_BOND_TYPE_MASK, 0])
    return out


def substruct_context_batch(num_graphs: int, seed: int) -> dict:
    """Config 3 (chem/pretrain_contextpred.py): substructure graphs + context graphs + overlap ids."""
    sub = zinc_batch(num_graphs, seed * 2 + 1, n_lo=12, n_hi=22)
    ctx = zinc_batch(num_graphs, seed * 2 + 2, n_lo=4, n_hi=14, tree_only=True)
    rng = np.random.default_rng(seed + 104729)
    sp, cp = sub["ptr"].numpy(), ctx["ptr"].numpy()
    center = sp[:-1] + np.array([rng.integers(0, sp[g + 1] - sp[g]) for g in range(num_graphs)], dtype=np.int64)
    ov, ovb = [], []
    for g in range(num_graphs):
        nc = int(cp[g + 1] - cp[g])
        k = int(rng.integers(1, min(4, nc) + 1))
        ov.append(cp[g] + rng.choice(nc, size=k, replace=False))
        ovb.append(np.full(k, g, dtype=np.int64))
    return dict(x_substruct=sub["x"], edge_index_substruct=sub["edge_index"], edge_attr_substruct=sub["edge_attr"],
                center_substruct_idx=_t(center),
                x_context=ctx["x"], edge_index_context=ctx["edge_index"], edge_attr_context=ctx["edge_attr"],
                overlap_context_substruct_idx=_t(np.concatenate(ov)),
                batch_overlapped_context=_t(np.concatenate(ovb)),
                overlapped_context_size=_t(np.array([len(o) for o in ov])), num_graphs=num_graphs)


def ppi_batch(num_graphs: int, seed: int, n_lo: int = 400, n_hi: int = 600, pairs_per_node: int = 5,
              num_tasks: int = 5000, target_rate: float = 0.05) -> dict:
    """Config 4 (bio/pretrain_supervised.py): PPI-ego-shaped graphs."""
    rng = np.random.default_rng(seed)
    eis, eas, bs, ns = [], [], [], []
    off = 0
    ptr = [0]
    for g in range(num_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        m = pairs_per_node * n
        a = rng.integers(0, n, size=m)
        b = (a + 1 + rng.integers(0, n - 1, size=m)) % n
        attr = np.zeros((m, 9), dtype=np.float32)
        attr[:, :7] = (rng.random((m, 7)) < 0.3)
        ei, ea = _directed(a, b, attr)
        eis.append(ei + off)
        eas.append(ea)
        bs.append(np.full(n, g, dtype=np.int64))
        ns.append(n)
        off += n
        ptr.append(off)
    ptr = np.array(ptr)
    y = (rng.random((num_graphs, num_tasks)) < target_rate).astype(np.int64)
    return dict(x=torch.ones(off, 1, dtype=torch.float32), edge_index=_t(np.concatenate(eis, axis=1)),
                edge_attr=_t(np.concatenate(eas), torch.float32), batch=_t(np.concatenate(bs)),
                center_node_idx=_t(ptr[:-1]), ptr=_t(ptr), go_target_pretrain=_t(y.reshape(-1)),
                num_graphs=num_graphs)


def split_graphs(batch: dict):
    """Per-graph (x [n,2], edge_index [2,e] graph-LOCAL, edge_attr [e,2]) numpy arrays of a zinc_batch: the form a dataset
    holds before collation (chem/loader.py:53-100 builds one such Data per molecule)."""
    ptr = batch["ptr"].numpy()
    x, ei, ea = batch["x"].numpy(), batch["edge_index"].numpy(), batch["edge_attr"].numpy()
    owner = np.searchsorted(ptr, ei[0], side="right") - 1  # edges are emitted graph by graph
    eptr = np.searchsorted(owner, np.arange(len(ptr)))
    return [(x[ptr[g]:ptr[g + 1]], ei[:, eptr[g]:eptr[g + 1]] - ptr[g], ea[eptr[g]:eptr[g + 1]]) for g in range(len(ptr) - 1)]


def one_direction_only(batch: dict, seed: int, keys=("edge_index", "edge_attr")) -> dict:
    """Asymmetric variant of a batch: of every bond's two adjacent directed edges (u,v),(v,u) keep exactly one, chosen at
    random.  The reference never feeds such a graph (chem/loader.py:83-86 always emits both directions), but only on it
    does a swapped target/source — aggregation onto edge_index[1] instead of edge_index[0] — change GIN/GCN/GraphSAGE
    results; the parity tests use it to pin the direction convention of every kernel."""
    rng = np.random.default_rng(seed + 15485863)
    out = dict(batch)
    ei = batch[keys[0]]
    m = ei.shape[1] // 2
    keep = _t(2 * np.arange(m) + rng.integers(0, 2, size=m))
    out[keys[0]] = ei[:, keep].contiguous()
    out[keys[1]] = batch[keys[1]][keep].contiguous()
    return out
