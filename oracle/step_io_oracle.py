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


def mask_edges_chem(edge_index, edge_attr, edge_off, masked_atom_indices, num_edge_type=5):
    """The mask_edge=True half of MaskAtom.__call__ (chem/util.py:243-272) graph by graph + BatchMasking's edge offset
    (chem/batch.py:41-42), on the collated batch.  -> (edge_attr_masked, connected_edge_indices, mask_edge_label, conn_off)"""
    ei = np.asarray(edge_index)
    ea = np.array(edge_attr, dtype=np.int64, copy=True)
    masked = set(int(i) for i in masked_atom_indices)
    conn, labels, off = [], [], [0]
    for g in range(len(edge_off) - 1):
        L = [j for j in range(int(edge_off[g]), int(edge_off[g + 1])) if int(ei[0, j]) in masked or int(ei[1, j]) in masked]   # :246-251
        sel = L[::2]                                                                                                                # :257, :268
        labels += [ea[j].copy() for j in sel]
        conn += sel
        for j in L:
            ea[j] = (num_edge_type, 0)                                                                                              # :263-265
        off.append(len(conn))
    lab = np.array(labels, dtype=np.int64).reshape(-1, 2)
    return ea, np.array(conn, dtype=np.int64), lab, np.array(off, dtype=np.int64)


def mask_edge_choice_bio(edge_off, rate, seed):
    """Per graph: the int(e/2 * rate + 1) bond pairs with the smallest keys splitmix64(seed, column id of the pair's first
    direction), ties by index, ascending; graph-LOCAL pair numbers (csrc/mask_edges.cu defines this draw)."""
    out = []
    for g in range(len(edge_off) - 1):
        e0, m = int(edge_off[g]), int(edge_off[g + 1] - edge_off[g]) // 2
        k = min(int(m * rate + 1), m) if m > 0 else 0            # bio/util.py:78-80
        order = sorted(range(m), key=lambda i: (splitmix64(seed & _M64, e0 + 2 * i), i))
        out.append(sorted(order[:k]))
    return out


def mask_edges_bio(edge_attr, edge_off, rate, seed):
    """MaskEdge.__call__(data, masked_edge_indices=[2 i ...]) graph by graph + the edge offset of bio/batch.py:95-96.
    -> (edge_attr_masked, masked_edge_idx, mask_edge_label, mask_off)"""
    ea = np.array(edge_attr, dtype=np.float32, copy=True)
    idx, off = [], [0]
    for g, local in enumerate(mask_edge_choice_bio(edge_off, rate, seed)):
        idx += [int(edge_off[g]) + 2 * i for i in local]
        off.append(len(idx))
    idx = np.array(idx, dtype=np.int64)
    labels = ea[idx].copy() if len(idx) else np.zeros((0, 9), np.float32)
    mask = np.array([0, 0, 0, 0, 0, 0, 0, 0, 1], dtype=np.float32)       # bio/util.py:98-102
    for j in idx:
        ea[j] = mask
        ea[j + 1] = mask
    return ea, idx, labels, np.array(off, dtype=np.int64)


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


# ---------------------------------------------------------------------------------------------------------------------
# ExtractSubstructureContextPair: restatement of chem/util.py:55-151 (+ chem/loader.py:146-221, the
# networkx round trip it goes through) and bio/util.py:123-205, each followed by BatchSubstructContext.from_data_list
# (chem/batch.py:141-210, bio/batch.py:196-265).
#
# What the reference computes per graph, root r:  d(v) = hop distance from r in the UNDIRECTED graph whose edges are the
# even-indexed columns of edge_index (chem/loader.py:169: `for j in range(0, num_bonds, 2)`; a pair already present is
# skipped, :173);  ball(c) = {v : d(v) <= c} with a negative cutoff (the constructor's 0 -> -1 quirk, chem/util.py:73-78)
# giving {r};  substructure = ball(k);  context = ball(l1) symmetric-difference ball(l2);  overlap = context & substructure.
# Both node sets are relabelled 0..m-1 and re-emitted through nx_to_graph_data_obj_simple: every kept undirected edge
# becomes two adjacent columns (i,j),(j,i) with the same attribute row.  A graph whose context is empty is DROPPED from the
# batch (chem/batch.py:168 "If there is no context, just skip!!").  bio: the substructure is the whole graph, the context is
# everything outside ball(l1) and every context node is an overlap node (bio/util.py:176-199).
#
# Ordering: the reference numbers the nodes of a set in networkx's subgraph-view iteration order (graph order when the
# set holds at least half of the nodes, CPython set order otherwise) and the edges in adjacency order; the overlap list is in
# set order.  None of that is defined by the reference's source and no consumer depends on it (row gathers, mean pools and
# permutation-equivariant encoders), so the order is DEFINED here: nodes ascending by original index, edges in the order of
# their first column in the source, overlap ascending.  tests/test_oracle_vs_reference.py undoes the reference's
# relabelling (it records `reset_idxes`' maps) and checks equality of node sets, per-node features, edge multisets,
# centre and overlap; the device kernels are then bit-exact against this restatement.
# ---------------------------------------------------------------------------------------------------------------------
_INF = 1 << 30


def first_pairs(ei):
    """[e/2] bool: pair p = columns (2p, 2p+1); True where {u,v} was not seen in an earlier pair (chem/loader.py:173)."""
    seen, out = set(), []
    for p in range(ei.shape[1] // 2):
        key = frozenset((int(ei[0, 2 * p]), int(ei[1, 2 * p])))
        out.append(key not in seen)
        seen.add(key)
    return np.array(out, dtype=bool)


def hop_distance(n, ei, root, first=None):
    d = np.full(n, _INF, dtype=np.int64)
    if n == 0:
        return d
    first = first_pairs(ei) if first is None else first
    adj = [[] for _ in range(n)]
    for p in np.nonzero(first)[0]:
        u, v = int(ei[0, 2 * p]), int(ei[1, 2 * p])
        adj[u].append(v)
        adj[v].append(u)
    d[root] = 0
    frontier, level = [root], 0
    while frontier:
        level += 1
        nxt = []
        for u in frontier:
            for v in adj[u]:
                if d[v] == _INF:
                    d[v] = level
                    nxt.append(v)
        frontier = nxt
    return d


def _ball(d, cutoff):
    return d <= max(int(cutoff), 0)


def _induced(ei, ea, first, member):
    """Columns / attribute rows of the relabelled induced subgraph: kept pairs in source order, both directions adjacent."""
    new = np.cumsum(member) - 1
    cols, rows = [], []
    for p in np.nonzero(first)[0]:
        u, v = int(ei[0, 2 * p]), int(ei[1, 2 * p])
        if member[u] and member[v]:
            cols += [(new[u], new[v]), (new[v], new[u])]
            rows += [ea[2 * p], ea[2 * p]]
    ei2 = np.array(cols, dtype=np.int64).T.reshape(2, -1) if cols else np.zeros((2, 0), np.int64)
    ea2 = np.array(rows) if rows else np.zeros((0,) + tuple(np.shape(ea)[1:]), dtype=np.asarray(ea).dtype)
    return new, ei2, ea2


def extract_pair(x, ei, ea, root, k, l1, l2, whole_graph=False):
    """One graph -> dict (graph-local numbering) or None when the context is empty."""
    n = len(x)
    first = first_pairs(ei)
    d = hop_distance(n, ei, root, first)
    in_s = np.ones(n, dtype=bool) if whole_graph else _ball(d, k)
    in_c = ~_ball(d, l1) if whole_graph else (_ball(d, l1) ^ _ball(d, l2))
    if not in_c.any():
        return None
    new_s, ei_s, ea_s = _induced(ei, ea, first, in_s)
    new_c, ei_c, ea_c = _induced(ei, ea, first, in_c)
    if whole_graph and ea_c.shape[0]:
        ea_c = ea_c.copy()
        ea_c[:, 7:] = 0          # bio/loader.py:60-62: nx_to_graph_data_obj re-emits w1..w7 and zeros for self-loop / mask
    return dict(x_substruct=np.asarray(x)[in_s], edge_index_substruct=ei_s, edge_attr_substruct=ea_s, center_substruct_idx=int(new_s[root]),
                x_context=np.asarray(x)[in_c], edge_index_context=ei_c, edge_attr_context=ea_c,
                overlap_context_substruct_idx=new_c[in_s & in_c].astype(np.int64), nodes_substruct=np.nonzero(in_s)[0], nodes_context=np.nonzero(in_c)[0])


def extract_pairs_batch(graphs, ids, roots, k, l1, l2, whole_graph=False):
    """graphs[g] = (x [n,F], edge_index [2,e] local, edge_attr [e,A]); roots[i] = graph-local root of batch slot i.
    -> the BatchSubstructContext fields (chem/batch.py:150-205) + `kept` (batch slots that survived)."""
    keys = ("x_substruct", "edge_index_substruct", "edge_attr_substruct", "x_context", "edge_index_context", "edge_attr_context")
    acc = {k_: [] for k_ in keys}
    center, overlap, seg, sizes, kept = [], [], [], [], []
    cs = cc = 0
    for slot, (g, r) in enumerate(zip(ids, roots)):
        x, ei, ea = graphs[g]
        p = extract_pair(np.asarray(x), np.asarray(ei), np.asarray(ea), int(r), k, l1, l2, whole_graph)
        if p is None:
            continue
        i = len(kept)
        kept.append(slot)
        acc["x_substruct"].append(p["x_substruct"])
        acc["edge_index_substruct"].append(p["edge_index_substruct"] + cs)
        acc["edge_attr_substruct"].append(p["edge_attr_substruct"])
        acc["x_context"].append(p["x_context"])
        acc["edge_index_context"].append(p["edge_index_context"] + cc)
        acc["edge_attr_context"].append(p["edge_attr_context"])
        center.append(p["center_substruct_idx"] + cs)
        overlap.append(p["overlap_context_substruct_idx"] + cc)
        seg.append(np.full(len(p["overlap_context_substruct_idx"]), i, dtype=np.int64))
        sizes.append(len(p["overlap_context_substruct_idx"]))
        cs += len(p["x_substruct"])
        cc += len(p["x_context"])
    x0, a0 = np.asarray(graphs[0][0]), np.asarray(graphs[0][2])
    empty = dict(x_substruct=np.zeros((0,) + x0.shape[1:], x0.dtype), x_context=np.zeros((0,) + x0.shape[1:], x0.dtype),
                 edge_index_substruct=np.zeros((2, 0), np.int64), edge_index_context=np.zeros((2, 0), np.int64),
                 edge_attr_substruct=np.zeros((0,) + a0.shape[1:], a0.dtype), edge_attr_context=np.zeros((0,) + a0.shape[1:], a0.dtype))
    out = {k_: (np.concatenate(v, axis=1 if k_.startswith("edge_index") else 0) if v else empty[k_]) for k_, v in acc.items()}
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    out.update(center_substruct_idx=np.array(center, dtype=np.int64), overlap_context_substruct_idx=cat(overlap),
               batch_overlapped_context=cat(seg), overlapped_context_size=np.array(sizes, dtype=np.int64), kept=np.array(kept, dtype=np.int64))
    return out


def draw_roots(node_counts, seed):
    """The root draw pgnn_extract_pairs defines when no roots are given: splitmix64(seed, batch slot) mod n (the reference
    uses random.sample(range(n), 1), chem/util.py:100-101: a uniform draw, RNG parity impossible by construction)."""
    return np.array([splitmix64(seed & _M64, i) % int(n) if n > 0 else 0 for i, n in enumerate(node_counts)], dtype=np.int64)
