"""CPU: host logic and oracles of the collation / optimizer rows."""
import importlib

import numpy as np
import pytest
import torch

from oracle import step_io_oracle as SO

syn = importlib.import_module("pretrain-gnns_b200.synthetic")
optim = importlib.import_module("pretrain-gnns_b200.optim")
data = importlib.import_module("pretrain-gnns_b200.data")


def test_collate_oracle_matches_literal_cat_loop():
    """The restatement against the reference's algorithm written out with torch.cat (chem/batch.py:33-52)."""
    graphs = syn.split_graphs(syn.zinc_batch(12, 5))
    ids = [7, 0, 3, 3, 11]
    got = SO.collate_chem(graphs, ids)
    xs, eis, eas, bs, cum = [], [], [], [], 0
    for i, g in enumerate(ids):
        x, ei, ea = (torch.from_numpy(np.ascontiguousarray(a)) for a in graphs[g])
        bs.append(torch.full((x.shape[0],), i, dtype=torch.long))
        xs.append(x); eis.append(ei + cum); eas.append(ea)
        cum += x.shape[0]
    assert np.array_equal(got["x"], torch.cat(xs, 0).numpy())
    assert np.array_equal(got["edge_index"], torch.cat(eis, -1).numpy())
    assert np.array_equal(got["edge_attr"], torch.cat(eas, 0).numpy())
    assert np.array_equal(got["batch"], torch.cat(bs, -1).numpy())
    assert got["node_off"][-1] == cum and got["edge_off"][-1] == got["edge_index"].shape[1]


def test_split_graphs_roundtrip():
    b = syn.zinc_batch(9, 2)
    got = SO.collate_chem(syn.split_graphs(b), list(range(9)))
    for k in ("x", "edge_index", "edge_attr", "batch"):
        assert np.array_equal(got[k], b[k].numpy()), k


def test_store_host_logic():
    graphs = syn.split_graphs(syn.zinc_batch(6, 1))
    D = [type("D", (), dict(x=g[0], edge_index=g[1], edge_attr=g[2])) for g in graphs]
    st = data.MoleculeStore.from_data_list(D, device="cpu")
    assert st.num_graphs == 6 and st.x.dtype == torch.uint8 and st.edge_index.dtype == torch.int32
    n, e = st.batch_sizes([5, 5, 0])
    assert n == 2 * graphs[5][0].shape[0] + graphs[0][0].shape[0] and e == 2 * graphs[5][1].shape[1] + graphs[0][1].shape[1]
    assert st.batch_sizes([]) == (0, 0)
    with pytest.raises(IndexError):
        st.batch_sizes([6])
    with pytest.raises(ValueError):
        data.MoleculeStore([0, 3], [0, 2], np.zeros((3, 2)), np.zeros((2, 3)), np.zeros((2, 2)), device="cpu")


def test_chunk_table():
    t = optim.chunk_table([(1000, 2000, 3000, 4000), (16, 32, 48, 64)], [10000, 5], chunk=4096)
    assert t["n"].tolist() == [4096, 4096, 1808, 5]
    assert t["param"].tolist() == [1000, 1000 + 4 * 4096, 1000 + 4 * 8192, 16]
    assert t["exp_avg_sq"].tolist()[1] == 4000 + 4 * 4096 and t.dtype.itemsize == 40


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_oracle_matches_torch(wd):
    """Pins the restatement (modern eps placement) to torch.optim.Adam of this image over several steps."""
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(257, generator=g, dtype=torch.float64)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-3, weight_decay=wd)
    q, m, v = p0.numpy().copy(), np.zeros(257), np.zeros(257)
    for step in range(1, 6):
        grad = torch.randn(257, generator=g, dtype=torch.float64) * (10.0 ** (step - 3))
        p.grad = grad.clone()
        opt.step()
        q, m, v = SO.adam_step(q, grad.numpy(), m, v, step, lr=1e-3, weight_decay=wd)
        assert np.allclose(q, p.detach().numpy(), rtol=1e-12, atol=1e-14)
    # legacy placement differs from the modern one only through eps
    a = SO.adam_step(q, grad.numpy(), m, v, 6, eps=0.0)[0]
    b = SO.adam_step(q, grad.numpy(), m, v, 6, eps=0.0, legacy_eps=True)[0]
    assert np.allclose(a, b, rtol=1e-13, atol=0)


def test_adam_ctor_errors_match_torch():
    with pytest.raises(ValueError):
        optim.Adam([])
    with pytest.raises(ValueError):
        optim.Adam([torch.nn.Parameter(torch.zeros(3))])  # CPU parameter: no fallback


def test_batch_stager_roundtrip_cpu():
    """Packing into one buffer, slot rotation and the views handed out (the CUDA side adds streams and events only)."""
    st = data.BatchStager("cpu")
    batches = []
    for seed in range(5):
        b = syn.mask_atoms(syn.zinc_batch(3 + seed, seed), seed)
        b = {k: v for k, v in b.items() if torch.is_tensor(v)}
        b["weights"] = torch.randn(7, 3)              # a second dtype and an odd byte length
        b["empty"] = torch.zeros(0, 2, dtype=torch.int64)
        batches.append(b)
    packed = [st.pack(b) for b in batches]
    assert all(p.nbytes % 16 == 0 or p.nbytes == 16 for p in packed)
    ticket = st.submit(packed[0])
    for i in range(5):
        got = st.take(ticket)
        ticket = st.submit(packed[(i + 1) % 5])      # prefetch: must not disturb the batch in use
        for k, v in batches[i].items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and torch.equal(got[k], v), (i, k)


def test_pair_first_flags_match_the_oracle_and_odd_edge_counts_are_rejected():
    """data._pair_first (host, once per store): first occurrence of every unordered bond per graph, as networkx keeps them
    (chem/loader.py:173), against oracle.first_pairs graph by graph; a graph with an odd number of edge columns is refused."""
    data = importlib.import_module("pretrain-gnns_b200.data")
    graphs = syn.split_graphs(syn.zinc_batch(30, 5))
    edge_ptr = np.concatenate([[0], np.cumsum([g[1].shape[1] for g in graphs])])
    ei = np.concatenate([g[1] for g in graphs], axis=1)
    flags = data._pair_first(edge_ptr, ei)
    ref = np.concatenate([SO.first_pairs(g[1]) for g in graphs]).astype(np.uint8)
    assert np.array_equal(flags, ref) and 0 < int((1 - flags).sum()) < len(flags) // 4    # the generator does repeat a few bonds
    with pytest.raises(ValueError):
        data._pair_first(np.array([0, 3]), ei[:, :3])


def test_transform_size_queries_on_the_host():
    """The host-side size functions of the f4 entry points (no GPU involved): pgnn_mask_edges_bio_count equals the oracle's draw
    sizes (int(e/2 * rate + 1) per graph, 0 for a graph without edges), the workspace queries grow with their arguments and
    refuse negative sizes."""
    import ctypes
    cabi = importlib.import_module("pretrain-gnns_b200._cabi")
    dll = cabi.lib.load()
    edge_off = np.array([0, 40, 40, 46, 1046], dtype=np.int64)
    n = dll.pgnn_mask_edges_bio_count(edge_off.ctypes.data_as(ctypes.c_void_p), 4, 0.15)
    assert n == sum(len(c) for c in SO.mask_edge_choice_bio(edge_off, 0.15, seed=1)) == 4 + 0 + 1 + 76
    assert dll.pgnn_extract_pairs_workspace_bytes(8, 200) < dll.pgnn_extract_pairs_workspace_bytes(8, 4000)
    assert dll.pgnn_extract_pairs_workspace_bytes(-1, 10) < 0 and dll.pgnn_mask_edges_chem_workspace_bytes(-1, 1) < 0
    assert dll.pgnn_mask_edges_chem_workspace_bytes(6000, 256) >= 6000 + 256 * 8
