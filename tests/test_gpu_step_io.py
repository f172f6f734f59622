"""GPU: device collation (bit-exact) and the single-launch Adam against their oracles."""
import importlib

import numpy as np
import pytest
import torch

from oracle import step_io_oracle as SO

pytestmark = pytest.mark.gpu
syn = importlib.import_module("pretrain-gnns_b200.synthetic")
optim = importlib.import_module("pretrain-gnns_b200.optim")
data = importlib.import_module("pretrain-gnns_b200.data")
chem = importlib.import_module("pretrain-gnns_b200.chem.model")
DEV = "cuda:0"


def _store(num, seed):
    b = syn.zinc_batch(num, seed)
    graphs = syn.split_graphs(b)
    D = [type("D", (), dict(x=g[0], edge_index=g[1], edge_attr=g[2])) for g in graphs]
    return data.MoleculeStore.from_data_list(D, device=DEV), graphs, b


def _store_from(graphs):
    D = [type("D", (), dict(x=g[0], edge_index=g[1], edge_attr=g[2])) for g in graphs]
    return data.MoleculeStore.from_data_list(D, device=DEV)


@pytest.mark.parametrize("ids", [[3], [0, 1, 2, 3], [9, 9, 2, 0, 9], list(range(39, -1, -1)), []])
def test_collate_bit_exact(ids):
    st, graphs, _ = _store(40, 3)
    got = st.collate(ids)
    ref = SO.collate_chem(graphs, ids)
    for k, r in ref.items():
        g = getattr(got, k).cpu().numpy()
        assert g.dtype == np.int64 and g.shape == r.shape and np.array_equal(g, r), k


def test_collate_large_batch_and_model_equivalence():
    """B > one scan tile (1024), random order with repeats; and the collated batch drives GNN.forward to the same
    node representations as the host-collated tensors."""
    st, graphs, _ = _store(300, 4)
    ids = np.random.default_rng(0).integers(0, 300, size=2500)
    got, ref = st.collate(ids), SO.collate_chem(graphs, ids)
    for k, r in ref.items():
        assert np.array_equal(getattr(got, k).cpu().numpy(), r), k
    ids = np.arange(64)
    got, ref = st.collate(ids), SO.collate_chem(graphs, ids)
    torch.manual_seed(0)
    gnn = chem.GNN(3, 300).to(DEV)
    a = gnn(got.x, got.edge_index, got.edge_attr)
    b = gnn(*(torch.from_numpy(ref[k]).to(DEV) for k in ("x", "edge_index", "edge_attr")))
    assert torch.equal(a, b)


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) for s in shapes]


SHAPES = [(120, 300), (600, 300), (600,), (7,), (3, 5), (1,), (4099,)]


@pytest.mark.parametrize("wd,scale", [(0.0, 1.0), (0.01, 1.0), (0.0, 0.25)])
def test_adam_matches_torch(wd, scale):
    """5 steps against torch.optim.Adam on CPU fp32 (gradient magnitudes spanning 1e-4..1e2, one tensor without a
    gradient, odd sizes for the unaligned tail).  fp32 elementwise arithmetic: 2e-6 relative on the state, and on the
    parameters 2e-6 relative plus 4 ulp of the step size (lr)."""
    lr = 1e-3
    P = _params(1, SHAPES)
    ref_p = [torch.nn.Parameter(p.clone()) for p in P]
    my_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in P]
    ref = torch.optim.Adam(ref_p, lr=lr, weight_decay=wd)
    mine = optim.Adam(my_p, lr=lr, weight_decay=wd, grad_scale=scale)
    for step in range(5):
        G = _params(10 + step, SHAPES)
        for i, (r, m, g) in enumerate(zip(ref_p, my_p, G)):
            if i == 3 and step < 4:
                r.grad = m.grad = None  # never stepped until the last iteration -> skipped like torch skips it
                continue
            g = g * 10.0 ** (step - 2)
            r.grad = (g * scale).clone()
            m.grad = g.to(DEV)
        ref.step()
        mine.step()
    for i, (r, m) in enumerate(zip(ref_p, my_p)):
        # tensor 3 is stepped once, at iteration 5: torch keeps a per-tensor step count and so does this optimizer
        err = (m.detach().cpu() - r.detach()).abs()
        assert bool((err <= 2e-6 * r.detach().abs() + 4e-7 * lr * 10).all()), (i, err.max().item())
        st = mine.state[m]
        for k in ("exp_avg", "exp_avg_sq"):
            a, b = st[k].cpu(), ref.state[r][k]
            assert bool(((a - b).abs() <= 2e-6 * b.abs() + 1e-6 * b.abs().max()).all()), (i, k)  # lerp cancels near zero


def test_adam_legacy_eps_vs_oracle():
    P = _params(2, [(513,)])[0]
    p = torch.nn.Parameter(P.clone().to(DEV))
    opt = optim.Adam([p], lr=1e-2, eps=1e-3, weight_decay=0.05, legacy_eps=True)
    q, m, v = P.double().numpy(), np.zeros(513), np.zeros(513)
    for step in range(1, 4):
        g = _params(20 + step, [(513,)])[0]
        p.grad = g.to(DEV)
        opt.step()
        q, m, v = SO.adam_step(q, g.double().numpy(), m, v, step, lr=1e-2, eps=1e-3, weight_decay=0.05, legacy_eps=True)
    assert np.allclose(p.detach().cpu().double().numpy(), q, rtol=3e-6, atol=1e-7)


def test_adam_state_dict_roundtrip_with_torch():
    P = _params(3, [(10, 4), (9,)])
    my_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in P]
    t_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in P]
    mine, ref = optim.Adam(my_p, lr=2e-3), torch.optim.Adam(t_p, lr=2e-3)
    for step in range(2):
        for a, b, g in zip(my_p, t_p, _params(30 + step, [(10, 4), (9,)])):
            a.grad, b.grad = g.to(DEV), g.to(DEV)
        mine.step(); ref.step()
    ref2 = torch.optim.Adam(t_p, lr=1.0)
    ref2.load_state_dict(mine.state_dict())          # our checkpoint loads into torch's optimizer
    assert ref2.param_groups[0]["lr"] == 2e-3
    mine2 = optim.Adam(my_p, lr=1.0)
    mine2.load_state_dict(ref.state_dict())          # and torch's into ours
    assert mine2._steps == [2, 2] and mine2.param_groups[0]["lr"] == 2e-3
    for a, b in zip(my_p, t_p):
        assert torch.allclose(mine2.state[a]["exp_avg"], ref.state[b]["exp_avg"], rtol=2e-6, atol=0)


def test_adam_param_groups_like_finetune():
    """chem/finetune.py:180-185: param-group dicts with a per-group lr (lr * lr_scale for the prediction head)."""
    P = _params(4, [(33, 7), (12,), (5, 5)])
    my_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in P]
    t_p = [torch.nn.Parameter(p.clone()) for p in P]
    groups = lambda ps: [{"params": ps[:2]}, {"params": ps[2:], "lr": 1e-3 * 7.0, "weight_decay": 0.0}]
    mine, ref = optim.Adam(groups(my_p), lr=1e-3, weight_decay=0.01), torch.optim.Adam(groups(t_p), lr=1e-3, weight_decay=0.01)
    for step in range(3):
        for a, b, g in zip(my_p, t_p, _params(40 + step, [(33, 7), (12,), (5, 5)])):
            a.grad, b.grad = g.to(DEV), g.clone()
        mine.step(); ref.step()
    for a, b in zip(my_p, t_p):
        assert torch.allclose(a.detach().cpu(), b.detach(), rtol=3e-6, atol=1e-7)
    assert [g["lr"] for g in mine.state_dict()["param_groups"]] == [1e-3, 7e-3]
    with pytest.raises(TypeError):
        optim.Adam([{"lr": 1.0}])


def test_adam_steps_fused_encoder_gradients():
    """The flat gradient buffer of the fused encoder is consumed in place (every p.grad is a view of it), twice in a row.
    The torch twin is stepped with copies of the SAME gradients: bias gradients in front of a BatchNorm are pure rounding
    noise, which Adam normalises to full-size steps, so two backward passes must not be compared through an optimizer."""
    b = syn.zinc_batch(16, 7)
    torch.manual_seed(1)
    gnn = chem.GNN(3, 300).to(DEV)
    twin = [torch.nn.Parameter(p.detach().clone()) for p in gnn.parameters()]
    mine, ref = optim.Adam(gnn.parameters(), lr=1e-3, weight_decay=1e-4), torch.optim.Adam(twin, lr=1e-3, weight_decay=1e-4)
    args = [b[k].to(DEV) for k in ("x", "edge_index", "edge_attr")]
    for _ in range(2):
        mine.zero_grad()
        gnn(*args).square().mean().backward()
        flat = gnn._fused_plan().last_flat_grad
        assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in gnn.parameters())
        for t, p in zip(twin, gnn.parameters()):
            t.grad = p.grad.clone()
        mine.step()
        ref.step()
    for (n, a), c in zip(gnn.named_parameters(), twin):
        assert torch.allclose(a, c, rtol=2e-6, atol=1e-8), n


def test_mask_atoms_device_vs_oracle():
    """pgnn_mask_atoms on a device-collated batch: bit-exact against oracle/step_io_oracle.mask_atoms (which is pinned to
    the reference's MaskAtom + BatchMasking in tests/test_oracle_vs_reference.py); B > 1024 exercises the scan's carry."""
    for G, B, seed in ((40, 16, 7), (300, 1500, 123456789)):
        store, graphs, _ = _store(G, 5)
        rng = np.random.default_rng(seed)
        ids = rng.integers(0, G, size=B)
        batch = store.collate(ids)
        off = store.node_offsets_host(ids)
        ref_x, ref_idx, ref_lab, ref_off = SO.mask_atoms(batch.x.cpu().numpy(), off, 0.15, seed)
        data.mask_atoms(batch, off, 0.15, seed)
        assert np.array_equal(batch.node_off.cpu().numpy(), off)
        assert np.array_equal(batch.mask_off.cpu().numpy(), ref_off)
        assert np.array_equal(batch.masked_atom_indices.cpu().numpy(), ref_idx)
        assert np.array_equal(batch.mask_node_label.cpu().numpy(), ref_lab)
        assert np.array_equal(batch.x.cpu().numpy(), ref_x)
        assert int((batch.x[:, 0] == 119).sum()) == len(ref_idx)


def test_substruct_context_and_bio_collate_device_vs_oracle():
    G = 24
    sub, ctx = syn.zinc_batch(G, 61, n_lo=12, n_hi=22), syn.zinc_batch(G, 62, n_lo=4, n_hi=14, tree_only=True)
    sg, cg = syn.split_graphs(sub), syn.split_graphs(ctx)
    rng = np.random.default_rng(5)
    center = [int(rng.integers(0, len(g[0]))) for g in sg]
    overlap = [sorted(rng.choice(len(g[0]), size=int(rng.integers(1, min(4, len(g[0])) + 1)), replace=False).tolist()) for g in cg]
    st = data.SubstructContextStore(_store_from(sg), _store_from(cg), center, overlap)
    ids = rng.integers(0, G, size=50)
    out = st.collate(ids)
    s, c = SO.collate_chem(sg, ids), SO.collate_chem(cg, ids)
    cen, _, _, _ = SO.collate_lists(np.arange(G + 1), np.array(center), ids, add=s["node_off"])
    ov, seg, sizes, _ = SO.collate_lists(np.cumsum([0] + [len(o) for o in overlap]), np.concatenate(overlap), ids, add=c["node_off"])
    for k, v in (("x_substruct", s["x"]), ("edge_index_substruct", s["edge_index"]), ("edge_attr_substruct", s["edge_attr"]),
                 ("center_substruct_idx", cen), ("x_context", c["x"]), ("edge_index_context", c["edge_index"]),
                 ("edge_attr_context", c["edge_attr"]), ("overlap_context_substruct_idx", ov), ("batch_overlapped_context", seg),
                 ("overlapped_context_size", sizes)):
        assert np.array_equal(getattr(out, k).cpu().numpy(), v), k
    # bio
    pb = syn.ppi_batch(6, 9, n_lo=20, n_hi=35, num_tasks=4)
    ptr = pb["ptr"].numpy()
    ei, ea = pb["edge_index"].numpy(), pb["edge_attr"].numpy()
    ea[::7, 8] = 1.0   # exercise the ninth (mask) bit as well
    owner = np.searchsorted(ptr, ei[0], side="right") - 1
    eptr = np.searchsorted(owner, np.arange(len(ptr)))
    graphs = [(int(ptr[g + 1] - ptr[g]), ei[:, eptr[g]:eptr[g + 1]] - ptr[g], ea[eptr[g]:eptr[g + 1]]) for g in range(6)]
    centers = [0, 3, 1, 0, 2, 5]
    bs = data.BioGraphStore([g[0] for g in graphs], [g[1] for g in graphs], [g[2] for g in graphs], centers)
    ids = np.array([5, 0, 0, 3])
    o = bs.collate(ids)
    ref = SO.collate_bio(graphs, ids)
    for k in ("x", "edge_index", "edge_attr", "batch", "node_off", "edge_off"):
        assert np.array_equal(getattr(o, k).cpu().numpy(), ref[k]), k
    cen, _, _, _ = SO.collate_lists(np.arange(7), np.array(centers), ids, add=ref["node_off"])
    assert np.array_equal(o.center_node_idx.cpu().numpy(), cen)


_PAIR_KEYS = ("x_substruct", "edge_index_substruct", "edge_attr_substruct", "center_substruct_idx", "x_context", "edge_index_context",
              "edge_attr_context", "overlap_context_substruct_idx", "batch_overlapped_context", "overlapped_context_size")


@pytest.mark.parametrize("k,l1,l2", [(5, 4, 7), (2, 1, 3), (1, 0, 2), (0, 0, 1), (2, 5, 3)])
def test_extract_substruct_context_pairs_device_vs_oracle(k, l1, l2):
    """ExtractSubstructureContextPair + BatchSubstructContext on the device (pgnn_extract_pairs / pgnn_extract_fill_chem) against
    oracle/step_io_oracle.extract_pairs_batch (itself checked against the reference's chem/util.py + chem/batch.py in
    tests/test_oracle_vs_reference.py): bit-exact, for the script's default radii, small radii, the 0 -> -1 quirk and l1 > l2;
    repeated graph ids, molecules with duplicate bonds, one-atom molecules (no context: dropped), given and drawn roots."""
    G = 40
    graphs = syn.split_graphs(syn.zinc_batch(G, 81))
    graphs.append((graphs[1][0][:1], np.zeros((2, 0), np.int64), np.zeros((0, 2), np.int64)))   # a one-atom molecule
    st = _store_from(graphs)
    rng = np.random.default_rng(11)
    ids = np.concatenate([rng.integers(0, G, size=70), [G, 3, G]])
    rng.shuffle(ids)
    n = np.array([len(graphs[g][0]) for g in ids])
    roots = (rng.integers(0, 1 << 30, size=len(ids)) % n).astype(np.int64)
    for given in (True, False):
        if not given:
            roots = SO.draw_roots(n, seed=4242)
        out = data.extract_substruct_context_pairs(st, ids, k, l1, l2, roots=roots if given else None, seed=4242)
        ref = SO.extract_pairs_batch(graphs, ids, roots, k, l1, l2)
        assert out.kept == len(ref["kept"]) and out.kept <= len(ids) - 2
        for key in _PAIR_KEYS:
            mine = getattr(out, key).cpu().numpy()
            assert mine.shape == ref[key].shape and np.array_equal(mine, ref[key]), (key, given)
        assert out.edge_index_substruct.is_contiguous() and out.edge_index_context.is_contiguous()
    assert not importlib.import_module("pretrain-gnns_b200.ops").device_errors()


def test_extracted_pairs_feed_the_contextpred_step():
    """The namespace extract_substruct_context_pairs returns is what ContextPredStep consumes (chem/pretrain_contextpred.py:50-97):
    one training step on pairs extracted on the device runs and gives the loss the CPU oracle computes on the oracle's extraction."""
    from oracle import steps_oracle as S
    ts = importlib.import_module("pretrain-gnns_b200.train_steps")
    graphs = syn.split_graphs(syn.zinc_batch(48, 83))
    st = _store_from(graphs)
    ids = np.arange(48)
    roots = SO.draw_roots([len(g[0]) for g in graphs], seed=7)
    out = data.extract_substruct_context_pairs(st, ids, 5, 4, 7, seed=7)
    ref = SO.extract_pairs_batch(graphs, ids, roots, 5, 4, 7)
    b = {key: torch.from_numpy(np.ascontiguousarray(ref[key])) for key in _PAIR_KEYS}
    b["num_graphs"] = int(len(ref["kept"]))
    step = ts.ContextPredStep(torch.device(DEV), batch_size=b["num_graphs"])
    P = S.make_params("contextpred", 3)
    step.load_state(P)
    d = {key: getattr(out, key) for key in _PAIR_KEYS}
    d["num_graphs"] = out.kept
    loss = step(d)
    l64 = S.grads_fp32_fp64(S.LOSSES["contextpred"], P, b)[3]
    assert abs(float(loss) - float(l64)) <= 2e-6 * max(1.0, abs(float(l64))), (float(loss), float(l64))


def test_extract_context_bio_device_vs_oracle():
    """bio ExtractSubstructureContextPair(l1, center=True) on the device against the oracle (bio/util.py:123-205): context = nodes
    further than l1 hops from the centre, every one an overlap node, self-loop / mask columns zeroed by the networkx round trip."""
    pb = syn.ppi_batch(6, 29, n_lo=60, n_hi=90, pairs_per_node=2, num_tasks=4)
    ptr = pb["ptr"].numpy()
    ei, ea = pb["edge_index"].numpy(), pb["edge_attr"].numpy().copy()
    ea[0:2, 8] = 1.0
    owner = np.searchsorted(ptr, ei[0], side="right") - 1
    eptr = np.searchsorted(owner, np.arange(len(ptr)))
    graphs = [(int(ptr[g + 1] - ptr[g]), ei[:, eptr[g]:eptr[g + 1]] - ptr[g], ea[eptr[g]:eptr[g + 1]]) for g in range(6)]
    centers = [0, 3, 1, 0, 2, 5]
    bs = data.BioGraphStore([g[0] for g in graphs], [g[1] for g in graphs], [g[2] for g in graphs], centers)
    ids = np.array([5, 0, 0, 3, 2])
    ograph = [(np.ones((g[0], 1), np.float32), g[1], g[2]) for g in graphs]
    for l1 in (1, 2):
        o = bs.extract_context(ids, l1)
        ref = SO.extract_pairs_batch(ograph, ids, [centers[g] for g in ids], 0, l1, 0, whole_graph=True)
        assert o.kept == len(ids)
        for key in ("x_context", "edge_index_context", "edge_attr_context", "overlap_context_substruct_idx", "batch_overlapped_context",
                    "overlapped_context_size"):
            mine = getattr(o, key).cpu().numpy()
            assert mine.shape == ref[key].shape and np.array_equal(mine, ref[key]), (key, l1)
        full = SO.collate_bio(graphs, ids)
        assert np.array_equal(o.edge_index_substruct.cpu().numpy(), full["edge_index"])
        # whole-graph mode keeps every edge column of the substructure side only when no bond repeats; the centre is offset per graph
        cen, _, _, _ = SO.collate_lists(np.arange(7), np.array(centers), ids, add=full["node_off"])
        assert np.array_equal(o.center_substruct_idx.cpu().numpy(), cen)


@pytest.mark.parametrize("paired", [True, False])
def test_mask_edges_chem_device_vs_oracle(paired):
    """MaskAtom(mask_edge=True) on the device (pgnn_mask_atoms + pgnn_mask_edges_chem) against the oracle (itself equal to the
    reference's MaskAtom + BatchMasking): bit-exact; also on a batch with one direction of each bond deleted, where L[::2] no longer
    means "one column per bond" and only the literal rule gives the reference's answer."""
    st, graphs, _ = _store(60, 17)
    rng = np.random.default_rng(2)
    ids = rng.integers(0, 60, size=90)
    b = st.collate(ids)
    if not paired:
        keep = torch.from_numpy(np.sort(rng.choice(b.edge_index.shape[1], size=b.edge_index.shape[1] * 2 // 3, replace=False))).to(DEV)
        cnt = torch.zeros(len(ids) + 1, dtype=torch.int64, device=DEV)
        owner = torch.searchsorted(b.edge_off, keep, right=True) - 1
        cnt[1:] = torch.bincount(owner, minlength=len(ids))
        b.edge_index, b.edge_attr, b.edge_off = b.edge_index[:, keep].contiguous(), b.edge_attr[keep].contiguous(), torch.cumsum(cnt, 0)
    ei0, ea0, eoff = b.edge_index.cpu().numpy(), b.edge_attr.cpu().numpy().copy(), b.edge_off.cpu().numpy()
    x0 = b.x.cpu().numpy().copy()
    data.mask_atoms(b, st.node_offsets_host(ids), 0.15, seed=99)
    data.mask_edges_chem(b, num_edge_type=5)
    _, idx, _, _ = SO.mask_atoms(x0, st.node_offsets_host(ids), 0.15, seed=99)
    ea2, conn, lab, off = SO.mask_edges_chem(ei0, ea0, eoff, idx)
    assert np.array_equal(b.edge_attr.cpu().numpy(), ea2)
    assert np.array_equal(b.connected_edge_indices.cpu().numpy(), conn) and np.array_equal(b.mask_edge_label.cpu().numpy(), lab)
    assert np.array_equal(b.connected_edge_off.cpu().numpy(), off) and len(conn) > 0


def test_mask_edges_bio_device_vs_oracle():
    """MaskEdge on the device (pgnn_mask_edges_bio) against the oracle (equal to the reference's MaskEdge + bio BatchMasking):
    bit-exact, incl. a graph larger than the kernel's shared-memory key cache (4096 pairs) and a graph without edges."""
    pb = syn.ppi_batch(4, 37, n_lo=60, n_hi=90, pairs_per_node=3, num_tasks=4)
    big = syn.ppi_batch(1, 38, n_lo=900, n_hi=900, pairs_per_node=5, num_tasks=4)
    graphs = []
    for src in (pb, big):
        ptr = src["ptr"].numpy()
        ei, ea = src["edge_index"].numpy(), src["edge_attr"].numpy()
        owner = np.searchsorted(ptr, ei[0], side="right") - 1
        eptr = np.searchsorted(owner, np.arange(len(ptr)))
        graphs += [(int(ptr[g + 1] - ptr[g]), ei[:, eptr[g]:eptr[g + 1]] - ptr[g], ea[eptr[g]:eptr[g + 1]]) for g in range(len(ptr) - 1)]
    graphs.append((3, np.zeros((2, 0), np.int64), np.zeros((0, 9), np.float32)))
    bs = data.BioGraphStore([g[0] for g in graphs], [g[1] for g in graphs], [g[2] for g in graphs], [0] * len(graphs))
    ids = np.array([4, 0, 5, 2, 2])
    o = bs.collate(ids)
    ea0, eoff = o.edge_attr.cpu().numpy().copy(), o.edge_off.cpu().numpy()
    assert (eoff[1] - eoff[0]) // 2 > 4096
    data.mask_edges_bio(o, eoff, 0.15, seed=5)
    ea2, idx, lab, off = SO.mask_edges_bio(ea0, eoff, 0.15, seed=5)
    assert np.array_equal(o.edge_attr.cpu().numpy(), ea2) and np.array_equal(o.masked_edge_idx.cpu().numpy(), idx)
    assert np.array_equal(o.mask_edge_label.cpu().numpy(), lab) and np.array_equal(o.mask_edge_off.cpu().numpy(), off)
