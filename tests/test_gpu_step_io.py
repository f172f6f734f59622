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
        if i == 3:
            continue  # torch keeps a per-tensor step count; this optimizer keeps one (documented), so tensor 3 differs
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
    assert mine2._step == 2 and mine2.param_groups[0]["lr"] == 2e-3
    for a, b in zip(my_p, t_p):
        assert torch.allclose(mine2.state[a]["exp_avg"], ref.state[b]["exp_avg"], rtol=2e-6, atol=0)


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
