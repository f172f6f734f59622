"""GPU, 2 ranks (skipped on a single-GPU box): the library's NVLink peer-memory all-reduce against NCCL on the real fused
encoder gradients -- mean of the two ranks' gradients, bit-identical on both ranks, several steps in a row, both the
write-in-place route and the copy-in route (parameters that still hold a gradient)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    syn = importlib.import_module("pretrain-gnns_b200.synthetic")
    chem = importlib.import_module("pretrain-gnns_b200.chem.model")
    pdist = importlib.import_module("pretrain-gnns_b200.dist")
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    gnn = chem.GNN(3, 300).to(dev).train()
    head = torch.nn.Linear(300, 7).to(dev)
    params = list(gnn.parameters()) + list(head.parameters())
    red = pdist.GradAllReducer(params, flat_sources=[pdist.encoder_flat_source(gnn)], backend="p2p")
    res = {"backend": red.backend, "steps": [], "state": {k: v.detach().cpu().clone() for k, v in gnn.state_dict().items()},
           "head": {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}}
    for step in range(4):
        b = syn.zinc_batch(8, 100 * step + rank)
        if step != 2:  # step 2 keeps the previous gradients: autograd accumulates, the encoder cannot write in place
            for p in params:
                p.grad = None
        head(gnn(*(b[k].to(dev) for k in ("x", "edge_index", "edge_attr")))).square().mean().backward()
        local = [p.grad.clone() for p in params]
        want = [g.clone() for g in local]
        for g in want:
            dist.all_reduce(g)
            g.div_(world)
        red.all_reduce_mean()
        torch.cuda.synchronize()
        res["steps"].append({"want": [g.cpu() for g in want], "got": [p.grad.detach().cpu().clone() for p in params],
                             "in_place": gnn._fused_plan().last_flat_grad.data_ptr() == red.regions[0].data_ptr()})
    red.close()
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def _oracle_shard_grads(state, head, rank, step):
    """Gradients of one rank's shard of one step from the CPU oracle (fp64): SURVEY.md 8(e) — the reduced gradient must
    equal the MEAN of the per-shard oracle gradients (BatchNorm statistics are per rank, like DDP without SyncBN)."""
    sys.path.insert(0, ROOT)
    from oracle import gnn_oracle as O
    syn = importlib.import_module("pretrain-gnns_b200.synthetic")
    b = syn.zinc_batch(8, 100 * step + rank)
    L = O.leaf_params(state, torch.float64)
    W, bias = head["weight"].double().requires_grad_(True), head["bias"].double().requires_grad_(True)
    rep = O.chem_gnn(L, b["x"], b["edge_index"], b["edge_attr"], 3, "gin", True)
    torch.nn.functional.linear(rep, W, bias).square().mean().backward()
    return [L[k].grad for k in state if k in L and L[k].requires_grad] + [W.grad, bias.grad]


def test_p2p_allreduce_matches_nccl_mean(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29100 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt")) for k in range(world)]
    assert r[0]["backend"] in ("p2p", "nvls")
    # step 0 (fresh gradients, parameters still at their initial values on both ranks): reduced gradient == mean of the
    # two shards' oracle gradients
    o0, o1 = (_oracle_shard_grads(r[0]["state"], r[0]["head"], k, 0) for k in range(world))
    gmax = max(float(((a + b) / 2).abs().max()) for a, b in zip(o0, o1))
    for a, b, got in zip(o0, o1, r[0]["steps"][0]["got"]):
        want = (a + b) / 2
        scale = max(float(want.abs().max()), 1e-3 * gmax)
        assert float((got.double() - want).abs().max()) <= 2e-4 * scale
    assert [s["in_place"] for s in r[0]["steps"]] == [True, True, False, True]
    for s0, s1 in zip(r[0]["steps"], r[1]["steps"]):
        for w, g0, g1 in zip(s0["want"], s0["got"], s1["got"]):
            assert torch.equal(g0, g1)  # every element was summed once, by one rank
            assert torch.allclose(g0, w, rtol=1e-6, atol=1e-9)  # (a + b) / 2 vs (a + b) * 0.5: same value up to the sum order
