"""CPU, gloo, world_size 2: the data-parallel host logic (graph sharding + flat gradient all-reduce).

Per SURVEY.md section 8(e): the reduced gradient must equal the mean of the per-shard gradients; BatchNorm
statistics stay per rank.  The per-shard gradients come from the CPU oracle here (the CUDA path needs a
GPU); what is under test is pretrain-gnns_b200/dist.py."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gnn_oracle as O
    syn = importlib.import_module("pretrain-gnns_b200.synthetic")
    pdist = importlib.import_module("pretrain-gnns_b200.dist")
    torch.set_num_threads(1)
    P = O.leaf_params(O.make_params("chem", "gin", 2, 300, seed=1))
    params = [v for v in P.values() if v.requires_grad]
    b = syn.zinc_batch(2, 50 + rank)
    rep = O.chem_gnn(P, b["x"], b["edge_index"], b["edge_attr"], 2, "gin", True)
    rep.square().mean().backward()
    local = [p.grad.clone() for p in params]
    red = pdist.GradAllReducer(params, overlap=False)
    red.all_reduce_mean()
    res = {"local": local, "reduced": [p.grad.clone() for p in params]}

    # second step through the overlapped route: the first 4 tensors stand in for the fused encoder (their gradients are
    # views of one flat buffer, reduced in place), the rest are packed and launched from the gradient hooks
    flat_params, rest = params[:4], params[4:]
    state = {"flat": None}
    red2 = pdist.GradAllReducer(params, flat_sources=[lambda: (state["flat"], flat_params)], overlap=True)
    for p in params:
        p.grad = None
    rep = O.chem_gnn(P, b["x"], b["edge_index"], b["edge_attr"], 2, "gin", True)
    (2.0 * rep.square().mean()).backward()
    assert red2._pending is not None  # launched from the hooks, before all_reduce_mean is called
    state["flat"] = torch.cat([p.grad.reshape(-1) for p in flat_params])
    for p, v in zip(flat_params, state["flat"].split([p.numel() for p in flat_params])):
        p.grad = v.view_as(p)
    res["local2"] = [p.grad.clone() for p in params]
    red2.all_reduce_mean()
    res["reduced2"] = [p.grad.clone() for p in params]
    red2.close()
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_is_mean_of_shards(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt")) for k in range(world)]
    for a, b, m0, m1 in zip(r[0]["local"], r[1]["local"], r[0]["reduced"], r[1]["reduced"]):
        assert torch.allclose(m0, (a + b) / 2, atol=1e-7, rtol=1e-6)
        assert torch.equal(m0, m1)
    for a, b, m0, m1 in zip(r[0]["local2"], r[1]["local2"], r[0]["reduced2"], r[1]["reduced2"]):
        assert torch.allclose(m0, (a + b) / 2, atol=1e-7, rtol=1e-6)
        assert torch.equal(m0, m1)


def test_shard_graphs_partitions_exactly():
    pdist = importlib.import_module("pretrain-gnns_b200.dist")
    for n in (0, 1, 7, 256, 513):
        for w in (1, 2, 4, 8):
            spans = [pdist.shard_graphs(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
