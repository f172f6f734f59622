"""The `train()` bodies of the reference's pre-training scripts on the B200 modules — one callable per BASELINE config.

Each class owns the modules the script builds, draws the synthetic batches of its config and, called
on one device-resident batch, runs exactly what the script runs between `batch.to(device)` and `optimizer.step()`:
zero the gradients, forward, the script's own loss head, `loss.backward()`.  The optimizer step is outside (SURVEY.md
8(d) defines the metric without it).  bench.py times these; the parity tests compare them with the oracle's
restatement of the same bodies (oracle/steps_oracle.py) and with the reference's own modules.

    MaskingStep        chem/pretrain_masking.py:46-70     GNN(5,300,gnn_type) + Linear(300,119), CE on fp64 logits
    ContextPredStep    chem/pretrain_contextpred.py:50-97 GNN(5,300) + GNN(3,300), cbow / mean pooling, BCE on fp64 scores
    BioSupervisedStep  bio/pretrain_supervised.py:25-42   bio GNN_graphpred(5,300,T=5000), BCE on fp64 logits
"""
from __future__ import annotations

import types

import torch

from . import ops, synthetic as syn
from .bio import model as bio
from .chem import model as chem

NUM_LAYER, EMB = 5, 300


def _fields(b, keys):
    return {k: b[k] for k in keys}


CONFIG_ID = {"masking": 2, "contextpred": 3, "bio_supervised": 4, "gcn": 2, "gat": 2, "graphsage": 2}
DEFAULT_BATCH = {"masking": 256, "contextpred": 128, "bio_supervised": 64, "gcn": 256, "gat": 256, "graphsage": 256}
MASKING_KEYS = ("x", "edge_index", "edge_attr", "masked_atom_indices")
CONTEXT_KEYS = ("x_substruct", "edge_index_substruct", "edge_attr_substruct", "center_substruct_idx", "x_context", "edge_index_context",
                "edge_attr_context", "overlap_context_substruct_idx", "batch_overlapped_context")
BIO_KEYS = ("x", "edge_index", "edge_attr", "batch", "center_node_idx", "go_target_pretrain")


def make_batches(config, rank, count, batch_size=None, num_tasks=5000):
    """Host (CPU tensor) batches of one config: seed = config-id * 1000 + 1000 * rank + batch index."""
    B = batch_size or DEFAULT_BATCH[config]
    out = []
    for i in range(count):
        seed = CONFIG_ID[config] * 1000 + 1000 * rank + i
        if config == "contextpred":
            out.append(_fields(syn.substruct_context_batch(B, seed), CONTEXT_KEYS))
        elif config == "bio_supervised":
            out.append(_fields(syn.ppi_batch(B, seed, num_tasks=num_tasks), BIO_KEYS))
        else:
            b = syn.mask_atoms(syn.zinc_batch(B, seed), seed)
            out.append(_fields(b, MASKING_KEYS) | {"labels": b["mask_node_label"][:, 0].contiguous()})
    return out


class _Step:
    modules: list

    def parameters(self):
        # the module set of a step is fixed: walk it once (nn.Module.parameters() re-traverses every submodule on each call,
        # ~200 us of host time per training step when zero_grad does it)
        ps = getattr(self, "_params", None)
        if ps is None:
            ps = self._params = [p for m in self.modules for p in m.parameters()]
        return ps

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    def flat_sources(self):
        """(module, ...) whose fused backward leaves one flat gradient buffer (dist.encoder_flat_source)."""
        return []

    def named_modules(self):
        return {}

    def load_state(self, P):
        """P: one flat dict keyed `<attribute>.<state_dict key>` ('model.gnns.0...', 'head.weight', ...)."""
        for name, m in self.named_modules().items():
            dev = next(m.parameters()).device
            m.load_state_dict({k[len(name) + 1:]: v for k, v in P.items() if k.startswith(name + ".")})
            m.to(dev)

    def named_parameters(self):
        return [(name + "." + k, p) for name, m in self.named_modules().items() for k, p in m.named_parameters()]


class MaskingStep(_Step):
    """chem/pretrain_masking.py:46-70 (mask_edge off, the script's default): node_rep = model(x, ei, ea);
    pred = linear_pred_atoms(node_rep[masked_atom_indices]); loss = CE(pred.double(), mask_node_label[:,0])."""
    def __init__(self, device, gnn_type="gin", batch_size=256):
        self.graphs_per_batch = batch_size
        self.config = "masking" if gnn_type == "gin" else gnn_type
        self.model = chem.GNN(NUM_LAYER, EMB, JK="last", drop_ratio=0, gnn_type=gnn_type).to(device).train()
        self.head = torch.nn.Linear(EMB, 119).to(device)
        self.modules = [self.model, self.head]
        self.workload = ("chem pretrain_masking 5-layer %s emb_dim=300 batch_size=%d (BASELINE configs[%d])"
                         % (gnn_type.upper() if gnn_type != "graphsage" else "GraphSAGE", batch_size, 1 if gnn_type == "gin" else 4))

    def make_batches(self, rank, count):
        return make_batches(self.config, rank, count, self.graphs_per_batch)

    def flat_sources(self):
        return [self.model]

    def named_modules(self):
        return {"model": self.model, "head": self.head}

    def __call__(self, b):
        self.zero_grad()
        rep = self.model(b["x"], b["edge_index"], b["edge_attr"])
        loss, _ = ops.masked_atom_loss(rep, b["masked_atom_indices"], b["labels"], self.head.weight, self.head.bias)
        loss.backward()
        return loss


class ContextPredStep(_Step):
    """chem/pretrain_contextpred.py:50-97 with the script's defaults: num_layer 5, csize 3 -> context encoder of
    l2 - l1 = 7 - 4 = 3 layers (:145-146,156-157), mode cbow, context_pooling mean, neg_samples 1."""
    def __init__(self, device, batch_size=128, neg_samples=1):
        self.graphs_per_batch, self.neg_samples = batch_size, neg_samples
        self.model_substruct = chem.GNN(NUM_LAYER, EMB, JK="last", drop_ratio=0, gnn_type="gin").to(device).train()
        self.model_context = chem.GNN(3, EMB, JK="last", drop_ratio=0, gnn_type="gin").to(device).train()
        self.modules = [self.model_substruct, self.model_context]
        self.concurrent, self._side = True, None   # context encoder on a second CUDA stream (see scores)
        self.workload = "chem pretrain_contextpred 5-layer GIN emb_dim=300 batch_size=%d, substruct + 3-layer context encoder (BASELINE configs[2])" % batch_size

    KEYS = CONTEXT_KEYS

    def make_batches(self, rank, count):
        return make_batches("contextpred", rank, count, self.graphs_per_batch)

    def flat_sources(self):
        return [self.model_substruct, self.model_context]

    def named_modules(self):
        return {"model_substruct": self.model_substruct, "model_context": self.model_context}

    def scores(self, b):
        B = b["center_substruct_idx"].shape[0]
        # The two encoders are independent until the dot products (the script runs them back to back, :54-57), and at
        # B = 128 neither fills the chip (17 and 10 row tiles of 128 nodes): the context encoder runs on a second stream,
        # forward and — autograd replays a node's backward on its forward stream — backward alike.
        main = torch.cuda.current_stream() if b["x_context"].is_cuda else None
        if main is not None and self.concurrent:
            if self._side is None:
                self._side = torch.cuda.Stream(b["x_context"].device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                ov = ops.row_gather(self.model_context(b["x_context"], b["edge_index_context"], b["edge_attr_context"]),
                                    b["overlap_context_substruct_idx"])
            ov.record_stream(main)
        else:
            ov = ops.row_gather(self.model_context(b["x_context"], b["edge_index_context"], b["edge_attr_context"]),
                                b["overlap_context_substruct_idx"])
        sub = ops.row_gather(self.model_substruct(b["x_substruct"], b["edge_index_substruct"], b["edge_attr_substruct"]),
                             b["center_substruct_idx"])
        if main is not None and self.concurrent:
            main.wait_stream(self._side)
        ctx = ops.global_mean_pool(ov, b["batch_overlapped_context"], B)   # one segment per graph of the batch
        pos = ops.shifted_rowdot(sub, ctx, 0)
        neg = torch.cat([ops.shifted_rowdot(sub, ctx, i + 1) for i in range(self.neg_samples)], dim=0)
        return pos, neg

    def __call__(self, b):
        self.zero_grad()
        pos, neg = self.scores(b)
        loss = ops.bce_with_logits_const(pos, 1.0) + self.neg_samples * ops.bce_with_logits_const(neg, 0.0)
        loss.backward()
        return loss


class BioSupervisedStep(_Step):
    """bio/pretrain_supervised.py:25-42: pred = model(batch); loss = BCEWithLogits(pred.double(), y.double()).
    drop_ratio = 0 (the script's default is 0.2; dropout has no RNG parity)."""
    def __init__(self, device, gnn_type="gin", batch_size=64, num_tasks=5000):
        self.graphs_per_batch, self.num_tasks = batch_size, num_tasks
        self.model = bio.GNN_graphpred(NUM_LAYER, EMB, num_tasks, JK="last", drop_ratio=0, graph_pooling="mean", gnn_type=gnn_type).to(device).train()
        self.modules = [self.model]
        self.workload = "bio pretrain_supervised 5-layer GIN emb_dim=300 PPI-ego-shaped graphs batch_size=%d T=%d (BASELINE configs[3])" % (batch_size, num_tasks)

    KEYS = BIO_KEYS

    def make_batches(self, rank, count):
        return make_batches("bio_supervised", rank, count, self.graphs_per_batch, self.num_tasks)

    def flat_sources(self):
        return [self.model.gnn]

    def named_modules(self):
        return {"model": self.model}

    def __call__(self, b):
        self.zero_grad()
        pred = self.model(types.SimpleNamespace(**b))
        loss = ops.bce_with_logits(pred, b["go_target_pretrain"].view(pred.shape))
        loss.backward()
        return loss


CONFIGS = {
    "masking": lambda dev: MaskingStep(dev, "gin"),
    "contextpred": lambda dev: ContextPredStep(dev),
    "bio_supervised": lambda dev: BioSupervisedStep(dev),
    "gcn": lambda dev: MaskingStep(dev, "gcn"),
    "gat": lambda dev: MaskingStep(dev, "gat"),
    "graphsage": lambda dev: MaskingStep(dev, "graphsage"),
}
