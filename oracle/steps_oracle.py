"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's train() bodies, one per BASELINE config.

Two flavours of each step:
  * `*_loss(L, b)`            on the oracle port (oracle/gnn_oracle.py) over `state_dict`-keyed leaf dictionaries: what the
                              parity tests differentiate in fp32 and fp64;
  * `Reference*Step(...)`     on the reference's OWN modules (oracle/reference_runner.py: chem/model.py, bio/model.py run
                              unmodified over the PyG stand-in), torch CPU: what `bench.py --impl reference` and the
                              `cpu_baseline` leg time (kind "reference"), and what the tests cross-check the port against.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as F

from . import gnn_oracle as O


def _cast(t, dtype):
    return t.to(dtype) if t.is_floating_point() else t


def sub(L, prefix):
    return {k[len(prefix):]: v for k, v in L.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------------------------
# oracle-port losses.  L: flat dict of leaves; keys as the product's modules name them under the given prefixes.
# ------------------------------------------------------------------------------------------------------------------
def masking_loss(L, b, gnn_type="gin", num_layer=5):
    """chem/pretrain_masking.py:48-52.  L: 'model.*' encoder, 'head.weight', 'head.bias'."""
    rep = O.chem_gnn(sub(L, "model."), b["x"], b["edge_index"], b["edge_attr"], num_layer, gnn_type, True)
    loss, logits = O.masking_loss(rep, b["masked_atom_indices"], b["labels"], L["head.weight"], L["head.bias"])
    return loss, dict(rep=rep, logits=logits)


def contextpred_loss(L, b, neg_samples=1):
    """chem/pretrain_contextpred.py:54-93 (cbow, mean).  L: 'model_substruct.*' (5 layers), 'model_context.*' (3 layers)."""
    B = b["center_substruct_idx"].shape[0]
    s = O.chem_gnn(sub(L, "model_substruct."), b["x_substruct"], b["edge_index_substruct"], b["edge_attr_substruct"], 5, "gin", True)
    c = O.chem_gnn(sub(L, "model_context."), b["x_context"], b["edge_index_context"], b["edge_attr_context"], 3, "gin", True)
    pos, neg = O.contextpred_scores(s[b["center_substruct_idx"]], c[b["overlap_context_substruct_idx"]], b["batch_overlapped_context"], B, neg_samples)
    return O.contextpred_loss(pos, neg, neg_samples), dict(pos=pos, neg=neg)


def bio_supervised_loss(L, b, gnn_type="gin"):
    """bio/pretrain_supervised.py:31-36.  L: 'model.gnn.*', 'model.graph_pred_linear.*'."""
    P = sub(L, "model.")
    dt = P["graph_pred_linear.bias"].dtype
    B = b["center_node_idx"].shape[0]
    pred = O.bio_graphpred(P, _cast(b["x"], dt), b["edge_index"], _cast(b["edge_attr"], dt), b["batch"], b["center_node_idx"], B, 5, gnn_type, True)
    y = b["go_target_pretrain"].view(pred.shape).to(torch.float64)
    return F.binary_cross_entropy_with_logits(pred.double(), y), dict(pred=pred)


def make_params(config, seed, randomize_bn=True, num_tasks=5000):
    """Seeded parameters of one config's modules as ONE flat dict, keyed `<module attribute>.<state_dict key>` — the
    naming both the product's train_steps classes and the Reference*Step classes load from."""
    g = torch.Generator().manual_seed(seed + 977)
    P = {}
    if config == "contextpred":
        P.update({"model_substruct." + k: v for k, v in O.make_params("chem", "gin", 5, 300, seed, randomize_bn=randomize_bn).items()})
        P.update({"model_context." + k: v for k, v in O.make_params("chem", "gin", 3, 300, seed + 1, randomize_bn=randomize_bn).items()})
    elif config == "bio_supervised":
        P.update({"model.gnn." + k: v for k, v in O.make_params("bio", "gin", 5, 300, seed, randomize_bn=randomize_bn).items()})
        P["model.graph_pred_linear.weight"] = torch.randn(num_tasks, 600, generator=g) * 0.03
        P["model.graph_pred_linear.bias"] = torch.randn(num_tasks, generator=g) * 0.03
    else:
        t = "gin" if config == "masking" else config
        P.update({"model." + k: v for k, v in O.make_params("chem", t, 5, 300, seed, randomize_bn=randomize_bn).items()})
        P["head.weight"] = torch.randn(119, 300, generator=g) * 0.05
        P["head.bias"] = torch.randn(119, generator=g) * 0.05
    return P


LOSSES = {"masking": masking_loss, "contextpred": contextpred_loss, "bio_supervised": bio_supervised_loss,
          "gcn": lambda L, b: masking_loss(L, b, "gcn"), "gat": lambda L, b: masking_loss(L, b, "gat"),
          "graphsage": lambda L, b: masking_loss(L, b, "graphsage")}


LAST_RELU_TRACE = None   # the fp64 run's ReLU pre-activations of the most recent grads_fp32_fp64 call (tests compare decisions)


def grads_fp32_fp64(loss_fn, P, b):
    """Differentiate `loss_fn` on fp32 and fp64 leaf copies of P.  -> (loss32, aux32, grads32, loss64, aux64, grads64, near_zero):
    near_zero = number of ReLU pre-activations of the fp64 run within rounding distance of the kink (gnn_oracle)."""
    res = []
    near = 0
    for dt in (torch.float32, torch.float64):
        L = O.leaf_params(P, dt)
        O.RELU_TRACE = [] if dt == torch.float64 else None
        try:
            loss, aux = loss_fn(L, b)
            if dt == torch.float64:
                global LAST_RELU_TRACE
                near = O.near_zero_preactivations(O.RELU_TRACE)
                LAST_RELU_TRACE = O.RELU_TRACE
        finally:
            O.RELU_TRACE = None
        loss.backward()
        res += [loss.detach(), {k: v.detach() for k, v in aux.items()}, {k: v.grad for k, v in L.items() if v.requires_grad}]
    return tuple(res) + (near,)


# ------------------------------------------------------------------------------------------------------------------
# the same bodies on the reference's own modules
# ------------------------------------------------------------------------------------------------------------------
class _RefStep:
    def parameters(self):
        return [p for m in self.modules for p in m.parameters()]

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    def load(self, P):
        """P: flat dict with the product's prefixes ('model.', 'head.', ...)."""
        for name, m in self.named.items():
            m.load_state_dict(sub(P, name + "."))


class ReferenceMaskingStep(_RefStep):
    def __init__(self, gnn_type="gin"):
        from . import reference_runner as R
        mod = R.load("chem")
        self.model = mod.GNN(5, 300, JK="last", drop_ratio=0, gnn_type=gnn_type).train()
        self.head = torch.nn.Linear(300, 119)
        self.named = {"model": self.model, "head": self.head}
        self.modules = list(self.named.values())

    def __call__(self, b):
        self.zero_grad()
        node_rep = self.model(b["x"], b["edge_index"], b["edge_attr"])
        pred_node = self.head(node_rep[b["masked_atom_indices"]])
        loss = F.cross_entropy(pred_node.double(), b["labels"])
        loss.backward()
        return loss


class ReferenceContextPredStep(_RefStep):
    def __init__(self):
        from . import reference_runner as R
        mod = R.load("chem")
        self.model_substruct = mod.GNN(5, 300, JK="last", drop_ratio=0, gnn_type="gin").train()
        self.model_context = mod.GNN(3, 300, JK="last", drop_ratio=0, gnn_type="gin").train()
        self.pool = mod.global_mean_pool
        self.named = {"model_substruct": self.model_substruct, "model_context": self.model_context}
        self.modules = list(self.named.values())

    def __call__(self, b):
        self.zero_grad()
        sub_rep = self.model_substruct(b["x_substruct"], b["edge_index_substruct"], b["edge_attr_substruct"])[b["center_substruct_idx"]]
        ov = self.model_context(b["x_context"], b["edge_index_context"], b["edge_attr_context"])[b["overlap_context_substruct_idx"]]
        ctx = self.pool(ov, b["batch_overlapped_context"])
        neg_ctx = ctx[O.cycle_rows(len(ctx), 1)]
        pos, neg = torch.sum(sub_rep * ctx, dim=1), torch.sum(sub_rep * neg_ctx, dim=1)
        loss = (F.binary_cross_entropy_with_logits(pos.double(), torch.ones(len(pos)).double())
                + F.binary_cross_entropy_with_logits(neg.double(), torch.zeros(len(neg)).double()))
        loss.backward()
        return loss


class ReferenceBioSupervisedStep(_RefStep):
    def __init__(self, num_tasks=5000, gnn_type="gin"):
        from . import reference_runner as R
        mod = R.load("bio")
        self.model = mod.GNN_graphpred(5, 300, num_tasks, JK="last", drop_ratio=0, graph_pooling="mean", gnn_type=gnn_type).train()
        self.named = {"model": self.model}
        self.modules = [self.model]

    def __call__(self, b):
        self.zero_grad()
        pred = self.model(types.SimpleNamespace(**b))
        y = b["go_target_pretrain"].view(pred.shape).to(torch.float64)
        loss = F.binary_cross_entropy_with_logits(pred.double(), y)
        loss.backward()
        return loss


REFERENCE_STEPS = {"masking": lambda: ReferenceMaskingStep("gin"), "contextpred": ReferenceContextPredStep,
                   "bio_supervised": ReferenceBioSupervisedStep, "gcn": lambda: ReferenceMaskingStep("gcn"),
                   "gat": lambda: ReferenceMaskingStep("gat"), "graphsage": lambda: ReferenceMaskingStep("graphsage")}


class PortStep:
    """The oracle-port flavour with the same calling convention as the Reference*Step classes (used when the reference's
    sources are not available on the box: cpu_baseline.kind = "port")."""

    def __init__(self, config, P):
        self.loss_fn = LOSSES[config]
        self.L = O.leaf_params(P)

    def parameters(self):
        return [v for v in self.L.values() if v.requires_grad]

    def __call__(self, b):
        for v in self.parameters():
            v.grad = None
        loss, _ = self.loss_fn(self.L, b)
        loss.backward()
        return loss
