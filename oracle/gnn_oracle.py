"""CPU oracle for the GNN message-passing hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A functional, plain-torch (CPU, fp32 or fp64) restatement of the reference's algorithm, written
against `state_dict`-keyed parameter dictionaries so that the shipped `.pth` files and the product's
modules can both feed it.  Each function cites the reference lines it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this;
nothing under `pretrain-gnns_b200/` does.

Pinning status: the reference ships NO golden vectors for this path.  The
oracle is therefore pinned against the reference's own `chem/model.py` / `bio/model.py` executed here,
unmodified, over `oracle/pyg103_standin` (a stand-in for the absent torch_geometric 1.0.3 /
torch_scatter 1.1.2 whose semantics are recalled, not verifiable): `tests/golden/make_golden.py`
generates the fixtures and `tests/test_oracle_vs_reference.py` re-runs the comparison whenever
`/root/reference` is present.  The part of the pin that rests on the stand-in (direction convention,
softmax epsilon) is "parity unpinned" in the strict sense and is flagged in DESIGN.md.

Conventions (SURVEY.md §8(c) [M]): `edge_index[0]` is the aggregation target, `edge_index[1]` the
source; self-loops are appended after the given edges; sums run in edge order (CPU `index_add_`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # torch.nn.BatchNorm1d default, chem/model.py:252
BN_MOMENTUM = 0.1
SELF_LOOP_BOND = 4  # chem/model.py:43
BIO_SELF_LOOP_COL = 7  # bio/model.py:43
GAT_SLOPE = 0.2     # chem/model.py:108
SOFTMAX_EPS = 1e-16  # torch_geometric 1.0.3 utils.softmax [M]


# --------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------
RELU_TRACE = None   # tests set this to a list to collect every ReLU's pre-activation (see near_zero_preactivations)


def _relu(x):
    if RELU_TRACE is not None:
        RELU_TRACE.append(x.detach())
    return torch.relu(x)


def near_zero_preactivations(trace, rel=4e-6):
    """How many ReLU inputs of a traced run lie within rounding distance (rel x the tensor's largest magnitude) of the
    kink.  An equally valid fp32 summation order can put such a unit on the other side, which changes its gradient
    mask: the parity tests grant their ReLU-boundary allowance only when this count is non-zero."""
    return int(sum(int((t.abs() <= rel * t.abs().max()).sum()) for t in trace if t.numel()))


def with_self_loops(edge_index: torch.Tensor, n: int) -> torch.Tensor:
    """chem/model.py:39 — N loops appended after the real edges."""
    loop = torch.arange(n, dtype=edge_index.dtype)
    return torch.cat([edge_index, torch.stack([loop, loop])], dim=1)


def chem_edge_rows(P, pre, edge_attr, n):
    """chem/model.py:42-47 — self-loop attr rows [4,0] appended, two-table lookup, add."""
    loops = torch.zeros(n, 2, dtype=edge_attr.dtype)
    loops[:, 0] = SELF_LOOP_BOND
    ea = torch.cat([edge_attr, loops], dim=0)
    return F.embedding(ea[:, 0], P[pre + "edge_embedding1.weight"]) + F.embedding(ea[:, 1], P[pre + "edge_embedding2.weight"])


def bio_edge_rows(P, pre, edge_attr, n):
    """bio/model.py:42-47 — self-loop rows one-hot at col 7, Linear(9, D or heads*D)."""
    loops = torch.zeros(n, 9, dtype=edge_attr.dtype)
    loops[:, BIO_SELF_LOOP_COL] = 1
    ea = torch.cat([edge_attr, loops], dim=0)
    return F.linear(ea, P[pre + "edge_encoder.weight"], P[pre + "edge_encoder.bias"])


def reduce_onto_target(msg, target, n, mean=False):
    """torch_geometric 1.0.3 scatter_('add'|'mean') onto edge_index[0] [M]."""
    out = torch.zeros((n,) + tuple(msg.shape[1:]), dtype=msg.dtype).index_add_(0, target, msg)
    if mean:
        cnt = torch.zeros(n, dtype=msg.dtype).index_add_(0, target, torch.ones(len(target), dtype=msg.dtype))
        out = out / cnt.clamp(min=1).view([-1] + [1] * (msg.dim() - 1))
    return out


def gcn_norm(ei, n, dtype):
    """chem/model.py:73-82 — degree over edge_index[0] of the loop-augmented edges."""
    deg = torch.zeros(n, dtype=dtype).index_add_(0, ei[0], torch.ones(ei.shape[1], dtype=dtype))
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return dis[ei[0]] * dis[ei[1]]


def segment_softmax(alpha, target, n):
    """torch_geometric 1.0.3 utils.softmax [M]: shifted by torch_scatter 1.1.2's scatter_max, whose output starts from
    its default fill_value = 0 (so the shift is max(0, segment max)), +1e-16 in the denominator."""
    idx = target.view(-1, 1).expand_as(alpha)
    mx = torch.zeros((n, alpha.shape[1]), dtype=alpha.dtype)
    mx = mx.scatter_reduce(0, idx, alpha.detach(), reduce="amax", include_self=True)
    ex = (alpha - mx[target]).exp()
    den = torch.zeros(n, alpha.shape[1], dtype=alpha.dtype).index_add_(0, target, ex)
    return ex / (den[target] + SOFTMAX_EPS)


def batch_norm(P, pre, h, training, new_stats=None):
    """torch.nn.BatchNorm1d (chem/model.py:269, bio/model.py:24) through the same functional the module
    calls.  Train: batch mean / biased var; running stats (momentum 0.1, unbiased var) are returned in
    `new_stats` instead of being updated in place, so the parameter dictionary stays immutable."""
    w, b = P[pre + "weight"], P[pre + "bias"]
    rm, rv = P[pre + "running_mean"].detach(), P[pre + "running_var"].detach()
    if not training:
        return F.batch_norm(h, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
    rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(h, rm, rv, w, b, True, BN_MOMENTUM, BN_EPS)
    if new_stats is not None:
        new_stats[pre + "running_mean"], new_stats[pre + "running_var"] = rm, rv
        if pre + "num_batches_tracked" in P:
            new_stats[pre + "num_batches_tracked"] = P[pre + "num_batches_tracked"] + 1
    return y


# --------------------------------------------------------------------------------------------
# conv layers.  `edge_rows` is [E', C] (already loop-augmented), `ei` is [2, E'].
# --------------------------------------------------------------------------------------------
def gin_conv_chem(P, pre, h, ei, edge_rows):
    """chem/model.py:37-55: aggr = sum(x_j + e); out = W2 relu(W1 aggr + b1) + b2."""
    aggr = reduce_onto_target(h.index_select(0, ei[1]) + edge_rows, ei[0], h.shape[0])
    z = _relu(F.linear(aggr, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"]))
    return F.linear(z, P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"])


def gin_conv_bio(P, pre, h, ei, edge_rows, training, new_stats=None):
    """bio/model.py:37-58: message = cat([x_j, e]); MLP = Linear(2D,2D) BN ReLU Linear(2D,D)."""
    aggr = reduce_onto_target(torch.cat([h.index_select(0, ei[1]), edge_rows], dim=1), ei[0], h.shape[0])
    z = F.linear(aggr, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"])
    z = _relu(batch_norm(P, pre + "mlp.1.", z, training, new_stats))
    return F.linear(z, P[pre + "mlp.3.weight"], P[pre + "mlp.3.bias"])


def gcn_conv(P, pre, h, ei, edge_rows):
    """chem/model.py:85-104, bio/model.py:92-114: Linear first, then sum(norm * (x_j + e))."""
    norm = gcn_norm(ei, h.shape[0], h.dtype)
    x = F.linear(h, P[pre + "linear.weight"], P[pre + "linear.bias"])
    return reduce_onto_target(norm.view(-1, 1) * (x.index_select(0, ei[1]) + edge_rows), ei[0], h.shape[0])


def sage_conv(P, pre, h, ei, edge_rows):
    """chem/model.py:182-202, bio/model.py:201-224: Linear, mean(x_j + e), L2-normalise rows."""
    x = F.linear(h, P[pre + "linear.weight"], P[pre + "linear.bias"])
    aggr = reduce_onto_target(x.index_select(0, ei[1]) + edge_rows, ei[0], h.shape[0], mean=True)
    return F.normalize(aggr, p=2, dim=-1)


def gat_conv(P, pre, h, ei, edge_rows, heads=2):
    """chem/model.py:134-165, bio/model.py:147-180: x_j += e; alpha = softmax_by_target(
    leaky_relu(<[x_i, x_j], att>)); out = mean_heads(sum x_j * alpha) + bias."""
    n, d = h.shape[0], P[pre + "bias"].shape[0]
    x = F.linear(h, P[pre + "weight_linear.weight"], P[pre + "weight_linear.bias"]).view(n, heads, d)
    xj = x.index_select(0, ei[1]) + edge_rows.view(-1, heads, d)
    xi = x.index_select(0, ei[0])
    alpha = (torch.cat([xi, xj], dim=-1) * P[pre + "att"]).sum(-1)
    alpha = segment_softmax(F.leaky_relu(alpha, GAT_SLOPE), ei[0], n)
    out = reduce_onto_target(xj * alpha.view(-1, heads, 1), ei[0], n)
    return out.mean(dim=1) + P[pre + "bias"]


# --------------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------------
def chem_gnn(P, x, edge_index, edge_attr, num_layer, gnn_type="gin", training=False, new_stats=None,
             pre="", keep=None):
    """chem/model.py:255-290 with JK='last', drop_ratio=0."""
    n = x.shape[0]
    h = F.embedding(x[:, 0], P[pre + "x_embedding1.weight"]) + F.embedding(x[:, 1], P[pre + "x_embedding2.weight"])
    ei = with_self_loops(edge_index, n)
    for l in range(num_layer):
        lp = f"{pre}gnns.{l}."
        rows = chem_edge_rows(P, lp, edge_attr, n)
        if gnn_type == "gin":
            h = gin_conv_chem(P, lp, h, ei, rows)
        elif gnn_type == "gcn":
            h = gcn_conv(P, lp, h, ei, rows)
        elif gnn_type == "graphsage":
            h = sage_conv(P, lp, h, ei, rows)
        elif gnn_type == "gat":
            h = gat_conv(P, lp, h, ei, rows)
        else:
            raise ValueError(gnn_type)
        if keep is not None:
            keep.append(h)
        h = batch_norm(P, f"{pre}batch_norms.{l}.", h, training, new_stats)
        if l != num_layer - 1:
            h = _relu(h)
    return h


def bio_gnn(P, x, edge_index, edge_attr, num_layer, gnn_type="gin", training=False, new_stats=None, pre=""):
    """bio/model.py:273-290 with JK='last', drop_ratio=0 (no outer BN)."""
    n = x.shape[0]
    ei = with_self_loops(edge_index, n)
    h = x
    for l in range(num_layer):
        lp = f"{pre}gnns.{l}."
        rows = bio_edge_rows(P, lp, edge_attr, n)
        if l == 0:
            h = F.embedding(h.to(torch.int64).view(-1), P[lp + "input_node_embeddings.weight"])  # bio/model.py:49-50
        if gnn_type == "gin":
            h = gin_conv_bio(P, lp, h, ei, rows, training, new_stats)
        elif gnn_type == "gcn":
            h = gcn_conv(P, lp, h, ei, rows)
        elif gnn_type == "graphsage":
            h = sage_conv(P, lp, h, ei, rows)
        elif gnn_type == "gat":
            h = gat_conv(P, lp, h, ei, rows)
        else:
            raise ValueError(gnn_type)
        if l != num_layer - 1:
            h = _relu(h)
    return h


# --------------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------------
def segment_mean(x, seg, num_seg):
    """global_mean_pool (chem/model.py:326,369) = scatter_mean: sum / count.clamp(min=1) [M]."""
    return reduce_onto_target(x, seg, num_seg, mean=True)


def chem_graphpred(P, x, edge_index, edge_attr, batch, num_graphs, num_layer, gnn_type="gin", training=False,
                   new_stats=None):
    """chem/model.py:358-369 with graph_pooling='mean'."""
    h = chem_gnn(P, x, edge_index, edge_attr, num_layer, gnn_type, training, new_stats, pre="gnn.")
    return F.linear(segment_mean(h, batch, num_graphs), P["graph_pred_linear.weight"], P["graph_pred_linear.bias"])


def bio_graphpred(P, x, edge_index, edge_attr, batch, center_node_idx, num_graphs, num_layer, gnn_type="gin",
                  training=False, new_stats=None):
    """bio/model.py:338-347: Linear(2D,T) on cat([mean_pool, node_rep[center]])."""
    h = bio_gnn(P, x, edge_index, edge_attr, num_layer, gnn_type, training, new_stats, pre="gnn.")
    rep = torch.cat([segment_mean(h, batch, num_graphs), h[center_node_idx]], dim=1)
    return F.linear(rep, P["graph_pred_linear.weight"], P["graph_pred_linear.bias"])


def masking_loss(node_rep, masked_atom_indices, labels, w, b):
    """chem/pretrain_masking.py:51-52: Linear(300,119) on gathered rows, CE on .double() logits."""
    logits = F.linear(node_rep[masked_atom_indices], w, b)
    return F.cross_entropy(logits.double(), labels), logits


def masking_edge_loss(node_rep, edge_index, connected_edge_indices, labels, w, b):
    """chem/pretrain_masking.py:58-61: rep[u]+rep[v] for the masked bonds, Linear(300,4), CE."""
    me = edge_index[:, connected_edge_indices]
    logits = F.linear(node_rep[me[0]] + node_rep[me[1]], w, b)
    return F.cross_entropy(logits.double(), labels), logits


def cycle_rows(num, shift):
    """chem/pretrain_contextpred.py:36-39: row r -> (r + shift) mod num."""
    return (torch.arange(num) + shift) % num


def contextpred_scores(substruct_rep, overlapped_rep, batch_overlapped, num_graphs, neg_samples=1):
    """chem/pretrain_contextpred.py:60-67 (cbow, mean pooling)."""
    ctx = segment_mean(overlapped_rep, batch_overlapped, num_graphs)
    neg = torch.cat([ctx[cycle_rows(num_graphs, i + 1)] for i in range(neg_samples)], dim=0)
    pos = (substruct_rep * ctx).sum(1)
    negs = (substruct_rep.repeat(neg_samples, 1) * neg).sum(1)
    return pos, negs


def contextpred_loss(pos, neg, neg_samples=1):
    """chem/pretrain_contextpred.py:86-93: BCE-with-logits in fp64, pos + neg_samples*neg."""
    lp = F.binary_cross_entropy_with_logits(pos.double(), torch.ones_like(pos, dtype=torch.float64))
    ln = F.binary_cross_entropy_with_logits(neg.double(), torch.zeros_like(neg, dtype=torch.float64))
    return lp + neg_samples * ln


# --------------------------------------------------------------------------------------------
# parameter construction (shapes per SURVEY.md §8(b); init mirrors the reference's)
# --------------------------------------------------------------------------------------------
def _lin(g, out_f, in_f, dtype):
    k = 1.0 / in_f ** 0.5
    return ((torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype), \
           ((torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype)


def _xavier(g, rows, cols, dtype):
    a = (6.0 / (rows + cols)) ** 0.5
    return ((torch.rand(rows, cols, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)


def _bn(P, pre, d, g, dtype, randomize):
    if randomize:  # non-trivial affine + running stats so eval-mode tests exercise them
        P[pre + "weight"] = (0.5 + torch.rand(d, generator=g, dtype=torch.float64)).to(dtype)
        P[pre + "bias"] = (torch.rand(d, generator=g, dtype=torch.float64) - 0.5).to(dtype)
        P[pre + "running_mean"] = (torch.rand(d, generator=g, dtype=torch.float64) - 0.5).to(dtype)
        P[pre + "running_var"] = (0.5 + torch.rand(d, generator=g, dtype=torch.float64)).to(dtype)
    else:
        P[pre + "weight"], P[pre + "bias"] = torch.ones(d, dtype=dtype), torch.zeros(d, dtype=dtype)
        P[pre + "running_mean"], P[pre + "running_var"] = torch.zeros(d, dtype=dtype), torch.ones(d, dtype=dtype)
    P[pre + "num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def make_params(domain: str, gnn_type: str, num_layer: int, emb_dim: int, seed: int, dtype=torch.float32,
                randomize_bn: bool = True, heads: int = 2):
    """Seeded state_dict-shaped parameters for `chem` or `bio` GNN (keys as in the shipped .pth)."""
    g = torch.Generator().manual_seed(seed)
    P, D = {}, emb_dim
    if domain == "chem":
        P["x_embedding1.weight"] = _xavier(g, 120, D, dtype)
        P["x_embedding2.weight"] = _xavier(g, 3, D, dtype)
    for l in range(num_layer):
        pre = f"gnns.{l}."
        c = heads * D if gnn_type == "gat" else D
        if gnn_type == "gin" and domain == "chem":
            P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"] = _lin(g, 2 * D, D, dtype)
            P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"] = _lin(g, D, 2 * D, dtype)
        elif gnn_type == "gin":
            P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"] = _lin(g, 2 * D, 2 * D, dtype)
            _bn(P, pre + "mlp.1.", 2 * D, g, dtype, randomize_bn)
            P[pre + "mlp.3.weight"], P[pre + "mlp.3.bias"] = _lin(g, D, 2 * D, dtype)
        elif gnn_type in ("gcn", "graphsage"):
            P[pre + "linear.weight"], P[pre + "linear.bias"] = _lin(g, D, D, dtype)
        elif gnn_type == "gat":
            P[pre + "weight_linear.weight"], P[pre + "weight_linear.bias"] = _lin(g, heads * D, D, dtype)
            P[pre + "att"] = _xavier(g, heads, 2 * D, dtype).view(1, heads, 2 * D)
            P[pre + "bias"] = ((torch.rand(D, generator=g, dtype=torch.float64) - 0.5) * 0.2).to(dtype)
        else:
            raise ValueError(gnn_type)
        if domain == "chem":
            P[pre + "edge_embedding1.weight"] = _xavier(g, 6, c, dtype)
            P[pre + "edge_embedding2.weight"] = _xavier(g, 3, c, dtype)
        else:
            P[pre + "edge_encoder.weight"], P[pre + "edge_encoder.bias"] = _lin(g, c, 9, dtype)
            if l == 0:
                P[pre + "input_node_embeddings.weight"] = _xavier(g, 2, D, dtype)
    if domain == "chem":
        for l in range(num_layer):
            _bn(P, f"batch_norms.{l}.", D, g, dtype, randomize_bn)
    return P


def is_float_param(k, v):
    return v.is_floating_point() and not k.endswith("running_mean") and not k.endswith("running_var")


def leaf_params(P, dtype=None):
    """Clone into autograd leaves (trainable entries only)."""
    out = {}
    for k, v in P.items():
        v = v.detach().clone()
        if dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        if is_float_param(k, v):
            v.requires_grad_(True)
        out[k] = v
    return out
