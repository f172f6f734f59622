import inspect
import torch
from torch_scatter import scatter_add, scatter_mean, scatter_max
from . import inits  # noqa: F401


def _reduce(kind, src, index, dim_size):
    if kind == "add":
        return scatter_add(src, index, 0, None, dim_size)
    if kind == "mean":
        return scatter_mean(src, index, 0, None, dim_size)
    if kind == "max":
        out, _ = scatter_max(src, index, 0, None, dim_size, fill_value=0)
        return out
    raise ValueError(kind)


class MessagePassing(torch.nn.Module):
    """1.0.x flow: edge_index[0] is the reduction target, edge_index[1] the source."""

    def __init__(self):
        super().__init__()
        self._msg_names = list(inspect.signature(self.message).parameters)
        self._upd_names = list(inspect.signature(self.update).parameters)[1:]

    def propagate(self, aggr, edge_index, **kwargs):
        kwargs["edge_index"] = edge_index
        n = None
        margs = []
        for name in self._msg_names:
            if name.endswith("_i"):
                t = kwargs[name[:-2]]
                n = t.size(0)
                margs.append(t[edge_index[0]])
            elif name.endswith("_j"):
                t = kwargs[name[:-2]]
                n = t.size(0)
                margs.append(t[edge_index[1]])
            else:
                margs.append(kwargs[name])
        msg = self.message(*margs)
        out = _reduce(aggr, msg, edge_index[0], n)
        return self.update(out, *[kwargs[k] for k in self._upd_names])

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


def _nb(batch, size):
    return int(batch.max().item()) + 1 if size is None else size


def global_add_pool(x, batch, size=None):
    return _reduce("add", x, batch, _nb(batch, size))


def global_mean_pool(x, batch, size=None):
    return _reduce("mean", x, batch, _nb(batch, size))


def global_max_pool(x, batch, size=None):
    return _reduce("max", x, batch, _nb(batch, size))


class GlobalAttention(torch.nn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn

    def forward(self, x, batch, size=None):
        raise NotImplementedError("out of scope (SURVEY.md §2 row 3)")


class Set2Set(torch.nn.Module):
    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()

    def forward(self, x, batch):
        raise NotImplementedError("out of scope (SURVEY.md §2 row 3)")
