"""Stand-in for torch_scatter==1.1.2 (test infrastructure; see ../README.md)."""
import torch


def _out_shape(src, dim, dim_size, index):
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    return shape


def scatter_add(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    dim = dim % src.dim()
    if out is None:
        out = src.new_full(_out_shape(src, dim, dim_size, index), fill_value)
    return out.index_add_(dim, index, src) if index.dim() == 1 else out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    total = scatter_add(src, index, dim, out, dim_size, fill_value)
    ones = torch.ones(index.shape[0], dtype=src.dtype, device=src.device)
    cnt = scatter_add(ones, index, 0, None, total.shape[dim % src.dim()]).clamp(min=1)
    view = [1] * total.dim()
    view[dim % src.dim()] = -1
    return total / cnt.view(view)


def scatter_max(src, index, dim=0, out=None, dim_size=None, fill_value=0):
    """1.1.2: the output is created with `fill_value` (default 0, not -inf) and the maximum is taken INTO it, so the
    result is max(fill_value, segment max) — a segment whose entries are all negative reports 0."""
    dim = dim % src.dim()
    assert dim == 0
    shape = _out_shape(src, dim, dim_size, index)
    res = src.new_full(shape, float("-inf") if fill_value is None else fill_value)
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    res = res.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    return res, None
