import torch
from torch_scatter import scatter_add, scatter_max


def add_self_loops(edge_index, num_nodes=None):
    n = int(edge_index.max().item()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loop, loop])], dim=1)


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max().item()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype, device=index.device)
    return out.index_add_(0, index, torch.ones(index.shape[0], dtype=out.dtype, device=index.device))


def softmax(src, index, num_nodes=None):
    n = int(index.max().item()) + 1 if num_nodes is None else num_nodes
    seg_max, _ = scatter_max(src, index, dim=0, dim_size=n, fill_value=0)
    ex = (src - seg_max[index]).exp()
    return ex / (scatter_add(ex, index, dim=0, dim_size=n)[index] + 1e-16)
