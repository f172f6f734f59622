"""Integer oracle for graph preparation — TEST INFRASTRUCTURE, NOT PRODUCT.

The reference never builds a CSR: PyG 1.0.3 gathers `x[edge_index[1]]` and scatter-adds onto
`edge_index[0]` edge by edge (chem/model.py:49 via MessagePassing.propagate [M]).  The product buckets
the same COO list by target (forward) and by source (backward) so that each output row is reduced by
one thread group in the ORIGINAL edge order; this file states that bucketing in numpy so the CUDA
prep can be checked bit-for-bit.  Self-loops stay implicit (they are appended last by
chem/model.py:39, so "after all real edges" is their position in every bucket).
"""
import numpy as np


def bucket(keys: np.ndarray, vals: np.ndarray, n: int):
    """Stable counting sort of edges by `keys`: returns rowptr[n+1], vals in bucket order, edge ids."""
    keys = np.asarray(keys, dtype=np.int64)
    order = np.argsort(keys, kind="stable").astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(keys, minlength=n), out=rowptr[1:])
    return rowptr, np.asarray(vals)[order].astype(np.int32), order


def graph_prep(edge_index: np.ndarray, n: int):
    """COO int64 [2,E] -> (by-target CSR, by-source CSR), each (rowptr, neighbour, edge_id) int32."""
    tgt, src = edge_index[0], edge_index[1]
    return bucket(tgt, src, n), bucket(src, tgt, n)


def chem_attr_codes(edge_attr: np.ndarray, order: np.ndarray) -> np.ndarray:
    """Pack [E,2] int64 bond (type, direction) into one byte `type | direction << 4`, bucket order."""
    ea = np.asarray(edge_attr)[order]
    return (ea[:, 0] | (ea[:, 1] << 4)).astype(np.uint8)


def segments(seg: np.ndarray, num_seg: int):
    """Node -> graph assignment (`batch`) as a CSR over graphs (stable), for global_mean_pool."""
    return bucket(seg, np.arange(len(seg)), num_seg)
