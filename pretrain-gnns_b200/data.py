"""Device-resident molecule store and on-device batch collation, and the staging of
host-collated batches (one pinned buffer, one asynchronous copy per batch) for callers that keep the reference's DataLoader.

The reference collates on the host, one Python loop of `torch.cat`s per batch (chem/batch.py:17-52 BatchMasking,
:141-210 BatchSubstructContext) inside DataLoader workers, then copies the batch to the GPU.  A B200 has 180 GB of
HBM: the whole pre-training set (ZINC15, 2M molecules x ~23 atoms: < 1 GB in the compact form below) stays resident and a
batch is one `pgnn_collate_chem` call on a list of graph ids -- no host work, no H2D copy per step.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from ._cabi import check, lib


class MoleculeStore:
    """Compact CSR-of-graphs for chem molecules (value ranges: chem/loader.py:22-51; all fit a byte).

    node_ptr / edge_ptr [G+1] int64; x [Nt,2] uint8; edge_index [2,Et] int32 graph-local; edge_attr [Et,2] uint8."""

    def __init__(self, node_ptr, edge_ptr, x, edge_index, edge_attr, device="cuda"):
        self.node_ptr_host = np.ascontiguousarray(node_ptr, dtype=np.int64)
        self.edge_ptr_host = np.ascontiguousarray(edge_ptr, dtype=np.int64)
        G = len(self.node_ptr_host) - 1
        if G < 0 or len(self.edge_ptr_host) != G + 1:
            raise ValueError("node_ptr and edge_ptr must both have num_graphs + 1 entries")
        Nt, Et = int(self.node_ptr_host[-1]), int(self.edge_ptr_host[-1])
        x, edge_index, edge_attr = np.asarray(x), np.asarray(edge_index), np.asarray(edge_attr)
        if x.shape != (Nt, 2) or edge_index.shape != (2, Et) or edge_attr.shape != (Et, 2):
            raise ValueError("store arrays do not match the prefix sums")
        if Nt and (x.min() < 0 or x.max() > 255) or Et and (edge_attr.min() < 0 or edge_attr.max() > 255):
            raise ValueError("atom / bond features must fit a byte")
        self.num_graphs, self.num_nodes, self.num_edges = G, Nt, Et
        self.device = torch.device(device)
        dev = self.device
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(dev)
        self.edge_ptr = torch.from_numpy(self.edge_ptr_host).to(dev)
        self.x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint8)).to(dev)
        self.edge_index = torch.from_numpy(np.ascontiguousarray(edge_index, dtype=np.int32)).to(dev)
        self.edge_attr = torch.from_numpy(np.ascontiguousarray(edge_attr, dtype=np.uint8)).to(dev)

    @classmethod
    def from_data_list(cls, data_list, device="cuda"):
        """data_list: objects with .x [n,2], .edge_index [2,e] (graph-local), .edge_attr [e,2] (the reference's Data)."""
        n = np.array([0] + [int(d.x.shape[0]) for d in data_list], dtype=np.int64)
        e = np.array([0] + [int(d.edge_index.shape[1]) for d in data_list], dtype=np.int64)
        cat = lambda xs, ax, shape: np.concatenate([np.asarray(a) for a in xs], axis=ax) if xs else np.zeros(shape, np.int64)
        return cls(np.cumsum(n), np.cumsum(e), cat([d.x for d in data_list], 0, (0, 2)),
                   cat([d.edge_index for d in data_list], 1, (2, 0)), cat([d.edge_attr for d in data_list], 0, (0, 2)), device)

    def batch_sizes(self, graph_ids_host):
        """(N, E) of the batch, from the host copy of the prefix sums (no device sync)."""
        ids = np.asarray(graph_ids_host, dtype=np.int64)
        if ids.size and (ids.min() < 0 or ids.max() >= self.num_graphs):
            raise IndexError("graph id out of range")
        return (int((self.node_ptr_host[ids + 1] - self.node_ptr_host[ids]).sum()),
                int((self.edge_ptr_host[ids + 1] - self.edge_ptr_host[ids]).sum()))

    def collate(self, graph_ids_host, graph_ids_dev=None):
        """-> namespace(x [N,2], edge_index [2,E], edge_attr [E,2], batch [N], node_off [B+1], edge_off [B+1], num_graphs),
        int64 CUDA tensors exactly as BatchMasking.from_data_list([dataset[i] for i in ids]) would hold them."""
        ids = np.ascontiguousarray(graph_ids_host, dtype=np.int64)
        N, E = self.batch_sizes(ids)
        B, dev = len(ids), self.device
        if graph_ids_dev is None:
            graph_ids_dev = torch.from_numpy(ids).to(dev, non_blocking=True)
        i64 = dict(dtype=torch.int64, device=dev)
        out = SimpleNamespace(x=torch.empty((N, 2), **i64), edge_index=torch.empty((2, E), **i64), edge_attr=torch.empty((E, 2), **i64),
                              batch=torch.empty((N,), **i64), node_off=torch.empty((B + 1,), **i64),
                              edge_off=torch.empty((B + 1,), **i64), num_graphs=B)
        check(lib.pgnn_collate_chem(self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(), self.x.data_ptr(), self.edge_index.data_ptr(),
                                    self.num_edges, self.edge_attr.data_ptr(), graph_ids_dev.data_ptr(), B, out.node_off.data_ptr(),
                                    out.edge_off.data_ptr(), out.x.data_ptr(), out.edge_index.data_ptr(), out.edge_attr.data_ptr(),
                                    out.batch.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "pgnn_collate_chem")
        return out


class PackedHostBatch:
    """A batch dict packed into ONE pinned host buffer (fields 16-byte aligned), built once per batch by BatchStager.pack."""

    def __init__(self, buf, fields):
        self.buf, self.fields = buf, fields  # fields: name -> (byte offset, byte length, dtype, shape)
        self.nbytes = int(buf.numel())


class BatchStager:
    """Host -> device staging of batches with the copy off the critical path.

    The reference moves a batch with one `.to(device)` per tensor right before the step (chem/pretrain_masking.py:47,
    `batch = batch.to(device)`), i.e. five small synchronous-looking copies in front of the first kernel.  Here a batch is ONE
    pinned buffer and ONE `cudaMemcpyAsync` on a side stream into one of `slots` device buffers, so the copy of batch i+1 runs
    under the kernels of batch i:

        t = stager.submit(packed[0])
        for i in range(steps):
            b = stager.take(t)                       # dict of device views; the compute stream waits for the copy
            loss = train_step(b)                      # enqueue the step
            t = stager.submit(packed[i + 1])         # prefetch under the step just enqueued
            log(loss.item())

    A slot is overwritten only after the work enqueued on the batch that last occupied it (everything up to the following
    `take`) has finished: the side stream waits on an event of the compute stream, so no host synchronisation is assumed."""

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.slots = int(slots)
        self.dev = [None] * self.slots          # device buffers, grown on demand
        self.ready = [None] * self.slots        # copy finished (side stream)
        self.release = [None] * self.slots      # consumers enqueued (compute stream)
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._next, self._last_taken = 0, None

    def pack(self, batch):
        fields, off = {}, 0
        for k, v in batch.items():
            if not torch.is_tensor(v):
                continue
            v = v.contiguous()
            nb = v.numel() * v.element_size()
            fields[k] = (off, nb, v.dtype, tuple(v.shape), v)
            off = (off + nb + 15) // 16 * 16
        buf = torch.empty(max(off, 16), dtype=torch.uint8)
        if self.cuda:
            buf = buf.pin_memory()
        out = {}
        for k, (o, nb, dt, shape, v) in fields.items():
            if nb:
                buf[o:o + nb].copy_(v.view(-1).view(torch.uint8))
            out[k] = (o, nb, dt, shape)
        return PackedHostBatch(buf, out)

    def submit(self, packed):
        """Start copying `packed` into the next slot; returns a ticket for take()."""
        s = self._next
        self._next = (s + 1) % self.slots
        if self.dev[s] is None or self.dev[s].numel() < packed.nbytes:
            self.dev[s] = torch.empty(max(packed.nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
            if self.cuda:  # the allocator may hand back a block whose previous user is still running on the compute stream
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.cuda:
            if s == self._last_taken:  # the slot's batch is still the one in use: everything enqueued so far must finish first
                self.release[s] = torch.cuda.Event()
                self.release[s].record(torch.cuda.current_stream(self.device))
                self._last_taken = None
            if self.release[s] is not None:
                self.stream.wait_event(self.release[s])
            with torch.cuda.stream(self.stream):
                self.dev[s][:packed.nbytes].copy_(packed.buf, non_blocking=True)
                self.ready[s] = torch.cuda.Event()
                self.ready[s].record(self.stream)
        else:
            self.dev[s][:packed.nbytes].copy_(packed.buf)
        return (s, packed)

    def take(self, ticket):
        """-> dict of device tensors (views of the slot).  Everything enqueued on the previous batch is now 'released'."""
        s, packed = ticket
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            if self._last_taken is not None and self._last_taken != s:
                self.release[self._last_taken] = torch.cuda.Event()
                self.release[self._last_taken].record(cur)
            cur.wait_event(self.ready[s])
        self._last_taken = s
        d = self.dev[s]
        return {k: (d[o:o + nb].view(dt).view(shape) if nb else torch.empty(shape, dtype=dt, device=self.device))
                for k, (o, nb, dt, shape) in packed.fields.items()}
