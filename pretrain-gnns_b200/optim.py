"""Single-launch Adam over every parameter tensor.

Mirror of what the reference's scripts construct: `optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.decay)`
(chem/pretrain_masking.py:134-136; one optimizer per module, stepped at :72-74).  Same constructor arguments,
`step()`, `zero_grad()`, `param_groups`, and a `state_dict()` in torch.optim.Adam's layout, so checkpoints move
either way.  Underneath: `pgnn_adam_step` (csrc/step_io.cu), one CTA per chunk of one tensor, all tensors in one
launch; exp_avg / exp_avg_sq live in two flat buffers.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from ._cabi import check, lib

_CHUNK = 4096  # elements per CTA
_CHUNK_DTYPE = np.dtype([("param", "<u8"), ("grad", "<u8"), ("exp_avg", "<u8"), ("exp_avg_sq", "<u8"), ("n", "<i8")])


def chunk_table(ptrs, numels, chunk=_CHUNK):
    """Host-side chunk table for pgnn_adam_step: rows (param, grad, exp_avg, exp_avg_sq, n) of at most `chunk` fp32
    elements each; `ptrs` = per tensor (param_ptr, grad_ptr, m_ptr, v_ptr), `numels` = its element count."""
    rows = []
    for (p, g, m, v), n in zip(ptrs, numels):
        for off in range(0, n, chunk):
            b = 4 * off
            rows.append((p + b, g + b, m + b, v + b, min(chunk, n - off)))
    return np.array(rows, dtype=_CHUNK_DTYPE)


class Adam:
    """torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) with one kernel launch per step.

    `grad_scale` multiplies every gradient as it is read (1/world_size turns an all-reduced SUM into the mean for free).
    `legacy_eps=True` reproduces torch 1.0.1's placement of eps (the version the reference pins, requirements.txt:2)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0, legacy_eps=False):
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("Invalid Adam hyper-parameter")
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("pretrain_gnns_b200.optim.Adam needs contiguous fp32 CUDA parameters")
        self.param_groups = [dict(params=params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.grad_scale, self.legacy_eps = float(grad_scale), bool(legacy_eps)
        self._step = 0
        dev = params[0].device
        sizes = [p.numel() for p in params]
        # each tensor's slice starts on a 16-byte boundary so the kernel's float4 path applies
        starts, tot = [], 0
        for n in sizes:
            starts.append(tot)
            tot += (n + 3) // 4 * 4
        self._m = torch.zeros(tot, dtype=torch.float32, device=dev)
        self._v = torch.zeros(tot, dtype=torch.float32, device=dev)
        self._slices = [(s, n) for s, n in zip(starts, sizes)]
        self._tables = {}

    # ---- torch.optim.Optimizer surface -----------------------------------------------------------------------------
    @property
    def state(self):
        out = {}
        for p, (s, n) in zip(self.param_groups[0]["params"], self._slices):
            out[p] = dict(step=torch.tensor(float(self._step)), exp_avg=self._m[s:s + n].view_as(p),
                          exp_avg_sq=self._v[s:s + n].view_as(p))
        return out

    def zero_grad(self, set_to_none=True):
        for p in self.param_groups[0]["params"]:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    def state_dict(self):
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self._slices)))
        st = {}
        if self._step:
            for i, (s, n) in enumerate(self._slices):
                p = self.param_groups[0]["params"][i]
                st[i] = dict(step=torch.tensor(float(self._step)), exp_avg=self._m[s:s + n].view_as(p).clone(),
                             exp_avg_sq=self._v[s:s + n].view_as(p).clone())
        return dict(state=st, param_groups=[g])

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g:
                self.param_groups[0][k] = tuple(g[k]) if k == "betas" else g[k]
        steps = set()
        for i, st in sd["state"].items():
            s, n = self._slices[int(i)]
            self._m[s:s + n].copy_(st["exp_avg"].reshape(-1))
            self._v[s:s + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; this optimizer keeps one")
        self._step = steps.pop() if steps else 0

    # ---- the step ---------------------------------------------------------------------------------------------------
    def _table(self, live):
        key = tuple((i, p.data_ptr(), p.grad.data_ptr()) for i, p in live)
        tab = self._tables.get(key)
        if tab is None:
            if len(self._tables) > 16:
                self._tables.clear()
            mb, vb = self._m.data_ptr(), self._v.data_ptr()
            host = chunk_table([(p.data_ptr(), p.grad.data_ptr(), mb + 4 * self._slices[i][0], vb + 4 * self._slices[i][0])
                                for i, p in live], [p.numel() for _, p in live])
            dev = torch.from_numpy(host.view(np.uint8).copy()).to(self._m.device)
            tab = (dev, len(host))
            self._tables[key] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        live = []
        for i, p in enumerate(g["params"]):
            if p.grad is None:
                continue
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous() or p.grad.device != p.device:
                raise ValueError("gradients must be contiguous fp32 tensors on the parameter's device")
            live.append((i, p))
        if not live:
            return loss
        self._step += 1
        dev, n = self._table(live)
        check(lib.pgnn_adam_step(dev.data_ptr(), n, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                                 self.grad_scale, self._step, int(self.legacy_eps),
                                 torch.cuda.current_stream(self._m.device).cuda_stream), "pgnn_adam_step")
        return loss
