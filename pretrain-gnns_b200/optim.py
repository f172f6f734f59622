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

    `params` is an iterable of tensors or of param-group dicts with per-group overrides, as chem/finetune.py:180-185 passes
    (`{"params": ..., "lr": lr * lr_scale}` for the head): one launch per group.  Step counts are kept per tensor like
    torch's `state[p]["step"]` (a parameter that receives no gradient on some step keeps its own bias correction); tensors
    of a group whose counts differ are stepped by separate launches.
    `grad_scale` multiplies every gradient as it is read (1/world_size turns an all-reduced SUM into the mean for free).
    `legacy_eps=True` reproduces torch 1.0.1's placement of eps (the version the reference pins, requirements.txt:2)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0, legacy_eps=False):
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        groups = params if isinstance(params[0], dict) else [dict(params=params)]
        self.param_groups = []
        for g in groups:
            if not isinstance(g, dict) or "params" not in g:
                raise TypeError("params must be an iterable of tensors or of dicts with a 'params' entry")
            ps = g["params"]
            ps = [ps] if torch.is_tensor(ps) else list(ps)
            grp = dict(defaults)
            grp.update({k: (tuple(v) if k == "betas" else v) for k, v in g.items() if k != "params"})
            grp["params"] = ps
            if (grp["lr"] < 0.0 or grp["eps"] < 0.0 or not 0.0 <= grp["betas"][0] < 1.0 or not 0.0 <= grp["betas"][1] < 1.0
                    or grp["weight_decay"] < 0.0):
                raise ValueError("Invalid Adam hyper-parameter")
            self.param_groups.append(grp)
        flat = [p for g in self.param_groups for p in g["params"]]
        if not flat:
            raise ValueError("optimizer got an empty parameter list")
        if len({id(p) for p in flat}) != len(flat):
            raise ValueError("some parameters appear in more than one parameter group")
        for p in flat:
            if not torch.is_tensor(p):
                raise TypeError("optimizer can only optimize Tensors, but one of the params is " + type(p).__name__)
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("pretrain_gnns_b200.optim.Adam needs contiguous fp32 CUDA parameters")
        self.grad_scale, self.legacy_eps = float(grad_scale), bool(legacy_eps)
        self._params = flat
        self._group_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        self._steps = [0] * len(flat)
        dev = flat[0].device
        sizes = [p.numel() for p in flat]
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
    def _state_of(self, i, clone=False):
        s, n = self._slices[i]
        p = self._params[i]
        m, v = self._m[s:s + n].view_as(p), self._v[s:s + n].view_as(p)
        return dict(step=torch.tensor(float(self._steps[i])), exp_avg=m.clone() if clone else m, exp_avg_sq=v.clone() if clone else v)

    @property
    def state(self):
        return {p: self._state_of(i) for i, p in enumerate(self._params)}

    def zero_grad(self, set_to_none=True):
        for p in self._params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    def state_dict(self):
        groups, at = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(at, at + len(g["params"])))
            at += len(g["params"])
            groups.append(d)
        st = {i: self._state_of(i, clone=True) for i in range(len(self._params)) if self._steps[i]}
        return dict(state=st, param_groups=groups)

    def load_state_dict(self, sd):
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for mine, g in zip(self.param_groups, sd["param_groups"]):
            if len(g["params"]) != len(mine["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in g:
                    mine[k] = tuple(g[k]) if k == "betas" else g[k]
        self._steps = [0] * len(self._params)
        self._m.zero_()
        self._v.zero_()
        for i, st in sd["state"].items():
            s, n = self._slices[int(i)]
            self._m[s:s + n].copy_(st["exp_avg"].reshape(-1))
            self._v[s:s + n].copy_(st["exp_avg_sq"].reshape(-1))
            self._steps[int(i)] = int(float(st["step"]))

    # ---- the step ---------------------------------------------------------------------------------------------------
    def _table(self, live):
        key = tuple((i, self._params[i].data_ptr(), self._params[i].grad.data_ptr()) for i in live)
        tab = self._tables.get(key)
        if tab is None:
            if len(self._tables) > 16:
                self._tables.clear()
            mb, vb = self._m.data_ptr(), self._v.data_ptr()
            host = chunk_table([(self._params[i].data_ptr(), self._params[i].grad.data_ptr(), mb + 4 * self._slices[i][0],
                                 vb + 4 * self._slices[i][0]) for i in live], [self._params[i].numel() for i in live])
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
        launches = {}  # (group, step count after this update) -> tensor indices
        for i, p in enumerate(self._params):
            if p.grad is None:
                continue
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous() or p.grad.device != p.device:
                raise ValueError("gradients must be contiguous fp32 tensors on the parameter's device")
            self._steps[i] += 1
            launches.setdefault((self._group_of[i], self._steps[i]), []).append(i)
        st = torch.cuda.current_stream(self._m.device).cuda_stream
        for (gi, step), live in launches.items():
            g = self.param_groups[gi]
            dev, n = self._table(tuple(live))
            check(lib.pgnn_adam_step(dev.data_ptr(), n, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                                     self.grad_scale, step, int(self.legacy_eps), st), "pgnn_adam_step")
        return loss
