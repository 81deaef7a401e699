"""Fused gradient clipping + Adam / AdamW: the step right after the hot path in
every reference training loop,

    nn.utils.clip_grad_norm_(model.parameters(), CLIP)     # others/realformer.py:314 etc.
    optimizer.step()                                        # optim.Adam :342 / optim.AdamW

``Adam`` / ``AdamW`` keep torch.optim's constructor arguments, ``param_groups`` (so
``ReduceLROnPlateau`` keeps working, others/realformer.py:343) and ``state_dict`` layout
(``step``, ``exp_avg``, ``exp_avg_sq``), and run the whole update as one libmmemo launch per 96
tensors.  Two ways to clip:

* drop-in: keep the reference's two calls, with ``clip_grad_norm_`` from this module (two launches:
  squared norm, scale; returns the total norm like torch's);
* fused: ``AdamW(..., max_grad_norm=CLIP)`` and no separate clip call — the coefficient is computed
  on the device from the squared norm and applied while the update reads the gradients (no extra
  pass over them, no host sync).

float32 CUDA parameters only (the master parameters of both precision modes are float32); there is
no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch
from torch import Tensor

from . import ops


def _ptrs(ts: List[Tensor]):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _ns(ts: List[Tensor]):
    return (C.c_int64 * len(ts))(*[t.numel() for t in ts])


def _check(ts: Iterable[Tensor], what: str) -> None:
    for t in ts:
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"mmemo_b200.optim needs contiguous float32 CUDA {what} "
                               "(no CPU fallback)")


def grad_sqnorm(grads: List[Tensor]) -> Tensor:
    """Device scalar  sum_i |g_i|^2  in one launch (per 96 tensors)."""
    _check(grads, "gradients")
    out = torch.empty(1, dtype=torch.float32, device=grads[0].device)
    ops._call("mmemo_grad_sqnorm_f32", len(grads), _ptrs(grads), _ns(grads), out.data_ptr(),
              ops._stream())
    return out


def clip_grad_norm_(parameters, max_norm: float) -> Tensor:
    """Fused ``torch.nn.utils.clip_grad_norm_(parameters, max_norm)`` (L2 norm): scales the
    gradients in place by ``min(1, max_norm / (norm + 1e-6))`` and returns the total norm as a
    0-dim device tensor.  No host synchronisation."""
    if isinstance(parameters, Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros(())
    sq = grad_sqnorm(grads)
    ops._call("mmemo_clip_grads_f32", len(grads), _ptrs(grads), _ns(grads), sq.data_ptr(),
              float(max_norm), ops._stream())
    return sq.sqrt().reshape(())


class _FusedAdam(torch.optim.Optimizer):
    _decoupled = False

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False,
                 max_grad_norm: Optional[float] = None):
        if amsgrad:
            raise ValueError("amsgrad is not used by the reference and not implemented")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        if max_grad_norm is not None and max_grad_norm <= 0:
            raise ValueError("max_grad_norm must be positive")
        self.max_grad_norm = max_grad_norm
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue       # e.g. the empty second group of Ren-MME/run.py:376-379
            _check(ps, "parameters")
            gs = [p.grad for p in ps]
            _check(gs, "gradients")
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)          # torch.optim keeps a CPU float tensor
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            groups.append((group, ps, gs))
        if not groups:
            return loss
        sq = None
        if self.max_grad_norm is not None:     # global norm over every group, like the reference
            sq = grad_sqnorm([g for _, _, gs in groups for g in gs])
        for group, ps, gs in groups:
            # parameters that joined later (frozen before) carry their own step count
            by_step = {}
            for i, p in enumerate(ps):
                by_step.setdefault(int(self.state[p]["step"]), []).append(i)
            for t, sel in sorted(by_step.items()):
                pp = [ps[i] for i in sel]
                gg = [gs[i] for i in sel]
                mm = [self.state[p]["exp_avg"] for p in pp]
                vv = [self.state[p]["exp_avg_sq"] for p in pp]
                b1, b2 = group["betas"]
                ops._call("mmemo_adam_step_f32", len(pp), _ptrs(pp), _ptrs(gg), _ptrs(mm),
                          _ptrs(vv), _ns(pp), float(group["lr"]), float(b1), float(b2),
                          float(group["eps"]), float(group["weight_decay"]),
                          int(self._decoupled), t + 1, None if sq is None else sq.data_ptr(),
                          float(self.max_grad_norm or 0.0), ops._stream())
                for p in pp:
                    self.state[p]["step"] += 1
                # the launch wrote the parameters through raw pointers: bump their version counters
                # so that everything keyed on them (the bf16 weight shadows of ops.shadow_bf16,
                # autograd's saved-tensor checks) sees the update like after an in-place torch op
                torch._C._increment_version(pp)
        return loss


class Adam(_FusedAdam):
    """``torch.optim.Adam`` (L2 weight decay folded into the gradient); others/realformer.py:342."""
    _decoupled = False


class AdamW(_FusedAdam):
    """``torch.optim.AdamW`` (decoupled weight decay, default 1e-2); cmu-mosei/run.py:398,
    Ren-MME/run.py:379, rencecps/run.py:202, robot_demo.py:502."""
    _decoupled = True

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False,
                 max_grad_norm: Optional[float] = None):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad, max_grad_norm)
