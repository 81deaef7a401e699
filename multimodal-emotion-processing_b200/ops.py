"""torch custom ops over the C ABI of libmmemo.so (include/mmemo.h).

Every op allocates its outputs with torch (device memory + stream plumbing only), passes raw
device pointers + the current CUDA stream to an ``extern "C"`` launcher, and registers an explicit
backward (``torch.library.register_autograd``) that calls the matching ``*_bwd`` launchers, so the
reference training loops (``loss.backward()``, ``clip_grad_norm_``, ``optimizer.step()``) work
unchanged.  No op has a CPU implementation: calling one without a CUDA tensor or without the
built library raises.

Precision: activations are float32 (parity mode) or bfloat16 (``bf16=True``).  Parameters stay
float32 ("master"); in bf16 mode GEMM weights are used through bf16 shadow copies refreshed when
the parameter's version counter changes; all parameter gradients are returned in float32.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

BF = torch.bfloat16
F32 = torch.float32
LN_EPS = 1e-5

# number of libmmemo kernel launches issued through this module (bench.py's gpu_launches)
launch_count = 0


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# Launch settings live with the STREAM in libmmemo (mmemo_stream_set_*): the first launch on a
# stream attaches a split-K scratch buffer of its own (launches of one stream are ordered and can
# share it; two streams never do), so concurrent streams - the ensemble's side streams, a capture
# stream next to the default stream - are independent.  The buffers live as long as the process:
# a captured CUDA graph keeps the pointer of the stream it was captured on.
_stream_ws: dict = {}
WORKSPACE_BYTES = 32 << 20
_PDL_OFF = os.environ.get("MMEMO_PDL", "1") == "0"     # debugging knob: fully serialised launches


def _ensure_stream(stream: int) -> None:
    key = (torch.cuda.current_device(), stream)
    if key in _stream_ws:
        return
    lib = _lib.load()
    buf = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=f"cuda:{key[0]}")
    _lib.check(lib.mmemo_stream_set_workspace(stream, buf.data_ptr(), buf.numel()), "stream_set_workspace")
    if _PDL_OFF:
        _lib.check(lib.mmemo_stream_set_pdl(stream, 0), "stream_set_pdl")
    _stream_ws[key] = buf


# Data-parallel overlap: while a bucket all-reduce is in flight its CTAs hold some SMs; a
# persistent (one CTA per SM) GEMM launched meanwhile would have that many CTAs serialised behind
# the others (measured: 12.9 -> 23.4 us per GEMM).  dp.GradReducer therefore lowers the SM budget of
# the next few GEMM launches of the compute stream after it issues an all-reduce; the countdown
# restores the full budget.
_budget_countdown = 0
_budget_stream = 0


def reserve_sms_for(n_launches: int, n_sms_reserved: int) -> None:
    global _budget_countdown, _budget_stream
    total = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    release_sms()
    _budget_stream = _stream()
    _lib.check(_lib.load().mmemo_stream_set_sm_budget(
        _budget_stream, max(2, (total - n_sms_reserved) // 2 * 2)), "sm_budget")
    _budget_countdown = n_launches


def release_sms() -> None:
    global _budget_countdown
    if _budget_countdown:
        _budget_countdown = 0
        _lib.check(_lib.load().mmemo_stream_set_sm_budget(_budget_stream, 0), "sm_budget")


def _call(name: str, *args) -> None:
    global launch_count, _budget_countdown
    _ensure_stream(_stream())
    launch_count += 1
    _lib.check(getattr(_lib.load(), name)(*args), name)
    if _budget_countdown and name.startswith("mmemo_linear"):
        _budget_countdown -= 1
        if _budget_countdown == 0:
            _lib.check(_lib.load().mmemo_stream_set_sm_budget(_budget_stream, 0), "sm_budget")


def _try_call(name: str, *args) -> bool:
    """Like ``_call`` for the grouped entry points: False when the library answers
    MMEMO_ERR_SHAPE (a problem of the group is outside the grouped kernel's limits; the caller then
    issues the problems one by one), raises on any other error."""
    global launch_count
    _ensure_stream(_stream())
    rc = getattr(_lib.load(), name)(*args)
    if rc == -2:
        return False
    _lib.check(rc, name)
    launch_count += 1
    return True


def _sfx(bf16: bool) -> str:
    return "bf16" if bf16 else "f32"


def _act_dtype(bf16: bool):
    return BF if bf16 else F32


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmemo_b200 ops run on CUDA tensors only (no CPU fallback)")


# ------------------------------------------------------------------------------------------------
# bf16 shadow copies of float32 master weights
# ------------------------------------------------------------------------------------------------
_shadow: dict = {}
_shadow_pad: dict = {}     # zero-padded (N, K rounded up to 8) shadows of the projection weights


def clear_shadow_cache() -> None:
    """Drop all bf16 weight shadows (call before CUDA-graph capture so the casts are captured)."""
    _shadow.clear()
    _shadow_pad.clear()


def shadow_bf16(*ws: Tensor) -> Tensor:
    """bf16 copy of one weight, or of several weights concatenated along dim 0 (e.g. [Wk; Wv])."""
    key = tuple((w.data_ptr(), w._version, tuple(w.shape)) for w in ws)
    hit = _shadow.get(key)
    if hit is not None:
        return hit
    # evict stale entries of the same storage(s)
    ptrs = {k[0] for k in key}
    for k in [k for k in _shadow if any(e[0] in ptrs for e in k) and len(k) == len(key)]:
        del _shadow[k]
    rows = sum(w.shape[0] for w in ws)
    cols = ws[0].numel() // ws[0].shape[0]
    out = torch.empty(rows, cols, dtype=BF, device=ws[0].device)
    r = 0
    for w in ws:
        wd = w.detach()
        if not wd.is_contiguous():
            wd = wd.contiguous()
        _call("mmemo_cast_f32_to_bf16", wd.data_ptr(), out[r:].data_ptr(), wd.numel(), _stream())
        r += w.shape[0]
    _shadow[key] = out
    return out


def shadow_bf16_block(groups: Sequence[Sequence[Tensor]]) -> List[Tensor]:
    """bf16 shadows of several weight groups (each group = weights concatenated along dim 0) with
    ONE cast launch; also fills the per-group cache used by ``shadow_bf16``."""
    keys = [tuple((w.data_ptr(), w._version, tuple(w.shape)) for w in g) for g in groups]
    if all(k in _shadow for k in keys):
        return [_shadow[k] for k in keys]
    flat = [w for g in groups for w in g]
    if len(flat) > 64:
        half = len(groups) // 2
        return shadow_bf16_block(groups[:half]) + shadow_bf16_block(groups[half:])
    outs, srcs, dsts, ns = [], [], [], []
    for g, k in zip(groups, keys):
        ptrs = {e[0] for e in k}
        for old in [o for o in _shadow if any(e[0] in ptrs for e in o) and len(o) == len(k)]:
            del _shadow[old]
        rows = sum(w.shape[0] for w in g)
        out = torch.empty(rows, g[0].numel() // g[0].shape[0], dtype=BF, device=g[0].device)
        r = 0
        for w in g:
            wd = w.detach()
            wd = wd if wd.is_contiguous() else wd.contiguous()
            srcs.append(wd)
            dsts.append(out[r:].data_ptr())
            ns.append(wd.numel())
            r += w.shape[0]
        _shadow[k] = out
        outs.append(out)
    _call("mmemo_cast_f32_to_bf16_multi", len(srcs), _arr(C.c_void_p, [t.data_ptr() for t in srcs]),
          _arr(C.c_void_p, dsts), _arr(C.c_int64, ns), _stream())
    return outs


def _weight(bf16: bool, *ws: Tensor) -> Tensor:
    if bf16:
        return shadow_bf16(*ws)
    if len(ws) == 1:
        w = ws[0].detach()
        w = w.reshape(w.shape[0], -1)
        return w if w.is_contiguous() else w.contiguous()
    return torch.cat([w.detach().reshape(w.shape[0], -1) for w in ws], 0)


def _rows(x: Tensor) -> Tuple[Tensor, int, int]:
    """View (..., K) as (M, K) with unit inner stride; returns (tensor, M, ld)."""
    K = x.shape[-1]
    if x.dim() == 2 and x.stride(1) == 1:
        return x, x.shape[0], x.stride(0)
    if not x.is_contiguous():
        x = x.contiguous()
    return x, x.numel() // K, K


# ------------------------------------------------------------------------------------------------
# gradient destinations: data-parallel training (dp.GradReducer) registers, per parameter, its slot
# in a flat all-reduce bucket; the fused block backward then lets the weight-gradient GEMMs write
# straight into the bucket (no per-parameter copy).  Only valid while p.grad is None at backward
# time (zero_grad(set_to_none=True)); the reducer disables it otherwise.
# ------------------------------------------------------------------------------------------------
_grad_dest: dict = {}
grad_dest_enabled = True
# set by dp.GradReducer.backward(): the bucket slots were zero-filled before this backward, so
# kernels that accumulate into them (split-K reduce-add) need no fill of their own
grad_dest_zeroed = False


# A slot can be WRITTEN by one op per backward: a parameter that is used twice (shared weights)
# gets a fresh buffer for its second gradient and autograd sums the two as usual.
_dest_claimed: set = set()


def register_grad_dest(param: Tensor, flat: Tensor, offset: int) -> None:
    _grad_dest[param.data_ptr()] = (flat, offset, param.numel())


def begin_backward() -> None:
    """Called by dp.GradReducer.backward(): every slot is writable again."""
    _dest_claimed.clear()


def _claim(*ws: Tensor) -> bool:
    keys = [w.data_ptr() for w in ws]
    if any(k in _dest_claimed for k in keys):
        return False
    _dest_claimed.update(keys)
    return True


def clear_grad_dest() -> None:
    _grad_dest.clear()
    _zbuf_dest.clear()


def unregister_grad_dests(flats: Sequence[Tensor]) -> None:
    """Forget every destination that lives in one of ``flats`` (a reducer that goes away must not
    leave entries keyed by parameter addresses a later model could reuse)."""
    ids = {id(f) for f in flats}
    for reg in (_grad_dest, _zbuf_dest):
        for k in [k for k, v in reg.items() if id(v[0]) in ids]:
            del reg[k]


# The fused full block accumulates ALL its parameter gradients (five weight gradients + the small
# "+=" outputs) into one zero-filled buffer ("zbuf", layout: full_block_grad_layout).  Under data
# parallelism the reducer gives every block a contiguous bucket region with that layout and
# registers it here (keyed by the block's Wq): the block's zbuf IS then a bucket slice, p.grad of
# all 15 parameters are views of the bucket, and nothing is copied or filled per block.
_zbuf_dest: dict = {}


def register_zbuf_dest(key: Tensor, flat: Tensor, offset: int, total: int) -> None:
    _zbuf_dest[key.data_ptr()] = (flat, offset, total)


def full_block_grad_layout(params: Sequence[Tensor]):
    """(key parameter, [(parameter, offset in floats)], total floats) of a full block's zbuf; the
    offsets are the views _block_full_backward hands to autograd."""
    wq, wk, wv, wo, n1w, n1b, n2w, n2b, f1w, f1b, f2w, f2b, ga, gb, gc = params
    d, dff = wq.shape[0], f1w.shape[0]
    n_small, sizes = _zbuf_layout(d, dff)
    p1, zo = 1 + 2 * d, 2 * (1 + 2 * d)
    lay = [(gb, 0), (n2w, 1), (n2b, 1 + d), (ga, p1), (n1w, p1 + 1), (n1b, p1 + 1 + d),
           (f2b, zo), (f1b, zo + d), (gc, zo + d + dff)]
    o = n_small
    for w, n in zip((wq, wk, wv, wo, f1w, f2w), (d * d, d * d, d * d, d * d, dff * d, d * dff)):
        lay.append((w, o))
        o += n
    assert o == n_small + sum(sizes)
    return wq, lay, o


def lite_zlayout(d: int):
    """Zero buffer of a lite block: [dpn (1+2d) | dc (1) | pad] then dWo (d,d), dWm (d,2d)."""
    n_small = (1 + 2 * d + 1 + 63) // 64 * 64
    return n_small, n_small + d * d + 2 * d * d


def lite_block_grad_layout(params: Sequence[Tensor]):
    """Like full_block_grad_layout for the lite block (params: Wo, Wm, norm weight, norm bias, c)."""
    wo, wm, nw, nb, c = params
    d = wo.shape[0]
    n_small, total = lite_zlayout(d)
    return wo, [(nw, 1), (nb, 1 + d), (c, 1 + 2 * d), (wo, n_small), (wm, n_small + d * d)], total


def claim_zbufs(keys: Sequence[Tensor], total: int) -> Optional[List[Tensor]]:
    """Bucket regions for the zero buffers of a GROUP of blocks (one grouped trunk layer): all of
    them or none (None: the caller allocates and fills its own buffer)."""
    if not (grad_dest_enabled and grad_dest_zeroed):
        return None
    zs = [_zbuf_dest.get(k.data_ptr()) for k in keys]
    if any(z is None or z[2] != total for z in zs):
        return None
    if any(k.data_ptr() in _dest_claimed for k in keys):
        return None
    _dest_claimed.update(k.data_ptr() for k in keys)
    return [z[0][z[1]:z[1] + z[2]] for z in zs]


def zbuf_region(key: Tensor) -> Tensor:
    z = _zbuf_dest[key.data_ptr()]
    return z[0][z[1]:z[1] + z[2]]


def _dest(w: Tensor) -> Optional[Tensor]:
    e = _grad_dest.get(w.data_ptr()) if grad_dest_enabled else None
    if e is None or e[2] != w.numel():
        return None
    return e[0][e[1]:e[1] + e[2]].view(w.shape)


def _dest_pair(wa: Tensor, wb: Tensor) -> Optional[Tensor]:
    """One (rows_a + rows_b, cols) buffer when the two parameters own adjacent bucket slots."""
    if not grad_dest_enabled:
        return None
    ea, eb = _grad_dest.get(wa.data_ptr()), _grad_dest.get(wb.data_ptr())
    if ea is None or eb is None or ea[0] is not eb[0] or ea[1] + ea[2] != eb[1]:
        return None
    return ea[0][ea[1]:ea[1] + ea[2] + eb[2]].view(wa.shape[0] + wb.shape[0], -1)


def _wgrad(w: Tensor, rows: int, cols: int):
    """(buffer to write dW into, tensor to return from the custom op)."""
    d = _dest(w)
    if d is not None and _claim(w):
        return d.view(rows, cols), torch.empty(0, device=w.device)
    buf = torch.empty(rows, cols, dtype=F32, device=w.device)
    return buf, buf


# ------------------------------------------------------------------------------------------------
# raw launch helpers (no autograd) -- shared by the fine-grained ops and the fused block ops
# ------------------------------------------------------------------------------------------------
def _linear_fwd(bf16, x, w, bias, pos, out, relu=False, accumulate=False, ldw=None, N=None, K=None):
    x, M, ldx = _rows(x)
    N = w.shape[0] if N is None else N
    K = x.shape[-1] if K is None else K
    ldw = w.stride(0) if ldw is None else ldw
    _call(f"mmemo_linear_fwd_{_sfx(bf16)}", x.data_ptr(), int(x.dtype == F32), ldx, w.data_ptr(),
          ldw, _p(bias), _p(pos), 0 if pos is None else pos.shape[0], out.data_ptr(),
          out.stride(-2), M, N, K, int(relu), int(accumulate), _stream())


def _linear_bwd_x(bf16, dy, w, dx, relu_src=None, accumulate=False, ldw=None, N=None, K=None):
    dy, M, lddy = _rows(dy)
    N = w.shape[0] if N is None else N
    K = w.shape[1] if K is None else K
    ldw = w.stride(0) if ldw is None else ldw
    _call(f"mmemo_linear_bwd_x_{_sfx(bf16)}", dy.data_ptr(), lddy, w.data_ptr(), ldw,
          dx.data_ptr(), dx.stride(-2), _p(relu_src),
          0 if relu_src is None else relu_src.stride(-2), M, N, K, int(accumulate), _stream())


def _linear_bwd_w(bf16, dy, x, dw, dbias=None, accumulate=False, N=None, K=None):
    dy, M, lddy = _rows(dy)
    x, _, ldx = _rows(x)
    N = dy.shape[-1] if N is None else N
    K = x.shape[-1] if K is None else K
    _call(f"mmemo_linear_bwd_w_{_sfx(bf16)}", dy.data_ptr(), lddy, x.data_ptr(),
          int(x.dtype == F32), ldx, dw.data_ptr(), dw.stride(0), _p(dbias), M, N, K,
          int(accumulate), _stream())


def _arr(ctype, vals):
    return (ctype * len(vals))(*vals)


MAX_GEMM_GROUP = 48     # problems per grouped tensor-core launch (csrc/gemm.h GEMM_TC_MAX_GROUP)


def _chunks(items, n=MAX_GEMM_GROUP):
    return [items[i:i + n] for i in range(0, len(items), n)]


def _linear_fwd_group(bf16, items):
    """items: [(x, w, bias, out2d, relu[, accumulate[, pos]])].  One grouped tensor-core launch (per
    48 problems) in bf16 mode; N, K and the weight's leading dimension come from ``w`` (a 2-D view,
    e.g. one half of a concat weight); ``pos`` = (L, N) position table added with period L (may be a
    column slice of a wider table: its row stride is passed along)."""
    items = [tuple(it) + (False, None)[len(it) - 5:] for it in items]
    if not bf16 or len(items) == 1:
        for x, w, b, out, relu, acc, pos in items:
            assert pos is None or pos.is_contiguous(), "a sliced position table needs the grouped path"
            _linear_fwd(bf16, x, w, b, pos, out, relu=relu, accumulate=acc, ldw=w.stride(0),
                        N=w.shape[0], K=w.shape[1])
        return
    for part in _chunks(items):
        xs = [_rows(it[0]) for it in part]
        _call("mmemo_linear_fwd_grouped_bf16", len(part),
              _arr(C.c_void_p, [x.data_ptr() for x, _, _ in xs]), _arr(C.c_int64, [ld for _, _, ld in xs]),
              _arr(C.c_void_p, [it[1].data_ptr() for it in part]),
              _arr(C.c_int64, [it[1].stride(0) for it in part]),
              _arr(C.c_void_p, [_p(it[2]) for it in part]),
              _arr(C.c_void_p, [it[3].data_ptr() for it in part]),
              _arr(C.c_int64, [it[3].stride(-2) for it in part]),
              _arr(C.c_int64, [M for _, M, _ in xs]), _arr(C.c_int64, [it[1].shape[0] for it in part]),
              _arr(C.c_int64, [it[1].shape[1] for it in part]),
              _arr(C.c_int, [int(it[4]) for it in part]), _arr(C.c_int, [int(it[5]) for it in part]),
              _arr(C.c_void_p, [_p(it[6]) for it in part]),
              _arr(C.c_int64, [0 if it[6] is None else it[6].shape[0] for it in part]),
              _arr(C.c_int64, [0 if it[6] is None else it[6].stride(0) for it in part]), _stream())


def _linear_bwd_x_group(bf16, items):
    """items: [(dy, w, dx2d, accumulate[, relu_src2d])] — outputs must not alias each other."""
    items = [it if len(it) == 5 else (*it, None) for it in items]
    if not bf16 or len(items) == 1:
        for dy, w, dx, acc, rs in items:
            _linear_bwd_x(bf16, dy, w, dx, relu_src=rs, accumulate=acc, ldw=w.stride(0),
                          N=w.shape[0], K=w.shape[1])
        return
    for part in _chunks(items):
        dys = [_rows(it[0]) for it in part]
        _call("mmemo_linear_bwd_x_grouped_bf16", len(part),
              _arr(C.c_void_p, [d.data_ptr() for d, _, _ in dys]), _arr(C.c_int64, [ld for _, _, ld in dys]),
              _arr(C.c_void_p, [it[1].data_ptr() for it in part]),
              _arr(C.c_int64, [it[1].stride(0) for it in part]),
              _arr(C.c_void_p, [it[2].data_ptr() for it in part]),
              _arr(C.c_int64, [it[2].stride(-2) for it in part]),
              _arr(C.c_void_p, [_p(it[4]) for it in part]),
              _arr(C.c_int64, [0 if it[4] is None else it[4].stride(-2) for it in part]),
              _arr(C.c_int64, [M for _, M, _ in dys]), _arr(C.c_int64, [it[1].shape[0] for it in part]),
              _arr(C.c_int64, [it[1].shape[1] for it in part]),
              _arr(C.c_int, [int(it[3]) for it in part]), _stream())


def _linear_bwd_w_group(bf16, items, zeroed=False):
    """items: [(dy, x, dw)] -> dw = dy^T x (float32), all in one grouped launch (per 48 problems) in
    bf16 mode.  ``zeroed``: every dw is already all zero (saves the split-K path its own fills)."""
    if not bf16 or len(items) == 1:
        for dy, x, dw in items:
            _linear_bwd_w(bf16, dy, x, dw)
        return
    for part in _chunks(items):
        dys = [_rows(it[0]) for it in part]
        xs = [_rows(it[1]) for it in part]
        _call("mmemo_linear_bwd_w_grouped_bf16", len(part),
              _arr(C.c_void_p, [d.data_ptr() for d, _, _ in dys]), _arr(C.c_int64, [ld for _, _, ld in dys]),
              _arr(C.c_void_p, [x.data_ptr() for x, _, _ in xs]), _arr(C.c_int64, [ld for _, _, ld in xs]),
              _arr(C.c_void_p, [it[2].data_ptr() for it in part]),
              _arr(C.c_int64, [it[2].stride(0) for it in part]),
              _arr(C.c_int64, [M for _, M, _ in dys]), _arr(C.c_int64, [d.shape[-1] for d, _, _ in dys]),
              _arr(C.c_int64, [x.shape[-1] for x, _, _ in xs]), 2 if zeroed else 0, _stream())


def _add_ln_fwd(bf16, res, x, gate, gamma, beta, relu=False):
    x2, M, ldx = _rows(x)
    d = x.shape[-1]
    if res is not None:
        res, _, ldres = _rows(res)
    else:
        ldres = 0
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    stat = torch.empty(2, M, dtype=F32, device=x.device)
    _call(f"mmemo_add_ln_fwd_{_sfx(bf16)}", _p(res), ldres, x2.data_ptr(), ldx, _p(gate),
          gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), d, stat[0].data_ptr(),
          stat[1].data_ptr(), M, d, LN_EPS, int(relu), _stream())
    return y, stat


def _add_ln_bwd(bf16, dy, res, x, gate, gamma, y, stat, relu, want_dres, dres_out=None, dpar=None,
                dxsum=None):
    """Returns (dres, dx, dparams) with dparams = float32 [1 + 2d] = (dgate | dgamma | dbeta);
    ``dxsum`` (float32 [d], zero-initialised by the caller) receives the column sums of dx."""
    dy, M, lddy = _rows(dy)
    x2, _, ldx = _rows(x)
    d = x.shape[-1]
    ldres = 0
    if res is not None:
        res, _, ldres = _rows(res)
    dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    dres = None
    if want_dres:
        dres = dres_out if dres_out is not None else torch.empty(x.shape, dtype=x.dtype,
                                                                 device=x.device)
    if dpar is None:
        dpar = torch.zeros(1 + 2 * d, dtype=F32, device=x.device)
    _call(f"mmemo_add_ln_bwd_{_sfx(bf16)}", dy.data_ptr(), lddy, _p(res), ldres, x2.data_ptr(),
          ldx, _p(gate), gamma.data_ptr(), _p(y) if relu else None, d, stat[0].data_ptr(),
          stat[1].data_ptr(), _p(dres), d, dx.data_ptr(), d, dpar.data_ptr(),
          dpar[1:].data_ptr(), dpar[1 + d:].data_ptr(), _p(dxsum), M, d, int(relu), _stream())
    return dres, dx, dpar


def _rowsum(bf16_in: bool, x, out, period=1):
    x, M, ldx = _rows(x)
    _call(f"mmemo_rowsum_{_sfx(bf16_in)}", x.data_ptr(), ldx, out.data_ptr(), M, x.shape[-1],
          period, _stream())


def _bld(t: Tensor, L: int) -> int:
    """Row stride of a (B, L, d) activation whose batch stride must be L*ld."""
    assert t.dim() == 3 and t.stride(2) == 1 and (t.shape[0] == 1 or t.stride(0) == L * t.stride(1)), \
        "activation must be (B, L, d) with unit inner stride and packed batch"
    return t.stride(1)


def _mask_strides(mask: Optional[Tensor], Lq: int, Lk: int) -> Tuple[int, int]:
    if mask is None:
        return 0, 0
    if mask.dim() == 2:
        return Lk, 0
    return Lq * Lk, Lk


def _attn_fwd(bf16, q, k, v, mask, s_prev, c, H, want_s):
    B, Lq, d = q.shape
    Lk = k.shape[1]
    hd = d // H
    dt = _act_dtype(bf16)
    o = torch.empty(B, Lq, d, dtype=dt, device=q.device)
    s = torch.empty(B, H, Lq, Lk, dtype=dt, device=q.device) if want_s else None
    stat = torch.empty(B, H, Lq, 2, dtype=F32, device=q.device)
    mbs, mrs = _mask_strides(mask, Lq, Lk)
    _call(f"mmemo_resattn_fwd_{_sfx(bf16)}", q.data_ptr(), _bld(q, Lq), k.data_ptr(), _bld(k, Lk),
          v.data_ptr(), _bld(v, Lk), _p(mask), mbs, mrs, _p(s_prev),
          _p(c) if s_prev is not None else None, _p(s), o.data_ptr(), d, stat.data_ptr(), B, H, Lq,
          Lk, hd, _stream())
    return o, s, stat


def _attn_bwd(bf16, do, q, k, v, mask, s, s_prev, c, ds_next, o, stat, H, dq, dk, dv, want_dsprev,
              dc_out=None):
    B, Lq, d = q.shape
    Lk = k.shape[1]
    hd = d // H
    dt = _act_dtype(bf16)
    ds_prev = None
    dc = None
    if s_prev is not None:
        # "+=" scalar: a zero-initialised slot of the caller, or our own
        dc = dc_out if dc_out is not None else torch.zeros(1, dtype=F32, device=q.device)
        if want_dsprev:
            ds_prev = torch.empty(B, H, Lq, Lk, dtype=dt, device=q.device)
    ws = None
    if bf16 and Lk > 128 and not _lib.load().mmemo_resattn_uses_tensor_cores(Lq, Lk, hd, d):
        ws = torch.empty(B, Lq, d, dtype=F32, device=q.device)
    mbs, mrs = _mask_strides(mask, Lq, Lk)
    _call(f"mmemo_resattn_bwd_{_sfx(bf16)}", do.data_ptr(), _bld(do, Lq), q.data_ptr(),
          _bld(q, Lq), k.data_ptr(), _bld(k, Lk), v.data_ptr(), _bld(v, Lk), _p(mask), mbs, mrs,
          _p(s), _p(s_prev), _p(c) if s_prev is not None else None, _p(ds_next), o.data_ptr(),
          _bld(o, Lq), stat.data_ptr(), dq.data_ptr(), _bld(dq, Lq), dk.data_ptr(), _bld(dk, Lk),
          dv.data_ptr(), _bld(dv, Lk), _p(ds_prev), _p(dc), _p(ws), B, H, Lq, Lk, hd, _stream())
    return ds_prev, dc


# ------------------------------------------------------------------------------------------------
# grouped residual attention (bf16): the problems of a fusion-trunk layer in ONE launch
# ------------------------------------------------------------------------------------------------
def score_stride(Lk: int) -> int:
    """Row stride of the score tensors the fused trunk allocates: padded to 8 elements so that the
    kernels move S / S_prev / dS as 16-byte vectors (the returned scores are the [..., :Lk] view)."""
    return (Lk + 7) // 8 * 8


def _attn_problem(q, k, v, mask, s_prev, c, s_out, o, stat, H, lds, d_o=None, s=None, ds_next=None,
                  dq=None, dk=None, dv=None, ds_prev=None, dc=None) -> "_lib.AttnProblem":
    B, Lq, d = q.shape
    Lk = k.shape[1]
    a = _lib.AttnProblem()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = _bld(q, Lq), _bld(k, Lk), _bld(v, Lk)
    a.mask, a.mask_bs = _p(mask), Lk
    a.s_prev, a.c = _p(s_prev), (_p(c) if s_prev is not None else None)
    a.s_out, a.lds = _p(s_out), lds
    a.o, a.ldo, a.lse = o.data_ptr(), _bld(o, Lq), stat.data_ptr()
    a.B, a.H, a.Lq, a.Lk, a.hd = B, H, Lq, Lk, d // H
    if d_o is not None:
        a.d_o, a.lddo = d_o.data_ptr(), _bld(d_o, Lq)
        a.s, a.ds_next = _p(s), _p(ds_next)
        a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
        a.lddq, a.lddk, a.lddv = _bld(dq, Lq), _bld(dk, Lk), _bld(dv, Lk)
        a.ds_prev, a.dc = _p(ds_prev), _p(dc)
    return a


def _attn_group_call(name: str, probs) -> bool:
    """One grouped launch; False when the mma kernels do not take one of the shapes."""
    arr = (_lib.AttnProblem * len(probs))(*probs)
    return _try_call(name, len(probs), C.cast(arr, C.c_void_p), _stream())


# ------------------------------------------------------------------------------------------------
# mmemo::linear  (Linear / Conv1d(k=1) with optional bias, fused position table, fused ReLU)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::linear", mutates_args=())
def linear_op(x: Tensor, w: Tensor, bias: Optional[Tensor], pos: Optional[Tensor], relu: bool,
              bf16: bool) -> Tensor:
    _need_cuda(x, w)
    N = w.shape[0]
    y = torch.empty(*x.shape[:-1], N, dtype=_act_dtype(bf16), device=x.device)
    if pos is not None:
        assert x.dim() == 3 and pos.shape[0] == x.shape[1], "position table length != seq length"
    _linear_fwd(bf16, x, _weight(bf16, w), bias, pos, y.view(-1, N), relu=relu)
    return y


@linear_op.register_fake
def _(x, w, bias, pos, relu, bf16):
    return x.new_empty(*x.shape[:-1], w.shape[0], dtype=_act_dtype(bf16))


@torch.library.custom_op("mmemo::linear_bwd", mutates_args=())
def linear_bwd_op(dy: Tensor, x: Tensor, w: Tensor, y: Optional[Tensor], has_bias: bool,
                  pos_len: int, need_dx: bool, bf16: bool) -> List[Tensor]:
    N, K = w.shape[0], w.numel() // w.shape[0]
    dev = x.device
    dy = dy.contiguous()
    if y is not None:  # fused ReLU: mask the incoming gradient once
        dy = torch.where(y > 0, dy, torch.zeros((), dtype=dy.dtype, device=dev))
    dw = torch.empty(N, K, dtype=F32, device=dev)
    dbias = torch.zeros(N, dtype=F32, device=dev) if has_bias else torch.empty(0, device=dev)
    _linear_bwd_w(bf16, dy.view(-1, N), x, dw, dbias if has_bias else None)
    dpos = torch.empty(0, device=dev)
    if pos_len > 0:
        dpos = torch.zeros(pos_len, N, dtype=F32, device=dev)
        _rowsum(bf16, dy.view(-1, N), dpos, period=pos_len)
    dx = torch.empty(0, device=dev)
    if need_dx:
        dxa = torch.empty(x.shape, dtype=_act_dtype(bf16), device=dev)
        _linear_bwd_x(bf16, dy.view(-1, N), _weight(bf16, w), dxa.view(-1, K))
        dx = dxa if dxa.dtype == x.dtype else dxa.to(x.dtype)
    return [dx, dw.view(w.shape), dbias, dpos]


def _linear_setup(ctx, inputs, output):
    x, w, bias, pos, relu, bf16 = inputs
    ctx.save_for_backward(x, w, output if relu else None)
    ctx.has_bias = bias is not None
    ctx.pos_len = 0 if pos is None else pos.shape[0]
    ctx.bf16 = bf16


def _linear_backward(ctx, dy):
    x, w, y = ctx.saved_tensors
    dx, dw, dbias, dpos = linear_bwd_op(dy, x, w, y, ctx.has_bias, ctx.pos_len,
                                        ctx.needs_input_grad[0], ctx.bf16)
    return (dx if ctx.needs_input_grad[0] else None, dw, dbias if ctx.has_bias else None,
            dpos if ctx.pos_len else None, None, None)


linear_op.register_autograd(_linear_backward, setup_context=_linear_setup)


def linear(x, w, bias=None, pos=None, relu=False, bf16=False):
    if w.dim() == 3:  # Conv1d(kernel_size=1) weight (out, in, 1)
        w = w.squeeze(-1)
    return linear_op(x, w, bias, pos, relu, bf16)


# ------------------------------------------------------------------------------------------------
# mmemo::add_ln   y = act(LN(res + gate*x))
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::add_ln", mutates_args=())
def add_ln_op(res: Optional[Tensor], x: Tensor, gate: Optional[Tensor], gamma: Tensor,
              beta: Tensor, relu: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda(x)
    return _add_ln_fwd(x.dtype == BF, res, x, gate, gamma, beta, relu)


@add_ln_op.register_fake
def _(res, x, gate, gamma, beta, relu):
    return torch.empty_like(x), x.new_empty(2, x.numel() // x.shape[-1], dtype=F32)


@torch.library.custom_op("mmemo::add_ln_bwd", mutates_args=())
def add_ln_bwd_op(dy: Tensor, res: Optional[Tensor], x: Tensor, gate: Optional[Tensor],
                  gamma: Tensor, y: Tensor, stat: Tensor, relu: bool) -> List[Tensor]:
    dres, dx, dpar = _add_ln_bwd(x.dtype == BF, dy.contiguous(), res, x, gate, gamma, y, stat, relu,
                                 res is not None)
    return [dres if dres is not None else torch.empty(0, device=x.device), dx, dpar]


def _add_ln_setup(ctx, inputs, output):
    res, x, gate, gamma, beta, relu = inputs
    y, stat = output
    ctx.save_for_backward(res, x, gate, gamma, y, stat)
    ctx.relu = relu
    ctx.set_materialize_grads(False)


def _add_ln_backward(ctx, dy, _dstat):
    res, x, gate, gamma, y, stat = ctx.saved_tensors
    d = x.shape[-1]
    dres, dx, dpar = add_ln_bwd_op(dy, res, x, gate, gamma, y, stat, ctx.relu)
    return (dres if res is not None else None, dx, dpar[0:1] if gate is not None else None,
            dpar[1:1 + d], dpar[1 + d:], None)


add_ln_op.register_autograd(_add_ln_backward, setup_context=_add_ln_setup)


def add_ln(res, x, gate, gamma, beta, relu=False):
    return add_ln_op(res, x, gate, gamma, beta, relu)[0]


# ------------------------------------------------------------------------------------------------
# mmemo::resattn  (the attention core on already-projected q/k/v; unit-testable piece of kernel a)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::resattn", mutates_args=())
def resattn_op(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor], s_prev: Optional[Tensor],
               c: Optional[Tensor], n_heads: int) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(q, k, v)
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o, s, stat = _attn_fwd(q.dtype == BF, q, k, v, mask, s_prev, c, n_heads, True)
    return o, s, stat


@torch.library.custom_op("mmemo::resattn_bwd", mutates_args=())
def resattn_bwd_op(do: Tensor, ds_next: Optional[Tensor], q: Tensor, k: Tensor, v: Tensor,
                   mask: Optional[Tensor], s: Tensor, s_prev: Optional[Tensor],
                   c: Optional[Tensor], o: Tensor, stat: Tensor, n_heads: int) -> List[Tensor]:
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ds_prev, dc = _attn_bwd(q.dtype == BF, do.contiguous(), q, k, v, mask, s, s_prev, c,
                            None if ds_next is None else ds_next.contiguous(), o, stat, n_heads,
                            dq, dk, dv, True)
    return [dq, dk, dv, ds_prev if ds_prev is not None else torch.empty(0, device=q.device),
            dc if dc is not None else torch.empty(0, device=q.device)]


def _resattn_setup(ctx, inputs, output):
    q, k, v, mask, s_prev, c, H = inputs
    o, s, stat = output
    ctx.save_for_backward(q, k, v, mask, s, s_prev, c, o, stat)
    ctx.H = H
    ctx.set_materialize_grads(False)


def _resattn_backward(ctx, do, ds, _dstat):
    q, k, v, mask, s, s_prev, c, o, stat = ctx.saved_tensors
    if do is None:
        do = torch.zeros_like(o)
    dq, dk, dv, dsp, dc = resattn_bwd_op(do, ds, q, k, v, mask, s, s_prev, c, o, stat, ctx.H)
    has_prev = s_prev is not None
    return (dq, dk, dv, None, dsp if has_prev else None,
            dc if (has_prev and c is not None) else None, None)


resattn_op.register_autograd(_resattn_backward, setup_context=_resattn_setup)


# ------------------------------------------------------------------------------------------------
# mmemo::block_full — the whole RealFormer block as ONE autograd node
#   (others/realformer.py:182-209 == robot_demo.py:347-374)
# params order: wq wk wv wo  n1w n1b n2w n2b  f1w f1b f2w f2b  a b c
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::block_full", mutates_args=())
def block_full_op(q: Tensor, kv: Tensor, mask: Optional[Tensor], s_prev: Optional[Tensor],
                  params: Sequence[Tensor], n_heads: int, bf16: bool, emit_s: bool,
                  same_qkv: bool) -> List[Tensor]:
    _need_cuda(q, kv)
    wq, wk, wv, wo, n1w, n1b, n2w, n2b, f1w, f1b, f2w, f2b, ga, gb, gc = params
    B, Lq, d = q.shape
    Lk = kv.shape[1]
    dt, dev = _act_dtype(bf16), q.device
    q = q.contiguous()
    kv = q if same_qkv else kv.contiguous()
    if bf16:   # all bf16 weight shadows of the block in one cast launch
        shadow_bf16_block([[wq], [wk, wv], [wo], [f1w], [f2w]])
    # Q projection and fused [K|V] projection (one GEMM, N = 2d)
    qp = torch.empty(B, Lq, d, dtype=dt, device=dev)
    kvp = torch.empty(B, Lk, 2 * d, dtype=dt, device=dev)
    _linear_fwd_group(bf16, [(q, _weight(bf16, wq), None, qp.view(-1, d), False),
                             (kv, _weight(bf16, wk, wv), None, kvp.view(-1, 2 * d), False)])
    kp, vp = kvp[..., :d], kvp[..., d:]
    o, s, stat = _attn_fwd(bf16, qp, kp, vp, mask, s_prev, gc, n_heads, emit_s)
    # output projection, gated residual + LN1
    x = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_fwd(bf16, o, _weight(bf16, wo), None, None, x.view(-1, d))
    h1, st1 = _add_ln_fwd(bf16, q, x, ga, n1w, n1b)
    # FFN (ReLU fused in the first GEMM's epilogue), gated residual + LN2
    dff = f1w.shape[0]
    f1 = torch.empty(B, Lq, dff, dtype=dt, device=dev)
    f2 = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_fwd(bf16, h1, _weight(bf16, f1w), f1b, None, f1.view(-1, dff), relu=True)
    _linear_fwd(bf16, f1, _weight(bf16, f2w), f2b, None, f2.view(-1, d))
    h2, st2 = _add_ln_fwd(bf16, h1, f2, gb, n2w, n2b)
    e = torch.empty(0, device=dev)
    return [h2, s if s is not None else e, qp, kvp, o, stat, x, h1, st1, f1, f2, st2]


def _zbuf_layout(d: int, dff: int):
    """(floats of the small "+=" outputs rounded to 256 B, [sizes of dWq, dWkv, dWo, dWf1, dWf2])."""
    n_small = (2 * (1 + 2 * d) + d + dff + 1 + 63) // 64 * 64
    return n_small, [d * d, 2 * d * d, d * d, dff * d, d * dff]


def _zbuf_weights(zbuf: Tensor, d: int, dff: int):
    n_small, sizes = _zbuf_layout(d, dff)
    shapes = [(d, d), (2 * d, d), (d, d), (dff, d), (d, dff)]
    out, o = [], n_small
    for n, sh in zip(sizes, shapes):
        out.append(zbuf[o:o + n].view(sh))
        o += n
    return out


@torch.library.custom_op("mmemo::block_full_bwd", mutates_args=())
def block_full_bwd_op(dh2: Tensor, ds_next: Optional[Tensor], q: Tensor, kv: Tensor,
                      mask: Optional[Tensor], s_prev: Optional[Tensor], s: Optional[Tensor],
                      saved: Sequence[Tensor], params: Sequence[Tensor], n_heads: int, bf16: bool,
                      same_qkv: bool, need_dsprev: bool) -> List[Tensor]:
    wq, wk, wv, wo, n1w, n1b, n2w, n2b, f1w, f1b, f2w, f2b, ga, gb, gc = params
    qp, kvp, o, stat, x, h1, st1, f1, f2, st2 = saved
    B, Lq, d = q.shape
    Lk = kv.shape[1]
    dff = f1w.shape[0]
    dt, dev = _act_dtype(bf16), q.device
    q = q.contiguous()
    kv = q if same_qkv else kv.contiguous()
    dh2 = dh2.contiguous()
    # one zero-filled buffer for every "+=" output of the block (LN params, biases, dc)
    # ... and, unless they have a data-parallel bucket slot, for the five weight gradients
    in_z = _dest(wq) is None
    n_small, w_sizes = _zbuf_layout(d, dff)
    zd = _zbuf_dest.get(wq.data_ptr()) if (grad_dest_enabled and grad_dest_zeroed) else None
    in_bucket = zd is not None and zd[2] == n_small + sum(w_sizes) and _claim(wq)
    if in_bucket:
        zbuf = zd[0][zd[1]:zd[1] + zd[2]]     # a bucket region, zero-filled by the reducer
        in_z = True
    else:
        zbuf = torch.zeros(n_small + (sum(w_sizes) if in_z else 0), dtype=F32, device=dev)
    z_dp2, z_dp1 = zbuf[:1 + 2 * d], zbuf[1 + 2 * d:2 * (1 + 2 * d)]
    zo = 2 * (1 + 2 * d)
    # LN2: h2 = LN(h1 + b*f2)
    # (the column sums of df2 = the bias gradient of the second FFN layer come out of the same pass)
    db_f2 = zbuf[zo:zo + d]
    dh1, df2, dp2 = _add_ln_bwd(bf16, dh2, h1, f2, gb, n2w, None, st2, False, True, dpar=z_dp2,
                                dxsum=db_f2)
    # FFN backward; df1 = (df2 W2) * (f1 > 0) fused in the GEMM epilogue
    df1 = torch.empty(B, Lq, dff, dtype=dt, device=dev)
    _linear_bwd_x(bf16, df2, _weight(bf16, f2w), df1.view(-1, dff), relu_src=f1.view(-1, dff))
    if in_z:
        dw_q, dw_kv, dw_o, dw_f1, dw_f2 = _zbuf_weights(zbuf, d, dff)
        r_q = r_kv = r_o = r_f1 = r_f2 = None       # returned through zbuf
    else:
        dw_f2, r_f2 = _wgrad(f2w, d, dff)
    _linear_bwd_x(bf16, df1, _weight(bf16, f1w), dh1.view(-1, d), accumulate=True)
    if not in_z:
        dw_f1, r_f1 = _wgrad(f1w, dff, d)
    db_f1 = zbuf[zo + d:zo + d + dff]
    _rowsum(bf16, df1.view(-1, dff), db_f1)
    # LN1: h1 = LN(q + a*x)
    dq, dx, dp1 = _add_ln_bwd(bf16, dh1, q, x, ga, n1w, None, st1, False, True, dpar=z_dp1)
    # output projection
    do = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_bwd_x(bf16, dx, _weight(bf16, wo), do.view(-1, d))
    if not in_z:
        dw_o, r_o = _wgrad(wo, d, d)
    # attention core
    dqp = torch.empty(B, Lq, d, dtype=dt, device=dev)
    dkvp = torch.empty(B, Lk, 2 * d, dtype=dt, device=dev)
    ds_prev, _ = _attn_bwd(bf16, do, qp, kvp[..., :d], kvp[..., d:], mask, s, s_prev, gc, ds_next,
                           o, stat, n_heads, dqp, dkvp[..., :d], dkvp[..., d:], need_dsprev,
                           dc_out=zbuf[zo + d + dff:zo + d + dff + 1])
    # projections: dq += dqp Wq ; dkv = dkvp [Wk;Wv]
    all_dest = False
    if not in_z:
        dw_q, r_q = _wgrad(wq, d, d)
        dw_kv = _dest_pair(wk, wv)
        if dw_kv is not None and not _claim(wk, wv):
            dw_kv = None
        r_kv = torch.empty(0, device=dev)
        if dw_kv is None:
            dw_kv = r_kv = torch.empty(2 * d, d, dtype=F32, device=dev)
        # every gradient of the group goes to a (zero-filled) bucket slot?
        all_dest = not any(r.numel() for r in (r_q, r_kv, r_o, r_f1, r_f2))
    if same_qkv:   # both input gradients accumulate into dq: keep them ordered
        _linear_bwd_x(bf16, dqp, _weight(bf16, wq), dq.view(-1, d), accumulate=True)
        _linear_bwd_x(bf16, dkvp, _weight(bf16, wk, wv), dq.view(-1, d), accumulate=True)
        dkv = torch.empty(0, device=dev)
    else:
        dkv = torch.empty(B, Lk, d, dtype=dt, device=dev)
        _linear_bwd_x_group(bf16, [(dqp, _weight(bf16, wq), dq.view(-1, d), True),
                                   (dkvp, _weight(bf16, wk, wv), dkv.view(-1, d), False)])
    # the five weight gradients of the block: one grouped launch (each alone fills < 1/4 of the GPU)
    _linear_bwd_w_group(bf16, [(df2.view(-1, d), f1.view(-1, dff), dw_f2),
                               (df1.view(-1, dff), h1.view(-1, d), dw_f1),
                               (dx.view(-1, d), o.view(-1, d), dw_o),
                               (dqp.view(-1, d), q.view(-1, d), dw_q),
                               (dkvp.view(-1, 2 * d), kv.view(-1, d), dw_kv)],
                        zeroed=in_z or (grad_dest_zeroed and all_dest))
    # (dc lives in zbuf; its output slot stays an empty placeholder — like the weight gradients
    # when they live in zbuf or in a bucket slot)
    ph = [torch.empty(0, device=dev) if r is None else r for r in (r_q, r_kv, r_o, r_f1, r_f2)]
    # (a zbuf that is a bucket region is not returned either: custom ops must not return views of
    # tensors they do not own; the autograd wrapper looks the region up again)
    return [dq, dkv, ds_prev if ds_prev is not None else torch.empty(0, device=dev),
            torch.empty(0, device=dev), ph[0], ph[1], ph[2],
            torch.empty(0, device=dev) if in_bucket else zbuf, ph[3], ph[4]]


def _block_full_setup(ctx, inputs, output):
    q, kv, mask, s_prev, params, H, bf16, emit_s, same_qkv = inputs
    h2, s = output[0], output[1]
    ctx.save_for_backward(q, kv, mask, s_prev, s if s.numel() else None, *output[2:], *params)
    ctx.cfg = (H, bf16, same_qkv, len(output) - 2)
    ctx.set_materialize_grads(False)


def _block_full_backward(ctx, grads):
    dh2, ds_next = grads[0], grads[1]
    H, bf16, same_qkv, nsaved = ctx.cfg
    t = ctx.saved_tensors
    q, kv, mask, s_prev, s = t[:5]
    saved, params = t[5:5 + nsaved], t[5 + nsaved:]
    if dh2 is None:
        dh2 = torch.zeros_like(q)
    d = q.shape[-1]
    need_dsprev = s_prev is not None and ctx.needs_input_grad[3]
    (dq, dkv, ds_prev, dc, dw_q, dw_kv, dw_o, zbuf, dw_f1,
     dw_f2) = block_full_bwd_op(dh2, ds_next, q, kv, mask, s_prev, s, saved, params, H, bf16,
                                same_qkv, need_dsprev)
    # the small "+=" outputs share one zero-initialised buffer: [dp2 | dp1 | db_f2 | db_f1 | dc]
    dff = params[8].shape[0]
    if zbuf.numel() == 0:                 # the block's zero buffer is a data-parallel bucket region
        zd = _zbuf_dest[params[0].data_ptr()]
        zbuf = zd[0][zd[1]:zd[1] + zd[2]]
    dp2, dp1 = zbuf[:1 + 2 * d], zbuf[1 + 2 * d:2 * (1 + 2 * d)]
    zo = 2 * (1 + 2 * d)
    db_f2, db_f1 = zbuf[zo:zo + d], zbuf[zo + d:zo + d + dff]
    dc = zbuf[zo + d + dff:zo + d + dff + 1]
    has_prev = s_prev is not None
    # weight gradients written straight into DP bucket slots come back as empty placeholders
    wq, wk, wv, wo, f1w, f2w = params[0], params[1], params[2], params[3], params[8], params[10]
    if zbuf.numel() > _zbuf_layout(d, dff)[0]:        # weight gradients inside the zero buffer
        dw_q, dw_kv, dw_o, dw_f1, dw_f2 = _zbuf_weights(zbuf, d, dff)
    else:
        dw_q = dw_q if dw_q.numel() else _dest(wq)
        dw_o = dw_o if dw_o.numel() else _dest(wo)
        dw_f1 = dw_f1 if dw_f1.numel() else _dest(f1w)
        dw_f2 = dw_f2 if dw_f2.numel() else _dest(f2w)
        if not dw_kv.numel():
            dw_kv = _dest_pair(wk, wv)
    pgrads = [dw_q, dw_kv[:d], dw_kv[d:], dw_o, dp1[1:1 + d], dp1[1 + d:], dp2[1:1 + d],
              dp2[1 + d:], dw_f1, db_f1, dw_f2, db_f2, dp1[0:1], dp2[0:1],
              dc if has_prev else None]
    return (dq, None if same_qkv else dkv, None, ds_prev if need_dsprev else None, pgrads,
            None, None, None, None)


block_full_op.register_autograd(_block_full_backward, setup_context=_block_full_setup)


# ------------------------------------------------------------------------------------------------
# mmemo::block_lite — concat -> minus -> LN block  (cmu-mosei/run.py:236-262, Ren-MME/run.py:188-214)
# params order: wo(proj) wm(minus, (d,2d)) nw nb c
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::block_lite", mutates_args=())
def block_lite_op(q: Tensor, kv: Tensor, mask: Optional[Tensor], s_prev: Optional[Tensor],
                  params: Sequence[Tensor], n_heads: int, bf16: bool, emit_s: bool) -> List[Tensor]:
    _need_cuda(q, kv)
    wo, wm, nw, nb, gc = params
    B, Lq, d = q.shape
    dt, dev = _act_dtype(bf16), q.device
    q, kv = q.contiguous(), kv.contiguous()
    o, s, stat = _attn_fwd(bf16, q, kv, kv, mask, s_prev, gc, n_heads, emit_s)
    x = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_fwd(bf16, o, _weight(bf16, wo), None, None, x.view(-1, d))
    # y = [q | x] Wm^T as two accumulating GEMMs over the halves of Wm (no concat copy)
    wms = _weight(bf16, wm)
    y = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_fwd(bf16, q, wms, None, None, y.view(-1, d), ldw=2 * d, N=d, K=d)
    _linear_fwd(bf16, x, wms[:, d:], None, None, y.view(-1, d), accumulate=True, ldw=2 * d, N=d, K=d)
    out, st = _add_ln_fwd(bf16, None, y, None, nw, nb)
    e = torch.empty(0, device=dev)
    return [out, s if s is not None else e, o, stat, x, y, st]


@torch.library.custom_op("mmemo::block_lite_bwd", mutates_args=())
def block_lite_bwd_op(dout: Tensor, ds_next: Optional[Tensor], q: Tensor, kv: Tensor,
                      mask: Optional[Tensor], s_prev: Optional[Tensor], s: Optional[Tensor],
                      saved: Sequence[Tensor], params: Sequence[Tensor], n_heads: int, bf16: bool,
                      need_dsprev: bool) -> List[Tensor]:
    wo, wm, nw, nb, gc = params
    o, stat, x, y, st = saved
    B, Lq, d = q.shape
    Lk = kv.shape[1]
    dt, dev = _act_dtype(bf16), q.device
    q, kv = q.contiguous(), kv.contiguous()
    _, dy, dpn = _add_ln_bwd(bf16, dout.contiguous(), None, y, None, nw, None, st, False, False)
    wms = _weight(bf16, wm)
    # minus: dq = dy Wm[:, :d], dx = dy Wm[:, d:], dWm = dy^T [q | x]
    dq = torch.empty(B, Lq, d, dtype=dt, device=dev)
    dx = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_bwd_x(bf16, dy, wms, dq.view(-1, d), ldw=2 * d, N=d, K=d)
    _linear_bwd_x(bf16, dy, wms[:, d:], dx.view(-1, d), ldw=2 * d, N=d, K=d)
    dwm = torch.empty(d, 2 * d, dtype=F32, device=dev)
    _linear_bwd_w(bf16, dy.view(-1, d), q.view(-1, d), dwm)
    _linear_bwd_w(bf16, dy.view(-1, d), x.view(-1, d), dwm[:, d:])
    # proj
    do = torch.empty(B, Lq, d, dtype=dt, device=dev)
    _linear_bwd_x(bf16, dx, _weight(bf16, wo), do.view(-1, d))
    dwo = torch.empty(d, d, dtype=F32, device=dev)
    _linear_bwd_w(bf16, dx.view(-1, d), o.view(-1, d), dwo)
    # attention core (Q = q, K = V = kv): dk and dv both flow into kv
    dqa = torch.empty(B, Lq, d, dtype=dt, device=dev)
    dk = torch.empty(B, Lk, d, dtype=dt, device=dev)
    dv = torch.empty(B, Lk, d, dtype=dt, device=dev)
    ds_prev, dc = _attn_bwd(bf16, do, q, kv, kv, mask, s, s_prev, gc, ds_next, o, stat, n_heads,
                            dqa, dk, dv, need_dsprev)
    dq = dq + dqa
    dkv = dk + dv
    return [dq, dkv, ds_prev if ds_prev is not None else torch.empty(0, device=dev),
            dc if dc is not None else torch.empty(0, device=dev), dwo, dwm, dpn]


def _block_lite_setup(ctx, inputs, output):
    q, kv, mask, s_prev, params, H, bf16, emit_s = inputs
    s = output[1]
    ctx.save_for_backward(q, kv, mask, s_prev, s if s.numel() else None, *output[2:], *params)
    ctx.cfg = (H, bf16, len(output) - 2)
    ctx.set_materialize_grads(False)


def _block_lite_backward(ctx, grads):
    dout, ds_next = grads[0], grads[1]
    H, bf16, nsaved = ctx.cfg
    t = ctx.saved_tensors
    q, kv, mask, s_prev, s = t[:5]
    saved, params = t[5:5 + nsaved], t[5 + nsaved:]
    if dout is None:
        dout = torch.zeros_like(q)
    d = q.shape[-1]
    need_dsprev = s_prev is not None and ctx.needs_input_grad[3]
    dq, dkv, ds_prev, dc, dwo, dwm, dpn = block_lite_bwd_op(dout, ds_next, q, kv, mask, s_prev, s,
                                                            saved, params, H, bf16, need_dsprev)
    has_prev = s_prev is not None
    pgrads = [dwo, dwm, dpn[1:1 + d], dpn[1 + d:], dc if has_prev else None]
    return (dq, dkv, None, ds_prev if need_dsprev else None, pgrads, None, None, None)


block_lite_op.register_autograd(_block_lite_backward, setup_context=_block_lite_setup)


# ------------------------------------------------------------------------------------------------
# mmemo::pool  — concat + mean||max pooling over a segment table (float32 output)
# segs: n_groups*n_slots tensors, group-major; group g tensors are (B, L_g, d)
# ------------------------------------------------------------------------------------------------
def _seg_table(segs: Sequence[Tensor], n_groups: int):
    n_slots = len(segs) // n_groups
    ptrs = (C.c_void_p * len(segs))(*[t.data_ptr() for t in segs])
    lens = (C.c_int64 * n_groups)(*[segs[g * n_slots].shape[1] for g in range(n_groups)])
    return ptrs, lens, n_slots


@torch.library.custom_op("mmemo::pool", mutates_args=())
def pool_op(segs: Sequence[Tensor], n_groups: int) -> Tuple[Tensor, Tensor]:
    _need_cuda(*segs)
    segs = [t.contiguous() for t in segs]
    ptrs, lens, n_slots = _seg_table(segs, n_groups)
    B, _, d = segs[0].shape
    bf16 = segs[0].dtype == BF
    out = torch.empty(B, 2 * n_slots * d, dtype=F32, device=segs[0].device)
    amax = torch.empty(B, n_slots * d, dtype=torch.int32, device=segs[0].device)
    _call(f"mmemo_pool_fwd_{_sfx(bf16)}", ptrs, lens, n_groups, n_slots, B, d, out.data_ptr(),
          amax.data_ptr(), _stream())
    return out, amax


@torch.library.custom_op("mmemo::pool_bwd", mutates_args=())
def pool_bwd_op(dout: Tensor, amax: Tensor, like: Sequence[Tensor], n_groups: int) -> List[Tensor]:
    dsegs = [torch.empty(t.shape, dtype=t.dtype, device=t.device) for t in like]
    ptrs, lens, n_slots = _seg_table(dsegs, n_groups)
    B, _, d = dsegs[0].shape
    _call(f"mmemo_pool_bwd_{_sfx(dsegs[0].dtype == BF)}", dout.contiguous().data_ptr(),
          amax.data_ptr(), ptrs, lens, n_groups, n_slots, B, d, _stream())
    return dsegs


def _pool_setup(ctx, inputs, output):
    segs, n_groups = inputs
    ctx.save_for_backward(output[1], *segs)
    ctx.n_groups = n_groups
    ctx.set_materialize_grads(False)


def _pool_backward(ctx, dout, _damax):
    amax, *segs = ctx.saved_tensors
    return pool_bwd_op(dout, amax, segs, ctx.n_groups), None


pool_op.register_autograd(_pool_backward, setup_context=_pool_setup)


def pool(segs: Sequence[Tensor], n_groups: int = 3) -> Tensor:
    return pool_op(list(segs), n_groups)[0]


# ------------------------------------------------------------------------------------------------
# heads and losses (float32)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::state_transfer", mutates_args=())
def state_transfer_op(feats: Tensor, trans: Tensor) -> Tensor:
    _need_cuda(feats)
    feats, trans = feats.float().contiguous(), trans.contiguous()
    B, P, C2 = feats.shape
    out = torch.empty(B, P, C2 // 2, dtype=F32, device=feats.device)
    _call("mmemo_state_transfer_fwd", feats.data_ptr(), trans.data_ptr(), out.data_ptr(), B, P,
          C2 // 2, _stream())
    return out


@torch.library.custom_op("mmemo::state_transfer_bwd", mutates_args=())
def state_transfer_bwd_op(dout: Tensor, feats: Tensor, trans: Tensor, out: Tensor) -> List[Tensor]:
    feats, trans = feats.float().contiguous(), trans.contiguous()
    B, P, C2 = feats.shape
    dfeats = torch.empty_like(feats)
    dtrans = torch.zeros_like(trans)
    _call("mmemo_state_transfer_bwd", dout.contiguous().data_ptr(), feats.data_ptr(),
          trans.data_ptr(), out.data_ptr(), dfeats.data_ptr(), dtrans.data_ptr(), B, P, C2 // 2,
          _stream())
    return [dfeats, dtrans]


def _st_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], output)


def _st_backward(ctx, dout):
    feats, trans, out = ctx.saved_tensors
    dfeats, dtrans = state_transfer_bwd_op(dout, feats, trans, out)
    return dfeats.to(feats.dtype), dtrans


state_transfer_op.register_autograd(_st_backward, setup_context=_st_setup)


@torch.library.custom_op("mmemo::bilinear_head", mutates_args=())
def bilinear_head_op(this_feat: Tensor, last_feat: Tensor, trans: Tensor, gamma: Tensor,
                     beta: Tensor, w: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    _need_cuda(this_feat)
    th, la = this_feat.float().contiguous(), last_feat.float().contiguous()
    B, Cn = th.shape
    out = torch.empty(B, Cn, dtype=F32, device=th.device)
    z = torch.empty(B, Cn, dtype=F32, device=th.device)
    _call("mmemo_bilinear_head_fwd", th.data_ptr(), la.data_ptr(), trans.contiguous().data_ptr(),
          gamma.data_ptr(), beta.data_ptr(), w.contiguous().data_ptr(), bias.data_ptr(),
          out.data_ptr(), z.data_ptr(), B, Cn, LN_EPS, _stream())
    return out, z


@torch.library.custom_op("mmemo::bilinear_head_bwd", mutates_args=())
def bilinear_head_bwd_op(dout: Tensor, this_feat: Tensor, last_feat: Tensor, trans: Tensor,
                         gamma: Tensor, beta: Tensor, w: Tensor, z: Tensor) -> List[Tensor]:
    th, la = this_feat.float().contiguous(), last_feat.float().contiguous()
    B, Cn = th.shape
    dev = th.device
    dth, dla = torch.empty_like(th), torch.empty_like(la)
    dT = torch.zeros(Cn, Cn, Cn, dtype=F32, device=dev)
    dg, db = torch.zeros(Cn, dtype=F32, device=dev), torch.zeros(Cn, dtype=F32, device=dev)
    dw = torch.zeros(Cn, 2 * Cn, dtype=F32, device=dev)
    dbias = torch.zeros(Cn, dtype=F32, device=dev)
    _call("mmemo_bilinear_head_bwd", dout.contiguous().data_ptr(), th.data_ptr(), la.data_ptr(),
          trans.contiguous().data_ptr(), gamma.data_ptr(), beta.data_ptr(),
          w.contiguous().data_ptr(), z.data_ptr(), dth.data_ptr(), dla.data_ptr(), dT.data_ptr(),
          dg.data_ptr(), db.data_ptr(), dw.data_ptr(), dbias.data_ptr(), B, Cn, LN_EPS, _stream())
    return [dth, dla, dT, dg, db, dw, dbias]


def _bh_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs[:6], output[1])
    ctx.set_materialize_grads(False)


def _bh_backward(ctx, dout, _dz):
    th, la, trans, gamma, beta, w, z = ctx.saved_tensors
    dth, dla, dT, dg, db, dw, dbias = bilinear_head_bwd_op(dout, th, la, trans, gamma, beta, w, z)
    return dth.to(th.dtype), dla.to(la.dtype), dT, dg, db, dw, dbias


bilinear_head_op.register_autograd(_bh_backward, setup_context=_bh_setup)


def bilinear_head(this_feat, last_feat, trans, gamma, beta, w, bias):
    return bilinear_head_op(this_feat, last_feat, trans, gamma, beta, w, bias)[0]


@torch.library.custom_op("mmemo::circle_loss", mutates_args=())
def circle_loss_op(logits: Tensor, labels: Tensor) -> Tensor:
    _need_cuda(logits)
    C_ = logits.shape[-1]
    lg = logits.float().contiguous()
    lb = labels.to(F32).contiguous()
    loss = torch.empty(logits.shape[:-1], dtype=F32, device=logits.device)
    _call("mmemo_circle_loss_fwd", lg.data_ptr(), lb.data_ptr(), loss.data_ptr(),
          lg.numel() // C_, C_, _stream())
    return loss


@torch.library.custom_op("mmemo::circle_loss_bwd", mutates_args=())
def circle_loss_bwd_op(dloss: Tensor, logits: Tensor, labels: Tensor) -> Tensor:
    C_ = logits.shape[-1]
    lg = logits.float().contiguous()
    lb = labels.to(F32).contiguous()
    dl = torch.empty_like(lg)
    _call("mmemo_circle_loss_bwd", dloss.to(F32).contiguous().data_ptr(), lg.data_ptr(),
          lb.data_ptr(), dl.data_ptr(), lg.numel() // C_, C_, _stream())
    return dl


def _cl_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _cl_backward(ctx, dloss):
    logits, labels = ctx.saved_tensors
    return circle_loss_bwd_op(dloss, logits, labels).to(logits.dtype), None


circle_loss_op.register_autograd(_cl_backward, setup_context=_cl_setup)


@torch.library.custom_op("mmemo::rdrop_kl", mutates_args=())
def rdrop_kl_op(logits: Tensor) -> Tensor:
    _need_cuda(logits)
    lg = logits.float().contiguous()
    out = torch.empty((), dtype=F32, device=logits.device)
    _call("mmemo_rdrop_kl_fwd", lg.data_ptr(), out.data_ptr(), lg.shape[0], lg.shape[1], _stream())
    return out


@torch.library.custom_op("mmemo::rdrop_kl_bwd", mutates_args=())
def rdrop_kl_bwd_op(dout: Tensor, logits: Tensor) -> Tensor:
    lg = logits.float().contiguous()
    dl = torch.empty_like(lg)
    _call("mmemo_rdrop_kl_bwd", dout.to(F32).contiguous().data_ptr(), lg.data_ptr(), dl.data_ptr(),
          lg.shape[0], lg.shape[1], _stream())
    return dl


def _rk_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])


def _rk_backward(ctx, dout):
    (logits,) = ctx.saved_tensors
    return rdrop_kl_bwd_op(dout, logits).to(logits.dtype)


rdrop_kl_op.register_autograd(_rk_backward, setup_context=_rk_setup)


# ------------------------------------------------------------------------------------------------
# mean(x^2) (the encoder benchmark's synthetic loss) — one launch per direction
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::sq_mean", mutates_args=())
def sq_mean_op(x: Tensor) -> Tensor:
    _need_cuda(x)
    x = x.contiguous()
    out = torch.zeros((), dtype=F32, device=x.device)
    _call(f"mmemo_sqmean_fwd_{_sfx(x.dtype == BF)}", x.data_ptr(), x.numel(), out.data_ptr(),
          _stream())
    return out


@torch.library.custom_op("mmemo::sq_mean_bwd", mutates_args=())
def sq_mean_bwd_op(dloss: Tensor, x: Tensor) -> Tensor:
    x = x.contiguous()
    dx = torch.empty_like(x)
    _call(f"mmemo_sqmean_bwd_{_sfx(x.dtype == BF)}", x.data_ptr(),
          dloss.to(F32).contiguous().data_ptr(), x.numel(), dx.data_ptr(), _stream())
    return dx


def _sq_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])


def _sq_backward(ctx, dloss):
    (x,) = ctx.saved_tensors
    return sq_mean_bwd_op(dloss, x)


sq_mean_op.register_autograd(_sq_backward, setup_context=_sq_setup)


def cast_bf16(x: Tensor) -> Tensor:
    """float32 -> bf16 copy of a tensor that needs no gradient (raw inputs), libmmemo's vector
    cast instead of ATen's generic copy kernel."""
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=BF, device=x.device)
    _call("mmemo_cast_f32_to_bf16", x.data_ptr(), y.data_ptr(), x.numel(), _stream())
    return y


# ------------------------------------------------------------------------------------------------
# dropout (counter-based; forward and backward are the same kernel with the same seed)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("mmemo::dropout", mutates_args=())
def dropout_op(x: Tensor, p: float, seed: int) -> Tensor:
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    _call(f"mmemo_dropout_{_sfx(x.dtype == BF)}", x.data_ptr(), y.data_ptr(), x.numel(), p,
          seed & 0xFFFFFFFFFFFFFFFF, _stream())
    return y


def _do_setup(ctx, inputs, output):
    ctx.p, ctx.seed = inputs[1], inputs[2]


def _do_backward(ctx, dy):
    return dropout_op(dy, ctx.p, ctx.seed), None, None


dropout_op.register_autograd(_do_backward, setup_context=_do_setup)

_dropout_counter = [0]
_dropout_step = {}     # device -> int64 scalar mixed into the grouped-dropout seeds at run time


def next_dropout_seed() -> int:
    """A fresh 63-bit seed per call, derived from torch's seed (``torch.manual_seed`` controls it)."""
    _dropout_counter[0] += 1
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _dropout_counter[0] * 0xD1B54A32D192ED03)
    return seed & 0x7FFFFFFFFFFFFFFF


def dropout(x: Tensor, p: float, training: bool) -> Tensor:
    if not training or p <= 0.0:
        return x
    return dropout_op(x, float(p), next_dropout_seed())


def dropout_step(dev) -> Tensor:
    """Device scalar added to every grouped-dropout seed when the kernel RUNS.  Python-side seeds
    are frozen when a training step is captured into a CUDA graph; a captured ``advance_dropout_step``
    (one increment per replay) makes every replay draw new masks."""
    dev = torch.device(dev)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    t = _dropout_step.get(dev)
    if t is None:
        t = _dropout_step[dev] = torch.zeros(1, dtype=torch.int64, device=dev)
    return t


def advance_dropout_step(dev="cuda") -> None:
    dropout_step(dev).add_(1)


def site_seeds(seed: int, G: int, site: int, n_sites: int = 2) -> List[int]:
    """One seed per (problem, dropout site) of a grouped layer."""
    return [(seed + (n_sites * g + site + 1) * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF
            for g in range(G)]


def _dropout_group(xs, ys, p: float, seeds) -> None:
    """ys[i] = dropout(xs[i]) for up to 64 tensors per launch; xs[i] is ys[i] = in place.  Forward
    and backward use the same seeds (same masks)."""
    if p <= 0.0 or not xs:
        return
    bf16 = xs[0].dtype == BF
    step = dropout_step(xs[0].device)
    for i in range(0, len(xs), 64):
        px, py, ps = xs[i:i + 64], ys[i:i + 64], seeds[i:i + 64]
        assert all(t.is_contiguous() for t in px) and all(t.is_contiguous() for t in py)
        _call(f"mmemo_dropout_multi_{_sfx(bf16)}", len(px),
              _arr(C.c_void_p, [t.data_ptr() for t in px]), _arr(C.c_void_p, [t.data_ptr() for t in py]),
              _arr(C.c_int64, [t.numel() for t in px]), _arr(C.c_uint64, list(ps)), float(p),
              step.data_ptr(), _stream())
