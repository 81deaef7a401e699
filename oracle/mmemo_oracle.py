"""CPU oracle for the multimodal-fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch on CPU, weights passed as a flat
``dict[str, Tensor]`` keyed by the reference's ``state_dict`` names) of the algorithm that the
reference implements inside its ``nn.Module`` classes.  It exists to *check* the CUDA path; the
product never imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``gpu_eager_baseline`` / ``--impl reference`` legs (the baselines the product is
measured AGAINST) may import it.

Pinning: the reference ships no tests or golden vectors for this path ("parity unpinned" by the
reference itself).  The oracle is therefore pinned against *outputs of the reference's own classes*
executed in the build container (``oracle/refload.py`` extracts them from ``/root/reference`` with
``ast``; ``tests/test_oracle.py`` compares, ``tests/golden/make_golden.py`` freezes the
reference outputs into ``tests/golden/*.pt`` so the comparison travels to boxes without
``/root/reference``).

Every function cites the reference lines it restates (paths relative to ``/root/reference``).
All arithmetic runs in the dtype of the inputs (fp32 / fp64 / bf16), like the reference does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

P = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------
# kernel (a): residual attention core
# --------------------------------------------------------------------------------------------


def _heads(x: torch.Tensor, n_heads: int) -> torch.Tensor:
    """(B, L, d) -> (B, H, L, hd).  others/realformer.py:172-177,189 (split_last + transpose)."""
    b, l, d = x.shape
    return x.reshape(b, l, n_heads, d // n_heads).permute(0, 2, 1, 3)


def resattn_core(q, k, v, mask, n_heads: int, c=None, s_prev=None):
    """Residual-attention core on *already projected* q/k/v.

    others/realformer.py:189-203, cmu-mosei/run.py:242-256, Ren-MME/run.py:194-208,
    robot_demo.py:354-368.  Order of operations is the reference's: scale, add ``c*S_prev``,
    subtract ``1e8*(1-mask)`` (in the working dtype), softmax, PV, merge heads.
    Returns (O (B,Lq,d), S (B,H,Lq,Lk) post-mask pre-softmax).
    """
    qh, kh, vh = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    hd = kh.shape[-1]
    s = qh @ kh.transpose(-2, -1) / math.sqrt(hd)
    if s_prev is not None:
        s = s + c * s_prev
    if mask is not None:
        if mask.dim() == 2:
            m = mask[:, None, None, :]
        else:  # (B, Lq, Lk) branch, others/realformer.py:197-199 (no caller uses it)
            m = mask[:, None, :, :]
        s = s - 1.0e8 * (1.0 - m)
    att = torch.softmax(s, dim=-1)
    o = (att @ vh).permute(0, 2, 1, 3)
    o = o.reshape(o.shape[0], o.shape[1], -1)
    return o, s


def _ln(x, w, b, eps: float = 1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def block_full(p: P, pre: str, q, k, v, mask, n_heads: int, s_prev=None):
    """Full RealFormer block.  others/realformer.py:182-209 == robot_demo.py:347-374."""
    qp = q @ p[pre + "w_qkv.0.weight"].t()
    kp = k @ p[pre + "w_qkv.1.weight"].t()
    vp = v @ p[pre + "w_qkv.2.weight"].t()
    o, s = resattn_core(qp, kp, vp, mask, n_heads, p[pre + "c"], s_prev)
    x = o @ p[pre + "proj.weight"].t()
    h1 = _ln(q + p[pre + "a"] * x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
    f = torch.relu(h1 @ p[pre + "ffn.0.weight"].t() + p[pre + "ffn.0.bias"])
    f = f @ p[pre + "ffn.2.weight"].t() + p[pre + "ffn.2.bias"]
    h2 = _ln(h1 + p[pre + "b"] * f, p[pre + "norm2.weight"], p[pre + "norm2.bias"])
    return h2, s


def block_lite(p: P, pre: str, q, k, v, mask, n_heads: int, s_prev=None, norm: str = "norm1"):
    """Lite block: no QKV projection, concat -> ``minus`` -> LN.
    cmu-mosei/run.py:236-262 (norm1), Ren-MME/run.py:188-214 (norm2)."""
    o, s = resattn_core(q, k, v, mask, n_heads, p[pre + "c"], s_prev)
    x = o @ p[pre + "proj.weight"].t()
    y = torch.cat([q, x], dim=-1) @ p[pre + "minus.weight"].t()
    return _ln(y, p[pre + norm + ".weight"], p[pre + norm + ".bias"]), s


def encoder_chain(p: P, pres: List[str], x, mask, n_heads: int):
    """BASELINE config 2: ``q, s = blk(q, x, x, mask, s)`` over full blocks."""
    q, s = x, None
    for pre in pres:
        q, s = block_full(p, pre, q, x, x, mask, n_heads, s)
    return q, s


# --------------------------------------------------------------------------------------------
# kernel (b1): modality projection / position embedding
# --------------------------------------------------------------------------------------------


def _w2d(w):
    return w[..., 0] if w.dim() == 3 else w  # Conv1d(k=1) weight (d, D, 1) == Linear weight (d, D)


def unify_realformer(p: P, pre: str, l, v, a):
    """others/realformer.py:133-143 (three bias-free k=1 convs == linears)."""
    return (l @ _w2d(p[pre + "linguistic.weight"]).t(),
            v @ _w2d(p[pre + "visual.weight"]).t(),
            a @ _w2d(p[pre + "acoustic.weight"]).t())


def unify_mosei(p: P, pre: str, l, v, a):
    """cmu-mosei/run.py:207-214."""
    return unify_realformer(p, pre, l, v, a)


def unify_renmme(p: P, pre: str, l, v, a):
    """Ren-MME/run.py:158-166: three linears + ONE shared LayerNorm."""
    w, b = p[pre + "norm1.weight"], p[pre + "norm1.bias"]
    return tuple(_ln(t, w, b) for t in unify_realformer(p, pre, l, v, a))


def unify_robot(p: P, pre: str, l, v256, v512, v1024, a):
    """robot_demo.py:293-311: five biased k=1 convs; v = cat(v256, v512, v1024) projections."""
    def lin(x, name):
        return x @ _w2d(p[pre + name + ".weight"]).t() + p[pre + name + ".bias"]
    v = torch.cat([lin(v256, "visual_256"), lin(v512, "visual_512"), lin(v1024, "visual_1024")], 2)
    return lin(l, "linguistic"), v, lin(a, "acoustic")


def add_pos(p: P, name: str, x):
    """others/realformer.py:145-152,225-227; robot_demo.py:314-321,392-394: x + E[arange(L)]."""
    return x + p[name + "position_embeddings.weight"][None, : x.shape[1], :]


# --------------------------------------------------------------------------------------------
# kernel (b3): the 9-chain fusion trunk + mean||max pooling
# --------------------------------------------------------------------------------------------

# (query modality, source modality) in the reference's fixed chain order
# others/realformer.py:232-257, cmu-mosei/run.py:278-313: ll lv la vv vl va aa al av
CHAINS = [("l", "l"), ("l", "v"), ("l", "a"), ("v", "v"), ("v", "l"), ("v", "a"),
          ("a", "a"), ("a", "l"), ("a", "v")]


def trunk(p: P, pre: str, feats, masks, n_heads: int, n_layers: int, kind: str, keep_all: bool,
          norm: str = "norm1"):
    """Nine chains x n_layers blocks; concat features then positions (l, a, v); mean||max pooling
    over ALL positions (mask-unaware).  others/realformer.py:228-262; cmu-mosei/run.py:274-318;
    Ren-MME/run.py:226-270; robot_demo.py:395-439."""
    outs = {"l": [], "v": [], "a": []}
    for ci, (qm, sm) in enumerate(CHAINS):
        q, s = feats[qm], None
        for i in range(n_layers):
            bp = f"{pre}multimodal_blocks.{n_layers * ci + i}."
            if kind == "full":
                q, s = block_full(p, bp, q, feats[sm], feats[sm], masks[sm], n_heads, s)
            else:
                q, s = block_lite(p, bp, q, feats[sm], feats[sm], masks[sm], n_heads, s, norm)
            if keep_all:
                outs[qm].append(q)
        if not keep_all:
            outs[qm].append(q)
    l = torch.cat(outs["l"], 2)
    v = torch.cat(outs["v"], 2)
    a = torch.cat(outs["a"], 2)
    x = torch.cat([l, a, v], 1)
    return torch.cat([x.mean(1), x.max(1)[0]], 1)


# --------------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------------


def realformer_multi_class(p: P, pre: str, l, v, a, lm, vm, am, n_heads: int, n_layers: int):
    """others/realformer.py:223-264."""
    l, v, a = unify_realformer(p, pre + "unify_dimension.", l, v, a)
    l = add_pos(p, pre + "linguistic_position.", l)
    v = add_pos(p, pre + "visual_position.", v)
    a = add_pos(p, pre + "acoustic_position.", a)
    x = trunk(p, pre, {"l": l, "v": v, "a": a}, {"l": lm, "v": vm, "a": am}, n_heads, n_layers,
              "full", keep_all=False)
    x = x @ p[pre + "fully_connected.weight"].t() + p[pre + "fully_connected.bias"]
    return torch.relu(_ln(x, p[pre + "normalization.weight"], p[pre + "normalization.bias"]))


def state_transfer_head(feats, trans):
    """Window recurrence.  others/realformer.py:274-286.  feats (B, P, 12) -> (B, P, 6)."""
    outs, prev_o, prev_g = [], None, None
    for i in range(feats.shape[1]):
        o, g = feats[:, i].chunk(2, 1)
        if i != 0:
            alpha = torch.sigmoid(g + prev_g)
            t0 = torch.tanh(prev_o @ trans)
            o = (1 - alpha) * o + alpha * t0
        outs.append(o.unsqueeze(1))
        prev_o, prev_g = o, g
    return torch.cat(outs, 1)


def realformer_state_transfer(p: P, l, v, a, lm, vm, am, n_heads: int, n_layers: int):
    """others/realformer.py:266-286.  Inputs carry a window dim: l (B,P,L,D), masks (B,P,L).
    The P windows are independent until the (B,P,6) recurrence, so they are folded into the batch
    (fold-vs-loop max-rel-diff = 0.0)."""
    b, w = l.shape[:2]
    fold = lambda t: t.reshape(b * w, *t.shape[2:])
    f = realformer_multi_class(p, "feature.", fold(l), fold(v), fold(a), fold(lm), fold(vm),
                               fold(am), n_heads, n_layers)
    f = f @ p["classifier.weight"].t() + p["classifier.bias"]
    return state_transfer_head(f.reshape(b, w, -1), p["trans"])


def bilinear_head(this, last, trans, ln_w, ln_b, out_w, out_b):
    """z[b,k] = sum_{j,m} this[b,j] last[b,m] T[j,m,k]; out = W [this || LN(z)] + b.
    cmu-mosei/run.py:332-339, Ren-MME/run.py:285-292, rencecps/run.py:141-148 (the reference's
    per-sample python loop is this einsum)."""
    z = torch.einsum("bj,bm,jmk->bk", this, last, trans)
    return torch.cat([this, _ln(z, ln_w, ln_b)], 1) @ out_w.t() + out_b


def mosei_multi_attn(p: P, pre: str, l, v, a, lm, vm, am, n_heads: int, n_layers: int):
    """cmu-mosei/run.py:272-319."""
    l, v, a = unify_mosei(p, pre + "unify_dimension.", l, v, a)
    x = trunk(p, pre, {"l": l, "v": v, "a": a}, {"l": lm, "v": vm, "a": am}, n_heads, n_layers,
              "lite", keep_all=True, norm="norm1")
    return x @ p[pre + "classifier.weight"].t()


def mosei_concat_trans(p: P, l, v, a, lm, vm, am, n_heads: int, n_layers: int):
    """cmu-mosei/run.py:329-339.  Index 0 = previous sentence (intensity), 1 = current."""
    last = mosei_multi_attn(p, "intensity.", l[:, 0], v[:, 0], a[:, 0], lm[:, 0], vm[:, 0], am[:, 0],
                            n_heads, n_layers)
    this = mosei_multi_attn(p, "stimulation.", l[:, 1], v[:, 1], a[:, 1], lm[:, 1], vm[:, 1],
                            am[:, 1], n_heads, n_layers)
    return bilinear_head(this, last, p["trans"], p["norm1.weight"], p["norm1.bias"],
                         p["out.weight"], p["out.bias"])


def renmme_multi_attn(p: P, pre: str, l, v, a, lm, vm, am, n_heads: int, n_layers: int):
    """Ren-MME/run.py:224-271."""
    l, v, a = unify_renmme(p, pre + "unify_dimension.", l, v, a)
    x = trunk(p, pre, {"l": l, "v": v, "a": a}, {"l": lm, "v": vm, "a": am}, n_heads, n_layers,
              "lite", keep_all=True, norm="norm2")
    return x @ p[pre + "classifier.weight"].t()


def renmme_base_model(p: P, pre_t, pre_tm, pro_t, pro_tm, pre_v, pre_vm, pro_v, pro_vm,
                      pre_a, pre_am, pro_a, pro_am, n_heads: int = 8, n_layers: int = 1):
    """Ren-MME/run.py:281-292 (12 positional tensors, same order)."""
    last = renmme_multi_attn(p, "intensity.", pre_t, pre_v, pre_a, pre_tm, pre_vm, pre_am,
                             n_heads, n_layers)
    this = renmme_multi_attn(p, "stimulation.", pro_t, pro_v, pro_a, pro_tm, pro_vm, pro_am,
                             n_heads, n_layers)
    return bilinear_head(this, last, p["trans"], p["norm3.weight"], p["norm3.bias"],
                         p["out.weight"], p["out.bias"])


def rencecps_concat_linear(p: P, feat):
    """rencecps/run.py:138-148.  feat (B, 2, dim)."""
    last = feat[:, 0] @ p["intensity.weight"].t()
    this = feat[:, 1] @ p["stimulation.weight"].t()
    return bilinear_head(this, last, p["trans"], p["norm.weight"], p["norm.bias"],
                         p["out.weight"], p["out.bias"])


def robot_multi_class(p: P, l, v256, v512, v1024, a, lm, vm, am, n_heads: int, n_layers: int):
    """robot_demo.py:390-441 (fully_connected / normalization exist but are unused, :440)."""
    l, v, a = unify_robot(p, "unify_dimension.", l, v256, v512, v1024, a)
    l = add_pos(p, "linguistic_position.", l)
    v = add_pos(p, "visual_position.", v)
    a = add_pos(p, "acoustic_position.", a)
    x = trunk(p, "", {"l": l, "v": v, "a": a}, {"l": lm, "v": vm, "a": am}, n_heads, n_layers,
              "full", keep_all=True)
    return x @ p["classifier.weight"].t() + p["classifier.bias"]


# --------------------------------------------------------------------------------------------
# losses (kernel b4 / row a10)
# --------------------------------------------------------------------------------------------


def multi_circle_loss(y_pred, y_true):
    """Per-row multi-label circle / ZLPR loss.  others/realformer.py:289-298 (== cmu-mosei/run.py
    :342-351, rencecps/run.py:151-160, robot_demo.py:444-453).  Literal +-1e12 masking."""
    y_true = y_true.to(y_pred.dtype)
    s = (1 - 2 * y_true) * y_pred
    neg = s - y_true * 1e12
    pos = s - (1 - y_true) * 1e12
    z = torch.zeros_like(s[..., :1])
    return (torch.logsumexp(torch.cat([neg, z], -1), -1)
            + torch.logsumexp(torch.cat([pos, z], -1), -1))


def multi_loss(y_pred, y_true):
    """Ren-MME/run.py:295-304: same, reduced with .mean()."""
    return multi_circle_loss(y_pred, y_true).mean()


def rdrop_kl(logits):
    """Symmetric sigmoid-KL between even / odd rows.  Ren-MME/run.py:332-334 (target NOT
    detached; 'batchmean' over B/2)."""
    ev, od = logits[::2], logits[1::2]
    kl0 = F.kl_div(F.logsigmoid(ev), torch.sigmoid(od), reduction="batchmean")
    kl1 = F.kl_div(F.logsigmoid(od), torch.sigmoid(ev), reduction="batchmean")
    return (kl0 + kl1) / 2


def window_masked_loss(logits, labels, wmask):
    """others/realformer.py:311-312: (circle_loss * window_mask).mean() over B*P."""
    return (multi_circle_loss(logits, labels) * wmask.to(logits.dtype)).mean()
