"""Drop-in modules for the reference script ``Ren-MME/run.py`` (3-modal Chinese TV data, 9 labels,
lite blocks, shared LayerNorm after projection, R-Drop consistency term)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .blocks import LiteAttentionBlock, as_mask, fusion_trunk_multi, is_bf16

# reference module constants (Ren-MME/run.py:21-39)
L_LEN, V_LEN, A_LEN = 40, 76, 275
L_DIM, V_DIM, A_DIM = 768, 640, 205
DIM, DROP, FFN, N_HEADS, N_LAYERS = 128, 0.1, 1, 8, 1


class Unify_Dimension(nn.Module):
    """Ren-MME/run.py:158-166: three bias-free linears followed by ONE shared LayerNorm."""

    def __init__(self, dim, l_dim=None, v_dim=None, a_dim=None):
        super().__init__()
        self.linguistic = nn.Linear(L_DIM if l_dim is None else l_dim, dim, bias=False)
        self.visual = nn.Linear(V_DIM if v_dim is None else v_dim, dim, bias=False)
        self.acoustic = nn.Linear(A_DIM if a_dim is None else a_dim, dim, bias=False)
        self.norm1 = nn.LayerNorm(dim)

    def forward(self, l, v, a):
        from .group_ops import project
        w, b = self.norm1.weight, self.norm1.bias
        ys = project([l, v, a], [self.linguistic.weight, self.visual.weight, self.acoustic.weight],
                     bf16=is_bf16())
        return tuple(ops.add_ln(None, y, None, w, b) for y in ys)


class Attention_Block(LiteAttentionBlock):
    """Ren-MME/run.py:169-214 (LayerNorm named ``norm2``, dropout from the module global)."""

    def __init__(self, dim, n_heads, ffn):
        super().__init__(dim, n_heads, ffn, DROP, norm_name="norm2")


class Multi_ATTN(nn.Module):
    """Ren-MME/run.py:217-271."""
    N_CLS = 9

    def __init__(self, dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim=None, v_dim=None,
                 a_dim=None):
        super().__init__()
        self.unify_dimension = Unify_Dimension(dim, l_dim, v_dim, a_dim)
        self.n_layers = n_layers
        self.multimodal_blocks = nn.ModuleList([Attention_Block(dim, n_heads, ffn)
                                                for _ in range(9 * n_layers)])
        self.classifier = nn.Linear(dim * 6 * n_layers, self.N_CLS, bias=False)

    def _tower(self, l, v, a, l_mask, v_mask, a_mask):
        """(blocks, projected features, masks) - the input of ``fusion_trunk_multi``."""
        l, v, a = self.unify_dimension(l, v, a)
        return (self.multimodal_blocks, {"l": l, "v": v, "a": a},
                {"l": as_mask(l_mask), "v": as_mask(v_mask), "a": as_mask(a_mask)})

    def forward(self, l, v, a, l_mask, v_mask, a_mask):
        x = fusion_trunk_multi([self._tower(l, v, a, l_mask, v_mask, a_mask)], self.n_layers,
                               keep_all=True)[0]
        return ops.linear(x, self.classifier.weight)


class Base_model(nn.Module):
    """Ren-MME/run.py:273-292.  Twelve positional tensors in the reference's order."""

    def __init__(self, dim=DIM, l_len=L_LEN, v_len=V_LEN, a_len=A_LEN, n_heads=N_HEADS,
                 n_layers=N_LAYERS, ffn=FFN, l_dim=None, v_dim=None, a_dim=None):
        super().__init__()
        self.intensity = Multi_ATTN(dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim, v_dim,
                                    a_dim)
        self.stimulation = Multi_ATTN(dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim,
                                      v_dim, a_dim)
        self.trans = nn.Parameter(torch.rand(9, 9, 9), requires_grad=True)
        self.norm3 = nn.LayerNorm(9)
        self.out = nn.Linear(18, 9)

    def forward(self, pre_text_feat, pre_text_mask, pro_text_feat, pro_text_mask, pre_video_feat,
                pre_video_mask, pro_video_feat, pro_video_mask, pre_audio_feat, pre_audio_mask,
                pro_audio_feat, pro_audio_mask):
        # both towers' trunks as ONE group per layer (they are independent until the bilinear head)
        ti, ts = self.intensity, self.stimulation
        pi, ps = fusion_trunk_multi(
            [ti._tower(pre_text_feat, pre_video_feat, pre_audio_feat, pre_text_mask, pre_video_mask,
                       pre_audio_mask),
             ts._tower(pro_text_feat, pro_video_feat, pro_audio_feat, pro_text_mask, pro_video_mask,
                       pro_audio_mask)], ti.n_layers, keep_all=True)
        last_feat = ops.linear(pi, ti.classifier.weight)
        this_feat = ops.linear(ps, ts.classifier.weight)
        return ops.bilinear_head(this_feat, last_feat, self.trans, self.norm3.weight,
                                 self.norm3.bias, self.out.weight, self.out.bias)


def multi_loss(y_pred, y_true):
    """Ren-MME/run.py:295-304: circle loss reduced with .mean()."""
    return ops.circle_loss_op(y_pred, y_true).mean()


def rdrop_kl(logits):
    """Ren-MME/run.py:332-334: (KL(even||odd) + KL(odd||even)) / 2, batchmean over B/2 (inline code
    in the reference's train(); exposed as a function here)."""
    return ops.rdrop_kl_op(logits)
