"""Drop-in modules for the reference script ``cmu-mosei/run.py`` (7 labels, lite blocks, two
towers + 7x7x7 bilinear transition head)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .blocks import LiteAttentionBlock, as_mask, fusion_trunk_multi, is_bf16

# reference module constants (cmu-mosei/run.py:24-42) read inside the classes
L_DIM, V_DIM, A_DIM = 300, 35, 74
DROP = 0.0


class Unify_Dimension(nn.Module):
    """cmu-mosei/run.py:207-214.  Input widths come from module globals in the reference; they are
    keyword arguments here with the same defaults."""

    def __init__(self, dim, l_dim=None, v_dim=None, a_dim=None):
        super().__init__()
        self.linguistic = nn.Linear(L_DIM if l_dim is None else l_dim, dim, bias=False)
        self.visual = nn.Linear(V_DIM if v_dim is None else v_dim, dim, bias=False)
        self.acoustic = nn.Linear(A_DIM if a_dim is None else a_dim, dim, bias=False)

    def forward(self, l, v, a):
        from .group_ops import project
        return tuple(project([l, v, a], [self.linguistic.weight, self.visual.weight,
                                         self.acoustic.weight], bf16=is_bf16()))


class Attention_Block(LiteAttentionBlock):
    """cmu-mosei/run.py:217-262."""

    def __init__(self, dim, n_heads, ffn):
        super().__init__(dim, n_heads, ffn, DROP, norm_name="norm1")


class Multi_ATTN(nn.Module):
    """cmu-mosei/run.py:265-319."""
    N_CLS = 7

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


class Concat_Trans(nn.Module):
    """cmu-mosei/run.py:321-339.  l (B,2,L,D): index 0 = previous sentence, 1 = current."""

    def __init__(self, dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim=None, v_dim=None,
                 a_dim=None):
        super().__init__()
        self.intensity = Multi_ATTN(dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim, v_dim,
                                    a_dim)
        self.stimulation = Multi_ATTN(dim, l_len, v_len, a_len, n_heads, n_layers, ffn, l_dim,
                                      v_dim, a_dim)
        self.trans = nn.Parameter(torch.rand(7, 7, 7), requires_grad=True)
        self.norm1 = nn.LayerNorm(7)
        self.out = nn.Linear(14, 7)

    def forward(self, l, v, a, l_mask, v_mask, a_mask):
        # both towers' trunks as ONE group per layer (they are independent until the bilinear head)
        ti, ts = self.intensity, self.stimulation
        pi, ps = fusion_trunk_multi(
            [ti._tower(l[:, 0], v[:, 0], a[:, 0], l_mask[:, 0], v_mask[:, 0], a_mask[:, 0]),
             ts._tower(l[:, 1], v[:, 1], a[:, 1], l_mask[:, 1], v_mask[:, 1], a_mask[:, 1])],
            ti.n_layers, keep_all=True)
        last_feat = ops.linear(pi, ti.classifier.weight)
        this_feat = ops.linear(ps, ts.classifier.weight)
        return ops.bilinear_head(this_feat, last_feat, self.trans, self.norm1.weight,
                                 self.norm1.bias, self.out.weight, self.out.bias)


def multi_circle_loss(y_pred, y_true):
    """cmu-mosei/run.py:342-351."""
    return ops.circle_loss_op(y_pred, y_true)
