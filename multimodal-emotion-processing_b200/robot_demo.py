"""Drop-in modules for the reference script ``robot_demo.py`` (dim 192, five biased input
projections, full RealFormer blocks, plain linear classifier; ensemble inference)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .blocks import FullAttentionBlock, as_act, as_mask, fusion_trunk, is_bf16

DROP = 0.1  # robot_demo.py:41


class Unify_Dimension_Conv1d(nn.Module):
    """robot_demo.py:293-311: five biased Conv1d(k=1); v = cat(v256, v512, v1024 projections)."""

    def __init__(self, dim, l_dim=768, dim_1024=1024, dim_512=512, dim_256=256, a_dim=40):
        super().__init__()
        self.linguistic = nn.Conv1d(l_dim, dim, kernel_size=1)
        self.visual_1024 = nn.Conv1d(dim_1024, dim // 3, kernel_size=1)
        self.visual_512 = nn.Conv1d(dim_512, dim // 3, kernel_size=1)
        self.visual_256 = nn.Conv1d(dim_256, dim // 3, kernel_size=1)
        self.acoustic = nn.Conv1d(a_dim, dim, kernel_size=1)
        self.drop = nn.Dropout(DROP)

    def forward(self, l, v_256, v_512, v_1024, a):
        bf = is_bf16()

        def proj(x, conv):
            return ops.dropout(ops.linear(x, conv.weight, conv.bias, bf16=bf), self.drop.p,
                               self.training)

        v = torch.cat((proj(v_256, self.visual_256), proj(v_512, self.visual_512),
                       proj(v_1024, self.visual_1024)), 2)
        return proj(l, self.linguistic), v, proj(a, self.acoustic)


class Position_Embedding(nn.Module):
    """robot_demo.py:314-321."""

    def __init__(self, max_len, dim):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_len, dim)
        self.len = max_len

    def forward(self, x):
        w = self.position_embeddings.weight
        return as_act(w)[None].expand(x.size(0), self.len, w.shape[1])


class Attention_Block(FullAttentionBlock):
    """robot_demo.py:324-374."""

    def __init__(self, dim, n_heads, ffn):
        super().__init__(dim, n_heads, ffn, DROP)


class Multi_class(nn.Module):
    """robot_demo.py:377-441 (``fully_connected`` / ``normalization`` exist but are unused)."""

    def __init__(self, dim, l_len, v_len, a_len, n_heads, n_layers, ffn):
        super().__init__()
        self.unify_dimension = Unify_Dimension_Conv1d(dim)
        self.linguistic_position = Position_Embedding(l_len, dim)
        self.visual_position = Position_Embedding(v_len, dim)
        self.acoustic_position = Position_Embedding(a_len, dim)
        self.n_layers = n_layers
        self.multimodal_blocks = nn.ModuleList([Attention_Block(dim, n_heads, ffn)
                                                for _ in range(9 * n_layers)])
        self.fully_connected = nn.Linear(dim * 6, dim)
        self.normalization = nn.LayerNorm(dim)
        self.drop = nn.Dropout(DROP)
        self.classifier = nn.Linear(dim * 6 * n_layers, 7)

    def forward(self, l, v_256, v_512, v_1024, a, l_mask, v_mask, a_mask):
        l, v, a = self.unify_dimension(l, v_256, v_512, v_1024, a)
        l = l + self.linguistic_position(l)
        v = v + self.visual_position(v)
        a = a + self.acoustic_position(a)
        x = fusion_trunk(self.multimodal_blocks, self.n_layers, {"l": l, "v": v, "a": a},
                         {"l": as_mask(l_mask), "v": as_mask(v_mask), "a": as_mask(a_mask)},
                         keep_all=True)
        return ops.linear(x, self.classifier.weight, self.classifier.bias)


def multi_circle_loss(y_pred, y_true):
    """robot_demo.py:444-453."""
    return ops.circle_loss_op(y_pred, y_true)


def sigmoids(pred, offset):
    """robot_demo.py:594-595 helper of demo_output: sigmoid with per-class logit offsets."""
    return torch.sigmoid(pred - offset)
