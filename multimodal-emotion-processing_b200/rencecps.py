"""Drop-in module for the reference script ``rencecps/run.py`` (text-only transition classifier)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

DIM = 768 * 3


class Concat_Linear(nn.Module):
    """rencecps/run.py:130-148.  feat (B, 2, dim): index 0 = previous sentence, 1 = current."""

    def __init__(self, dim):
        super().__init__()
        self.intensity = nn.Linear(dim, 9, bias=False)
        self.stimulation = nn.Linear(dim, 9, bias=False)
        self.trans = nn.Parameter(torch.rand(9, 9, 9), requires_grad=True)
        self.norm = nn.LayerNorm(9)
        self.out = nn.Linear(18, 9)

    def forward(self, feat):
        last_feat = ops.linear(feat[:, 0], self.intensity.weight)
        this_feat = ops.linear(feat[:, 1], self.stimulation.weight)
        return ops.bilinear_head(this_feat, last_feat, self.trans, self.norm.weight, self.norm.bias,
                                 self.out.weight, self.out.bias)


def multi_circle_loss(y_pred, y_true):
    """rencecps/run.py:151-160."""
    return ops.circle_loss_op(y_pred, y_true)
