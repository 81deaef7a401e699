"""BASELINE config 2: a standalone RealFormer residual-attention encoder, i.e. a chain of
``others/realformer.py`` ``Attention_Block``s called the way ``Multi_class.forward`` chains them
(others/realformer.py:232-233): the query stream evolves, keys/values stay the chain's source
sequence, and each layer receives the previous layer's pre-softmax scores."""
from __future__ import annotations

import torch
import torch.nn as nn

from .blocks import FullAttentionBlock, as_act


class ResidualEncoder(nn.Module):
    def __init__(self, dim: int = 512, n_heads: int = 8, n_layers: int = 6, ffn: int = 2,
                 drop: float = 0.0):
        super().__init__()
        self.blocks = nn.ModuleList([FullAttentionBlock(dim, n_heads, ffn, drop)
                                     for _ in range(n_layers)])

    def forward(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        x = as_act(x)                        # one cast for the whole chain (bf16 mode)
        q, s = x, None
        n = len(self.blocks)
        for i, blk in enumerate(self.blocks):
            # the last layer's scores feed nothing: do not write them
            q, s = blk(q, x, x, mask, s, emit_scores=i + 1 < n)
        return q
