"""Device-side batch assembly.  The reference data loaders build every batch on
the host, sample by sample: slice / stride the variable-length feature arrays, ``np.concatenate``
zero padding, scrub NaN/Inf with Python loops (others/realformer.py:72-82), stack, and finally
``torch.cuda.FloatTensor(list_of_arrays)`` (others/realformer.py:308).  Here the raw sequences of
a batch are packed once into ONE pinned buffer (a single H2D copy) and one kernel
(``mmemo_assemble_batch_f32``) writes the padded ``(N, m_len, D)`` batch and its mask.

    rb = RaggedBatch.pack([seq_0, seq_1, None, ...])      # None / empty = 'no_name' slot
    x, mask = rb.cuda().assemble(m_len=50, mode="tail")   # realformer;  mode="stride": robot_demo
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops

MODES = {"tail": 0, "head": 1, "stride": 2}
SCRUB_VALUE = -71.0        # others/realformer.py:81, cmu-mosei/run.py:110


class RaggedBatch:
    """Variable-length float32 sequences of one modality, concatenated row-wise."""

    def __init__(self, flat: torch.Tensor, row_start: torch.Tensor, n_rows: torch.Tensor):
        self.flat, self.row_start, self.n_rows = flat, row_start, n_rows

    @classmethod
    def pack(cls, seqs: Sequence[Optional[np.ndarray]], dim: Optional[int] = None,
             pin: bool = True) -> "RaggedBatch":
        """Concatenate (T_i, D) arrays (``None`` or empty = a missing utterance) into pinned host
        memory.  ``dim`` is only needed when every sequence is missing."""
        arrs = [None if (s is None or len(s) == 0) else np.asarray(s) for s in seqs]
        dims = {a.shape[1] for a in arrs if a is not None}
        if len(dims) > 1:
            raise ValueError(f"sequences disagree on the feature width: {sorted(dims)}")
        D = dims.pop() if dims else dim
        if D is None:
            raise ValueError("all sequences are empty: pass dim=")
        lens = [0 if a is None else a.shape[0] for a in arrs]
        total = sum(lens)
        pin = pin and torch.cuda.is_available()
        flat = torch.empty(max(total, 1), D, dtype=torch.float32, pin_memory=pin)
        view, pos = flat.numpy(), 0
        starts = []
        for a, n in zip(arrs, lens):
            starts.append(pos)
            if n:
                view[pos:pos + n] = a          # converts float64 features to float32 like FloatTensor
                pos += n
        return cls(flat, torch.tensor(starts, dtype=torch.int64), torch.tensor(lens, dtype=torch.int64))

    def cuda(self, device=None, non_blocking: bool = True) -> "RaggedBatch":
        dev = torch.device("cuda" if device is None else device)
        return RaggedBatch(self.flat.to(dev, non_blocking=non_blocking),
                           self.row_start.to(dev, non_blocking=non_blocking),
                           self.n_rows.to(dev, non_blocking=non_blocking))

    def __len__(self) -> int:
        return self.n_rows.numel()

    def assemble(self, m_len: int, mode: str = "tail", scrub: Optional[float] = SCRUB_VALUE,
                 lead_shape: Optional[Sequence[int]] = None):
        """-> ``(x (N, m_len, D), mask (N, m_len))`` float32 on the GPU; ``lead_shape`` reshapes the
        sample dimension, e.g. ``(B, P)`` for realformer's windows.  ``scrub=None`` keeps NaN/Inf."""
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        if not self.flat.is_cuda:
            raise RuntimeError("RaggedBatch.assemble runs on the GPU: call .cuda() first "
                               "(mmemo_b200 has no CPU fallback)")
        N, D = len(self), self.flat.shape[1]
        out = torch.empty(N, m_len, D, dtype=torch.float32, device=self.flat.device)
        mask = torch.empty(N, m_len, dtype=torch.float32, device=self.flat.device)
        ops._call("mmemo_assemble_batch_f32", self.flat.data_ptr(), self.row_start.data_ptr(),
                  self.n_rows.data_ptr(), out.data_ptr(), mask.data_ptr(), N, m_len, D,
                  MODES[mode], int(scrub is not None), float(scrub or 0.0), ops._stream())
        if lead_shape is not None:
            out = out.view(*lead_shape, m_len, D)
            mask = mask.view(*lead_shape, m_len)
        return out, mask

    # ---- cmu-mosei/run.py:104-151: statistics rows + head / tail views -----------------------------
    def two_views(self, m_len: int) -> torch.Tensor:
        """bool (N,): samples long enough (>= m_len - 3 rows) to yield a head AND a tail view
        (cmu-mosei/run.py:137; the loader emits two training samples when the current sentence's
        text does, :181-188)."""
        return self.n_rows >= (m_len - 3)

    def assemble_stats(self, m_len: int, view: str = "head", scrub: Optional[float] = None,
                       lead_shape: Optional[Sequence[int]] = None):
        """cmu-mosei ``masking`` (non-BERT branch): rows 0-2 = column-wise max / min / mean over the
        whole sample, then ``m_len - 3`` body rows — ``view="head"``: the first ones (``feat[0]`` of
        the reference), ``view="tail"``: the last ones (``feat[-1]``; equals the head view for
        samples with a single view).  ``scrub=-71`` for the acoustic features (is_audio=True)."""
        if view not in ("head", "tail"):
            raise ValueError("view must be 'head' or 'tail'")
        if m_len < 3:
            raise ValueError("m_len must be >= 3 (three statistics rows)")
        if not self.flat.is_cuda:
            raise RuntimeError("RaggedBatch.assemble_stats runs on the GPU: call .cuda() first "
                               "(mmemo_b200 has no CPU fallback)")
        N, D = len(self), self.flat.shape[1]
        out = torch.empty(N, m_len, D, dtype=torch.float32, device=self.flat.device)
        mask = torch.empty(N, m_len, dtype=torch.float32, device=self.flat.device)
        ops._call("mmemo_assemble_stats_batch_f32", self.flat.data_ptr(), self.row_start.data_ptr(),
                  self.n_rows.data_ptr(), out.data_ptr(), mask.data_ptr(), N, m_len, D,
                  int(view == "tail"), int(scrub is not None), float(scrub or 0.0), ops._stream())
        if lead_shape is not None:
            out = out.view(*lead_shape, m_len, D)
            mask = mask.view(*lead_shape, m_len)
        return out, mask
