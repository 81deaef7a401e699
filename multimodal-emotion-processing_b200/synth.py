"""Seeded synthetic inputs shaped like the reference's batches.

No real dataset is available (the reference reads CMU-MOSEI ``.csd`` / private ``.npy`` files from
hard-coded paths, others/realformer.py:41-42, Ren-MME/run.py:42).  Features are N(0,1) fp32, masks
are prefix masks ``arange(L) < len_b`` (every reference loader pads at the end:
others/realformer.py:72-77, cmu-mosei/run.py:126,146, Ren-MME/run.py:63, robot_demo.py:93) and
labels are Bernoulli with the MOSEI class rates computed from cmu-mosei/labels.txt.
Everything is generated on CPU from one ``torch.Generator`` so that the oracle, the golden
fixtures and the CUDA path see identical bytes.
"""
from __future__ import annotations

import torch

# happy, sad, angry, disgust, surprise, fear, neutral  (cmu-mosei/labels.txt, 23 248 rows)
MOSEI_RATES = (0.536, 0.258, 0.215, 0.176, 0.100, 0.082, 0.150)


def gen(seed: int = 1234) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def feats(g: torch.Generator, *shape: int) -> torch.Tensor:
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def prefix_mask(g: torch.Generator, lead: tuple, L: int, min_len: int = 1) -> torch.Tensor:
    """float32 0/1 mask of shape (*lead, L) with a random valid prefix of >= min_len steps."""
    n = 1
    for s in lead:
        n *= s
    lens = torch.randint(min_len, L + 1, (n,), generator=g)
    m = (torch.arange(L)[None, :] < lens[:, None]).to(torch.float32)
    return m.reshape(*lead, L)


def labels(g: torch.Generator, lead: tuple, n_cls: int) -> torch.Tensor:
    """int64 multi-hot labels; class rates cycle through the MOSEI rates."""
    rates = torch.tensor([MOSEI_RATES[i % len(MOSEI_RATES)] for i in range(n_cls)])
    u = torch.rand(*lead, n_cls, generator=g)
    return (u < rates).to(torch.int64)


def realformer_batch(seed=1234, B=32, P=6, L=(50, 50, 50), D=(300, 35, 74), n_cls=6,
                     empty_windows=False):
    """others/realformer.py batch: l (B,P,L,300) v (B,P,L,35) a (B,P,L,74), masks (B,P,L),
    labels (B,P,6), window mask (B,P)."""
    g = gen(seed)
    out = {
        "l": feats(g, B, P, L[0], D[0]), "v": feats(g, B, P, L[1], D[1]),
        "a": feats(g, B, P, L[2], D[2]),
        "l_mask": prefix_mask(g, (B, P), L[0]), "v_mask": prefix_mask(g, (B, P), L[1]),
        "a_mask": prefix_mask(g, (B, P), L[2]),
        "label": labels(g, (B, P), n_cls),
        "wmask": torch.ones(B, P, dtype=torch.int64),
    }
    if empty_windows:  # 'no_name' context slots: zero features, all-zero masks, zero loss weight
        dead = torch.rand(B, P, generator=g) < 0.25
        dead[:, 0] = False  # the first slot of a group is never 'no_name' (realformer.py:65)
        for k in ("l", "v", "a", "l_mask", "v_mask", "a_mask", "label"):
            out[k][dead] = 0
        out["wmask"][dead] = 0
    return out


def mosei_batch(seed=1234, B=32, L=(20, 100, 200), D=(300, 35, 74), n_cls=7):
    """cmu-mosei/run.py batch: index 0 = previous sentence, 1 = current."""
    g = gen(seed)
    return {
        "l": feats(g, B, 2, L[0], D[0]), "v": feats(g, B, 2, L[1], D[1]),
        "a": feats(g, B, 2, L[2], D[2]),
        "l_mask": prefix_mask(g, (B, 2), L[0]), "v_mask": prefix_mask(g, (B, 2), L[1]),
        "a_mask": prefix_mask(g, (B, 2), L[2]),
        "label": labels(g, (B,), n_cls),
    }


def renmme_batch(seed=1234, B=256, L=(40, 76, 275), D=(768, 640, 205), n_cls=9, rdrop_pairs=True):
    """Ren-MME/run.py batch: 12 tensors in Base_model.forward order (Ren-MME/run.py:281-282);
    rows 2i / 2i+1 carry identical inputs (R-Drop pairs, Ren-MME/run.py:143-146)."""
    g = gen(seed)
    n = B // 2 if rdrop_pairs else B

    def rep(t):
        return t.repeat_interleave(2, 0).contiguous() if rdrop_pairs else t

    t = {}
    for mod, Lm, Dm in (("text", L[0], D[0]), ("video", L[1], D[1]), ("audio", L[2], D[2])):
        for tower in ("pre", "pro"):
            t[f"{tower}_{mod}_feat"] = rep(feats(g, n, Lm, Dm))
            t[f"{tower}_{mod}_mask"] = rep(prefix_mask(g, (n,), Lm))
    order = ["pre_text_feat", "pre_text_mask", "pro_text_feat", "pro_text_mask",
             "pre_video_feat", "pre_video_mask", "pro_video_feat", "pro_video_mask",
             "pre_audio_feat", "pre_audio_mask", "pro_audio_feat", "pro_audio_mask"]
    return {"inputs": [t[k] for k in order], "label": rep(labels(g, (n,), n_cls)).float()}


def rencecps_batch(seed=1234, B=128, dim=2304, n_cls=9):
    g = gen(seed)
    return {"feat": feats(g, B, 2, dim), "label": labels(g, (B,), n_cls)}


def robot_batch(seed=1234, B=1, L=(25, 100, 100), n_cls=7):
    """robot_demo.py batch: l (B,25,768), v256/v512/v1024 (B,100,*), a (B,100,40)."""
    g = gen(seed)
    return {
        "l": feats(g, B, L[0], 768), "v_256": feats(g, B, L[1], 256),
        "v_512": feats(g, B, L[1], 512), "v_1024": feats(g, B, L[1], 1024),
        "a": feats(g, B, L[2], 40),
        "l_mask": prefix_mask(g, (B,), L[0]), "v_mask": prefix_mask(g, (B,), L[1]),
        "a_mask": prefix_mask(g, (B,), L[2]),
        "label": labels(g, (B,), n_cls),
    }


def encoder_batch(seed=1234, B=64, L=128, d=512):
    """BASELINE config 2: x (B, L, d), mask (B, L); dy = a fixed random cotangent so that parity
    tests can use loss = mean(out * dy) (mean(out^2) of a LayerNorm output is ~constant and has
    vanishing gradients)."""
    g = gen(seed)
    return {"x": feats(g, B, L, d), "mask": prefix_mask(g, (B,), L), "dy": feats(g, B, L, d)}


def randomize_gates(state: dict, seed: int = 1) -> dict:
    """The ReZero gates a, b, c are initialised to 0 (others/realformer.py:169-171) which makes
    every attention/FFN weight gradient exactly 0; parity tests draw them from U(-0.5, 0.5)."""
    g = gen(seed)
    for k, v in state.items():
        if k.endswith((".a", ".b", ".c")) or k in ("a", "b", "c"):
            state[k] = (torch.rand(v.shape, generator=g) - 0.5).to(v.dtype)
    return state
