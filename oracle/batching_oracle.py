"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's host-side batch assembly, used
by tests/ to check the device kernel (csrc/assemble.cu).  Never imported by the product.

Pinned against the reference: tests/test_oracle.py runs these functions against the reference's own
``masking`` (others/realformer.py:72-82) extracted from /root/reference when it is present, and
against committed literal cases otherwise.
"""
import math

import numpy as np


def masking_tail(m, m_len, scrub=-71.0):
    """others/realformer.py:72-82 applied to ``features[-m_len:]`` (call sites :100-102): keep the
    last m_len rows, zero-pad at the end, mask = 1 on real rows, NaN/Inf -> -71."""
    m = np.asarray(m)[-m_len:]
    if len(m) >= m_len:
        m_mask = np.ones(m_len)
    else:
        m_mask = np.concatenate((np.ones(len(m)), np.zeros(m_len - len(m))))
    m = np.concatenate([m, np.zeros([m_len] + list(m.shape[1:]))], axis=0)[:m_len, ...]
    m = m.copy()
    for i in range(len(m)):
        for j in range(len(m[i])):
            if math.isinf(m[i][j]) or math.isnan(m[i][j]):
                m[i][j] = scrub
    return m, m_mask


def features_stride(feat, m_len):
    """robot_demo.py:119-135 (audio) / :137-150 (text) / :86-99 (video) on an in-memory array:
    empty -> zeros; shorter -> zero-pad; else rows 0, gap, 2 gap, ... (gap = T // m_len), first
    m_len of them, mask all ones."""
    feat = np.asarray(feat)
    D = feat.shape[1]
    if len(feat) == 0:
        return np.zeros((m_len, D)), np.zeros(m_len)
    if len(feat) < m_len:
        pad = m_len - len(feat)
        return (np.concatenate([feat, np.zeros((pad, D))], axis=0),
                np.concatenate((np.ones(len(feat)), np.zeros(pad)), axis=0))
    gap = len(feat) // m_len
    rows = [feat[i][None] for i in range(0, len(feat), gap)]
    return np.concatenate(rows[:m_len], axis=0), np.ones(m_len)


def head(feat, m_len):
    """First m_len rows, zero-padded (first view of cmu-mosei/run.py:139 without the statistics rows)."""
    feat = np.asarray(feat)[:m_len]
    pad = m_len - len(feat)
    return (np.concatenate([feat, np.zeros((pad, feat.shape[1]))], axis=0),
            np.concatenate((np.ones(len(feat)), np.zeros(pad))))


def mosei_masking(m, m_len, is_audio=False):
    """cmu-mosei/run.py:104-151, the ``is_bert=False`` branch (the only one its data_loader calls,
    :169-180): NaN/Inf -> -71 for audio; three statistics rows (column max / min / mean over ALL
    rows) in front; >= m_len-3 rows -> two views (head ``m[:m_len-3]``, tail ``m[len-m_len+3:]``)
    with all-ones masks, else one zero-padded view with mask = 1 on len+3 rows.
    Returns (list of views, list of masks) like the reference."""
    m = np.array(m, dtype=np.float64, copy=True)
    feat, feat_mask = [], []
    if is_audio:
        for i in range(len(m)):
            for j in range(len(m[i])):
                if math.isinf(m[i][j]) or math.isnan(m[i][j]):
                    m[i][j] = -71.
    stats = np.stack([m.max(axis=0), m.min(axis=0), m.mean(axis=0)], axis=0)
    if len(m) >= m_len - 3:
        for body in (m[:m_len - 3], m[len(m) - m_len + 3:]):
            feat.append(np.concatenate((stats, body), axis=0))
            feat_mask.append(np.ones(m_len))
    else:
        mm = np.concatenate((stats, m), axis=0)
        mm = np.concatenate([mm, np.zeros([m_len] + list(mm.shape[1:]))], axis=0)[:m_len, ...]
        feat.append(mm)
        feat_mask.append(np.concatenate((np.ones(len(m) + 3), np.zeros(m_len - len(m) - 3))))
    return feat, feat_mask
