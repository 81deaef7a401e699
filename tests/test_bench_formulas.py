"""bench.py derives each kernel's ALGORITHMIC flops / bytes (the numerator of `roofline.achieved`)
from the C-ABI arguments of the launch.  These tests tie the argument positions it reads to the
prototypes in mmemo_b200/_lib.py (= include/mmemo.h) and the formulas to SURVEY.md §8(d)."""
import ctypes as C

import bench
from mmemo_b200 import _lib

INT_TYPES = (C.c_int64, C.c_int)


def _args(name, **at):
    """A call tuple for `name`: pointers non-null, integers 1, then the given {index: value}."""
    sig = _lib.SIGNATURES[name]
    a = [(1 if t in INT_TYPES else (0.0 if t is C.c_float else 0xdead)) for t in sig]
    for i, v in at.items():
        i = int(i)
        if isinstance(v, int) and v is not None and not isinstance(v, bool):
            assert sig[i] in INT_TYPES, (name, i, "bench reads a size from a non-integer argument")
        a[i] = v
    return tuple(a)


def test_linear_formulas_and_positions():
    M, N, K = 8192, 1024, 512
    fl, by, tag = bench._algorithmic("mmemo_linear_fwd_bf16", _args("mmemo_linear_fwd_bf16", **{
        "1": 0, "10": M, "11": N, "12": K}))
    assert fl == 2 * M * N * K and by == 2 * (M * K + N * K + M * N) and tag == f"{M}x{N}x{K}"
    fl, by, _ = bench._algorithmic("mmemo_linear_bwd_x_bf16", _args("mmemo_linear_bwd_x_bf16", **{
        "6": None, "8": M, "9": N, "10": K}))
    assert fl == 2 * M * N * K and by == 2 * (M * N + N * K + M * K)
    fl, by, _ = bench._algorithmic("mmemo_linear_bwd_w_bf16", _args("mmemo_linear_bwd_w_bf16", **{
        "8": M, "9": N, "10": K}))
    assert fl == 2 * M * N * K and by == 2 * (M * N + M * K) + 4 * N * K      # fp32 dW


def test_grouped_linear_formulas():
    Ms, Ns, Ks = [8192, 8192], [512, 1024], [512, 512]
    name = "mmemo_linear_fwd_grouped_bf16"   # ..., M, N, K, relu, accumulate, pos, period, ldpos, stream
    assert len(_lib.SIGNATURES[name]) == 17
    a = [2] + [None] * 16
    a[8], a[9], a[10] = Ms, Ns, Ks
    fl, by, tag = bench._algorithmic(name, tuple(a))
    assert fl == sum(2 * m * n * k for m, n, k in zip(Ms, Ns, Ks))
    assert tag == "8192x512x512+8192x1024x512"
    name = "mmemo_linear_bwd_x_grouped_bf16"      # ..., relu_src, ldrelu, M, N, K, accumulate, stream
    assert len(_lib.SIGNATURES[name]) == 14
    a = [2] + [None] * 13
    a[9], a[10], a[11] = Ms, Ns, Ks
    fl, by, _ = bench._algorithmic(name, tuple(a))
    assert fl == sum(2 * m * n * k for m, n, k in zip(Ms, Ns, Ks))
    name = "mmemo_linear_bwd_w_grouped_bf16"
    assert len(_lib.SIGNATURES[name]) == 12
    a = [2] + [None] * 11
    a[7], a[8], a[9] = Ms, Ns, Ks
    fl, by, _ = bench._algorithmic(name, tuple(a))
    assert by == sum(2 * (m * n + m * k) + 4 * n * k for m, n, k in zip(Ms, Ns, Ks))
    # large groups (a fusion-trunk layer): one tag entry per distinct shape
    a[7], a[8], a[9] = [9600] * 9, [96] * 9, [96] * 9
    a[0] = 9
    _, _, tag = bench._algorithmic(name, tuple(a))
    assert tag == "G9:9x9600x96x96"


def test_grouped_attention_and_layernorm_formulas():
    """The grouped entry points take a table: bench reads it back through the ctypes struct."""
    B, H, Lq, Lk, hd = 192, 6, 50, 50, 16
    d, S = H * hd, B * H * Lq * Lk
    probs = (_lib.AttnProblem * 2)()
    for i in range(2):
        q = probs[i]
        q.q, q.k, q.v = 0x1000, 0x2000, 0x3000
        q.B, q.H, q.Lq, q.Lk, q.hd = B, H, Lq, Lk, hd
        q.s_out = 0x4000
    probs[1].s_prev = 0x5000
    fl, by, tag = bench._algorithmic("mmemo_resattn_fwd_grouped_bf16",
                                     (2, C.cast(probs, C.c_void_p), None))
    assert fl == 2 * 4 * B * Lq * Lk * d
    assert by == 2 * (2 * B * d * 3 * Lq + 4 * B * Lk + 2 * B * Lq * d + 2 * S) + 2 * S
    assert tag.startswith("G2:B192H6hd16:")
    probs[0].v = probs[0].k            # lite block: K = V read once
    _, by2, _ = bench._algorithmic("mmemo_resattn_fwd_grouped_bf16",
                                   (2, C.cast(probs, C.c_void_p), None))
    assert by - by2 == 2 * B * d * Lk
    Ms = [9600] * 9
    name = "mmemo_add_ln_fwd_grouped_bf16"
    a = [9] + [None] * 13
    a[1], a[9], a[10] = [1] * 9, Ms, 96
    _, by, _ = bench._algorithmic(name, tuple(a))
    assert by == 3 * 2 * sum(Ms) * 96
    name = "mmemo_add_ln_bwd_grouped_bf16"
    a = [9] + [None] * 16
    a[2], a[14], a[15] = [1] * 9, Ms, 96
    _, by, _ = bench._algorithmic(name, tuple(a))
    assert by == 5 * 2 * sum(Ms) * 96


def test_attention_formulas_match_survey_8d():
    B, H, L, hd = 64, 8, 128, 64
    d, S = H * hd, B * H * L * L
    name = "mmemo_resattn_fwd_bf16"
    fl, by, tag = bench._algorithmic(name, _args(name, **{"15": B, "16": H, "17": L, "18": L,
                                                          "19": hd}))
    # 4·B·Lq·Lk·d flops; q,k,v in + o out + mask + S_prev read + S written
    assert fl == 4 * B * L * L * d
    assert by == 2 * B * d * 3 * L + 4 * B * L + 2 * B * L * d + 2 * S + 2 * S
    assert tag.endswith("+prev+S")
    fl, by, tag = bench._algorithmic(name, _args(name, **{"9": None, "11": None, "15": B, "16": H,
                                                          "17": L, "18": L, "19": hd}))
    assert by == 2 * B * d * 3 * L + 4 * B * L + 2 * B * L * d and "+prev" not in tag
    name = "mmemo_resattn_bwd_bf16"
    fl, by, tag = bench._algorithmic(name, _args(name, **{"27": B, "28": H, "29": L, "30": L,
                                                          "31": hd}))
    # (a'): q,k,v,dO + dq,dk,dv + o, S re-read, dS_next, S_prev read + dS_prev written; 8·B·Lq·Lk·d
    assert fl == 8 * B * L * L * d
    assert by == 2 * B * d * 4 * L + 2 * B * d * 3 * L + 2 * B * L * d + 4 * 2 * S
    fl, _, tag = bench._algorithmic(name, _args(name, **{"11": None, "27": B, "28": H, "29": L,
                                                         "30": L, "31": hd}))
    assert fl == 10 * B * L * L * d and "+recompute" in tag


def test_layernorm_and_colsum_formulas():
    M, d = 8192, 512
    name = "mmemo_add_ln_fwd_bf16"
    _, by, _ = bench._algorithmic(name, _args(name, **{"11": M, "12": d}))
    assert by == 3 * 2 * M * d                       # res, x read; y written
    name = "mmemo_add_ln_bwd_bf16"
    _, by, _ = bench._algorithmic(name, _args(name, **{"20": M, "21": d}))
    assert by == 5 * 2 * M * d                       # dy, res, x read; dres, dx written
    name = "mmemo_rowsum_bf16"
    _, by, _ = bench._algorithmic(name, _args(name, **{"3": M, "4": 1024}))
    assert by == 2 * M * 1024


def test_every_timed_launcher_has_a_formula():
    """Kernels that appear in the bench step must not silently report 0 bytes and 0 flops."""
    for name in ("mmemo_linear_fwd_bf16", "mmemo_linear_bwd_x_bf16", "mmemo_linear_bwd_w_bf16",
                 "mmemo_resattn_fwd_bf16", "mmemo_resattn_bwd_bf16", "mmemo_add_ln_fwd_bf16",
                 "mmemo_add_ln_bwd_bf16", "mmemo_rowsum_bf16"):
        fl, by, _ = bench._algorithmic(name, _args(name))
        assert by > 0, name
