"""CPU tests of the host logic with the kernel launches stubbed out.

The C launchers cannot run here (no GPU), so ``ops._call`` is replaced by a recorder.  Outputs
are uninitialised memory, but everything the Python side is responsible for is exercised: op
schemas, autograd wiring of every custom op, output shapes/dtypes, which parameters receive a
gradient (must equal the reference's set, e.g. first-layer ``c`` gets none — SURVEY §7), the
sequence of launcher names, and state_dict compatibility with the reference fixtures.
"""
import pytest
import torch

import mmemo_b200
from mmemo_b200 import ops
from tests import cases


class _FakeLib:
    def mmemo_resattn_uses_tensor_cores(self, *a):
        return 0

    def mmemo_resattn_uses_mma(self, *a):
        return 1


@pytest.fixture()
def stub(monkeypatch):
    calls = []
    monkeypatch.setattr(ops, "_call", lambda name, *a: calls.append(name))
    monkeypatch.setattr(ops, "_try_call", lambda name, *a: calls.append(name) or True)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops._lib, "load", lambda: _FakeLib())
    mmemo_b200.robot_demo.DROP = 0.0
    mmemo_b200.ren_mme.DROP = 0.0
    ops.clear_shadow_cache()
    yield calls
    ops.clear_shadow_cache()


class _Loss:
    multi_circle_loss = staticmethod(lambda p, t: ops.circle_loss_op(p, t))
    multi_loss = staticmethod(lambda p, t: ops.circle_loss_op(p, t).mean())
    rdrop_kl = staticmethod(lambda p: ops.rdrop_kl_op(p))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(cases.CASES))
def test_module_wiring_matches_reference_grad_set(stub, name, mode):
    c = cases.CASES[name]
    g = torch.load(cases.golden_path(name))
    model = c.our_model(mmemo_b200).train()
    assert model.load_state_dict(g["state"]).missing_keys == []
    with mmemo_b200.precision(mode):
        logits, loss, grads, igrads = cases.run_module_with_grads(model, c, g["batch"], _Loss)
    assert logits.shape == g["logits"].shape
    assert set(grads) == set(g["grads"]), sorted(set(grads) ^ set(g["grads"]))
    for k, v in grads.items():
        assert v.shape == g["grads"][k].shape and v.dtype == torch.float32, k
    for k, v in igrads.items():
        assert v.shape == g["input_grads"][k].shape
    assert any(n.startswith("mmemo_resattn_fwd") for n in stub) or name == "rencecps_concat_linear"
    assert any(n.endswith("_bf16") for n in stub) == (mode == "bf16" and name != "rencecps_concat_linear")


def test_state_dict_keys_equal_reference(stub):
    for name, c in cases.CASES.items():
        g = torch.load(cases.golden_path(name))
        model = c.our_model(mmemo_b200)
        assert list(model.state_dict().keys()) == list(g["state"].keys()), name
        for k, v in model.state_dict().items():
            assert v.shape == g["state"][k].shape, (name, k)


def test_unfused_paths_when_k_is_not_v(stub):
    blk = mmemo_b200.realformer.Attention_Block(16, 2)
    q, k, v = torch.randn(2, 5, 16, requires_grad=True), torch.randn(2, 7, 16), torch.randn(2, 7, 16)
    out, s = blk(q, k, v, torch.ones(2, 7))
    assert out.shape == (2, 5, 16) and s.shape == (2, 2, 5, 7)
    out.sum().backward()
    assert q.grad.shape == q.shape and blk.w_qkv[2].weight.grad is not None
    lite = mmemo_b200.cmu_mosei.Attention_Block(16, 2, 1)
    out, s = lite(q, k, v, torch.ones(2, 7))
    assert out.shape == (2, 5, 16)


def test_trunk_skips_unused_score_writes(stub, monkeypatch):
    """The last layer of a chain must not write its scores (nothing consumes them) - in the grouped
    trunk (one op per layer over the nine chains) and in the per-block path - and the drop-in blocks
    themselves keep returning scores afterwards (emit_scores is a per-call argument, not module
    state)."""
    from mmemo_b200 import blocks, group_ops
    emitted, groups = [], []
    real, real_g = ops.block_lite_op, group_ops.trunk_lite_op

    def spy(q, kv, mask, s_prev, params, H, bf16, emit_s):
        emitted.append(emit_s)
        return real(q, kv, mask, s_prev, params, H, bf16, emit_s)

    def spy_g(qs, kvs, masks, s_prevs, params, H, bf16, emit_s, drop_p, drop_seed):
        groups.append((len(qs), len(s_prevs), emit_s))
        return real_g(qs, kvs, masks, s_prevs, params, H, bf16, emit_s, drop_p, drop_seed)

    monkeypatch.setattr(ops, "block_lite_op", spy)
    monkeypatch.setattr(group_ops, "trunk_lite_op", spy_g)
    m = mmemo_b200.cmu_mosei.Multi_ATTN(16, 4, 5, 6, 2, 2, 1, l_dim=8, v_dim=6, a_dim=7)
    args = (torch.randn(2, 4, 8), torch.randn(2, 5, 6), torch.randn(2, 6, 7), torch.ones(2, 4),
            torch.ones(2, 5), torch.ones(2, 6))
    out_g = m(*args)
    assert groups == [(9, 0, True), (9, 9, False)] and emitted == []
    monkeypatch.setattr(blocks, "GROUPED_TRUNK", False)
    out_b = m(*args)
    assert emitted == [True, False] * 9 and out_b.shape == out_g.shape
    blk = m.multimodal_blocks[1]                  # a last-layer block, called directly
    x = torch.randn(2, 4, 16)
    out, s = blk(x, x, x, torch.ones(2, 4))
    assert s is not None and s.shape == (2, 2, 4, 4)


def test_grouped_trunk_launch_count(stub):
    """One grouped launch per kernel kind and layer: a State_Transfer step (2 layers x 9 chains,
    fwd+bwd) stays under 60 libmmemo launches in bf16 mode (~350 before)."""
    m = mmemo_b200.realformer.State_Transfer(l_dim=16, v_dim=8, a_dim=8, dim=16, l_len=4, v_len=4,
                                             a_len=4, n_heads=2, n_layers=2, ffn=2).train()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith((".a", ".b", ".c")):
                p.fill_(0.3)
    b = dict(l=torch.randn(2, 3, 4, 16), v=torch.randn(2, 3, 4, 8), a=torch.randn(2, 3, 4, 8))
    ones = torch.ones(2, 3, 4)
    with mmemo_b200.precision("bf16"):
        out = m(b["l"], b["v"], b["a"], ones, ones, ones)
        ops.circle_loss_op(out, torch.zeros(2, 3, 6)).mean().backward()
    assert len(stub) <= 60, (len(stub), stub)
    assert sum(n.startswith("mmemo_resattn") for n in stub) == 4      # 2 layers x (fwd + bwd)


def test_shadow_cache_tracks_parameter_version(stub):
    w = torch.nn.Parameter(torch.randn(8, 4))
    a = ops.shadow_bf16(w)
    assert ops.shadow_bf16(w) is a
    with torch.no_grad():
        w.add_(1.0)
    assert ops.shadow_bf16(w) is not a
    assert len(ops._shadow) == 1


def test_no_cpu_fallback():
    """Without the stub, a CPU tensor must raise instead of silently computing on the host."""
    with pytest.raises(RuntimeError):
        ops.linear(torch.randn(2, 3), torch.randn(4, 3))


def test_position_length_mismatch_raises(stub):
    m = mmemo_b200.realformer.Multi_class(8, 6, 7, 16, 4, 5, 6, 2, 1, 2)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 9, 8), torch.randn(2, 5, 6), torch.randn(2, 6, 7), torch.ones(2, 9),
          torch.ones(2, 5), torch.ones(2, 6))


def test_grad_reducer_on_encoder_skips_undefined_grads(stub):
    """The first-layer ``c`` receives an UNDEFINED gradient (its hook still fires): it must stay out
    of the buckets; all other parameters end up as views of the flat buckets."""
    import socket
    import torch.distributed as dist
    from mmemo_b200 import dp, synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        model = mmemo_b200.ResidualEncoder(32, 4, 3)
        red = dp.GradReducer(model, world_size=2, bucket_bytes=16 << 10)
        b = synth.encoder_batch(B=2, L=16, d=32)
        for _ in range(3):
            model.zero_grad(set_to_none=True)
            red.backward((model(b["x"], b["mask"]).float() ** 2).mean())
        assert model.blocks[0].c.grad is None
        assert len(red.buckets) >= 2
        flat_ptrs = {bk.flat.untyped_storage().data_ptr() for bk in red.buckets}
        for n, p in model.named_parameters():
            if n != "blocks.0.c":
                assert p.grad.untyped_storage().data_ptr() in flat_ptrs, n
    finally:
        dist.destroy_process_group()


def test_fused_adamw_host_logic(monkeypatch):
    """optim.AdamW: torch.optim-compatible param_groups / state layout, one launch per step-count
    group, global norm over all groups, empty groups skipped (Ren-MME/run.py:376-379)."""
    from mmemo_b200 import optim as mo
    calls = []
    monkeypatch.setattr(ops, "_call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(mo, "_check", lambda ts, what: None)
    w1, w2 = torch.nn.Parameter(torch.ones(5, 3)), torch.nn.Parameter(torch.ones(7))
    frozen = torch.nn.Parameter(torch.ones(2))
    opt = mo.AdamW([{"params": [w1, w2, frozen], "lr": 1e-3}, {"params": [], "lr": 1e-3}],
                   max_grad_norm=1.0)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.ones(1))])
    assert set(opt.param_groups[0]) >= {"lr", "betas", "eps", "weight_decay"}
    assert opt.param_groups[0]["weight_decay"] == ref.param_groups[0]["weight_decay"] == 1e-2
    w1.grad, w2.grad = torch.ones_like(w1), torch.ones_like(w2)
    opt.step()
    names = [n for n, _ in calls]
    assert names == ["mmemo_grad_sqnorm_f32", "mmemo_adam_step_f32"]
    a = calls[1][1]
    assert a[0] == 2 and a[11] == 1 and a[12] == 1        # two tensors, decoupled, step 1
    assert a[13] is not None and a[14] == 1.0             # fused clipping
    assert set(opt.state[w1]) == {"step", "exp_avg", "exp_avg_sq"} and frozen not in opt.state
    # a parameter that starts receiving gradients later gets its own bias-correction step
    calls.clear()
    frozen.grad = torch.ones_like(frozen)
    opt.step()
    steps = sorted(a[12] for n, a in calls if n == "mmemo_adam_step_f32")
    assert steps == [1, 2]
    assert float(opt.state[w1]["step"]) == 2 and float(opt.state[frozen]["step"]) == 1
    # ReduceLROnPlateau-style lr edits are picked up
    opt.param_groups[0]["lr"] = 5e-4
    calls.clear()
    opt.step()
    assert all(abs(a[6] - 5e-4) < 1e-12 for n, a in calls if n == "mmemo_adam_step_f32")
    with pytest.raises(ValueError):
        mo.Adam([w1], amsgrad=True)
    assert mo.Adam([w1]).param_groups[0]["weight_decay"] == 0.0


def test_block_gradient_regions_are_claimed_once_and_forgotten_with_their_reducer():
    """ops.claim_zbufs: the bucket regions of a GROUP of blocks are handed out all-or-nothing, once
    per backward, only while the reducer has zero-filled them; unregister_grad_dests drops every
    destination that lives in a removed reducer's buckets."""
    k1, k2 = torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(2, 2))
    other = torch.nn.Parameter(torch.zeros(3))
    flat_a, flat_b = torch.zeros(256), torch.zeros(256)
    ops.clear_grad_dest()
    ops.register_zbuf_dest(k1, flat_a, 0, 100)
    ops.register_zbuf_dest(k2, flat_a, 128, 100)
    ops.register_grad_dest(other, flat_b, 32)
    try:
        ops.grad_dest_enabled, ops.grad_dest_zeroed = True, False
        ops.begin_backward()
        assert ops.claim_zbufs([k1, k2], 100) is None                  # buckets not zero-filled
        ops.grad_dest_zeroed = True
        assert ops.claim_zbufs([k1, other], 100) is None               # one block has no region
        assert ops.claim_zbufs([k1, k2], 96) is None                   # layout mismatch
        z = ops.claim_zbufs([k1, k2], 100)
        assert [t.data_ptr() for t in z] == [flat_a.data_ptr(), flat_a[128:].data_ptr()]
        assert all(t.numel() == 100 for t in z)
        assert ops.claim_zbufs([k1, k2], 100) is None                  # second use in this backward
        ops.begin_backward()
        assert ops.claim_zbufs([k2], 100) is not None
        ops.unregister_grad_dests([flat_a])
        ops.begin_backward()
        assert ops.claim_zbufs([k2], 100) is None and ops._dest(other) is not None
        ops.unregister_grad_dests([flat_b])
        assert ops._dest(other) is None
    finally:
        ops.grad_dest_zeroed = False
        ops.clear_grad_dest()
        ops.begin_backward()


def test_bucket_slot_is_written_once_per_backward():
    """A weight that is used twice in one forward (shared module) must not have both gradients
    written into the same data-parallel bucket slot: the second use gets a fresh buffer and
    autograd sums the two."""
    w = torch.nn.Parameter(torch.zeros(4, 3))
    flat = torch.zeros(64)
    ops.clear_grad_dest()
    ops.register_grad_dest(w, flat, 16)
    ops.grad_dest_enabled = True
    try:
        ops.begin_backward()
        buf1, ret1 = ops._wgrad(w, 4, 3)
        buf2, ret2 = ops._wgrad(w, 4, 3)
        assert ret1.numel() == 0 and buf1.data_ptr() == flat[16:].data_ptr()      # the slot itself
        assert ret2.numel() == 12 and buf2.data_ptr() != buf1.data_ptr()          # a fresh buffer
        ops.begin_backward()                                                        # next step
        buf3, ret3 = ops._wgrad(w, 4, 3)
        assert ret3.numel() == 0 and buf3.data_ptr() == buf1.data_ptr()
    finally:
        ops.clear_grad_dest()
        ops.begin_backward()


def test_ensemble_emotion_scores_follow_demo_output():
    """robot_demo.py:594-595,609,615-622: score = 1 / (1 + exp(-x + t)) with the per-emotion
    offsets happy .1, sad .1, angry -.1, disgust 0, surprise .1, fear 0, rounded to 2 places."""
    import math

    from mmemo_b200.robot_demo import Ensemble
    ens = Ensemble([torch.nn.Linear(2, 2)])
    pred = torch.tensor([[0.3, -1.2, 2.0, 0.0, -0.1, 5.0, 9.9]])
    offs = [0.1, 0.1, -0.1, 0.0, 0.1, 0.0]
    exp = [round(1 / (1 + math.exp(-float(pred[0][i]) + offs[i])), 2) for i in range(6)]
    got = ens.emotions(pred)
    assert list(got) == ["happy", "sad", "angry", "disgust", "surprise", "fear"]
    assert all(abs(g - e) <= 0.011 for g, e in zip(got.values(), exp)), (got, exp)
    with pytest.raises(ValueError):
        Ensemble([])
    with pytest.raises(RuntimeError):          # no CPU fallback
        ens(*[torch.zeros(1, 2)] * 8)


def test_bucket_length_is_sliceable_for_any_world():
    """The all-reduce kernel needs bucket_len % (4 * world) == 0; layouts of the tested worlds
    (2, 4, 8) keep the 1024-float granularity."""
    from mmemo_b200 import dp
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 96 * 96, 1, 33)]
    for world in range(1, 17):
        offs, n = dp._Bucket.layout(ps, world)
        assert n % (4 * world) == 0 and n >= offs[-1] + 33 and all(o % 32 == 0 for o in offs)
        if world in (1, 2, 4, 8, 16):
            assert n == dp._Bucket.layout(ps)[1] and n % 1024 == 0


def test_training_dropout_stays_on_the_grouped_path(stub, monkeypatch):
    """Ren-MME trains with dropout 0.1 (Ren-MME/run.py:36).  The grouped trunk keeps its launch
    count: the two dropout sites of the 18 chains are ONE grouped launch each per direction, not
    36 per-tensor kernels (and not the per-op fallback path)."""
    monkeypatch.setattr(mmemo_b200.ren_mme, "DROP", 0.1)
    m = mmemo_b200.ren_mme.Base_model(16, 4, 5, 6, 2, 1, 1, l_dim=8, v_dim=6, a_dim=7).train()
    g = torch.Generator().manual_seed(0)
    inputs = []
    for L, D in ((4, 8), (5, 6), (6, 7)):
        for _ in range(2):
            inputs += [torch.randn(2, L, D, generator=g), torch.ones(2, L)]
    with mmemo_b200.precision("bf16"):
        out = m(*inputs)
        ops.circle_loss_op(out, torch.zeros(2, 9)).mean().backward()
    assert sum(n == "mmemo_dropout_multi_bf16" for n in stub) == 4     # 2 sites x (fwd + bwd)
    assert not any(n in ("mmemo_dropout_bf16", "mmemo_dropout_f32") for n in stub)
    assert sum(n.startswith("mmemo_resattn") for n in stub) == 2
    assert len(stub) <= 60, (len(stub), stub)
