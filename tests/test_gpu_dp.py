"""GPU tests of the data-parallel gradient exchange: libmmemo's symmetric-memory all-reduce kernel
(csrc/allreduce.cu, NVLS multicast and plain peer path) and GradReducer on top of it.  The
single-GPU cases run a world of one; the two-rank cases need two GPUs and skip otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _allreduce_worker(rank, world, port, q):
    import torch.distributed._symmetric_memory as symm_mem

    from mmemo_b200 import ops
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        n = 1024 * 300
        flat = symm_mem.empty(n, dtype=torch.float32, device=dev)
        h = symm_mem.rendezvous(flat, dist.group.WORLD)
        src = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
        ref = src.clone()
        dist.all_reduce(ref)
        out = {"multicast": bool(h.multicast_ptr)}
        for name, mc in (("nvls", h.multicast_ptr), ("peer", 0)):
            if name == "nvls" and not mc:
                continue
            flat.copy_(src)
            torch.cuda.synchronize()
            dist.barrier()
            off, cnt = 2048, 1024 * 200
            ops._call("mmemo_allreduce_sum_f32", mc or None, h.buffer_ptrs_dev,
                      h.signal_pad_ptrs_dev, h.signal_pad_size // 8, off, cnt, rank, world, 4,
                      torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            dist.barrier()
            out[name] = (
                float((flat[off:off + cnt] - ref[off:off + cnt]).abs().max()),
                bool(torch.equal(flat[:off], src[:off])
                     and torch.equal(flat[off + cnt:], src[off + cnt:])))
        q.put((rank, out))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    os._exit(0)


def _reducer_worker(rank, world, port, q):
    from mmemo_b200 import dp, synth
    from mmemo_b200.encoder import ResidualEncoder
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        torch.manual_seed(0)
        model = ResidualEncoder(64, 4, 2)
        model.load_state_dict(synth.randomize_gates(model.state_dict(), seed=3))
        model = model.to(dev)
        batch = synth.encoder_batch(seed=11, B=8, L=16, d=64)
        x, mask, dy = batch["x"], batch["mask"], batch["dy"]
        sl = dp.shard_bounds(8, rank, world)
        grads = {}
        for transport in ("nccl", "symm"):
            red = dp.GradReducer(model, world, bucket_bytes=64 << 10, transport=transport)
            for _ in range(3):                      # first step builds the buckets, then overlapped
                model.zero_grad(set_to_none=True)
                out = model(x[sl].to(dev), mask[sl].to(dev))
                red.backward((out * dy[sl].to(dev)).mean())
            torch.cuda.synchronize()
            grads[transport] = {k: p.grad.detach().cpu().clone()
                                for k, p in model.named_parameters() if p.grad is not None}
            red.remove()
        def rel(k):
            return float((grads["nccl"][k] - grads["symm"][k]).abs().max()
                         / grads["nccl"][k].abs().max().clamp_min(1e-20))
        # scalar gates are one atomics-ordered sum with cancellation: two backward passes of the
        # SAME transport differ by ~1e-5 there, so they are compared at a tenth of the weight
        worst = max(rel(k) / (10.0 if grads["nccl"][k].numel() == 1 else 1.0) for k in grads["nccl"])
        q.put((rank, worst))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    os._exit(0)


def _run(worker, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    return res


@pytest.mark.parametrize("world", [1, 2])
def test_symmetric_allreduce_kernel(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    res = _run(_allreduce_worker, world)
    for rank, out in res.items():
        assert isinstance(out, dict), out
        for name in ("nvls", "peer"):
            if name in out:
                err, untouched = out[name]
                assert err <= 1e-5, (rank, name, err)
                assert untouched, (rank, name)


def test_grad_reducer_symm_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(_reducer_worker, 2)
    for rank, worst in res.items():
        assert isinstance(worst, float), worst
        assert worst <= 1e-5, (rank, worst)


def test_full_block_gradients_live_in_the_bucket_single_gpu():
    """The reducer's bucket plumbing without any collective (no_comm, a pretended world of two, one
    GPU): the fused full block's gradient buffer is a bucket region (``mmemo_grad_unit``), so after
    the first step no parameter gradient is copied, the buckets hold exactly the local gradients /
    world, and nothing in the step fills or copies per block."""
    import mmemo_b200
    from mmemo_b200 import dp as mdp, ops, synth

    dev = torch.device("cuda")
    mmemo_b200.set_precision("bf16")
    try:
        torch.manual_seed(0)
        model = mmemo_b200.ResidualEncoder(128, 2, 3, 2)
        sd = synth.randomize_gates({k: v.detach().clone() for k, v in model.state_dict().items()})
        model.load_state_dict(sd)
        model = model.to(dev).train()
        b = synth.encoder_batch(seed=5, B=4, L=128, d=128)
        x, m = b["x"].to(dev), b["mask"].to(dev)

        def loss_fn():
            return ops.sq_mean_op(model(x, m))

        model.zero_grad(set_to_none=True)
        loss_fn().backward()
        ref = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

        red = mdp.GradReducer(model, world_size=2, transport="nccl", bucket_bytes=1 << 19)
        red.no_comm = True
        for step in range(3):
            model.zero_grad(set_to_none=True)
            red.backward(loss_fn())
            if step == 0:
                continue
            assert len(red.buckets) >= 2 and sum(len(bk.units) for bk in red.buckets) == 3
            for bk in red.buckets:
                lo = bk.flat.data_ptr()
                hi = lo + 4 * bk.flat.numel()
                for p in bk.params:
                    assert lo <= p.grad.data_ptr() < hi           # a view of the bucket: zero-copy
            for n, p in model.named_parameters():
                if n in ref:
                    err = (2 * p.grad - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-12)
                    assert err < 2e-2, (n, float(err))            # bf16 run-to-run (atomics order)
        n0 = ops.launch_count
        model.zero_grad(set_to_none=True)
        red.backward(loss_fn())
        model.zero_grad(set_to_none=True)
        plain = ops.launch_count
        loss_fn().backward()
        assert ops.launch_count - plain == plain - n0            # same libmmemo launches as plain
        red.remove()
        ops.clear_grad_dest()
    finally:
        mmemo_b200.set_precision("fp32")


def test_grouped_trunk_gradients_live_in_the_bucket_single_gpu():
    """Same for the grouped fusion trunk (lite blocks of cmu-mosei / Ren-MME): every block's zero
    buffer of a trunk layer is a bucket region; gradients equal the plain backward's / world."""
    import mmemo_b200
    from mmemo_b200 import dp as mdp, ops, synth

    dev = torch.device("cuda")
    mmemo_b200.set_precision("bf16")
    try:
        torch.manual_seed(1)
        ct = mmemo_b200.cmu_mosei.Concat_Trans(32, 12, 20, 28, 2, 2, 1, l_dim=24, v_dim=16, a_dim=8)
        sd = synth.randomize_gates({k: v.detach().clone() for k, v in ct.state_dict().items()})
        ct.load_state_dict(sd)
        ct = ct.to(dev).train()
        mb = synth.mosei_batch(seed=5, B=4, L=(12, 20, 28), D=(24, 16, 8))
        args = [mb[k].to(dev) for k in ("l", "v", "a", "l_mask", "v_mask", "a_mask")]
        label = mb["label"].to(dev)

        def loss_fn():
            return ops.circle_loss_op(ct(*args), label).mean()

        ct.zero_grad(set_to_none=True)
        loss_fn().backward()
        ref = {n: p.grad.detach().clone() for n, p in ct.named_parameters() if p.grad is not None}
        red = mdp.GradReducer(ct, world_size=2, transport="nccl", bucket_bytes=1 << 16)
        red.no_comm = True
        for step in range(3):
            ct.zero_grad(set_to_none=True)
            red.backward(loss_fn())
        n_units = sum(len(bk.units) for bk in red.buckets)
        n_blocks = sum(1 for m in ct.modules() if hasattr(m, "mmemo_grad_unit"))
        assert n_units == n_blocks and n_units >= 18
        in_bucket = 0
        for bk in red.buckets:
            lo = bk.flat.data_ptr()
            hi = lo + 4 * bk.flat.numel()
            for p in bk.params:
                assert lo <= p.grad.data_ptr() < hi
                in_bucket += 1
        assert in_bucket == len(ref)
        for n, p in ct.named_parameters():
            if n in ref:
                err = (2 * p.grad - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-12)
                assert err < 3e-2, (n, float(err))
        red.remove()
        ops.clear_grad_dest()
    finally:
        mmemo_b200.set_precision("fp32")
