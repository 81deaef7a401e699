"""world_size-2 gloo tests (CPU) of the data-parallel host logic in mmemo_b200/dp.py: sharding, the
bucket layout, unused-parameter handling and that overlapped bucketed all-reduce reproduces the
single-process full-batch gradients (tolerance 1e-5)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmemo_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(12, 40)
        self.b = torch.nn.Linear(40, 40)
        self.unused = torch.nn.Parameter(torch.zeros(3))      # never receives a gradient
        self.c = torch.nn.Linear(40, 5)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


class UnitNet(Net):
    """``b``'s gradients are declared as one contiguous unit (the protocol of the fused blocks,
    ``mmemo_grad_unit``): [pad 7 | bias 40 | pad 17 | weight 1600 | pad 8], key = weight."""

    def mmemo_grad_unit(self):
        return self.b.weight, [(self.b.bias, 7), (self.b.weight, 64), (self.unused, 3)], 1672


def _loss(model, x, y):
    return ((model(x) - y) ** 2).mean()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = (UnitNet if os.environ.get("MMEMO_TEST_UNITS") == "1" else Net)()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(16, 12, generator=g), torch.randn(16, 5, generator=g)
    red = dp.GradReducer(model, world, bucket_bytes=4096)
    shard = dp.shard_batch({"x": x, "y": y}, rank, world, align=2)
    layouts, grads = [], None
    for step in range(3):                       # step 0 builds the buckets, 1-2 overlap
        model.zero_grad(set_to_none=True)
        red.backward(_loss(model, shard["x"], shard["y"]))
        layouts.append(red.bucket_layout())
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    if isinstance(model, UnitNet):       # the unit sits in ONE bucket with the declared offsets
        bk = [b for b in red.buckets if b.units]
        assert len(bk) == 1 and len(bk[0].units) == 1
        unit, start = bk[0].units[0]
        assert unit.total == 1672 and start % 32 == 0
        base = bk[0].flat.data_ptr()
        assert model.b.bias.grad.data_ptr() == base + 4 * (start + 7)
        assert model.b.weight.grad.data_ptr() == base + 4 * (start + 64)
        assert all(p is not model.unused for p in bk[0].params)     # no gradient: left out
    q.put((rank, layouts[-1], {k: v.tolist() for k, v in grads.items()}))
    dist.destroy_process_group()


@pytest.mark.parametrize("units", [False, True])
def test_bucketed_allreduce_matches_full_batch(units, monkeypatch):
    monkeypatch.setenv("MMEMO_TEST_UNITS", "1" if units else "0")   # inherited by the spawned ranks
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    assert res[0][1] == res[1][1] and len(res[0][1]) >= 2          # same multi-bucket layout
    torch.manual_seed(0)
    model = Net()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(16, 12, generator=g), torch.randn(16, 5, generator=g)
    _loss(model, x, y).backward()
    for k, p in model.named_parameters():
        if p.grad is None:
            assert k not in res[0][2] and k == "unused"
            continue
        for r in res:
            got = torch.tensor(r[2][k])
            assert torch.allclose(got, p.grad, rtol=1e-5, atol=1e-7), k


def test_shard_bounds_keep_rdrop_pairs_together():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            s = dp.shard_bounds(256, r, world, align=2)
            assert s.start % 2 == 0 and (s.stop - s.start) == 256 // world
            seen += list(range(s.start, s.stop))
        assert seen == list(range(256))
    with pytest.raises(ValueError):
        dp.shard_bounds(6, 0, 4, align=2)


def test_shard_batch_nested():
    b = {"inputs": [torch.arange(8), torch.arange(16).view(8, 2)], "label": torch.arange(8)}
    s = dp.shard_batch(b, 1, 2, align=2)
    assert s["inputs"][0].tolist() == [4, 5, 6, 7] and s["inputs"][1].shape == (4, 2)


def test_grad_reducer_refuses_gradient_accumulation():
    """after a step p.grad is a view of its all-reduce bucket (already averaged); a
    second backward() without zero_grad(set_to_none=True) would re-reduce the old contribution.
    The reducer raises instead of training on silently wrong gradients."""
    import socket
    import pytest
    import torch
    import torch.distributed as dist
    from mmemo_b200 import dp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        model = torch.nn.Linear(4, 3)
        red = dp.GradReducer(model, world_size=2, bucket_bytes=1 << 10)
        x = torch.randn(5, 4)
        red.backward(model(x).sum())              # first step: discovers the buckets
        model.zero_grad(set_to_none=True)
        red.backward(model(x).sum())
        with pytest.raises(RuntimeError, match="zero_grad"):
            red.backward(model(x).sum())          # grads of the previous step still attached
        model.zero_grad(set_to_none=True)
        red.backward(model(x).sum())              # fine again
    finally:
        dist.destroy_process_group()
