"""torchrun --nproc-per-node N tools/test_allreduce.py — checks libmmemo's symmetric-memory
all-reduce (NVLS multicast and plain peer path) against NCCL and times both."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmemo_b200 import ops  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    n = 12_800_000 // 1024 * 1024            # ~51 MB of float32
    flat = symm_mem.empty(n, dtype=torch.float32, device=dev)
    h = symm_mem.rendezvous(flat, dist.group.WORLD)
    if rank == 0:
        print(f"world {world} multicast_ptr {h.multicast_ptr:#x} pad {h.signal_pad_size}", flush=True)
    slot_base = h.signal_pad_size // 8
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    src = torch.randn(n, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize()
    dist.barrier()
    s = torch.cuda.current_stream().cuda_stream

    def ar(mc, off, cnt, blocks):
        ops._call("mmemo_allreduce_sum_f32", mc or None, h.buffer_ptrs_dev, h.signal_pad_ptrs_dev,
                  slot_base, off, cnt, rank, world, blocks, s)

    for name, mc in (("nvls", h.multicast_ptr), ("peer", 0)):
        if name == "nvls" and not mc:
            continue
        flat.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        ar(mc, 0, n, 16)
        torch.cuda.synchronize(); dist.barrier()
        err = (flat - ref).abs().max().item()
        # sub-range: only [off, off+cnt) may change
        flat.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        off, cnt = 4096, 1024 * 64
        ar(mc, off, cnt, 4)
        torch.cuda.synchronize(); dist.barrier()
        ok_in = (flat[off:off + cnt] - ref[off:off + cnt]).abs().max().item()
        ok_out = torch.equal(flat[:off], src[:off]) and torch.equal(flat[off + cnt:], src[off + cnt:])
        print(f"[rank {rank}] {name}: full max|err| {err:.3e}  sub {ok_in:.3e} untouched {ok_out}",
              flush=True)
        for blocks in (4, 8, 16, 32):
            for size in (2 << 20, 8 << 20, n):       # elements: 8 MB, 32 MB, 51 MB
                size = min(size, n)
                for _ in range(3):
                    ar(mc, 0, size, blocks)
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    ar(mc, 0, size, blocks)
                e1.record()
                torch.cuda.synchronize()
                if rank == 0:
                    print(f"  {name} blocks {blocks:2d} {size * 4 / 1e6:6.1f} MB: "
                          f"{e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
    for size in (2 << 20, 8 << 20, n):
        size = min(size, n)
        t = flat[:size]
        for _ in range(3):
            dist.all_reduce(t)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(t)
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"  nccl {size * 4 / 1e6:6.1f} MB: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
