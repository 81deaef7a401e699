"""Not a test: kernel list of ONE eager cfg-2 step with and without the data-parallel reducer
(torch.profiler, rank 0), to see what the bucket plumbing adds.  Run on the GPU box:
  python tools/dp_profile.py                                   # single process, no reducer
  MMEMO_DP_NO_COMM=1 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      tools/dp_profile.py                                      # reducer, collectives off
"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mmemo_b200
    from mmemo_b200 import ops, synth
    from mmemo_b200 import dp as mdp
    mmemo_b200.set_precision("bf16")
    torch.manual_seed(0)
    model = mmemo_b200.ResidualEncoder(512, 8, 6, 2).to(dev).train()
    b = synth.encoder_batch(seed=1234 + rank, B=64, L=128, d=512)
    x, m = b["x"].to(dev), b["mask"].to(dev)
    red = mdp.GradReducer(model, world) if world > 1 else None

    def step():
        model.zero_grad(set_to_none=True)
        loss = ops.sq_mean_op(model(x, m))
        if red is not None:
            red.backward(loss)
        else:
            loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        agg = defaultdict(lambda: [0, 0.0])
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                a = agg[e.name[:100]]
                a[0] += 1
                a[1] += e.device_time
        tot = sum(v[1] for v in agg.values())
        print(f"world {world}: {sum(v[0] for v in agg.values())} device activities, {tot:.0f} us")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {v[0]:4d} x {v[1] / v[0]:8.1f} us = {v[1]:8.1f}  {k}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
