"""Not a test: kernel timeline of one DP step (torch.profiler / CUPTI) to see whether the bucket
all-reduces overlap backward.  Run under torchrun on the GPU box from the repo root."""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import mmemo_b200
from mmemo_b200 import ops, synth, dp
mmemo_b200.set_precision("bf16")
torch.manual_seed(0)
model = mmemo_b200.ResidualEncoder(512, 8, 6, 2)
model.load_state_dict(synth.randomize_gates({k: v.detach().clone() for k, v in model.state_dict().items()}))
model = model.to(dev).train()
b = synth.encoder_batch(seed=1 + rank, B=64, L=128, d=512)
x, m = b["x"].to(dev), b["mask"].to(dev)
red = dp.GradReducer(model, world, bucket_bytes=int(float(os.environ.get("MMEMO_BUCKET_MB", "8")) * (1 << 20)))
def step():
    model.zero_grad(set_to_none=True)
    out = model(x, m)
    red.backward((out.float() ** 2).mean())
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
ops.clear_shadow_cache(); model.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
for _ in range(3): g.replay()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay(); g.replay()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    half = len(ev) // 2
    ev = ev[half:]                      # second replay
    t0 = ev[0].time_range.start
    tot = ev[-1].time_range.end - t0
    nccl = [e for e in ev if "nccl" in e.name.lower()]
    print(f"kernels in one step: {len(ev)}, span {tot:.0f} us, nccl kernels {len(nccl)}")
    for e in nccl:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        ov = [k for k in ev if k is not e and "nccl" not in k.name.lower() and k.time_range.start < e.time_range.end and k.time_range.end > e.time_range.start]
        busy = sum(min(k.time_range.end, e.time_range.end) - max(k.time_range.start, e.time_range.start) for k in ov)
        print(f"  nccl start {s:8.0f} us dur {d:7.0f} us  overlapping compute kernels {len(ov):3d} (busy {busy:6.0f} us)  {e.name[:50]}")
    # slowest compute kernels during vs outside nccl
    import collections
    agg = collections.defaultdict(lambda: [0, 0.0, 0, 0.0])
    for k in ev:
        if "nccl" in k.name.lower(): continue
        during = any(k.time_range.start < e.time_range.end and k.time_range.end > e.time_range.start for e in nccl)
        a = agg[k.name[:40]]
        dur = k.time_range.end - k.time_range.start
        if during: a[2] += 1; a[3] += dur
        else: a[0] += 1; a[1] += dur
    for n, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][3]))[:8]:
        print(f"  {n:40s} alone n={a[0]:3d} avg {a[1]/max(a[0],1):6.1f} us | during nccl n={a[2]:3d} avg {a[3]/max(a[2],1):6.1f} us")
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
