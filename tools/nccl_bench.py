"""Not a test: all-reduce timing at the gradient-bucket sizes (run under torchrun on the GPU box)."""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (0.5, 2, 8.4, 25, 50):
    x = torch.ones(int(mb * (1 << 20) / 4), device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        dist.all_reduce(x)
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) / 20 * 1e3
    n = dist.get_world_size()
    if rank == 0:
        print(f"allreduce {mb:5.1f} MB x{n}: {t:7.1f} us  busbw {2*(n-1)/n*mb*1.048576/t*1e3:6.1f} GB/s  (NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')})", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
