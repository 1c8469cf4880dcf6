"""Probe: what the host of an N-GPU box can move.  Every rank copies pinned host memory up and down at the same time
(two streams), all ranks together; prints per-rank and aggregate GB/s.  Run under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30   # floats: 4 GiB each way
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out.fill_(0.0)
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def go(up, down):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = 3 * 4 * n * (int(up) + int(down)) / dt / 1e9
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t)
    if rank == 0: print("up=%d down=%d: rank0 %.1f GB/s, all ranks %.1f GB/s" % (up, down, gbs, float(t)), flush=True)
go(1, 0); go(0, 1); go(1, 1)
if world > 1: dist.destroy_process_group()
