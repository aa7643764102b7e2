#!/usr/bin/env python
"""Aggregate host-to-device ceiling of one box: every rank copies its own pinned 1.455 GB fp32 buffer (the bench
input of one step) to its GPU, all ranks at once.  torchrun --nproc-per-node N tools/h2d_probe_mp.py"""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, F = 38400, 9472
h = torch.empty(N, F, dtype=torch.float32).pin_memory()
d = torch.empty(N, F, dtype=torch.float32, device="cuda")
d.copy_(h, non_blocking=True); torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gb = world * 5 * N * F * 4 / 1e9
    print(f"h2d ceiling: {world} GPUs, {gb / t.item():.1f} GB/s aggregate ({gb / t.item() / world:.1f} per GPU); "
          f"as info Gbit/s of fp32 [N][F] input for J15_L30_Z1280: {gb / t.item() * 1e9 / (N * 4) * 19200 / 1e9:.2f}", flush=True)
if world > 1: dist.destroy_process_group()
