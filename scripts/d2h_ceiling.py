"""Raw device->host ceiling of the box: every rank copies one step's observations (86.7 MB at 4096 envs) from HBM into
pinned host memory, all ranks at once. Launch: torchrun --nproc-per-node N scripts/d2h_ceiling.py  (or plain python)."""
import json
import os
import time

import torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
nbytes = 4096 * 84 * 84 * 3
src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev)
out = {}
for label, bufs in (("one pinned buffer", 1), ("two pinned buffers, alternating", 2)):
    dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(bufs)]
    for i in range(5):
        dst[i % bufs].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    reps = 60
    t0 = time.perf_counter()
    for i in range(reps):
        dst[i % bufs].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    out[label] = {"GB/s total": world * reps * nbytes / secs / 1e9, "GB/s per GPU": reps * nbytes / secs / 1e9,
                  "M env-steps/s if obs-bound": world * reps * 4096 / secs / 1e6}
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "cpu_count": os.cpu_count(), "results": out}))
if dist is not None:
    dist.destroy_process_group()
