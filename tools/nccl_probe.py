"""Minimal NCCL sanity probe under torchrun: init, one tiny all-reduce, barrier, timings."""
import os, sys, time, datetime
import torch, torch.distributed as dist
t0 = time.time()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
t1 = time.time()
x = torch.ones(31, dtype=torch.float64, device=dev) * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
t2 = time.time()
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
y = torch.ones(31, dtype=torch.float64, device=dev)
dist.all_reduce(y)
s.synchronize()
dist.barrier()
t3 = time.time()
if rank == 0:
    print("nccl probe ok: world %d sum %.0f ; init %.1fs first allreduce %.1fs second+barrier %.1fs NVLS=%s" %
          (world, x[0].item(), t1 - t0, t2 - t1, t3 - t2, os.environ.get("NCCL_NVLS_ENABLE")), flush=True)
dist.destroy_process_group()
