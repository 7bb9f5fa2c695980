"""Diagnostic: which part of the end-to-end step stalls with >= 4 NCCL ranks."""
import os, sys, time, datetime, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
import fitter_mc_b200 as fmc
import datasets, oracle_lib as oracle
variant = os.environ.get("V", "full")
faulthandler.dump_traceback_later(int(os.environ.get("WD", "45")), exit=True)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
model, x, y, err = datasets.dataset("c2")
start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
ctx = fmc.Context(lr); ctx.set_data(model, x, y, err)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
acc = torch.zeros(ctx.accumulator_len(), dtype=torch.float64, device=dev); ctx.set_accumulator_dev(acc.data_ptr())
nres = 200000
t0 = time.time()
for k in range(12):
    if variant in ("full", "nocpu"):
        ctx.set_data(model, x, y, err)
    ctx.launch(start, nres, first_counter=(k * world + rank) * nres)
    dist.all_reduce(acc)
    if variant in ("full", "noset"):
        h = acc.cpu()
    else:
        stream.synchronize()
    ctx.wait()
    if rank == 0: print(variant, "step", k, "ok %.2fs" % (time.time() - t0), flush=True)
dist.barrier()
if rank == 0: print(variant, "DONE", flush=True)
dist.destroy_process_group()
