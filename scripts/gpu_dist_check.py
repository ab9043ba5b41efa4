"""torchrun --nproc-per-node G scripts/gpu_dist_check.py [nx ns]: channel-sharded f-k filter over G
GPUs (NCCL all-to-all) against the single-GPU result of the same matrix; prints error and timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import das4whales_b200 as dw
from das4whales_b200 import synth, dist as d4wdist
from das4whales_b200.fk import FkFilter
rank, world = dist.get_rank(), dist.get_world_size()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 24000
mask = dw.dsp.fk_filter_design((nx, ns), [0, nx, 1], 2.0419046878814697, 200.0)
x = synth.synth_strain(nx, ns, seed=7)                     # same seed on every rank -> same matrix
cpr = nx // world
be = d4wdist.CudaBackend(mask, nx, ns, world)
flt = d4wdist.ShardedFkFilter(nx, ns, be)
xl = x[rank * cpr:(rank + 1) * cpr].contiguous()
y_sh = flt(xl, tapering=True)
y_ref = FkFilter(mask)(x, tapering=True)[rank * cpr:(rank + 1) * cpr]
err = float((y_sh - y_ref).abs().max() / y_ref.abs().max())
for _ in range(2):
    flt(xl)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    flt(xl)
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / reps
t = torch.tensor([err, dt], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"sharded fk {nx}x{ns} over {world} GPUs: max rel diff vs single GPU {t[0].item():.2e}, "
          f"{t[1].item()*1e3:.2f} ms/step -> {nx / t[1].item():.0f} channels/s", flush=True)
dist.destroy_process_group()
