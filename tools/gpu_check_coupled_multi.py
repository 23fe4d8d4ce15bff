"""DEV TOOL (gpurun --gpus N, under torchrun): the centralised micro-grid problem by price coordination
(DewhFleet.coupled_step) with the fleet SHARDED over the ranks of one box -- the per-iteration all-reduce of the
[Nt + 2] sums over NCCL, the best-response descent with its block all-reduces.  Every rank must end with the same
bounds; rank 0 prints them with the wall time and writes gpurun_out/coupled_multi.json.
    torchrun --nproc-per-node N tools/gpu_check_coupled_multi.py [N_h] [N_p] [iters]"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_coupled import _case
from pyhybridcontrol_b200 import distributed
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_h = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N_p = int(sys.argv[2]) if len(sys.argv) > 2 else 48
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
params, T0, dem, price, P, p_other = _case(N_h, N_p, seed=0)
lo, hi = distributed.shard_range(N_h, rank, world)
fleet = DewhFleet(params[lo:hi], N_p, device=dev)
fleet.build()
res = {}
for name, kw in (("dual iterations only", dict(response_passes=0)), ("with best-response descent", dict())):
    fleet.coupled_step(T0[lo:hi], dem[lo:hi], price, p_other, iters=10, rel_gap=1e-2, **kw)       # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = fleet.coupled_step(T0[lo:hi], dem[lo:hi], price, p_other, iters=iters, rel_gap=1e-2, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    vals = torch.tensor([out["lower_bound"], out["upper_bound"], float(out["iterations"])], dtype=torch.float64, device=dev)
    allv = [torch.empty_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    same = all(bool(torch.equal(a, allv[0])) for a in allv)
    res[name] = dict(world=world, agents=N_h, N_p=N_p, lower_bound=out["lower_bound"], upper_bound=out["upper_bound"],
                     gap=out["gap"], iterations=out["iterations"], response_solves=out.get("response_solves"),
                     wall_ms=1e3 * dt, identical_on_all_ranks=same)
    if rank == 0:
        print(name, res[name], flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/coupled_multi_%d.json" % world, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
