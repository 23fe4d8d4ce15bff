"""DEV TOOL (gpurun --gpus N, torchrun): the peer-store aggregate exchange (distributed.PeerExchange, csrc/aggregate.cu)
on REAL peers: every rank publishes random [B, Nt] plans for 200 steps, gathers with lag 0 and lag 1, and compares with
an NCCL all-reduce of the same local sums (bit-identical: the gather adds in rank order, so does the check)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.distributed import PeerExchange

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Nt, B = 49, 100 + rank
ex = PeerExchange(Nt, dev)
rng = np.random.default_rng(10 + rank)
P = torch.as_tensor(rng.uniform(2700, 3300, B)).to(dev)
bad = 0
prev_ref = None
ts = []
for step in range(200):
    u = torch.as_tensor((rng.random((B, Nt)) > 0.5).astype(float)).to(dev)
    local_sum = cabi.aggregate_power(u, P)
    allv = [torch.empty_like(local_sum) for _ in range(world)]
    dist.all_gather(allv, local_sum)
    ref = torch.zeros_like(local_sum)
    for a in allv:
        ref = ref + a
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ex.publish(u, P)
    lagged = ex.gather(lag=1)
    e1.record()
    now = ex.gather(lag=0)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
    if not torch.equal(now, ref):
        bad += 1
    if prev_ref is not None and not torch.equal(lagged, prev_ref):
        bad += 1
    prev_ref = ref
print("rank %d: %d mismatches in 200 steps, error word %d, publish + lagged gather %.1f us (median)" % (
    rank, bad, ex.error(), float(np.median(ts))), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
