"""DEV TOOL (gpurun): where the stage-DP solve spends its time on the bench workload -- per-agent device times of the
table kernel's phases (load + set-up, sweep) and of the search, from the kernels' own %globaltimer stamps (stats
columns 1, 2, 4), next to CUDA-event times of the whole call, for a few option sets."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn

dev = torch.device("cuda:0")
N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
wl = syn.dewh_batch(B, N_p, seed=1)
Nt = wl["Nt"]
d = cabi.make_dims(B, Nt, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
mats = {k: torch.tensor(v, dtype=torch.float64, device=dev) for k, v in wl["mats"].items()}
mats["C"] = torch.ones((1, 1, 1), dtype=torch.float64, device=dev)
evo = cabi.condense(d, mats)
x0 = torch.tensor(wl["x0"], dtype=torch.float64, device=dev)
w = torch.tensor(wl["omega"], dtype=torch.float64, device=dev)
rhs = cabi.constraint_rhs(d, evo, x0, w)
nvt = d.nv * Nt
cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = wl["q_u"]; cost[:, :, 1:] = wl["q_mu"][:, None, :]
cost_t = torch.tensor(cost.reshape(B, nvt), dtype=torch.float64, device=dev)
lb = torch.zeros(nvt, dtype=torch.float64, device=dev)
ub = torch.tensor(np.tile([1.0, np.inf, np.inf], Nt), dtype=torch.float64, device=dev)
isb = torch.tensor(np.tile([1, 0, 0], Nt).astype(np.uint8), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for name, kw in [("auto 4096", dict(cells=4096)), ("fused 8192", dict(cells=8192, fuse_search=1)), ("two kernels 8192", dict(cells=8192, fuse_search=0)),
                 ("fused 4096", dict(cells=4096, fuse_search=1)), ("fused 2048", dict(cells=2048, fuse_search=1)),
                 ("fused 8192 fp32", dict(cells=8192, fuse_search=1, table_fp64=0)),
                 ("fused linear 4096", dict(cells=4096, fuse_search=1, bound=1))]:
    o = cabi.stage_dp_default_opts(**kw)
    ts = []
    for rep in range(6):
        if not os.environ.get("DP_NOFLUSH"):
            flush.fill_(rep)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        v, obj, status, stats = cabi.stage_dp_solve(d, mats, rhs, cost_t, lb, ub, isb, o)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    st = stats.cpu().numpy().astype(float)
    ob = obj.cpu().numpy()
    if ref is None:
        ref = ob
    print("%-20s events (incl. launch) min %.1f us | per agent: set-up %.1f (max %.1f)  sweep %.1f (max %.1f)  search %.1f (max %.1f) us | "
          "nodes mean %.1f max %d | status %s | max |obj - ref| %.2e" % (
              name, min(ts), st[:, 1].mean() / 10, st[:, 1].max() / 10, st[:, 2].mean() / 10, st[:, 2].max() / 10,
              st[:, 4].mean() / 10, st[:, 4].max() / 10, st[:, 0].mean(), st[:, 0].max(),
              np.bincount(status.cpu().numpy(), minlength=3).tolist(), np.abs(ob - ref).max()), flush=True)
