"""DEV TOOL (gpurun): K1 on 8192 DEWH agents (the four H matrices, 1.27 GB per launch) -- the launch the bench's
roofline_condense_large_batch times; meant to be run under ncu."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
dev = torch.device("cuda:0")
B, N_p = 8192, 48
wl = syn.dewh_batch(256, N_p, seed=1)
Nt = wl["Nt"]
d = cabi.make_dims(B, Nt, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
mats = {k: torch.tensor(np.concatenate([v] * 32, axis=0)[:B], dtype=torch.float64, device=dev) for k, v in wl["mats"].items()}
mats["C"] = torch.ones((1, 1, 1), dtype=torch.float64, device=dev)
want = ("H_x", "H_v", "H_omega", "H_5")
for rep in range(3):
    evo = cabi.condense(d, mats, want=want)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for rep in range(5):
    evo = cabi.condense(d, mats, want=want)
e1.record(); torch.cuda.synchronize()
nbytes = sum(evo[k].numel() * 8 for k in want)
ms = e0.elapsed_time(e1) / 5
print("K1 %d agents: %.3f ms, %.0f GB/s written (%d bytes)" % (B, ms, nbytes / ms / 1e6, nbytes))
