"""DEV TOOL: trajectory of the price coordination on the GPU (bounds every few iterations) for one synthetic case."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_coupled import _case
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

N_h, N_p = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
params, T0, dem, price, P, p_other = _case(N_h, N_p, seed=0)
fleet = DewhFleet(params, N_p, device="cuda")
fleet.build()
for every in (25,):
    torch.cuda.synchronize(); t = time.time()
    out = fleet.coupled_step(T0, dem, price, p_other, iters=iters, rel_gap=1e-2, check_every=every)
    torch.cuda.synchronize()
    print("check_every", every, {k: v for k, v in out.items() if k not in ("u", "plan", "lam")}, "%.1f ms" % ((time.time() - t) * 1e3))
