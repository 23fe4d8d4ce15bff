"""DEV TOOL (gpurun): BASELINE configs[3] (10,000 DEWHs, 96 closed-loop steps) with the per-step device time of the
batch solve next to the wall time of the loop."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

dev = torch.device("cuda:0")
N_p, Nt, total, steps = 48, 49, int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 96
base = [syn.dewh_agent_params(a) for a in range(256)]
fleet = DewhFleet([base[b % 256] for b in range(total)], N_p, device=dev)
T0 = np.random.default_rng(1).integers(55, 65, size=total).astype(float)
prof = np.stack([syn.dhw_demand_profile(steps + Nt, seed=b) for b in range(256)])
demand = prof[np.arange(total) % 256]
price = syn.price_profile(steps + Nt, seed=2)
fleet.closed_loop(T0, demand, price, 2)
torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    log = fleet.closed_loop(T0, demand, price, steps)
    e1.record()
    torch.cuda.synchronize()
    sm = log["solve_ms"].cpu().numpy()
    print("rep %d: events %.1f ms, host wall %.1f ms | solve_ms sum %.1f mean %.2f max %.2f at step %d | not optimal %d" % (
        rep, e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, sm.sum(), sm.mean(), sm.max(), int(sm.argmax()),
        int((log["status"] != 0).sum())))
    print("   solve_ms by step:", " ".join("%.1f" % x for x in sm))
