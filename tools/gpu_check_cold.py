"""DEV TOOL: stage-DP search effort when slack penalties are unavoidable (cold start below T_min)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.batch import BatchMpc
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
B, N_p = 100, 48
wl = syn.dewh_batch(B, N_p, seed=1)
Nt = wl["Nt"]
cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = wl["q_u"]; cost[:, :, 1:] = wl["q_mu"][:, None, :]
for dT in (0.0, -8.0, -12.0, -20.0):
    x0 = wl["x0"] * 0 + 50.0 + dT if dT else wl["x0"]
    for solver in ("stage_dp", "bnc"):
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device="cuda:0", solver=solver,
                      dp_opts=cabi.stage_dp_default_opts(max_nodes=300000), opts=cabi.default_opts(max_nodes=20000, max_pivots=200000))
        bm.build()
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = bm.solve(x0, wl["omega"], cost_v=cost.reshape(B, -1)); e1.record(); torch.cuda.synchronize()
        st = res["stats"].cpu().numpy(); s = res["status"].cpu().numpy()
        print("x0 = %s %-8s: %.3f ms, status %s, nodes mean %.1f max %d, obj sum %.6f" % (
            "nominal" if not dT else "T_min%+.0f" % dT, solver, e0.elapsed_time(e1), np.bincount(s, minlength=4).tolist(),
            st[:, 0].mean(), st[:, 0].max(), res["obj"][torch.isfinite(res["obj"])].sum().item()))
