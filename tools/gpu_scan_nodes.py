"""DEV TOOL: distribution of stage-DP search effort over agent blocks (what each rank of an 8-GPU run gets)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.batch import BatchMpc
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
B, N_p = 100, 48
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for blk in range(8):
    allnodes, times = [], []
    for k0 in range(0, 32, 4):
        wl = syn.dewh_batch(B, N_p, seed=1, k0=k0, first_agent=blk * B)
        Nt = wl["Nt"]
        cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = wl["q_u"]; cost[:, :, 1:] = wl["q_mu"][:, None, :]
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device="cuda:0", solver="stage_dp", dp_opts=cabi.stage_dp_default_opts(cells=cells))
        bm.build()
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1)); e1.record(); torch.cuda.synchronize()
        allnodes.append(res["stats"][:, 0].cpu().numpy()); times.append(e0.elapsed_time(e1))
    n = np.concatenate(allnodes)
    per_step_max = [int(a.max()) for a in allnodes]
    print("agents %3d-%3d: expansions mean %.1f p99 %.0f max %d; per-instant max %s; solve ms %s" % (
        blk * B, blk * B + B - 1, n.mean(), np.percentile(n, 99), n.max(), per_step_max, ["%.2f" % t for t in times]))
