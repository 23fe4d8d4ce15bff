"""DEV TOOL: stage-DP search effort when the tariff is piecewise constant (exact ties between heating slots)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.batch import BatchMpc
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import TOU_LEVELS

B, N_p = 100, 48
wl = syn.dewh_batch(B, N_p, seed=1)
Nt = wl["Nt"]
k = np.arange(Nt) % 96
hour = k / 4.0
level = np.full(Nt, TOU_LEVELS["low_off_peak"])
stnd = ((hour >= 6) & (hour < 7)) | ((hour >= 10) & (hour < 18)) | ((hour >= 20) & (hour < 22))
peak = ((hour >= 7) & (hour < 10)) | ((hour >= 18) & (hour < 20))
level[stnd] = TOU_LEVELS["low_stnd"]; level[peak] = TOU_LEVELS["low_peak"]
price = level / 3600.0 / 100.0 / 1000.0 * 900.0
P = np.array([p["P_h_Nom"] for p in wl["params"]])
q_u = price[None, :] * P[:, None]
cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = q_u; cost[:, :, 1] = 10 * q_u.sum(1)[:, None]; cost[:, :, 2] = q_u.sum(1)[:, None]
for gap, fp64 in ((0.0, 0), (0.0, 1), (1e-5, 0)):
    for cells in (8192,):
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device="cuda:0", solver="stage_dp",
                      dp_opts=cabi.stage_dp_default_opts(cells=cells, mip_rel_gap=gap, max_nodes=2000000, table_fp64=fp64))
        bm.build()
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1)); e1.record(); torch.cuda.synchronize()
        st = res["stats"].cpu().numpy(); s = res["status"].cpu().numpy()
        print("flat TOU tariff: table %s gap %.0e cells %d: %.3f ms, status %s, nodes mean %.1f p50 %.0f max %d, obj sum %.9f" % (
            "fp64" if fp64 else "fp32", gap, cells, e0.elapsed_time(e1), np.bincount(s, minlength=3).tolist(), st[:, 0].mean(), np.median(st[:, 0]), st[:, 0].max(),
            res["obj"].sum().item()))
