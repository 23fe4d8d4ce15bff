"""DEV TOOL (gpurun): BASELINE config 5 with the reference's mpc_sb_full semantics -- ONE solve per agent with the
robust row-min right-hand side over the whole horizon (controller_base.py:442-444) -- for 1000 DEWHs x 32 demand
scenarios at N_p = 24 / 48 / 96, under each value-table bound of the stage-DP kernels (constant cells / linear cells).
Prints and writes gpurun_out/robust.json: ms per step, proven-optimal count, search expansions (mean / max), the largest
certified gap among the unfinished agents, and the agreement of the objectives between the bounds."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi  # noqa: E402
from pyhybridcontrol_b200.batch import BatchMpc  # noqa: E402
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("ROBUST_AGENTS", "1000"))
S = 32
horizons = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["24", "48", "96"])]
max_nodes = int(os.environ.get("ROBUST_MAX_NODES", "4000000"))
gap = float(os.environ.get("ROBUST_GAP", "0"))
out = {}


def ev_time(fn, reps=2):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), r


for N_p in horizons:
    Nt = N_p + 1
    rng = np.random.default_rng(1000 + N_p)
    wl = syn.dewh_batch(256, N_p, seed=5)
    rep = lambda a: np.concatenate([a] * (B // 256 + 1), axis=0)[:B]  # noqa: E731
    mats = {k_: rep(v) for k_, v in wl["mats"].items()}
    scen = rep(wl["omega"])[:, :, None] * rng.uniform(0.5, 1.8, size=(B, Nt, S))
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = rep(wl["q_u"])
    cost[:, :, 1:] = rep(wl["q_mu"])[:, None, :]
    x0 = torch.as_tensor(rep(wl["x0"])).to(dev)
    om = torch.as_tensor(rep(wl["omega"])).to(dev)
    sc = torch.as_tensor(scen).to(dev)
    cst = torch.as_tensor(cost.reshape(B, -1)).to(dev)
    ref_obj = None
    for name, bound in (("nominal_constant", "constant"), ("constant", "constant"), ("linear", "linear")):
        bm = BatchMpc(mats, N_p, nu_l=1, device=dev, dp_bound=bound)
        bm.dp_opts.max_nodes = max_nodes
        if os.environ.get("ROBUST_CELLS"):
            bm.dp_opts.cells = min(int(os.environ["ROBUST_CELLS"]), cabi.stage_dp_max_cells(bm.dims, bm.dp_opts))
        bm.dp_opts.mip_rel_gap = gap
        bm.build(want=("H_x", "H_v", "H_omega", "H_5"))
        if name.startswith("nominal"):
            fn = lambda: bm.solve(x0, om, cost_v=cst)  # noqa: E731
        else:
            fn = lambda: bm.solve(x0, om, cost_v=cst, scenarios=sc)  # noqa: E731
        ms, res = ev_time(fn)
        st = res["status"].cpu().numpy()
        stats = res["stats"].cpu().numpy()
        obj = res["obj"].cpu().numpy()
        key = "N%d_%s" % (N_p, name)
        out[key] = dict(agents=B, scenarios=S, N_p=N_p, cells=int(bm.dp_opts.cells), ms_per_step=ms,
                        optimal=int((st == 0).sum()), node_limit=int((st == 2).sum()),
                        nodes_mean=float(stats[:, 0].mean()), nodes_p99=float(np.percentile(stats[:, 0], 99)),
                        nodes_max=int(stats[:, 0].max()), max_gap=float(stats[:, 6].max()) * 1e-9,
                        obj_sum=float(obj[st == 0].sum()))
        line = "N_p=%d %-17s cells %5d: %9.3f ms, optimal %d/%d, nodes mean %.1f p99 %.0f max %d, worst certified gap %.2e" % (
            N_p, name, bm.dp_opts.cells, ms, out[key]["optimal"], B, out[key]["nodes_mean"], out[key]["nodes_p99"],
            out[key]["nodes_max"], out[key]["max_gap"])
        if name == "constant":
            ref_obj, ref_st = obj, st
        if name == "linear" and ref_obj is not None:
            both = (st == 0) & (ref_st == 0)
            d = np.abs(obj[both] - ref_obj[both]) / np.maximum(1.0, np.abs(ref_obj[both]))
            out[key]["max_rel_obj_diff_vs_constant"] = float(d.max()) if both.any() else None
            # an unfinished search still returns its incumbent: it can only be worse than a proven optimum
            worse = (obj[ref_st == 0] - ref_obj[ref_st == 0]).min() if (ref_st == 0).any() else 0.0
            line += "; vs constant: max rel diff %.2e on %d agents, min (lin - const) %.2e" % (
                d.max() if both.any() else 0.0, int(both.sum()), worse)
        print(line, flush=True)
        del bm
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/robust.json", "w"), indent=1)
