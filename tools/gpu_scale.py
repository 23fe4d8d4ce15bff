"""DEV TOOL: large-batch behaviour of K1 (HBM roofline) and of the stage-DP solve (gpurun)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn

dev = torch.device("cuda:0")
N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
Bs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["100", "1000", "10000"])]
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
base = syn.dewh_batch(256, N_p, seed=1)
Nt = base["Nt"]
out = {}
for B in Bs:
    rep = (B + 255) // 256
    tile = lambda a: np.concatenate([a] * rep, axis=0)[:B]
    mats = {k: torch.tensor(tile(v), dtype=torch.float64, device=dev) for k, v in base["mats"].items()}
    mats["C"] = torch.ones((1, 1, 1), dtype=torch.float64, device=dev)
    d = cabi.make_dims(B, Nt, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
    res = {}
    for label, want in (("all12", cabi.EVO_NAMES), ("H4", ("H_x", "H_v", "H_omega", "H_5"))):
        evo = cabi.condense(d, mats, want=want)
        nbytes = sum(evo[k].numel() * 8 for k in want)
        ts = []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); cabi.condense(d, mats, want=want, out=evo); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        res[label] = dict(ms=ms, GBs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peaks["hbm_gbs"])
        print("K1 %s B=%d: %.3f ms, %.1f GB/s = %.3f of measured HBM peak" % (label, B, ms, res[label]["GBs"], res[label]["frac"]))
        if label == "all12":
            del evo
    x0 = torch.tensor(tile(base["x0"]), dtype=torch.float64, device=dev)
    w = torch.tensor(tile(base["omega"]), dtype=torch.float64, device=dev)
    rhs = cabi.constraint_rhs(d, evo, x0, w)
    nvt = 3 * Nt
    cost = np.zeros((256, Nt, 3)); cost[:, :, 0] = base["q_u"]; cost[:, :, 1:] = base["q_mu"][:, None, :]
    cost_t = torch.tensor(tile(cost.reshape(256, nvt)), dtype=torch.float64, device=dev)
    lb = torch.zeros(nvt, dtype=torch.float64, device=dev)
    ub = torch.tensor(np.tile([1.0, np.inf, np.inf], Nt), dtype=torch.float64, device=dev)
    isb = torch.tensor(np.tile([1, 0, 0], Nt).astype(np.uint8), device=dev)
    o = cabi.stage_dp_default_opts(cells=cells, table_fp64=int(os.environ.get("TABLE_FP64", "0")))
    ts = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); v, obj, st, stats = cabi.stage_dp_solve(d, mats, rhs, cost_t, lb, ub, isb, o); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    ok = int((st == 0).sum())
    kf = float(stats[:, 7].double().sum()) * 1024 * 2
    res["stage_dp"] = dict(ms=ms, solves_per_s=B / ms * 1e3, optimal=ok, tflops=kf / ms / 1e9)
    print("stage_dp B=%d G=%d: %.3f ms -> %.0f solves/s, optimal %d/%d, %.2f algorithmic TFLOP/s, table write %.1f GB/s" % (
        B, cells, ms, B / ms * 1e3, ok, B, kf / ms / 1e9, B * (Nt - 1) * cells * 4 / ms / 1e6))
    out[B] = res
    del evo, mats
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_scale_%d.json" % N_p, "w"), indent=1)
