"""DEV TOOL: first-light check of every kernel against the oracle on a real B200 (run under gpurun)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from oracle import mld as omld, condense as oc, assemble as oa, solve as osv

dev = torch.device("cuda:0")
print(cabi.device_info())
N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else B
wl = syn.dewh_batch(B, N_p, seed=1)
Nt = wl["Nt"]
d = cabi.make_dims(B, Nt, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
mats = {k: torch.tensor(v, dtype=torch.float64, device=dev) for k, v in wl["mats"].items()}
mats["C"] = torch.ones((1, 1, 1), dtype=torch.float64, device=dev)
evo = cabi.condense(d, mats)
torch.cuda.synchronize()
# --- condense parity
worst = 0.0
for b in range(min(B, 8)):
    full, dd, vt = omld.complete({k: v[b] for k, v in wl["mats"].items()}, nu_l=1)
    ref = oc.condense(full, dd, Nt)
    for k, r in ref.items():
        g = evo[k][b].cpu().numpy()
        worst = max(worst, float(np.abs(g - r).max() / max(1.0, np.abs(r).max())))
print("condense max rel err", worst)
x0 = torch.tensor(wl["x0"], dtype=torch.float64, device=dev)
w = torch.tensor(wl["omega"], dtype=torch.float64, device=dev)
rhs = cabi.constraint_rhs(d, evo, x0, w)
nvt = d.nv * Nt
cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = wl["q_u"]; cost[:, :, 1] = wl["q_mu"][:, None, 0]; cost[:, :, 2] = wl["q_mu"][:, None, 1]
cost_t = torch.tensor(cost.reshape(B, nvt), dtype=torch.float64, device=dev)
lb = torch.zeros(nvt, dtype=torch.float64, device=dev)
ub = torch.tensor(np.tile([1.0, np.inf, np.inf], Nt), dtype=torch.float64, device=dev)
isb = torch.tensor(np.tile([1, 0, 0], Nt).astype(np.uint8), device=dev)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v, obj, status, stats = cabi.milp_solve(cost_t, evo["H_v"], rhs, lb, ub, isb)
    e1.record(); torch.cuda.synchronize()
    print("milp_solve B=%d N_p=%d: %.3f ms" % (B, N_p, e0.elapsed_time(e1)))
st = stats.cpu().numpy(); sta = status.cpu().numpy(); objn = obj.cpu().numpy(); vn = v.cpu().numpy()
print("status counts", np.bincount(sta, minlength=6), "pivots mean/max", st[:, 1].mean(), st[:, 1].max(), "nodes mean/max", st[:, 0].mean(), st[:, 0].max(), "cuts mean", st[:, 2].mean(), "max_rows max", st[:, 4].max())
bad = 0; diffu = 0
t0 = time.perf_counter()
for b in range(ncheck):
    full, dd, vt = omld.complete({k: vv[b] for k, vv in wl["mats"].items()}, nu_l=1)
    ref = oc.condense(full, dd, Nt)
    prob = oa.build_problem(ref, dd, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
    s, o, vr = osv.solve_milp(prob)
    if abs(objn[b] - o) > 1e-6 * max(1, abs(o)):
        bad += 1; print("OBJ MISMATCH", b, objn[b], o, sta[b], st[b])
    elif not np.array_equal(np.round(vn[b][prob.is_bin]), np.round(vr[prob.is_bin])):
        diffu += 1; print("DECISION DIFF", b, objn[b], o)
print("checked %d: obj mismatches %d decision diffs %d  (HiGHS %.1f ms/solve)" % (ncheck, bad, diffu, (time.perf_counter() - t0) * 1e3 / max(1, ncheck)))
print("fp64 peak TFLOP/s", cabi.fp64_peak_tflops())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(N_p=N_p, B=B, bad=bad, diffu=diffu, pivots=st[:, 1].tolist(), nodes=st[:, 0].tolist()), open("gpurun_out/gpu_check_%d_%d.json" % (N_p, B), "w"))
