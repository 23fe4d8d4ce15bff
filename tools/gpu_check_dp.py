"""DEV TOOL: stage-DP solver (csrc/stage_dp.cu) against the branch-and-cut kernel and the HiGHS oracle (gpurun)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from oracle import mld as omld, condense as oc, assemble as oa, solve as osv

dev = torch.device("cuda:0")
N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else B
cells = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["2048", "4096", "8192", "16384"])]
k0 = int(sys.argv[5]) if len(sys.argv) > 5 else 0
wl = syn.dewh_batch(B, N_p, seed=1, k0=k0)
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
torch.cuda.synchronize()
res = {}
for G in cells:
    o = cabi.stage_dp_default_opts(cells=G)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        v, obj, status, stats = cabi.stage_dp_solve(d, mats, rhs, cost_t, lb, ub, isb, o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    st = stats.cpu().numpy()
    print("stage_dp G=%d B=%d N_p=%d: %.3f ms  status %s nodes mean %.1f p50 %.0f max %d max_open %d" % (
        G, B, N_p, ms, np.bincount(status.cpu().numpy(), minlength=6).tolist(), st[:, 0].mean(), np.median(st[:, 0]), st[:, 0].max(), st[:, 4].max()))
    res[G] = (v.cpu().numpy(), obj.cpu().numpy(), status.cpu().numpy(), st, ms)
# all cell counts must agree bit-for-bit on the decisions
ref_v, ref_obj = res[cells[-1]][0], res[cells[-1]][1]
for G in cells[:-1]:
    print("G=%d vs G=%d: max |obj diff| %.3e, decisions equal %s" % (G, cells[-1], np.abs(res[G][1] - ref_obj).max(),
          np.array_equal(np.round(res[G][0].reshape(B, Nt, 3)[:, :, 0]), np.round(ref_v.reshape(B, Nt, 3)[:, :, 0]))))
# feasibility + objective consistency against the condensed matrices
H = evo["H_v"].cpu().numpy(); rh = rhs.cpu().numpy()
viol = np.max(np.einsum("bmn,bn->bm", H, ref_v) - rh)
print("max H v - rhs %.3e ; max |obj - c'v| %.3e" % (viol, np.abs((cost.reshape(B, nvt) * ref_v).sum(1) - ref_obj).max()))
if os.environ.get("WITH_BNC", "1") == "1" and N_p <= 48:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v2, obj2, status2, stats2 = cabi.milp_solve(cost_t, evo["H_v"], rhs, lb, ub, isb)
    e1.record(); torch.cuda.synchronize()
    ok = status2.cpu().numpy() == 0
    print("bnc: %.1f ms, optimal %d/%d; max rel obj diff vs stage_dp %.3e; decisions equal on %d/%d" % (
        e0.elapsed_time(e1), ok.sum(), B, (np.abs(obj2.cpu().numpy() - ref_obj) / np.maximum(1, np.abs(ref_obj)))[ok].max(),
        sum(np.array_equal(np.round(v2[b].cpu().numpy()[::3]), np.round(ref_v[b][::3])) for b in range(B) if ok[b]), ok.sum()))
bad = diffu = 0
t0 = time.perf_counter()
for b in range(ncheck):
    full, dd, vt = omld.complete({k: vv[b] for k, vv in wl["mats"].items()}, nu_l=1)
    ref = oc.condense(full, dd, Nt)
    prob = oa.build_problem(ref, dd, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
    s, o_, vr = osv.solve_milp(prob)
    if abs(ref_obj[b] - o_) > 1e-6 * max(1, abs(o_)):
        bad += 1; print("OBJ MISMATCH", b, ref_obj[b], o_)
    elif not np.array_equal(np.round(ref_v[b][prob.is_bin]), np.round(vr[prob.is_bin])):
        diffu += 1; print("DECISION DIFF", b, ref_obj[b], o_)
print("HiGHS check %d agents: obj mismatches %d decision diffs %d (%.1f ms/solve)" % (ncheck, bad, diffu, (time.perf_counter() - t0) * 1e3 / max(1, ncheck)))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({str(G): dict(ms=res[G][4], nodes=res[G][3][:, 0].tolist()) for G in cells}, open("gpurun_out/gpu_check_dp_%d_%d.json" % (N_p, B), "w"))
