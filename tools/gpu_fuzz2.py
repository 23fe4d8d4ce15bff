"""DEV TOOL (gpurun): (1) the general branch-and-cut kernel on random MLDs with a vector state and continuous auxiliaries
against HiGHS; (2) stage-DP with random convex quadratic / L1 terms (MIQP) against enumeration with HiGHS QPs."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_stage_dp import random_scalar_mld
from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
from pyhybridcontrol_b200.batch import BatchMpc

dev = torch.device("cuda:0")
rng = np.random.default_rng(77)

# ---- (1) general MLDs: nx = 2, one binary input + one continuous input, one binary delta, one continuous z, slacks
B, Nt = 40, 6
nx, nu, nd, nz, nw, ny, nc = 2, 2, 1, 1, 1, 1, 4
m = dict(A=rng.uniform(-0.5, 0.9, (B, nx, nx)), B1=rng.uniform(-1, 1, (B, nx, nu)), B2=rng.uniform(-1, 1, (B, nx, nd)),
         B3=rng.uniform(-0.5, 0.5, (B, nx, nz)), B4=rng.uniform(-1, 1, (B, nx, nw)), b5=rng.uniform(-0.2, 0.2, (B, nx, 1)),
         C=rng.uniform(-1, 1, (B, ny, nx)), D1=rng.uniform(-0.3, 0.3, (B, ny, nu)),
         E=rng.uniform(-1, 1, (B, nc, nx)), F1=rng.uniform(-0.5, 0.5, (B, nc, nu)), F2=rng.uniform(-0.5, 0.5, (B, nc, nd)),
         F3=rng.uniform(-1, 1, (B, nc, nz)), F4=rng.uniform(-0.3, 0.3, (B, nc, nw)), f5=rng.uniform(1.0, 3.0, (B, nc, 1)),
         G=rng.uniform(-0.3, 0.3, (B, nc, ny)), Psi=-np.tile(np.eye(nc), (B, 1, 1)))
nv = nu + nd + nz + nc
x0 = rng.uniform(-1, 1, (B, nx)); om = rng.uniform(-1, 1, (B, Nt * nw))
cost = rng.uniform(-0.5, 1.0, (B, Nt, nv)); cost[:, :, nu + nd + nz:] = rng.uniform(5, 20, (B, 1, nc))
bm = BatchMpc(m, Nt - 1, Nt, nu_l=1, device=dev)          # the LAST input is binary (nu_l = 1), delta binary, z continuous
assert not bm.stage_dp_ok
# box the continuous columns so that the LP is bounded
lbs, ubs = bm.lb_v.copy(), bm.ub_v.copy()
cont = (bm.is_bin_v == 0) & ~np.isfinite(ubs) & (np.tile(np.arange(nv), Nt) < nu + nd + nz)
lbs[cont], ubs[cont] = -2.0, 2.0
bm.lb_v, bm.ub_v = lbs, ubs
bm.build()
t0 = time.time()
r = bm.solve(x0, om, cost_v=cost.reshape(B, -1))
torch.cuda.synchronize()
obj, st, v = r["obj"].cpu().numpy(), r["status"].cpu().numpy(), r["v"].cpu().numpy()
bad = 0
for b in range(0 if not os.environ.get('FUZZ2_DEBUG') else B, B):
    full, d, vt = omld.complete({k: a[b] for k, a in m.items()}, nu_l=1)
    prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, x0[b], om[b], atoms=None)
    prob.c = cost[b].ravel().copy(); prob.lb, prob.ub = lbs.copy(), ubs.copy()
    s_, o_, v_ = osv.solve_milp(prob, polish=True)
    if s_ == osv.INFEASIBLE:
        ok = st[b] == 1
    else:
        ok = st[b] == 0 and abs(obj[b] - o_) <= 1e-6 * max(1.0, abs(o_))
    if not ok:
        bad += 1; print("  general MLD agent", b, "gpu", obj[b], st[b], "highs", o_, s_)
print("(1) branch-and-cut on %d random vector-state MLDs (n = %d, %d binaries, m = %d): mismatches %d, status %s" % (
    B, nv * Nt, int(bm.is_bin_v.sum()), nc * Nt, bad, np.bincount(st, minlength=6).tolist()))

# ---- (2) MIQP: random convex terms on the stage-DP class (the enumeration oracle solves 2^(nb*Nt) QPs per agent: keep small)
B, Nt = int(os.environ.get('FUZZ2_MIQP_B', '6')), 6
bad2 = 0
for case, (nu, nd_, nc_, ny_) in enumerate([(1, 0, 2, 1), (2, 0, 3, 2), (1, 1, 2, 1)]):
    mm = random_scalar_mld(rng, B, nu, nd_, nc_, ny_, True, 0)
    nb = nu + nd_
    x0 = rng.uniform(-1.5, 1.5, (B, 1)); om = rng.uniform(-1, 1, (B, Nt))
    cost = np.zeros((B, Nt, nb + nc_)); cost[:, :, :nb] = rng.uniform(-0.3, 0.8, (B, Nt, nb)); cost[:, :, nb:] = rng.uniform(2, 15, (B, 1, nc_))
    wx2, wx1 = rng.uniform(0, 0.6, (1, Nt, 1)), rng.uniform(0, 0.5, (1, Nt, 1))
    wy2, wmu2 = rng.uniform(0, 0.6, (1, Nt, ny_)), rng.uniform(0, 3.0, (1, Nt, nc_))
    qx = rng.uniform(-1.0, 0.5, (Nt,))
    bm = BatchMpc(mm, Nt - 1, Nt, nu_l=nu, device=dev, solver="stage_dp")
    bm.build()
    r = bm.solve(x0, om, cost_v=cost.reshape(B, -1), w_x=np.tile(qx, (B, 1)), quad=dict(x2=wx2, x1=wx1, y2=wy2, mu2=wmu2))
    obj, st = r["obj"].cpu().numpy(), r["status"].cpu().numpy()
    vg = r["v"].cpu().numpy()
    dbg = os.environ.get("FUZZ2_DEBUG")
    for b in (range(B) if not dbg else []):
        full, d, vt = omld.complete({k: a[b] for k, a in mm.items()}, nu_l=nu)
        atoms = {"q_x": qx, "q_L22_x": np.sqrt(wx2.ravel()), "q_L1_x": wx1.ravel(), "q_L22_y": np.sqrt(wy2.reshape(-1)),
                 "q_L22_mu": np.sqrt(wmu2.reshape(-1))}
        prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, x0[b], om[b], atoms=atoms)
        prob.c[:(nb + nc_) * Nt] += cost[b].ravel()
        if dbg:
            vfull = np.zeros(prob.n); vfull[:vg.shape[1]] = vg[b]
            # epigraph columns of the L1 atom: t = |A v + a0| at the GPU point
            ne = prob.n - vg.shape[1]
            if ne:
                rows = prob.H[-2 * ne:-ne, :vg.shape[1]] @ vg[b] - prob.rhs[-2 * ne:-ne]
                vfull[vg.shape[1]:] = np.abs(rows)
            print("  DEBUG case", case, "agent", b, "gpu obj", obj[b], "oracle objective at the gpu point", prob.objective(vfull),
                  "max row violation", float(np.max(prob.H @ vfull - prob.rhs)))
            st_q, o_q, v_q = osv._continuous_subproblem(prob, np.round(vg[b][prob.is_bin[:vg.shape[1]]])), None, None
            print("      QP with the gpu binaries fixed:", st_q[0])
            continue
        s_, o_, v_, second = osv.solve_enumerate(prob)
        ok = st[b] == 0 and abs(obj[b] - o_) <= 1e-6 * max(1.0, abs(o_))
        if not ok:
            bad2 += 1; print("  MIQP case", case, "agent", b, "gpu", obj[b], st[b], "enum", o_)
print("(2) stage-DP MIQP on %d random problems: mismatches %d" % (3 * B, bad2))
