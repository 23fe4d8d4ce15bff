"""DEV TOOL: bounds / iteration counts of the price coordination on the GPU for synthetic cases.
    python tools/gpu_check_coupled.py N_h N_p [iters]        surplus-hump case of tests/test_gpu_coupled.py
    python tools/gpu_check_coupled.py day [iters]            first instant of the closed-loop test case (24 agents)"""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_coupled import _case
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

if sys.argv[1] == "day":
    N_h, N_p, steps = 24, 16, 8
    Nt = N_p + 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    params = [syn.dewh_agent_params(700 + b) for b in range(N_h)]
    T0 = np.array([syn.dewh_initial_state(700 + b) for b in range(N_h)])
    dem = np.stack([syn.dhw_demand_profile(steps + Nt, seed=700 + b) for b in range(N_h)])[:, :Nt]
    price = syn.price_profile(steps + Nt, seed=7)[:Nt]
    P = np.array([p["P_h_Nom"] for p in params])
    k = np.arange(steps + Nt)
    p_other = (-0.7 * P.sum() * np.clip(np.sin(k / 12 * np.pi), 0, None) + 0.05 * P.sum())[:Nt]
else:
    N_h, N_p = int(sys.argv[1]), int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    params, T0, dem, price, P, p_other = _case(N_h, N_p, seed=0)
fleet = DewhFleet(params, N_p, device="cuda")
fleet.build()
for kw in (dict(response_passes=0), dict()):
    torch.cuda.synchronize(); t = time.time()
    out = fleet.coupled_step(T0, dem, price, p_other, iters=iters, rel_gap=1e-2, **kw)
    torch.cuda.synchronize()
    print(kw, {k: v for k, v in out.items() if k not in ("u", "plan", "lam")}, "%.1f ms" % ((time.time() - t) * 1e3))
    print("   lam/price", np.round((out["lam"].cpu().numpy() / price), 3).tolist())
if os.environ.get("TRACE"):
    from pyhybridcontrol_b200 import cabi
    dev = fleet.device
    Nt = N_p + 1
    pr = torch.as_tensor(price, device=dev).contiguous(); po = torch.as_tensor(p_other, device=dev).contiguous()
    cost = fleet.cost_from_prices(pr)
    lam = [pr.clone(), torch.empty_like(pr)]
    state, sums = cabi.coupling_state(dev), torch.empty(Nt + 2, dtype=torch.float64, device=dev)
    a_lo, a_hi = torch.zeros(Nt, dtype=torch.float64, device=dev), fleet.P_nom.sum().expand(Nt).contiguous()
    x0 = torch.as_tensor(T0, device=dev).reshape(-1, 1)
    for it in range(int(os.environ["TRACE"])):
        cur, nxt = lam[it & 1], lam[1 - (it & 1)]
        cabi.coupling_price_cost(cur, fleet.P_nom, cost, 3, 0)
        res = fleet.control_step(x0, dem, cost)
        cabi.coupling_sums(res["u"], fleet.P_nom, res["obj"].contiguous(), res["status"], sums)
        cabi.coupling_dual_step(sums, po, pr, a_lo, a_hi, 1.0, cur, nxt, state)
        st = state.cpu().numpy()
        print(it, "dual %.5f primal %.4f LB %.5f g2 %.3e" % (st[2], st[3], st[0], st[6]),
              "lam/price", np.round((cur / pr).cpu().numpy(), 3).tolist(),
              "agg/Ptot", np.round((sums[:Nt] / fleet.P_nom.sum()).cpu().numpy(), 2).tolist())
