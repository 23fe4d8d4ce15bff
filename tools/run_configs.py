"""DEV TOOL (gpurun): BASELINE.json configs 3-5 on ONE B200 -- the per-GPU share of the multi-GPU configurations,
through the same public path (DewhFleet / BatchMpc) the tests use.  Writes gpurun_out/configs.json."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhybridcontrol_b200 import cabi, distributed
from pyhybridcontrol_b200.batch import BatchMpc
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

dev = torch.device("cuda:0")
out = {}


def tile_params(B):
    base = [syn.dewh_agent_params(a) for a in range(256)]
    return [base[b % 256] for b in range(B)]


def ev_time(fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), r


# ---- config 3: residential micro-grid, 1000 DEWHs + PV + demand + grid agent (one GPU's shard of it)
B, N_p = 1000, 48
Nt = N_p + 1
fleet = DewhFleet(tile_params(B), N_p, device=dev)
rng = np.random.default_rng(0)
T0 = torch.as_tensor(rng.integers(55, 65, size=B).astype(float)).to(dev)
demand = torch.as_tensor(np.stack([syn.dhw_demand_profile(Nt, seed=b % 256) for b in range(B)])).to(dev)
price = syn.price_profile(Nt, seed=1)
k = np.arange(Nt)
p_pv = torch.as_tensor(-3000.0 * B * np.clip(np.sin((k / 96.0) * 2 * np.pi - 0.5 * np.pi), 0, None)).to(dev)   # -P_pv_max * units * omega_pv
p_res = torch.as_tensor(1200.0 * B * (1.0 + 0.3 * np.sin(k / 96.0 * 4 * np.pi))).to(dev)                        # P_res_ave * units * omega_res


def microgrid_step():
    fleet.build()
    res = fleet.control_step(T0.reshape(B, 1), demand, fleet.cost_from_prices(price))
    p_dev = fleet.aggregate_power(res["u"])            # + NCCL all-reduce when several ranks run
    grid = distributed.grid_evaluate(p_dev, p_pv, p_res)
    return res, grid


ms, (res, grid) = ev_time(microgrid_step)
cost_grid = float((grid["p_imp"] * torch.as_tensor(price).to(dev)).sum())
out["config3"] = dict(agents=B, N_p=N_p, ms_per_step=ms, solves_per_s=B / ms * 1e3, optimal=int((res["status"] == 0).sum()),
                      grid_import_cost=cost_grid)
print("config 3 (1 GPU shard): %d DEWHs + PV + demand + grid evaluation: %.3f ms/step, %.0f solves/s, optimal %d/%d" % (
    B, ms, B / ms * 1e3, out["config3"]["optimal"], B))
del fleet
torch.cuda.empty_cache()

# ---- config 4: closed-loop 24 h (96 steps of 15 min) for a 10,000-agent fleet on one GPU
B, steps = int(os.environ.get("C4_AGENTS", "10000")), 96
fleet = DewhFleet(tile_params(B), N_p, device=dev)
T0 = rng.integers(55, 65, size=B).astype(float)
prof = np.stack([syn.dhw_demand_profile(steps + Nt, seed=b) for b in range(256)])
demand = prof[np.arange(B) % 256]
price = syn.price_profile(steps + Nt, seed=2)
torch.cuda.synchronize()
t0 = time.perf_counter()
log = fleet.closed_loop(T0, demand, price, steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
st = log["status"].cpu().numpy()
Tl = log["T"].cpu().numpy()
out["config4"] = dict(agents=B, sim_steps=steps, wall_s=dt, ms_per_control_step=1e3 * dt / steps,
                      solves_per_s=B * steps / dt, not_optimal=int((st != 0).sum()),
                      T_min=float(Tl.min()), T_max=float(Tl.max()), heater_duty=float(log["u"].mean()))
print("config 4 (1 GPU): %d agents x %d closed-loop steps: %.2f s wall, %.2f ms per control step, %.0f solves/s, "
      "not optimal %d, T in [%.1f, %.1f], duty %.3f" % (B, steps, dt, 1e3 * dt / steps, B * steps / dt, (st != 0).sum(),
                                                      Tl.min(), Tl.max(), float(log["u"].mean())))
del fleet, log
torch.cuda.empty_cache()

# ---- config 5: scenario-based MPC sweep, 1000 DEWHs x 32 demand scenarios, N_p in {24, 48, 96}
#   (a) reference semantics, mpc_sb_full: ONE solve per agent with the robust row-min right-hand side over the
#       whole horizon (controller_base.py:442-444);  (b) mpc_sb_reduced: nominal constraints + the scenario set on the
#       first N_sb_reduced = 8 steps only (micro_grid_control_simulation.py:200-213);  (c) B x S independent solves,
#       one per (agent, scenario) pair, in chunks of 8000.
B, S = 1000, 32
for N_p in (24, 48, 96):
    Nt = N_p + 1
    wl = syn.dewh_batch(256, N_p, seed=5)
    rep = lambda a: np.concatenate([a] * 4, axis=0)[:B]
    mats = {k_: rep(v) for k_, v in wl["mats"].items()}
    bm = BatchMpc(mats, N_p, nu_l=1, device=dev)
    bm.build(want=("H_x", "H_v", "H_omega", "H_5"))
    scen = rep(wl["omega"])[:, :, None] * rng.uniform(0.5, 1.8, size=(B, Nt, S))
    cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = rep(wl["q_u"]); cost[:, :, 1:] = rep(wl["q_mu"])[:, None, :]
    x0 = torch.as_tensor(rep(wl["x0"])).to(dev); om = torch.as_tensor(rep(wl["omega"])).to(dev)
    sc = torch.as_tensor(scen).to(dev); cst = torch.as_tensor(cost.reshape(B, -1)).to(dev)
    variants = (("a_robust_full_horizon", lambda: bm.solve(x0, om, cost_v=cst, scenarios=sc)),
                ("b_sb_reduced_8_steps", lambda: bm.solve(x0, om, cost_v=cst, extra_constraints=[dict(omega_scenarios_k=sc, N_tilde=8)])))
    for name, fn in variants:
        ms, res = ev_time(fn, reps=2)
        key = "config5_N%d_%s" % (N_p, name)
        out[key] = dict(agents=B, scenarios=S, N_p=N_p, ms_per_step=ms, solves_per_s=B / ms * 1e3,
                        optimal=int((res["status"] == 0).sum()), nodes_mean=float(res["stats"][:, 0].double().mean()))
        print("config 5%s (1 GPU): %d agents x %d scenarios, N_p=%d: %.3f ms, %.0f solves/s, optimal %d/%d, nodes mean %.1f" % (
            name[0], B, S, N_p, ms, B / ms * 1e3, out[key]["optimal"], B, out[key]["nodes_mean"]))
    del bm
    torch.cuda.empty_cache()
    # (c) independent solves: every (agent, scenario) pair is its own MILP
    chunk = 8000
    Bc = chunk
    pairs = B * S
    idx_a = np.repeat(np.arange(B), S); idx_s = np.tile(np.arange(S), B)
    bmc = None
    tot_ms, n_opt, nodes = 0.0, 0, 0.0
    for c0 in range(0, pairs, chunk):
        ia, isc = idx_a[c0:c0 + chunk], idx_s[c0:c0 + chunk]
        if len(ia) < chunk:
            break
        bmc = BatchMpc({k_: v[ia] for k_, v in mats.items()}, N_p, nu_l=1, device=dev)
        bmc.build(want=("H_x", "H_v", "H_omega", "H_5"))
        xo = torch.as_tensor(rep(wl["x0"])[ia]).to(dev); oo = torch.as_tensor(scen[ia, :, isc]).to(dev)
        cc = torch.as_tensor(cost.reshape(B, -1)[ia]).to(dev)
        ms, res = ev_time(lambda: bmc.solve(xo, oo, cost_v=cc), reps=2)
        tot_ms += ms; n_opt += int((res["status"] == 0).sum()); nodes += float(res["stats"][:, 0].double().sum())
        del bmc
        torch.cuda.empty_cache()
    done = (pairs // chunk) * chunk
    key = "config5_N%d_c_independent" % N_p
    out[key] = dict(pairs=done, N_p=N_p, ms_total=tot_ms, solves_per_s=done / tot_ms * 1e3, optimal=n_opt, nodes_mean=nodes / done)
    print("config 5c (1 GPU): %d independent (agent, scenario) solves, N_p=%d: %.1f ms, %.0f solves/s, optimal %d/%d, nodes mean %.1f" % (
        done, N_p, tot_ms, done / tot_ms * 1e3, n_opt, done, nodes / done))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
