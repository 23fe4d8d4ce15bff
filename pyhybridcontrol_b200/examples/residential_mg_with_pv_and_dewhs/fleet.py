"""A fleet of DEWH agents on one GPU rank: the batched counterpart of the reference's per-device loop
(examples/residential_mg_with_pv_and_dewhs/modelling/micro_grid_agents.py:699-700, 714-735, 739-740).

Per control step and per agent the reference does  build() -> solve() -> sim_step_k();  here the whole shard does
  K1 condense -> K2 rhs -> K3/K4 mixed-integer solve -> K5 sim step (re-parametrised DEWH model) -> K6 aggregate
with every buffer resident in HBM and one NCCL all-reduce of the [Nt] aggregate power when several ranks run.
"""
import numpy as np
import torch

from ... import cabi, distributed
from ...batch import BatchMpc


class DewhFleet(object):
    def __init__(self, params, N_p, device="cuda", opts=None):
        """params: list of per-agent DEWH parameter dicts (parameters.dewh_param_struct keys)."""
        self.device = torch.device(device)
        self.B = len(params)
        self.N_p = int(N_p)
        self.Nt = self.N_p + 1
        self.params = torch.as_tensor(cabi.pack_dewh_params(params), dtype=torch.float64).to(self.device)
        self.P_nom = self.params[:, 6].contiguous()
        self.band = torch.as_tensor([[p.get("T_h_max_sub_T_h_on", 12.0), p.get("T_h_max_sub_T_h_off", 4.0)]
                                     for p in params], dtype=torch.float64).to(self.device)
        B = self.B
        one = torch.ones((1, 1, 1), dtype=torch.float64, device=self.device)
        self._mats = dict(
            A=torch.empty((B, 1, 1), dtype=torch.float64, device=self.device),
            B1=torch.empty((B, 1, 1), dtype=torch.float64, device=self.device),
            B4=torch.empty((B, 1, 1), dtype=torch.float64, device=self.device),
            b5=torch.empty((B, 1, 1), dtype=torch.float64, device=self.device),
            C=one,
            E=torch.tensor([[[1.0], [-1.0]]], dtype=torch.float64, device=self.device),
            Psi=torch.tensor([[[-1.0, 0.0], [0.0, -1.0]]], dtype=torch.float64, device=self.device),
            f5=torch.stack([self.params[:, 8], -self.params[:, 7]], dim=1).reshape(B, 2, 1).contiguous(),
        )
        self.refresh_control_model()
        self.batch = BatchMpc(self._mats, self.N_p, nu_l=1, B=B, device=self.device, opts=opts)
        self._want = ("H_x", "H_v", "H_omega", "H_5")
        self.soft_top_mult, self.soft_bot_mult = 10.0, 1.0

    def refresh_control_model(self):
        """const_heat=True control model -> A, B1, B4, b5 blocks (micro_grid_models.py:37-57)."""
        model = cabi.dewh_control_model(self.params)
        for i, name in enumerate(("A", "B1", "B4", "b5")):
            self._mats[name].copy_(model[:, i].reshape(self.B, 1, 1))

    def build(self, full=False):
        self.batch.mats.update({k: v for k, v in self._mats.items()})
        return self.batch.build(want=cabi.EVO_NAMES if full else self._want, reuse=not full)

    def cost_from_prices(self, price):
        """price [Nt] or [B, Nt] (currency per W per step) -> cost on v~ [B, 3*Nt]:
        q_u = price * P_h_Nom, q_mu = [10, 1] * sum(q_u)  (micro_grid_control_simulation.py:193-198)."""
        price = torch.as_tensor(price, dtype=torch.float64).to(self.device)
        if price.dim() == 1:
            price = price.unsqueeze(0).expand(self.B, -1)
        q_u = price * self.P_nom[:, None]
        tot = q_u.sum(dim=1, keepdim=True)
        cost = torch.stack([q_u, (self.soft_top_mult * tot).expand(-1, self.Nt),
                            (self.soft_bot_mult * tot).expand(-1, self.Nt)], dim=2)
        return cost.reshape(self.B, 3 * self.Nt).contiguous()

    def control_step(self, x0, omega_forecast, cost_v, extra_constraints=(), rhs=None):
        """-> dict(v, obj, status, stats, u [B, Nt] view)."""
        res = self.batch.solve(x0, omega_forecast, cost_v=cost_v, extra_constraints=extra_constraints, rhs=rhs)
        res["u"] = res["v"].view(self.B, self.Nt, 3)[:, :, 0]
        return res

    def sim_step(self, T, u0, D_h):
        """Advance every tank one step with the re-parametrised simulation model (micro_grid_agents.py:389-408)."""
        T1, _, cons = cabi.dewh_sim_step(self.params, T, u0, D_h)
        return T1, cons

    def aggregate_power(self, u):
        """sum_b P_nom[b] u[b, k] over this rank's agents, then over ranks -> [Nt]."""
        return distributed.allreduce_aggregate(cabi.aggregate_power(u, self.P_nom))

    def coupled_step(self, x0, omega_forecast, price, p_other, iters=300, theta=1.0, rel_gap=1e-2, check_every=25,
                     grid_limits=None, extra_constraints=(), response_passes=12, response_groups=8, warm_iters=50):
        """One control instant of the CENTRALISED micro-grid problem (micro_grid_agents.py:691-735: the price sits on
        the grid import z_k = max(0, sum_i P_i u_i,k + p_other_k), q_z, and the devices only bring their slack
        penalties) by price coordination: the batched exact agent solve at the internal price lambda, the aggregate,
        one dual step -- repeated on the device until the certified gap (upper bound - lower bound) / |upper bound|
        is below ``rel_gap`` (the reference accepts MIPGap = 1e-2, micro_grid_control_simulation.py:232) or ``iters``
        is reached.  The state is read back every ``check_every`` iterations only.

        price [Nt]; p_other [Nt] = PV + residential demand outputs (W, PV negative); grid_limits = (P_g_min, P_g_max)
        bounds on y_k (micro_grid_models.py:145-150), None = not binding.  Across ranks the sums are all-reduced, so
        every rank walks the same lambda.
        The coordination's own plans are lumpy (agents with the same preferences flip together), so its best plan is
        then improved by a best-response descent on the true cost (``response_passes`` passes over blocks of agents,
        ``response_groups`` blocks at first; 0 passes = off): the upper bound only ever goes down.
        -> dict(u [B, Nt] plan, plan (v, obj = every agent's penalty, status, stats), lam [Nt] best internal price,
                lower_bound, upper_bound, dual_upper_bound (before the descent), gap, iterations, skipped,
                response_solves, response_accepted)"""
        dev, B, Nt = self.device, self.B, self.Nt
        price = torch.as_tensor(price, dtype=torch.float64).to(dev).reshape(Nt).contiguous()
        p_other = torch.as_tensor(p_other, dtype=torch.float64).to(dev).reshape(Nt).contiguous()
        P_tot = distributed.allreduce_aggregate(self.P_nom.sum().reshape(1).clone())
        a_lo = torch.zeros(Nt, dtype=torch.float64, device=dev)
        a_hi = P_tot.expand(Nt).contiguous()
        if grid_limits is not None:
            a_lo = torch.maximum(a_lo, float(grid_limits[0]) - p_other)
            a_hi = torch.minimum(a_hi, float(grid_limits[1]) - p_other)
        cost = self.cost_from_prices(price)                      # slack penalties as in the decentralised problem
        lam = [price.clone(), torch.empty_like(price)]
        state, sums = cabi.coupling_state(dev), torch.empty(Nt + 2, dtype=torch.float64, device=dev)
        u_best = torch.zeros((B, Nt), dtype=torch.float64, device=dev)
        lam_best = price.clone()
        x0 = torch.as_tensor(x0, dtype=torch.float64).to(dev).reshape(B, 1)
        st, rhs, it = None, None, 0
        out = dict(response_solves=0, response_accepted=0)
        best_total, best_plan = float("inf"), None

        def dual_iterations(n):
            nonlocal st, rhs, it
            for _ in range(n):
                cur, nxt = lam[it & 1], lam[1 - (it & 1)]
                cabi.coupling_price_cost(cur, self.P_nom, cost, 3, 0)
                res = self.control_step(x0, omega_forecast, cost, extra_constraints=extra_constraints, rhs=rhs)
                rhs = res.get("rhs")
                cabi.coupling_sums(res["u"], self.P_nom, res["obj"].contiguous(), res["status"], sums)
                distributed.allreduce_aggregate(sums)
                cabi.coupling_dual_step(sums, p_other, price, a_lo, a_hi, theta, cur, nxt, state)
                cabi.coupling_keep_best(res["u"], cur, state, u_best, lam_best)
                it += 1
                if it % int(check_every) == 0:
                    st = state.cpu().numpy()
                    if np.isfinite(st[1]) and st[1] - st[0] <= rel_gap * abs(st[1]):
                        break
            st = state.cpu().numpy()

        def plan_at_best_price():
            cabi.coupling_price_cost(lam_best, self.P_nom, cost, 3, 0)
            return self.control_step(x0, omega_forecast, cost, extra_constraints=extra_constraints, rhs=rhs)

        # phase 1: a first stretch of dual iterations; phase 2: descent from its plan; phase 3: the dual iterations go
        # on with the better upper bound in Polyak's step length until the gap closes
        dual_iterations(min(int(iters), int(warm_iters)) if response_passes else int(iters))
        dual_ub = float(st[1])
        plan = plan_at_best_price()
        closed = np.isfinite(st[1]) and st[1] - st[0] <= rel_gap * abs(st[1])
        ran_more = False
        if response_passes and plan.get("rhs") is not None and not closed:
            br = self._best_response(plan, cost, price, p_other, a_lo, a_hi, response_passes, response_groups)
            out.update(response_solves=br["solves"], response_accepted=br["accepted"])
            best_total = br["total"]
            best_plan = dict(v=br["v"], u=br["v"].view(B, Nt, 3)[:, :, 0], obj=br["pen"], status=plan["status"],
                             stats=plan["stats"])
            state[1].clamp_(max=best_total)
            st = state.cpu().numpy()
            if not (st[1] - st[0] <= rel_gap * abs(st[1])):
                dual_iterations(int(iters) - it)
                ran_more = True
        elif not closed and it < int(iters):
            dual_iterations(int(iters) - it)
            ran_more = True
        if best_plan is None or st[1] < best_total:          # the dual iterates found something better still
            best_plan = plan_at_best_price() if ran_more else plan
            best_plan["u"] = u_best                          # identical to the plan's own u (deterministic solve)
            best_total = float(st[1])
        ub = best_total
        out.update(lam=lam_best, lower_bound=float(st[0]), upper_bound=ub, dual_upper_bound=dual_ub,
                   gap=float((ub - st[0]) / abs(ub)) if np.isfinite(ub) and ub != 0 else float("inf"),
                   iterations=int(st[5]), skipped=int(st[7]), u=best_plan["u"], plan=best_plan)
        return out

    def coupled_step_exact(self, x0, omega_forecast, price, p_other, grid_limits=None, mip_rel_gap=0.0,
                           max_nodes=2000000):
        """The CENTRALISED micro-grid problem of one instant as ONE mixed-integer program, solved to optimality by the
        general branch-and-cut kernel -- for fleets small enough for it (the reference itself runs 20 heaters;
        micro_grid_agents.py:691-735 builds the same monolithic problem for cvxpy).  Decision vector: every heater's
        v~ = [u; mu] over the horizon, then the grid import z_k.  With a non-negative price the grid MLD's
        delta / z pair (micro_grid_models.py:145-168: z = [y >= 0] y) is equivalent to  z_k >= y_k, z_k >= 0  at
        the optimum, so the import is carried by one continuous column per step and the big-M rows are not needed;
        the grid limits, when given, bound the aggregate directly.  Single rank only (the problem does not shard).
        -> dict(u [B, Nt], v [B, 3 Nt], z [Nt], obj, status, stats)"""
        if distributed.is_distributed():
            raise NotImplementedError("the monolithic problem lives on one GPU; use coupled_step() across ranks")
        dev, B, Nt = self.device, self.B, self.Nt
        price = torch.as_tensor(price, dtype=torch.float64).to(dev).reshape(Nt)
        if bool((price < 0).any()):
            raise NotImplementedError("a negative import price needs the grid MLD's delta / z rows")
        p_other = torch.as_tensor(p_other, dtype=torch.float64).to(dev).reshape(Nt)
        x0 = torch.as_tensor(x0, dtype=torch.float64).to(dev).reshape(B, 1)
        omega = torch.as_tensor(omega_forecast, dtype=torch.float64).to(dev).reshape(B, Nt)
        nva = 3 * Nt
        n = B * nva + Nt
        H_a, rhs_a = self.batch.constraint_rows(x0, omega)                  # [B, 2 Nt, 3 Nt], [B, 2 Nt]
        ma = H_a.shape[1]
        rows = B * ma + Nt + (2 * Nt if grid_limits is not None else 0)
        H = torch.zeros((1, rows, n), dtype=torch.float64, device=dev)
        rhs = torch.zeros((1, rows), dtype=torch.float64, device=dev)
        for b in range(B):
            H[0, b * ma:(b + 1) * ma, b * nva:(b + 1) * nva] = H_a[b]
            rhs[0, b * ma:(b + 1) * ma] = rhs_a[b]
        r0 = B * ma
        k = torch.arange(Nt, device=dev)
        for b in range(B):                                                   #  sum_i P_i u_i,k - z_k <= -p_other_k
            H[0, r0 + k, b * nva + 3 * k] = self.P_nom[b]
        H[0, r0 + k, B * nva + k] = -1.0
        rhs[0, r0:r0 + Nt] = -p_other
        if grid_limits is not None:
            r1 = r0 + Nt
            for b in range(B):
                H[0, r1 + k, b * nva + 3 * k] = self.P_nom[b]
                H[0, r1 + Nt + k, b * nva + 3 * k] = -self.P_nom[b]
            rhs[0, r1:r1 + Nt] = float(grid_limits[1]) - p_other
            rhs[0, r1 + Nt:r1 + 2 * Nt] = p_other - float(grid_limits[0])
        cost = self.cost_from_prices(price).reshape(B, Nt, 3).clone()
        cost[:, :, 0] = 0.0                                                  # the energy price sits on the import z
        c = torch.cat([cost.reshape(-1), price]).reshape(1, n)
        lb_a, ub_a, isb_a = self.batch._bounds_dev()
        lb = torch.cat([lb_a.repeat(B), torch.zeros(Nt, dtype=torch.float64, device=dev)])
        ub = torch.cat([ub_a.repeat(B), torch.full((Nt,), float("inf"), dtype=torch.float64, device=dev)])
        isb = torch.cat([isb_a.repeat(B), torch.zeros(Nt, dtype=torch.uint8, device=dev)])
        opts = cabi.default_opts(mip_rel_gap=float(mip_rel_gap), max_nodes=int(max_nodes))
        v, obj, status, stats = cabi.milp_solve(c, H, rhs, lb, ub, isb, opts)
        va = v[0, :B * nva].reshape(B, nva)
        return dict(u=va.view(B, Nt, 3)[:, :, 0], v=va, z=v[0, B * nva:], obj=float(obj[0]), status=int(status[0]),
                    stats=stats[0])

    def _best_response(self, plan, cost, price, p_other, a_lo, a_hi, passes, groups):
        """Descent on the true centralised cost from the coordination's plan (hmpc.h: hmpc_coupling_response_cost_f64
        ..): blocks of agents answer their marginal price in turn, a block's answer is kept only if the total went
        down.  The block count doubles (up to one agent per block) whenever a whole pass brings nothing."""
        dev, B, Nt = self.device, self.B, self.Nt
        batch, d = self.batch, self.batch.dims
        lb, ub, isb = batch._bounds_dev()
        rhs = plan["rhs"]
        v_cur = torch.zeros((B, 3 * Nt), dtype=torch.float64, device=dev)
        pen_cur = torch.zeros(B, dtype=torch.float64, device=dev)
        v_bak, pen_bak = torch.empty_like(v_cur), torch.empty_like(pen_cur)
        sums_cur = torch.zeros(Nt + 2, dtype=torch.float64, device=dev)
        sums_cand = torch.empty_like(sums_cur)
        brs = torch.tensor([float("inf"), 0, 0, 0], dtype=torch.float64, device=dev)

        def evaluate(lo, hi):
            cabi.coupling_sums(v_cur.view(B, Nt, 3)[:, :, 0], self.P_nom, pen_cur, None, sums_cand)
            distributed.allreduce_aggregate(sums_cand)
            cabi.coupling_accept(sums_cand, p_other, price, a_lo, a_hi, sums_cur, brs)
            if hi > lo:
                cabi.coupling_restore(lo, hi, brs, v_bak, pen_bak, v_cur, pen_cur)

        # load the starting plan: the whole fleet is one block, accepted against +inf
        cabi.coupling_merge(0, B, Nt, 3, plan["v"], plan["obj"].contiguous(), plan["status"], cost, v_cur, pen_cur, v_bak,
                            pen_bak)
        evaluate(0, B)
        # every block ends in an all-reduce, so the number of blocks per pass must be the same on every rank: it is
        # capped by the LARGEST shard (shards differ by one agent), and a rank whose block is empty still evaluates
        B_max = distributed.allreduce_max_int(B, dev)
        G = max(1, min(int(groups), B_max))
        solves, last = 0, float(brs[0].cpu())
        for _ in range(int(passes)):
            for lo, hi in distributed.response_blocks(B, G):
                if hi > lo:
                    cabi.coupling_response_cost(sums_cur, v_cur, self.P_nom, p_other, price, cost, 3, 0)
                    sub = cabi.make_dims(hi - lo, d.Nt, nx=d.nx, nu=d.nu, ndelta=d.ndelta, nz=d.nz, nmu=d.nmu,
                                         nomega=d.nomega, ny=d.ny, nc=d.nc)
                    mats = {k: (m[lo:hi] if m.shape[0] == B and B > 1 else m) for k, m in batch.mats.items()}
                    v_new, obj_new, status_new, _ = cabi.stage_dp_solve(sub, mats, rhs[lo:hi], cost[lo:hi], lb, ub, isb,
                                                                        batch.dp_opts)
                    solves += 1
                    cabi.coupling_merge(lo, hi, Nt, 3, v_new, obj_new, status_new, cost, v_cur, pen_cur, v_bak, pen_bak)
                evaluate(lo, hi)
            now = float(brs[0].cpu())                                    # one read-back per pass (the same on every rank)
            if not now < last:
                if G >= B_max:
                    break
                G = min(B_max, 2 * G)
            last = now
        st = brs.cpu().numpy()
        return dict(v=v_cur, pen=pen_cur, total=float(st[0]), solves=solves, accepted=int(st[2]) - 1)

    CONTROLLERS = ("mpc_pb", "mpc_ce", "mpc_sb_reduced", "mpc_sb_full", "mpc_minmax", "thermo")

    def thermostat(self, T, u_prev):
        """Rule-based input of the non-predictive controller (theromstat_control.py:38-62) for every tank."""
        return cabi.dewh_thermostat(self.params, self.band, T, u_prev)

    def _extra_sets(self, controller, k, scenarios, demand_minmax, N_sb_reduced):
        """Constraint sets a controller variant adds to the standard one at instant k
        (micro_grid_control_simulation.py:200-227); every set shares the rows of H_v."""
        Nt = self.Nt
        if controller in ("mpc_sb_reduced", "mpc_sb_full"):
            if scenarios is None:
                raise ValueError("controller '%s' needs `scenarios`" % controller)
            sc = scenarios(k) if callable(scenarios) else scenarios[:, k:k + Nt, :]
            sc = torch.as_tensor(sc, dtype=torch.float64).to(self.device).reshape(self.B, Nt, -1).contiguous()
            if controller == "mpc_sb_reduced":
                return [dict(omega_scenarios_k=sc, N_tilde=min(int(N_sb_reduced), Nt))]
            return [dict(omega_scenarios_k=sc)]
        if controller == "mpc_minmax":
            if demand_minmax is None:
                raise ValueError("controller 'mpc_minmax' needs `demand_minmax` = (min profile, max profile)")
            out = []
            for prof in demand_minmax:
                prof = torch.as_tensor(prof, dtype=torch.float64).to(self.device)
                win = prof[k:k + Nt] if prof.dim() == 1 else prof[:, k:k + Nt]
                out.append(dict(omega_tilde_k=win.expand(self.B, Nt).contiguous()))
            return out
        return []

    def closed_loop(self, T0, demand, price, sim_steps, demand_actual=None, controller="mpc_ce", scenarios=None,
                    N_sb_reduced=8, demand_minmax=None, u_init=None, coupling=None):
        """Closed-loop simulation of the shard (reference loop: examples/.../micro_grid_control_simulation.py:184-236
        with the per-device work of micro_grid_agents.py:699-700, 733-740): at every step k the forecast window
        demand[:, k:k+Nt] and price[k:k+Nt] give the MPC problem, the first control is applied to the re-parametrised
        simulation model with the actual draw, and the aggregate power is exchanged.  Everything stays in HBM; the
        log comes back as device tensors.

        controller (the campaign's six variants, micro_grid_control_simulation.py:144-152):
          mpc_pb          forecast = the actual draw (deterministic; needs demand_actual of sim_steps + Nt columns)
          mpc_ce          forecast = `demand` (certainty equivalent)
          mpc_sb_reduced  mpc_ce + scenario set (row-wise min over S scenarios) on the first N_sb_reduced steps
          mpc_sb_full     mpc_ce + scenario set on the whole horizon
          mpc_minmax      mpc_ce + one set for the min-draw and one for the max-draw profile
          thermo          thermostat rule, no optimisation (obj is NaN, P_agg has one column)

        T0 [B] initial temperatures; demand [B, sim_steps + Nt] (L/s); price [sim_steps + Nt] or [B, sim_steps + Nt];
        demand_actual [B, >= sim_steps] defaults to demand; scenarios [B, sim_steps + Nt, S] or callable k -> [B, Nt, S];
        demand_minmax = (min, max) profiles [sim_steps + Nt] or [B, sim_steps + Nt]; u_init [B] input before step 0
        (thermostat only).  coupling = dict(p_other [sim_steps + Nt] PV + residential-demand forecast in W, plus any
        ``coupled_step`` option): the reference's centralised operation -- the price applies to the grid import, every
        instant is solved by ``coupled_step`` (obj is then every agent's penalty part, and the log gains
        coupled_upper_bound / coupled_gap [sim_steps]).
        -> dict(T [sim_steps + 1, B], u [sim_steps, B], obj [sim_steps, B], status [sim_steps, B],
                P_agg [sim_steps, Nt], cons [sim_steps, B, 2], mu_hat [sim_steps, B, 2] (first-step slacks of the
                plan), omega / omega_hat [sim_steps, B] (actual draw / first forecast value), solve_ms [sim_steps]
                (device time of the batch solve))"""
        if controller not in self.CONTROLLERS:
            raise ValueError("controller must be one of %s" % (self.CONTROLLERS,))
        dev, B, Nt = self.device, self.B, self.Nt
        T = torch.as_tensor(T0, dtype=torch.float64).to(dev).reshape(B).clone()
        demand = torch.as_tensor(demand, dtype=torch.float64).to(dev)
        price = torch.as_tensor(price, dtype=torch.float64).to(dev)
        actual = demand if demand_actual is None else torch.as_tensor(demand_actual, dtype=torch.float64).to(dev)
        forecast = demand
        if controller == "mpc_pb":
            if actual.shape[1] < sim_steps + Nt:
                raise ValueError("mpc_pb needs the actual draw over sim_steps + N_tilde steps")
            forecast = actual
        log = dict(T=[T.clone()], u=[], obj=[], status=[], P_agg=[], cons=[], mu_hat=[], omega=[], omega_hat=[])
        events = []
        if controller == "thermo":
            u_prev = (torch.zeros(B, dtype=torch.float64, device=dev) if u_init is None
                      else torch.as_tensor(u_init, dtype=torch.float64).to(dev).reshape(B).clone())
            nan = torch.full((B,), float("nan"), dtype=torch.float64, device=dev)
            ok = torch.zeros(B, dtype=torch.int32, device=dev)
            for k in range(sim_steps):
                u0 = self.thermostat(T, u_prev)
                T, cons = self.sim_step(T, u0, actual[:, k].contiguous())
                log["T"].append(T.clone()); log["u"].append(u0); log["obj"].append(nan); log["status"].append(ok)
                log["P_agg"].append(self.aggregate_power(u0.reshape(B, 1))); log["cons"].append(cons)
                log["mu_hat"].append(nan.reshape(B, 1).expand(B, 2)); log["omega"].append(actual[:, k])
                log["omega_hat"].append(forecast[:, k])
                u_prev = u0
            out = {k: torch.stack(v) for k, v in log.items()}
            out["solve_ms"] = torch.zeros(sim_steps, dtype=torch.float64, device=dev)
            return out
        self.build()                                   # the control model does not change along the run
        if coupling is not None:
            if price.dim() != 1:
                raise ValueError("the centralised problem has one import price per step")
            copts = dict(coupling)
            p_other = torch.as_tensor(copts.pop("p_other"), dtype=torch.float64).to(dev)
            coupled = dict(ub=[], gap=[])
        for k in range(sim_steps):
            pk = price[k:k + Nt] if price.dim() == 1 else price[:, k:k + Nt]
            extra = self._extra_sets(controller, k, scenarios, demand_minmax, N_sb_reduced)
            events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            events[-1][0].record()
            if coupling is not None:
                cs = self.coupled_step(T.reshape(B, 1), forecast[:, k:k + Nt].contiguous(), pk, p_other[k:k + Nt],
                                       extra_constraints=extra, **copts)
                res = cs["plan"]
                coupled["ub"].append(cs["upper_bound"]); coupled["gap"].append(cs["gap"])
            else:
                res = self.control_step(T.reshape(B, 1), forecast[:, k:k + Nt].contiguous(), self.cost_from_prices(pk),
                                        extra_constraints=extra)
            events[-1][1].record()
            u0 = res["u"][:, 0].contiguous()
            T, cons = self.sim_step(T, u0, actual[:, k].contiguous())
            log["T"].append(T.clone()); log["u"].append(u0); log["obj"].append(res["obj"])
            log["status"].append(res["status"]); log["P_agg"].append(self.aggregate_power(res["u"])); log["cons"].append(cons)
            log["mu_hat"].append(res["v"].view(B, Nt, 3)[:, 0, 1:]); log["omega"].append(actual[:, k])
            log["omega_hat"].append(forecast[:, k])
        out = {k: torch.stack(v) for k, v in log.items()}
        torch.cuda.synchronize(dev)
        out["solve_ms"] = torch.tensor([a.elapsed_time(b) for a, b in events], dtype=torch.float64, device=dev)
        if coupling is not None:
            out["coupled_upper_bound"] = torch.tensor(coupled["ub"], dtype=torch.float64, device=dev)
            out["coupled_gap"] = torch.tensor(coupled["gap"], dtype=torch.float64, device=dev)
        return out

    def grid_log(self, log, price, pv=None, resd=None, grid_params=None, device_ids=None):
        """Grid-agent view of a closed-loop log (micro_grid_agents.py:625-646, 736-756): the stacked device powers
        omega = [P_h_Nom u of every DEWH ordered by device id, PV y, demand y], the grid MLD step
        (micro_grid_models.py:137-172) with delta = [y >= 0], z = delta y through K5, and the planned counterparts
        (``*_hat``) from the first step of the aggregate plan.  pv / resd: dict(omega [steps] actual, omega_hat
        [steps] forecast, gain) with gain = -P_pv_max P_pv_units / P_res_ave P_res_units.  Single-rank view (the
        per-device columns are this rank's agents).  -> dict of device tensors for results.grid_log_blocks."""
        from .parameters import grid_param_struct
        from .models import grid_mld_matrices
        gp = dict(grid_param_struct if grid_params is None else grid_params)
        dev, B = self.device, self.B
        u = log["u"].to(dev)
        steps = u.shape[0]
        order = np.argsort(np.asarray(device_ids)) if device_ids is not None else np.arange(B)
        order = torch.as_tensor(order, device=dev)
        cols, cols_hat = [(u * self.P_nom[None, :])[:, order]], [(u * self.P_nom[None, :])[:, order]]
        for src in (pv, resd):
            if src is not None:
                w = torch.as_tensor(src["omega"], dtype=torch.float64).to(dev).reshape(steps, 1)
                wh = torch.as_tensor(src.get("omega_hat", src["omega"]), dtype=torch.float64).to(dev).reshape(steps, 1)
                cols.append(float(src["gain"]) * w)
                cols_hat.append(float(src["gain"]) * wh)
        out = {}
        mats = {k: torch.as_tensor(v, dtype=torch.float64).to(dev).unsqueeze(0)
                for k, v in grid_mld_matrices(gp, sum(c.shape[1] for c in cols)).items()}
        for tag, parts in (("", cols), ("_hat", cols_hat)):
            omega = torch.cat(parts, dim=1).contiguous()
            d = cabi.make_dims(steps, 1, nx=0, nu=0, ndelta=1, nz=1, nmu=0, nomega=omega.shape[1], ny=1, nc=6)
            zero = torch.zeros((steps, 1), dtype=torch.float64, device=dev)
            _, y, _ = cabi.lsim_step(d, mats, None, None, zero, zero, omega)    # y = D4 omega does not need delta / z
            delta = (y >= 0).to(torch.float64)                                  # unique feasible auxiliaries (a12)
            z = delta * y
            _, y2, cons = cabi.lsim_step(d, mats, None, None, delta, z, omega)
            out.update({"omega" + tag: omega, "y" + tag: y2.reshape(steps), "delta" + tag: delta.reshape(steps),
                        "z" + tag: z.reshape(steps)})
            if not tag:
                out["cons"] = cons
        price = torch.as_tensor(price, dtype=torch.float64).to(dev)
        out["price"] = price[:steps] if price.dim() == 1 else price[0, :steps]
        return out

    def campaign(self, controllers, T0, demand, price, sim_steps, **kwargs):
        """Every named controller variant from the same initial state and data, one after the other
        (micro_grid_control_simulation.py:157-236 runs them side by side) -> {name: closed_loop log}."""
        return {name: self.closed_loop(T0, demand, price, sim_steps, controller=name, **kwargs) for name in controllers}
