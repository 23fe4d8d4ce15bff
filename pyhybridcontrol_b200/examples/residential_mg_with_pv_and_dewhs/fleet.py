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

    def control_step(self, x0, omega_forecast, cost_v, extra_constraints=()):
        """-> dict(v, obj, status, stats, u [B, Nt] view)."""
        res = self.batch.solve(x0, omega_forecast, cost_v=cost_v, extra_constraints=extra_constraints)
        res["u"] = res["v"].view(self.B, self.Nt, 3)[:, :, 0]
        return res

    def sim_step(self, T, u0, D_h):
        """Advance every tank one step with the re-parametrised simulation model (micro_grid_agents.py:389-408)."""
        T1, _, cons = cabi.dewh_sim_step(self.params, T, u0, D_h)
        return T1, cons

    def aggregate_power(self, u):
        """sum_b P_nom[b] u[b, k] over this rank's agents, then over ranks -> [Nt]."""
        return distributed.allreduce_aggregate(cabi.aggregate_power(u, self.P_nom))

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
                    N_sb_reduced=8, demand_minmax=None, u_init=None):
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
        (thermostat only).
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
        for k in range(sim_steps):
            pk = price[k:k + Nt] if price.dim() == 1 else price[:, k:k + Nt]
            extra = self._extra_sets(controller, k, scenarios, demand_minmax, N_sb_reduced)
            events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            events[-1][0].record()
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
