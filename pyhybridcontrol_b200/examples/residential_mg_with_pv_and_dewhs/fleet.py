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

    def control_step(self, x0, omega_forecast, cost_v):
        """-> dict(v, obj, status, stats, u [B, Nt] view)."""
        res = self.batch.solve(x0, omega_forecast, cost_v=cost_v)
        res["u"] = res["v"].view(self.B, self.Nt, 3)[:, :, 0]
        return res

    def sim_step(self, T, u0, D_h):
        """Advance every tank one step with the re-parametrised simulation model (micro_grid_agents.py:389-408)."""
        T1, _, cons = cabi.dewh_sim_step(self.params, T, u0, D_h)
        return T1, cons

    def aggregate_power(self, u):
        """sum_b P_nom[b] u[b, k] over this rank's agents, then over ranks -> [Nt]."""
        return distributed.allreduce_aggregate(cabi.aggregate_power(u, self.P_nom))

    def closed_loop(self, T0, demand, price, sim_steps, demand_actual=None):
        """Closed-loop simulation of the shard (reference loop: examples/.../micro_grid_control_simulation.py:184-236
        with the per-device work of micro_grid_agents.py:699-700, 733-740): at every step k the forecast window
        demand[:, k:k+Nt] and price[k:k+Nt] give the MPC problem, the first control is applied to the re-parametrised
        simulation model with the actual draw, and the aggregate power is exchanged.  Everything stays in HBM; the
        log comes back as device tensors.

        T0 [B] initial temperatures; demand [B, sim_steps + Nt] (L/s); price [sim_steps + Nt] or [B, sim_steps + Nt];
        demand_actual [B, sim_steps] defaults to demand[:, :sim_steps].
        -> dict(T [sim_steps + 1, B], u [sim_steps, B], obj [sim_steps, B], status [sim_steps, B],
                P_agg [sim_steps, Nt], cons [sim_steps, B, 2])"""
        dev, B, Nt = self.device, self.B, self.Nt
        T = torch.as_tensor(T0, dtype=torch.float64).to(dev).reshape(B).clone()
        demand = torch.as_tensor(demand, dtype=torch.float64).to(dev)
        price = torch.as_tensor(price, dtype=torch.float64).to(dev)
        actual = demand[:, :sim_steps] if demand_actual is None else torch.as_tensor(demand_actual, dtype=torch.float64).to(dev)
        log = dict(T=[T.clone()], u=[], obj=[], status=[], P_agg=[], cons=[])
        self.build()                                   # the control model does not change along the run
        for k in range(sim_steps):
            pk = price[k:k + Nt] if price.dim() == 1 else price[:, k:k + Nt]
            res = self.control_step(T.reshape(B, 1), demand[:, k:k + Nt].contiguous(), self.cost_from_prices(pk))
            u0 = res["u"][:, 0].contiguous()
            T, cons = self.sim_step(T, u0, actual[:, k].contiguous())
            log["T"].append(T.clone()); log["u"].append(u0); log["obj"].append(res["obj"])
            log["status"].append(res["status"]); log["P_agg"].append(self.aggregate_power(res["u"])); log["cons"].append(cons)
        return {k: torch.stack(v) for k, v in log.items()}
