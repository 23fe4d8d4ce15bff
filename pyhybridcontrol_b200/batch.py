"""Batched front door of the GPU hot path: condense -> assemble -> mixed-integer solve for B agents at once.

``BatchMpc`` is what the per-agent ``MpcController`` objects are thin views over (B = 1) and what the
benchmark and the closed-loop driver use directly.  All arithmetic happens in the CUDA kernels behind
``cabi``; torch tensors are device buffers only.

Reference call stack being replaced (per agent, serially, in the reference):
  MpcController.build  (controllers/mpc_controller.py:76-101)  -> MldEvoMatrices.update (K1)
  ConstraintSolvedController.gen_evo_constraints (controllers/controller_base.py:411-456)  (K2)
  ConstraintSolvedController.solve (controllers/controller_base.py:491-540) -> cvxpy -> Gurobi  (K3/K4)
"""
import numpy as np
import torch

from . import cabi


def _dev_tensor(a, device, dtype=torch.float64):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device)


class BatchMpc(object):
    def __init__(self, mats, N_p, N_tilde=None, nu_l=0, nmu_l=0, B=None, device="cuda", opts=None, solver="auto",
                 dp_opts=None, dp_bound="auto"):
        """mats: name -> array [B|1, r, c] (or [r, c]); missing blocks are zero, C defaults to I (nx == ny).

        solver: "auto" -- the exact stage-DP kernels (csrc/stage_dp.cu) when the MLD is in their class (scalar
        state, binary inputs, one slack per row: every DEWH of the reference example), else the general
        branch-and-cut kernel (csrc/milp_bnc.cu); "bnc" / "stage_dp" force one of them.
        dp_bound: "constant" (one value per table cell), "linear" (a line per cell: the bound to use when slack
        penalties are active along the whole trajectory -- full-horizon robust constraint sets) or "auto" (default:
        linear for a solve whose constraint sets carry scenarios over more than half of the horizon, else constant)."""
        self.device = torch.device(device)
        self.N_p = int(N_p)
        self.Nt = int(N_tilde) if N_tilde is not None else self.N_p + 1
        m = {}
        Bs = set()
        for k, v in mats.items():
            if v is None:
                continue
            t = _dev_tensor(v, self.device)
            if t.dim() == 2:
                t = t.unsqueeze(0)
            if t.numel() == 0:
                continue
            m[k] = t
            Bs.add(t.shape[0])
        Bs.discard(1)
        if len(Bs) > 1:
            raise ValueError("inconsistent batch sizes in mats: %s" % sorted(Bs))
        self.B = int(B) if B is not None else (Bs.pop() if Bs else 1)

        def rows(names):
            return max([m[n].shape[1] for n in names if n in m] + [0])

        def cols(names):
            return max([m[n].shape[2] for n in names if n in m] + [0])
        nx = rows(("A", "B1", "B2", "B3", "B4", "b5"))
        if "C" not in m and nx:
            m["C"] = torch.eye(nx, dtype=torch.float64, device=self.device).unsqueeze(0)
        ny = rows(("C", "D1", "D2", "D3", "D4", "d5"))
        nc = rows(("E", "F1", "F2", "F3", "F4", "f5", "G", "Psi"))
        self.dims = cabi.make_dims(self.B, self.Nt, nx=nx, nu=cols(("B1", "D1", "F1")), ndelta=cols(("B2", "D2", "F2")),
                                   nz=cols(("B3", "D3", "F3")), nmu=cols(("Psi",)), nomega=cols(("B4", "D4", "F4")),
                                   ny=ny, nc=nc)
        self.mats = m
        d = self.dims
        self.nv = d.nv
        self.nvt = d.nv * self.Nt
        self.nwt = d.nomega * self.Nt
        self.mrows = d.nc * self.Nt
        self.nu_l, self.nmu_l = int(nu_l), int(nmu_l)
        step_bin = [0] * (d.nu - self.nu_l) + [1] * self.nu_l + [1] * d.ndelta + [0] * d.nz + \
                   [0] * (d.nmu - self.nmu_l) + [1] * self.nmu_l
        self.is_bin_step = np.array(step_bin, dtype=np.uint8)
        lb_step = np.full(d.nv, -np.inf)
        ub_step = np.full(d.nv, np.inf)
        lb_step[self.is_bin_step == 1] = 0.0
        ub_step[self.is_bin_step == 1] = 1.0
        lb_step[d.nu + d.ndelta + d.nz:] = np.maximum(lb_step[d.nu + d.ndelta + d.nz:], 0.0)   # mu >= 0
        self.lb_v = np.tile(lb_step, self.Nt)
        self.ub_v = np.tile(ub_step, self.Nt)
        self.is_bin_v = np.tile(self.is_bin_step, self.Nt)
        self.opts = opts if opts is not None else cabi.default_opts()
        if solver not in ("auto", "bnc", "stage_dp"):
            raise ValueError("solver must be 'auto', 'bnc' or 'stage_dp'")
        self.solver = solver
        # table resolution: the optimum does not depend on it; since the grid follows the violation-free band stage by
        # stage (round 2) 4096 cells resolve what 8192 cells of one global window did, at two thirds of the sweep time
        if dp_bound not in ("constant", "linear", "auto"):
            raise ValueError("dp_bound must be 'constant', 'linear' or 'auto'")
        self.dp_bound = dp_bound
        if dp_opts is None:
            dp_opts = cabi.stage_dp_default_opts()
            if dp_bound == "linear":
                dp_opts.bound = cabi.DP_BOUND_LINEAR
            if cabi.stage_dp_supported(d):       # the two stage buffers must fit shared memory in this format
                dp_opts.cells = min(dp_opts.cells, cabi.stage_dp_max_cells(d, dp_opts))
        else:
            self.dp_bound = "linear" if dp_opts.bound == cabi.DP_BOUND_LINEAR else "constant"
        self.dp_opts = dp_opts
        self.stage_dp_ok = self._stage_dp_class()
        if solver == "stage_dp" and not self.stage_dp_ok:
            raise ValueError("this MLD is outside the stage-DP class (needs nx == 1, nz == 0, binary inputs, "
                             "Psi = -diag(d >= 0), A > 0)")
        self.evo = None
        self._lb_dev = self._ub_dev = self._bin_dev = None
        self.disable_soft_constraints = False

    def _dp_opts_for(self, robust):
        """the stage-DP options of one solve: under dp_bound="auto" a robust solve (scenario sets over most of the
        horizon) gets linear cells, with the cell count its stage buffers allow"""
        if self.dp_bound != "auto" or not robust or self.dims.nc != 2 or self.dims.nu + self.dims.ndelta != 1:
            return self.dp_opts
        o = cabi.stage_dp_default_opts(bound=cabi.DP_BOUND_LINEAR, mip_rel_gap=self.dp_opts.mip_rel_gap,
                                       max_nodes=self.dp_opts.max_nodes, feas_tol=self.dp_opts.feas_tol,
                                       fuse_search=self.dp_opts.fuse_search)
        o.cells = min(self.dp_opts.cells, cabi.stage_dp_max_cells(self.dims, o))
        return o

    def _stage_dp_class(self):
        """One-time host check that every agent of the batch is in the class of hmpc_stage_dp_solve_f64."""
        d = self.dims
        if not cabi.stage_dp_supported(d) or self.nu_l != d.nu or self.nmu_l != 0:
            return False
        A = self.mats.get("A")
        if A is None or not bool((A > 0).all()):
            return False
        aN = A.reshape(-1) ** self.Nt
        if not bool(((aN > 1e-3) & (aN < 1e3)).all()):
            return False
        Psi = self.mats.get("Psi")
        if d.nmu:
            if Psi is None:
                return False
            off = Psi - torch.diag_embed(torch.diagonal(Psi, dim1=1, dim2=2))
            if bool((off != 0).any()) or bool((torch.diagonal(Psi, dim1=1, dim2=2) > 0).any()):
                return False
        return True

    # ---- index helpers (v(k) = [u; delta; z; mu], reference: controllers/components/variables.py:233-241)
    def var_slices(self):
        d = self.dims
        o = 0
        out = {}
        for name, dim in (("u", d.nu), ("delta", d.ndelta), ("z", d.nz), ("mu", d.nmu)):
            out[name] = (o, o + dim)
            o += dim
        return out

    def var_index(self, name):
        a, b = self.var_slices()[name]
        return (np.arange(self.Nt)[:, None] * self.nv + np.arange(a, b)[None, :]).ravel()

    # ---- K1
    def build(self, want=cabi.EVO_NAMES, reuse=False):
        """K1.  reuse=True writes into the tensors of the previous build (no allocation in the control loop)."""
        self.evo = cabi.condense(self.dims, self.mats, want=want, out=self.evo if (reuse and self.evo) else None)
        return self.evo

    def _bounds_dev(self):
        lb, ub = self.lb_v.copy(), self.ub_v.copy()
        if self.disable_soft_constraints and self.dims.nmu:
            idx = self.var_index("mu")
            lb[idx] = 0.0
            ub[idx] = 0.0
        key = (self.disable_soft_constraints,)
        if self._lb_dev is None or self._bkey != key:
            self._lb_dev = _dev_tensor(lb, self.device)
            self._ub_dev = _dev_tensor(ub, self.device)
            self._bin_dev = torch.as_tensor(self.is_bin_v, dtype=torch.uint8).to(self.device)
            self._bkey = key
        return self._lb_dev, self._ub_dev, self._bin_dev

    # ---- K2 + cost + K3/K4
    def constraint_rows(self, x0, omega=None, scenarios=None, N_tilde=None):
        """(H rows view, rhs) of one evolution-constraint set (reference: controller_base.py:411-456)."""
        rows = self.mrows if N_tilde is None else int(N_tilde) * self.dims.nc
        rhs = cabi.constraint_rhs(self.dims, self.evo, x0, omega, scenarios=scenarios, rows=rows)
        H = self.evo["H_v"][:, :rows, :]
        return H, rhs

    def linear_cost(self, w_v=None, w_x=None, w_y=None, x0=None, omega=None):
        """Linear atoms -> (c [B,nvt], c0 [B]); weights on x / y are pulled back through Gamma_v / L_v."""
        d = self.dims
        xc = yc = None
        if w_x is not None:
            xc = cabi.predict(self.evo["Phi_x"], None, self.evo["Gamma_omega"], self.evo["Gamma_5"].reshape(d.B, -1),
                              x0, None, omega)
        if w_y is not None:
            yc = cabi.predict(self.evo["L_x"], None, self.evo["L_omega"], self.evo["L_5"].reshape(d.B, -1), x0, None,
                              omega)
        if w_v is None and w_x is None and w_y is None:
            z = torch.zeros((d.B, self.nvt), dtype=torch.float64, device=self.device)
            return z, torch.zeros((d.B,), dtype=torch.float64, device=self.device)
        return cabi.linear_cost(d.B, self.nvt, w_v=w_v, w_x=w_x, Gamma_v=self.evo.get("Gamma_v"), xc=xc, w_y=w_y,
                                L_v=self.evo.get("L_v"), yc=yc)

    def _mat_B(self, name, rows, cols):
        t = self.mats.get(name)
        if t is None:
            return torch.zeros((self.B, rows, cols), dtype=torch.float64, device=self.device)
        return t.expand(self.B, rows, cols)

    def _stage_terms(self, quad, x0, omega):
        """Quadratic / L1 atoms -> hmpc_stage_terms tensors + the part that folds into the linear cost.

        quad keys (all optional, weights >= 0, [B|1, Nt, dim] or [Nt, dim]):  x2, x1 (state, squared / absolute),
        y2, y1 (outputs), mu2 (slacks, squared), u2 / u1, delta2 / delta1 (binaries: b^2 = |b| = b, linear)."""
        d, B, Nt = self.dims, self.B, self.Nt
        nb = d.nu + d.ndelta

        def W(key, dim):
            t = quad.get(key)
            if t is None:
                return None
            t = _dev_tensor(t, self.device).reshape(-1, Nt, dim)
            if bool((t < 0).any()):
                raise ValueError("quadratic / L1 weight '%s' must be non-negative (convex cost)" % key)
            return t.expand(B, Nt, dim)
        fold = torch.zeros((B, Nt, self.nv), dtype=torch.float64, device=self.device)
        for key, off, dim in (("u2", 0, d.nu), ("u1", 0, d.nu), ("delta2", d.nu, d.ndelta), ("delta1", d.nu, d.ndelta)):
            t = W(key, dim) if dim else None
            if t is not None:
                fold[:, :, off:off + dim] += t
        h, ga, r, wq, w1 = [], [], [], [], []
        zero = torch.zeros((B, Nt), dtype=torch.float64, device=self.device)
        x2, x1 = W("x2", d.nx), W("x1", d.nx)
        if x2 is not None or x1 is not None:
            xfree = cabi.predict(self.evo["Phi_x"], None, self.evo["Gamma_omega"], self.evo["Gamma_5"].reshape(B, -1),
                                 x0, None, omega).reshape(B, Nt)
            h.append(torch.ones(B, dtype=torch.float64, device=self.device))
            ga.append(torch.zeros((B, nb), dtype=torch.float64, device=self.device))
            r.append(xfree)
            wq.append(zero if x2 is None else x2[:, :, 0])
            w1.append(zero if x1 is None else x1[:, :, 0])
        y2, y1 = W("y2", d.ny), W("y1", d.ny)
        if y2 is not None or y1 is not None:
            yfree = cabi.predict(self.evo["L_x"], None, self.evo["L_omega"], self.evo["L_5"].reshape(B, -1), x0, None,
                                 omega).reshape(B, Nt, d.ny)
            Cm, Dm = self._mat_B("C", d.ny, d.nx), torch.cat([self._mat_B("D1", d.ny, d.nu), self._mat_B("D2", d.ny, d.ndelta)], dim=2)
            for rr in range(d.ny):
                h.append(Cm[:, rr, 0]); ga.append(Dm[:, rr, :]); r.append(yfree[:, :, rr])
                wq.append(zero if y2 is None else y2[:, :, rr])
                w1.append(zero if y1 is None else y1[:, :, rr])
        terms = {}
        if h:
            if len(h) > 4:
                raise NotImplementedError("at most 4 state / output terms per stage")
            terms = dict(h=torch.stack(h, dim=1), ga=torch.stack(ga, dim=1), r=torch.stack(r, dim=2),
                         wq=torch.stack(wq, dim=2), w1=torch.stack(w1, dim=2))
        mu2 = W("mu2", d.nmu) if d.nmu else None
        if mu2 is not None:
            terms["qmu"] = mu2
        return terms, fold.reshape(B, -1)

    def solve(self, x0, omega, cost_v=None, w_x=None, w_y=None, scenarios=None, extra_constraints=(),
              with_std_constraints=True, quad=None, rhs=None):
        """One control step for the whole batch.  Returns dict(v, obj, status, stats, c0) of device tensors.
        ``quad``: convex quadratic / L1 weights (see _stage_terms) -- an MIQP, stage-DP path only.
        ``rhs``: the folded right-hand side a previous call returned (same x0, omega and constraint sets, new
        cost) -- stage-DP path only; skips K2."""
        if self.evo is None:
            raise RuntimeError("build() must be called before solve()")
        d = self.dims
        x0 = _dev_tensor(x0, self.device).reshape(d.B, d.nx) if d.nx else None
        omega = _dev_tensor(omega, self.device).reshape(d.B, self.nwt) if self.nwt else None
        if scenarios is not None:
            scenarios = _dev_tensor(scenarios, self.device)
        cost_v = None if cost_v is None else _dev_tensor(cost_v, self.device).reshape(-1, self.nvt)
        w_x = None if w_x is None else _dev_tensor(w_x, self.device).reshape(d.B, -1)
        w_y = None if w_y is None else _dev_tensor(w_y, self.device).reshape(d.B, -1)
        if w_x is None and w_y is None and cost_v is not None:
            c, c0 = cost_v, torch.zeros((d.B,), dtype=torch.float64, device=self.device)
        else:
            c, c0 = self.linear_cost(cost_v, w_x, w_y, x0, omega)
        Hs, rs = [], []
        use_dp = self.solver == "stage_dp" or (self.solver == "auto" and self.stage_dp_ok)
        robust = scenarios is not None and with_std_constraints
        for ec in extra_constraints:
            if ec.get("omega_scenarios_k") is not None and (ec.get("N_tilde") is None or 2 * int(ec["N_tilde"]) > self.Nt):
                robust = True
        if rhs is not None:
            if not use_dp or quad:
                raise NotImplementedError("a cached right-hand side is taken on the stage-DP MILP path only")
            lb, ub, isb = self._bounds_dev()
            v, obj, status, stats = cabi.stage_dp_solve(d, self.mats, rhs, c.contiguous(), lb, ub, isb, self.dp_opts)
            return dict(v=v, obj=obj + c0, status=status, stats=stats, c0=c0, solver="stage_dp", rhs=rhs)
        if d.nc and with_std_constraints:
            H, r = self.constraint_rows(x0, omega, scenarios=scenarios)
            Hs.append(H)
            rs.append(r)
        for ec in extra_constraints:
            w2 = ec.get("omega_tilde_k")
            sc = ec.get("omega_scenarios_k")
            xk = ec.get("x_k")                 # a set may carry its own state (controller_base.py:411-456)
            w2 = omega if w2 is None else _dev_tensor(w2, self.device).reshape(d.B, self.nwt)
            sc = None if sc is None else _dev_tensor(sc, self.device).reshape(d.B, self.nwt, -1)
            xk = x0 if xk is None or not d.nx else _dev_tensor(xk, self.device).reshape(d.B, d.nx)
            H, r = self.constraint_rows(xk, w2, scenarios=sc, N_tilde=ec.get("N_tilde"))
            Hs.append(H)
            rs.append(r)
        lb, ub, isb = self._bounds_dev()
        terms = None
        if quad:
            if not (use_dp and rs):
                raise NotImplementedError("quadratic / L1 cost atoms run on the stage-DP path only (scalar-state MLDs); "
                                          "the branch-and-cut kernel is an MILP solver")
            terms, fold = self._stage_terms(quad, x0, omega)
            c = c.expand(d.B, self.nvt) + fold
        if use_dp and rs:
            # every constraint set shares the rows of H_v (prefixes for reduced horizons), so the sets fold into
            # the row-wise minimum of their right-hand sides
            rhs = torch.full((d.B, self.mrows), float("inf"), dtype=torch.float64, device=self.device)
            for r in rs:
                rhs[:, :r.shape[1]] = torch.minimum(rhs[:, :r.shape[1]], r)
            if len(rs) == 1 and rs[0].shape[1] == self.mrows:
                rhs = rs[0]
            v, obj, status, stats = cabi.stage_dp_solve(d, self.mats, rhs, c.contiguous(), lb, ub, isb,
                                                        self._dp_opts_for(robust and terms is None), terms=terms)
            return dict(v=v, obj=obj + c0, status=status, stats=stats, c0=c0, solver="stage_dp", rhs=rhs)
        if len(Hs) == 1 and Hs[0].shape[1] == self.mrows:
            H, rhs = self.evo["H_v"], rs[0]
        elif Hs:
            H, rhs = torch.cat(Hs, dim=1).contiguous(), torch.cat(rs, dim=1).contiguous()
        else:
            H = torch.zeros((d.B, 0, self.nvt), dtype=torch.float64, device=self.device)
            rhs = torch.zeros((d.B, 0), dtype=torch.float64, device=self.device)
        v, obj, status, stats = cabi.milp_solve(c, H, rhs, lb, ub, isb, self.opts)
        return dict(v=v, obj=obj + c0, status=status, stats=stats, c0=c0, solver="bnc")

    def predictions(self, v, x0, omega):
        """x~ [B, nx*Nt], y~ [B, ny*Nt] for a given v~ (reference: variables.py:245-286)."""
        d = self.dims
        x0 = _dev_tensor(x0, self.device).reshape(d.B, d.nx) if d.nx else None
        omega = _dev_tensor(omega, self.device).reshape(d.B, self.nwt) if self.nwt else None
        xt = yt = None
        if d.nx:
            xt = cabi.predict(self.evo["Phi_x"], self.evo["Gamma_v"], self.evo["Gamma_omega"],
                              self.evo["Gamma_5"].reshape(d.B, -1), x0, v, omega)
        if d.ny:
            yt = cabi.predict(self.evo["L_x"], self.evo["L_v"], self.evo["L_omega"], self.evo["L_5"].reshape(d.B, -1),
                              x0, v, omega)
        return xt, yt
