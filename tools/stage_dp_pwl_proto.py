"""DEV TOOL: research prototype (round-1 notes, profiles/r1_notes.md) -- a nodal piecewise-linear lower bound for the
stage-DP table, built on the numpy twin.  Valid by construction (chord deficiency at every convex kink is subtracted),
tighter than the constant-per-cell bound where slack penalties are active, looser in the nominal regime (it loses
the exact identity map of the no-input action); max(constant, piecewise-linear) is the candidate for round 2."""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
from stage_dp_proto import StageDp, from_dewh_problem, _round_down_f32

class StageDpPwl(StageDp):
    """nodal piecewise-linear lower bound: LB_k(s) = interpolation of L_k[j] at nodes s_j = S0 + j w, j = 0..G"""
    def _table(self):
        Nt, G, w, S0 = self.Nt, self.G, self.w, self.S0
        nodes = S0 + np.arange(G + 1) * w
        self.L = np.zeros((Nt + 1, G + 1), dtype=np.float32)
        cur = np.zeros(G + 1)
        def interp(vals, s, out):
            x = (s - S0) / w
            j = np.floor(x).astype(int)
            inside = (x >= 0) & (x <= G)
            j = np.clip(j, 0, G - 1)
            t = x - j
            v = vals[j] + (vals[j + 1] - vals[j]) * t
            return np.where(inside, v, out)
        for k in range(Nt - 1, 0, -1):
            out = self.tailmin[k + 1]
            akk = self.ak[k]
            def F(s):   # min over actions of stage cost + interpolated next bound, exact at given states
                pen = np.zeros_like(s)
                for i in range(self.nc):
                    pen = pen + self.q[k, i] * np.maximum(0.0, self.e[i] * akk * s - self.rhs[k, i])
                stay = interp(cur, s, out)
                move = self.c[k] + interp(cur, s + self.shift[k], out)
                return pen + np.minimum(stay, move)
            Fn = F(nodes)
            # interior convex-kink candidates of each cell: hinge kinks, and the point whose shifted image is a node
            defic = np.zeros(G)
            cand = []
            for i in range(self.nc):
                if self.e[i] != 0:
                    cand.append(np.full(G, self.rhs[k, i] / (self.e[i] * akk)))      # s where row i switches
            # shifted node: s + shift = node  ->  s = node_m - shift; within cell j: the unique such point
            fr = (self.shift[k] / w) % 1.0
            cand.append(nodes[:-1] + (1.0 - fr) * w if fr > 0 else nodes[:-1])
            lo, hi = nodes[:-1], nodes[1:]
            for t in cand:
                t = np.clip(t, lo, hi)
                lam = (t - lo) / w
                chord = Fn[:-1] + (Fn[1:] - Fn[:-1]) * lam
                defic = np.maximum(defic, chord - F(t))
            defic = np.where(np.isfinite(defic), defic, 0.0)
            Ln = Fn.copy()
            Ln[:-1] -= defic; 
            Ln[1:] = np.minimum(Ln[1:], Fn[1:] - defic)
            # window edges: beyond the window the bound is `tailmin`; keep continuity conservative
            self.L[k] = _round_down_f32(Ln - 1e-12 * np.abs(Ln))
            cur = self.L[k].astype(np.float64)
    def bound(self, k, s):
        if k >= self.Nt: return 0.0
        x = (s - self.S0) / self.w
        if x < 0 or x > self.G: return float(self.tailmin[k])
        j = min(int(np.floor(x)), self.G - 1)
        t = x - j
        Lk = self.L[k]
        return float(Lk[j]) + (float(Lk[j + 1]) - float(Lk[j])) * t - 1e-12

if __name__ == '__main__':
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p=int(sys.argv[1]) if len(sys.argv)>1 else 48
    S=32
    wl = syn.dewh_batch(64, N_p, seed=5); Nt=wl['Nt']
    rng=np.random.default_rng(123)
    scen = wl["omega"][:, :, None] * rng.uniform(0.5, 1.8, size=(64, Nt, S))
    for b in [62,45,17,13,52,63,0,1]:
        mats = {k: v[b] for k, v in wl["mats"].items()}
        full, d, vt = omld.complete(mats, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]), omega_scenarios=scen[b])
        args = from_dewh_problem(mats, prob, Nt)
        st, oref, vref = osv.solve_milp(prob, polish=True)
        line = "agent %2d ref %.6f |" % (b, oref - prob.c0)
        for cls, G in ((StageDp, 8192), (StageDpPwl, 8192), (StageDpPwl, 2048), (StageDpPwl, 512)):
            dp = cls(*args, cells=G)
            obj, u, nodes = dp.solve(max_nodes=300000)
            line += " %s G=%d: %.6f nodes %d |" % (cls.__name__[7:] or "const", G, obj, nodes)
        print(line)
