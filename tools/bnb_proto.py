"""DEV TOOL (not imported by the product, not an oracle): numpy twin of the GPU branch-and-bound kernel
(pyhybridcontrol_b200/csrc/milp_bnb.cu).  Used to choose branching/pivot rules by counting pivots and nodes
on the synthetic DEWH workload before spending GPU time.  One dense tableau per problem, bounded dual
simplex, depth-first search that re-uses the SAME tableau for every node (a node is just a set of bounds).
"""
import numpy as np

INF = np.inf


class DualSimplexBnB(object):
    def __init__(self, c, H, rhs, lb, ub, is_bin, big=1e7, ptol=1e-9, dtol=1e-9, itol=1e-6, branch="mostfrac",
                 flip=True):
        self.m, self.n = H.shape
        m, n = self.m, self.n
        self.c = np.concatenate([c, np.zeros(m)])
        self.lo0 = np.concatenate([np.where(np.isfinite(lb), lb, -big), np.zeros(m)])
        self.hi0 = np.concatenate([np.where(np.isfinite(ub), ub, np.where(c[:n] < 0, big, INF) if True else INF), np.full(m, INF)])
        # free columns get an artificial box so that the slack basis is dual feasible
        self.T = np.array(H, dtype=float)          # m x n, B^-1 N
        self.bbar = np.array(rhs, dtype=float)     # B^-1 rhs
        self.basis = np.arange(n, n + m)
        self.nb = np.arange(n)
        self.d = np.array(c, dtype=float)
        self.bin_idx = np.nonzero(is_bin)[0]
        self.is_bin = np.concatenate([is_bin, np.zeros(m, bool)])
        self.ptol, self.dtol, self.itol = ptol, dtol, itol
        self.pivots = 0
        self.flips = 0
        self.nodes = 0
        self.branch = branch
        self.flip = flip
        self.xN = np.zeros(n)
        self.zB = 0.0  # c_B' bbar tracked lazily (recomputed)

    # ---- node set-up: place nonbasics on the bound their reduced cost asks for, recompute basics
    def load_bounds(self, lo, hi):
        self.lo, self.hi = lo, hi
        v = self.nb
        l, h = lo[v], hi[v]
        atl = self.d >= 0
        x = np.where(atl, l, h)
        x = np.where(np.isfinite(x), x, np.where(np.isfinite(l), l, h))
        self.xN = x
        self.xB = self.bbar - self.T @ self.xN

    def objective(self):
        return float(self.c[self.basis] @ self.xB + self.c[self.nb] @ self.xN)

    def solve_lp(self, cutoff=INF, maxit=5000):
        """-> 'opt' | 'inf' | 'cut'"""
        T = self.T
        for _ in range(maxit):
            lo_b, hi_b = self.lo[self.basis], self.hi[self.basis]
            viol_lo = lo_b - self.xB
            viol_hi = self.xB - hi_b
            viol = np.maximum(viol_lo, viol_hi)
            r = int(np.argmax(viol))
            if viol[r] <= self.ptol:
                return "opt"
            below = viol_lo[r] > viol_hi[r]
            row = T[r]
            l, h = self.lo[self.nb], self.hi[self.nb]
            movable = h > l
            at_lower = self.xN <= l
            at_upper = self.xN >= h
            # need x_Br to increase (below) : contribution -T_rj * t_j > 0
            if below:
                cand = movable & (((row < -1e-9) & at_lower) | ((row > 1e-9) & at_upper))
                target = lo_b[r]
            else:
                cand = movable & (((row > 1e-9) & at_lower) | ((row < -1e-9) & at_upper))
                target = hi_b[r]
            if not cand.any():
                return "inf"
            ratios = np.where(cand, np.abs(self.d) / np.maximum(np.abs(row), 1e-300), INF)
            if self.flip:
                # bound-flipping ratio test: walk break points while the slope stays positive
                delta = abs(self.xB[r] - target)
                order = np.argsort(ratios)
                q = -1
                for j in order:
                    if not np.isfinite(ratios[j]):
                        break
                    rng = h[j] - l[j]
                    if np.isfinite(rng) and delta - abs(row[j]) * rng > self.ptol:
                        # flip j to its other bound, keep going
                        newx = h[j] if at_lower[j] else l[j]
                        dx = newx - self.xN[j]
                        self.xB -= T[:, j] * dx
                        self.xN[j] = newx
                        delta -= abs(row[j]) * rng
                        self.flips += 1
                        continue
                    q = int(j)
                    break
                if q < 0:
                    return "inf"
            else:
                q = int(np.argmin(ratios))
            piv = row[q]
            t = (self.xB[r] - target) / piv
            # primal update
            self.xB -= T[:, q] * t
            xq_new = self.xN[q] + t
            # pivot
            colq = T[:, q].copy()
            rowr = row / piv
            dq = self.d[q]
            bq = self.bbar[r] / piv
            T -= np.outer(colq, rowr)
            self.bbar -= colq * bq
            T[r] = rowr
            self.bbar[r] = bq
            T[:, q] = -colq / piv
            T[r, q] = 1.0 / piv
            self.d -= dq * rowr
            self.d[q] = -dq / piv
            leaving = self.basis[r]
            self.basis[r] = self.nb[q]
            self.nb[q] = leaving
            self.xB[r] = xq_new
            self.xN[q] = target
            self.pivots += 1
            if cutoff < INF and self.objective() >= cutoff:
                return "cut"
        return "lim"

    def solution(self):
        x = np.zeros(self.c.size)
        x[self.basis] = self.xB
        x[self.nb] = self.xN
        return x

    def run(self, max_nodes=100000, gap=0.0):
        lo, hi = self.lo0.copy(), self.hi0.copy()
        best, best_x = INF, None
        stack = [(lo, hi, -INF)]
        while stack and self.nodes < max_nodes:
            lo, hi, bound = stack.pop()
            if bound >= best - 1e-9 * max(1, abs(best)):
                continue
            self.nodes += 1
            self.load_bounds(lo, hi)
            st = self.solve_lp(cutoff=best)
            if st != "opt":
                continue
            obj = self.objective()
            if obj >= best - 1e-9 * max(1, abs(best)):
                continue
            x = self.solution()
            xb = x[self.bin_idx]
            frac = np.abs(xb - np.round(xb))
            if frac.max(initial=0) <= self.itol:
                best, best_x = obj, x.copy()
                continue
            if self.branch == "mostfrac":
                j = self.bin_idx[int(np.argmax(frac))]
            elif self.branch == "first":
                j = self.bin_idx[np.nonzero(frac > self.itol)[0][0]]
            else:
                j = self.bin_idx[np.nonzero(frac > self.itol)[0][-1]]
            first = 1.0 if x[j] >= 0.5 else 0.0
            for val in (1.0 - first, first):
                l2, h2 = lo.copy(), hi.copy()
                l2[j] = h2[j] = val
                stack.append((l2, h2, obj))
        return best, best_x


if __name__ == "__main__":
    import sys, time
    sys.path.insert(0, ".")
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    wl = syn.dewh_batch(B, N_p, seed=1)
    Nt = wl["Nt"]
    for branch in ("mostfrac", "first", "last"):
        for flip in (False, True):
            tot_p = tot_n = tot_f = 0
            ok = True
            for b in range(B):
                mats = {k: v[b] for k, v in wl["mats"].items()}
                full, d, vt = omld.complete(mats, nu_l=1)
                evo = oc.condense(full, d, Nt)
                prob = oa.build_problem(evo, d, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
                s = DualSimplexBnB(prob.c, prob.H, prob.rhs, prob.lb, prob.ub, prob.is_bin, branch=branch, flip=flip)
                obj, x = s.run()
                st, oref, vref = osv.solve_milp(prob)
                good = abs(obj + prob.c0 - oref) <= 1e-6 * max(1, abs(oref))
                ok &= good
                if not good:
                    print("MISMATCH", b, obj + prob.c0, oref)
                tot_p += s.pivots; tot_n += s.nodes; tot_f += s.flips
            print(f"N_p={N_p} branch={branch:8s} flip={flip}: ok={ok} pivots/prob={tot_p/B:.0f} nodes/prob={tot_n/B:.0f} flips/prob={tot_f/B:.0f}")
