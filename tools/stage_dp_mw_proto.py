"""DEV TOOL (not imported by the product, not an oracle): numpy twin of the MOVING-WINDOW value table of
pyhybridcontrol_b200/csrc/stage_dp.cu (round 2).

The grid of stage k covers the absolute cells [o_k, o_k + G) of a lattice of width w that is shared by all stages (an
absolute cell a holds the scaled states s in [a w, (a + 1) w)), so the no-input action still maps a cell onto a cell
exactly; o_k follows the violation-free band of the stage.  Everything below / above the window of a stage is ONE
semi-infinite cell each, with a bound that is valid for all of its states (penalty at its favourable edge, minimum
over every next-stage cell its image can touch) -- so leaving the window never falls back to the trivial bound.
`lin=True` stores a line per cell (value at the left / right edge) instead of a constant: see
tools/stage_dp_lin_proto.py for the construction."""
import numpy as np

from stage_dp_proto import StageDp, EDGE_EPS, from_dewh_problem  # noqa: F401
from stage_dp_lin_proto import hull_line


class StageDpMw(StageDp):
    def __init__(self, a, b, e, rhs, c, q, cells=2048, lin=False, below=2.0, above=1.0):
        self.lin, self.below, self.above = lin, below, above
        StageDp.__init__(self, a, b, e, rhs, c, q, cells)

    def _window(self):
        Nt, G = self.Nt, self.G
        rlo = rhi = 0.0
        margin = np.abs(self.shift).max()
        a_k, b_k = np.zeros(Nt), np.zeros(Nt)
        E = U = 0.0
        for k in range(Nt):
            lo_k, hi_k = -np.inf, np.inf
            for i in range(self.nc):
                if self.e[i] == 0.0:
                    continue
                lim = self.rhs[k, i] / self.e[i] / self.ak[k]
                if self.e[i] > 0:
                    hi_k = min(hi_k, lim)
                else:
                    lo_k = max(lo_k, lim)
            lo_c = min(max(lo_k, rlo), rhi)
            hi_c = max(min(hi_k, rhi), rlo)
            # slew-limited envelopes: the band edge a trajectory can actually follow (the state climbs by at most one
            # shift per stage and never falls faster than the most negative shift)
            smax, smin = max(0.0, self.shift[k - 1]) if k else 0.0, min(0.0, self.shift[k - 1]) if k else 0.0
            E = lo_c if k == 0 else min(lo_c, E + smax)
            U = hi_c if k == 0 else max(hi_c, U + smin)
            a_k[k] = max(min(E, U) - self.below * margin, rlo - 0.01 * margin)
            b_k[k] = min(max(E, U) + self.above * margin, rhi + 0.01 * margin)
            b_k[k] = max(b_k[k], a_k[k])
            rlo += min(0.0, self.shift[k])
            rhi += max(0.0, self.shift[k])
        W = max((b_k - a_k).max(), 5.4 * margin * (1.0 + 16.0 / G))
        self.w = W / G
        self.S0 = 0.0
        self.o = np.floor(a_k / self.w).astype(np.int64)
        self.tailmin = np.concatenate([np.cumsum(np.minimum(self.c, 0.0)[::-1])[::-1], [0.0]])

    def _pen_edge(self, k, s, rows):
        p = self.ak[k] * s
        return sum(self.q[k, i] * max(0.0, self.e[i] * p - self.rhs[k, i]) for i in rows)

    KB = 8          # one-shift-wide bands on each side of the window; band KB = everything beyond them

    def _band_index(self, idx):
        """local cell index outside [0, G) -> (side, band)"""
        if idx < 0:
            return 0, min((-idx - 1) // self.Wb, self.KB)
        return 1, min((idx - self.G) // self.Wb, self.KB)

    def _table(self):
        Nt, G, w, KB = self.Nt, self.G, self.w, self.KB
        self.Wb = Wb = int(np.ceil(np.abs(self.shift).max() / w)) + 1
        cells = np.arange(G)
        self.YL = np.zeros((Nt + 1, G))
        self.YR = np.zeros((Nt + 1, G))
        self.band = np.zeros((Nt + 1, 2, KB + 1))     # [stage, side (0 below / 1 above), band]
        o = np.concatenate([self.o, [self.o[-1]]])
        curL, curR = np.zeros(G), np.zeros(G)
        for k in range(Nt - 1, 0, -1):
            bn = self.band[k + 1]
            d0 = int(o[k] - o[k + 1])

            def at(arr, idx):
                below = np.minimum((-idx - 1) // Wb, KB)
                above = np.minimum((idx - G) // Wb, KB)
                v = np.where(idx < 0, bn[0][np.clip(below, 0, KB)], bn[1][np.clip(above, 0, KB)])
                ok = (idx >= 0) & (idx < G)
                v[ok] = arr[idx[ok]]
                return v
            r = self.shift[k] / w
            i0 = int(np.floor(r))
            fr = r - i0
            c = self.c[k]
            akk = self.ak[k]
            edge = fr < EDGE_EPS or fr > 1.0 - EDGE_EPS
            sL, sR = at(curL, cells + d0), at(curR, cells + d0)
            if not self.lin or edge:
                lo_i, hi_i = (i0 - 1, i0 + 1) if fr < EDGE_EPS else ((i0, i0 + 2) if fr > 1.0 - EDGE_EPS else (i0, i0 + 1))
                mv = np.full(G, np.inf)
                for d in range(lo_i, hi_i + 1):
                    mv = np.minimum(mv, np.minimum(at(curL, cells + d0 + d), at(curR, cells + d0 + d)))
            if not self.lin:
                lo_e = akk * ((o[k] + cells - EDGE_EPS) * w)
                hi_e = akk * ((o[k] + cells + 1.0 + EDGE_EPS) * w)
                pen = np.zeros(G)
                for i in range(self.nc):
                    ed = lo_e if self.e[i] >= 0 else hi_e
                    pen += self.q[k, i] * np.maximum(0.0, self.e[i] * ed - self.rhs[k, i])
                best = pen + np.minimum(np.minimum(sL, sR), c + mv)
                oL = oR = best
            else:
                if edge:
                    y_lo = np.minimum(sL, c + mv)
                    y_hi = np.minimum(sR, c + mv)
                    oL, oR = y_lo, y_hi
                else:
                    mL, mR = at(curL, cells + d0 + i0), at(curR, cells + d0 + i0)
                    nL, nR = at(curL, cells + d0 + i0 + 1), at(curR, cells + d0 + i0 + 1)
                    A_lo = mL + fr * (mR - mL)
                    B_hi = nL + fr * (nR - nL)
                    y_lo = np.minimum(sL, c + A_lo)
                    y_hi = np.minimum(sR, c + B_hi)
                    y_t = c + np.minimum(mR, nL)
                    oL, oR = hull_line(y_lo, y_t, y_hi, 1.0 - fr)
                for i in range(self.nc):
                    sl = self.e[i] * akk * w
                    vL = self.e[i] * akk * ((o[k] + cells) * w) - self.rhs[k, i]
                    act = np.minimum(vL, vL + sl) >= 0.0
                    oL = oL + np.where(act, self.q[k, i] * vL, 0.0)
                    oR = oR + np.where(act, self.q[k, i] * (vL + sl), 0.0)
                slop = 4e-15 * (np.abs(oL) + np.abs(oR))
                oL, oR = oL - slop, oR - slop
            self.YL[k], self.YR[k] = oL, oR
            # bands outside the window: interval of local cells [L, R] (R / L infinite for the last band)
            cmin = np.minimum(curL, curR)
            big = 1 << 40

            def range_min(j0, j1):     # min of the next stage over local cells [j0, j1], bands included
                v = np.inf
                if j0 < 0:
                    m0, m1 = min((-min(j1, -1) - 1) // Wb, KB), min((-j0 - 1) // Wb, KB)
                    v = min(v, bn[0][m0:m1 + 1].min())
                if j1 >= G:
                    m0, m1 = min((max(j0, G) - G) // Wb, KB), min((j1 - G) // Wb, KB)
                    v = min(v, bn[1][m0:m1 + 1].min())
                a, b = max(j0, 0), min(j1, G - 1)
                if a <= b:
                    v = min(v, cmin[a:b + 1].min())
                return v
            for side in (0, 1):
                for m in range(KB + 1):
                    if side == 0:
                        L, R = (-(m + 1) * Wb if m < KB else -big), -m * Wb - 1
                    else:
                        L, R = G + m * Wb, (G + (m + 1) * Wb - 1 if m < KB else big)
                    vals = []
                    for ca, sh, frac in ((0.0, 0, False), (c, i0, True)):
                        j0 = L + d0 + sh - (1 if frac else 0)
                        j1 = R + d0 + sh + (2 if frac else 0)
                        vals.append(ca + range_min(max(j0, -big), min(j1, big)))
                    # penalty at the favourable edge of the band
                    pen = 0.0
                    for i in range(self.nc):
                        if self.e[i] < 0:
                            pos = (o[k] + R + 1 + EDGE_EPS) * w if R < big else np.inf
                        elif self.e[i] > 0:
                            pos = (o[k] + L - EDGE_EPS) * w if L > -big else -np.inf
                        else:
                            pos = 0.0
                        if np.isfinite(pos):
                            pen += self.q[k, i] * max(0.0, self.e[i] * akk * pos - self.rhs[k, i])
                    self.band[k, side, m] = min(vals) + pen
            curL, curR = self.YL[k], self.YR[k]

    def bound(self, k, s):
        if k >= self.Nt:
            return 0.0
        x = s / self.w - self.o[k]
        fl = np.floor(x)
        if fl < 0:
            return float(self.band[k, 0, min(int(-fl - 1) // self.Wb, self.KB)])
        if fl >= self.G:
            return float(self.band[k, 1, min(int(fl - self.G) // self.Wb, self.KB)])
        j = int(fl)
        fr = x - fl
        L, R = self.YL[k], self.YR[k]
        if not self.lin:
            return float(L[j])
        v = L[j] + fr * (R[j] - L[j])
        if fr < EDGE_EPS:
            v = min(v, R[j - 1] if j > 0 else self.band[k, 0, 0])
        if fr > 1.0 - EDGE_EPS:
            v = min(v, L[j + 1] if j + 1 < self.G else self.band[k, 1, 0])
        return float(v)

    # -- exact search with a rising cost threshold (iterative deepening on the bound): depth-first as in StageDp.solve,
    #    but no node whose bound exceeds T is opened while no solution below T is known; T starts just above the root
    #    bound and doubles its distance to it whenever the tree below T is exhausted without a solution.  The final
    #    pass (the one that finds a solution under T and then proves it) opens only nodes with bound < optimum +
    #    (T - optimum), so a poor first dive can no longer trap the search in a bad subtree.
    def solve_ida(self, max_nodes=2000000, delta0=1e-3, grow=4.0):
        Nt = self.Nt
        best, best_u, nodes = np.inf, None, 0
        root_lb = min(self.stage_cost(0, 0.0, act) + self.bound(1, self.shift[0] * act) for act in (0.0, 1.0))
        delta = max(delta0 * max(1.0, abs(root_lb)), 1e-9)
        passes = 0
        while nodes < max_nodes:
            passes += 1
            T = root_lb + delta
            u = np.zeros(Nt)
            stack = [(0, 0.0, 0.0, -1.0, -np.inf)]
            cutoff_hit = False
            while stack and nodes < max_nodes:
                k, s, cost, uprev, bd = stack.pop()
                tol = 1e-11 * max(1.0, abs(best)) if np.isfinite(best) else 0.0
                lim = min(best - tol, T)
                if bd >= lim:
                    if bd < best - tol:
                        cutoff_hit = True
                    continue
                if k > 0:
                    u[k - 1] = uprev
                nodes += 1
                if k == Nt:
                    best, best_u = cost, u.copy()
                    continue
                p = self.ak[k] * s
                kids = []
                for act in (0.0, 1.0):
                    c2 = cost + self.stage_cost(k, p, act)
                    s2 = s + self.shift[k] * act
                    b2 = c2 + self.bound(k + 1, s2)
                    if b2 < min(best - tol, T):
                        kids.append((b2, k + 1, s2, c2, act))
                    elif b2 < best - tol:
                        cutoff_hit = True
                for b2, k2, s2, c2, act in sorted(kids, key=lambda t: -t[0]):
                    stack.append((k2, s2, c2, act, b2))
            if not cutoff_hit or (np.isfinite(best) and best <= T):
                break
            delta *= grow
        self.passes = passes
        return best, best_u, nodes

    # -- the team kernel's split of one hard agent over several CTAs (csrc/stage_dp.cu, dp_search with part / nparts):
    #    the depth-`depth` prefixes below the root (the 32 "children" one expansion produces) are ranked by their
    #    bounds and dealt to `nparts` parts, rank r to part r mod nparts; every part searches its own subtrees with the
    #    threshold passes of solve_ida -- steps of at least a quarter of the known gap once an incumbent exists --
    #    and prunes with the best plan ANY part has found.  On the GPU the parts run side by side and exchange the
    #    incumbent through one atomic key; here they run one after the other, which is one of the orders the GPU may
    #    produce.  Returns (best, plan, expansions).
    def solve_split(self, nparts=8, depth=5, max_nodes=2000000, delta0=1e-3, grow=4.0):
        Nt = self.Nt
        D = min(depth, Nt)
        kids = []
        for code in range(1 << D):
            s, cost, acts = 0.0, 0.0, []
            for t in range(D):
                act = float(code >> t & 1)
                cost += self.stage_cost(t, self.ak[t] * s, act)
                s += self.shift[t] * act
                acts.append(act)
            kids.append((cost + self.bound(D, s), code, s, cost, acts))
        order = sorted(range(len(kids)), key=lambda i: (kids[i][0], i))
        root_lb = min(kd[0] for kd in kids)
        best, best_u, nodes = np.inf, None, 0
        for part in range(nparts):
            mine = [kids[i] for r, i in enumerate(order) if r % nparts == part]
            delta = max(delta0 * max(1.0, abs(root_lb)), 1e-9)
            while nodes < max_nodes:
                if np.isfinite(best) and best > root_lb:
                    delta = max(delta, 0.25 * (best - root_lb))
                T = root_lb + delta
                u = np.zeros(Nt)
                cutoff_hit = False
                stack = [(D, s, cost, acts, bd) for bd, code, s, cost, acts in sorted(mine, key=lambda kd: -kd[0])]
                while stack and nodes < max_nodes:
                    k, s, cost, setter, bd = stack.pop()
                    tol = 1e-11 * max(1.0, abs(best)) if np.isfinite(best) else 0.0
                    if bd >= min(best - tol, T):
                        if bd < best - tol:
                            cutoff_hit = True
                        continue
                    if isinstance(setter, list):
                        u[:k] = setter
                    else:
                        u[k - 1] = setter
                    nodes += 1
                    if k == Nt:
                        best, best_u = cost, u.copy()
                        continue
                    p = self.ak[k] * s
                    nxt = []
                    for act in (0.0, 1.0):
                        c2 = cost + self.stage_cost(k, p, act)
                        s2 = s + self.shift[k] * act
                        b2 = c2 + self.bound(k + 1, s2)
                        if b2 < min(best - tol, T):
                            nxt.append((b2, k + 1, s2, c2, act))
                        elif b2 < best - tol:
                            cutoff_hit = True
                    for b2, k2, s2, c2, act in sorted(nxt, key=lambda t_: -t_[0]):
                        stack.append((k2, s2, c2, act, b2))
                if not cutoff_hit or (np.isfinite(best) and best <= T):
                    break
                delta *= grow
        return best, best_u, nodes
