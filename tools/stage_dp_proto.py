"""DEV TOOL (not imported by the product, not an oracle): numpy twin of the stage-DP kernels
(pyhybridcontrol_b200/csrc/stage_dp.cu) with the same grid, the same widened cells and the same FP32 round-down of
the stored table, so that the two properties the CUDA path relies on can be checked on the CPU
(tests/test_stage_dp_proto.py):

  1. VALIDITY  -- the table is a lower bound of the true cost-to-go at every state of every cell;
  2. EXACTNESS -- depth-first search over the binary sequence, pruned with that bound, returns the optimum
                  (compared with HiGHS / enumeration by the test).

Problem (one agent, one binary input, rows with their own slack -- the DEWH shape):
    min sum_k c_k u_k + sum_i q_k,i max(0, e_i p_k - rhs_k,i),   p_0 = 0,  p_k+1 = a p_k + b u_k
"""
import numpy as np

EDGE_EPS = 1e-9


def _round_down_f32(x):
    """largest float32 <= x (what __double2float_rd does)"""
    f = np.asarray(x, dtype=np.float64).astype(np.float32)
    up = f.astype(np.float64) > x
    f[up] = np.nextafter(f[up], np.float32(-np.inf))
    return f


class StageDp(object):
    def __init__(self, a, b, e, rhs, c, q, cells=2048):
        self.a, self.b = float(a), float(b)
        self.e = np.asarray(e, dtype=float)
        self.rhs = np.asarray(rhs, dtype=float)          # [Nt, nc]
        self.c = np.asarray(c, dtype=float)              # [Nt]
        self.q = np.asarray(q, dtype=float)              # [Nt, nc]
        self.Nt, self.nc = self.rhs.shape
        self.G = int(cells)
        Nt = self.Nt
        self.ak = self.a ** np.arange(Nt + 2)
        self.shift = self.b / self.ak[1:Nt + 1]          # translation of s = p / a^k when the input is on
        self._window()
        self._table()

    # -- grid window: hull of the violation-free band (clamped to what is reachable), one shift of margin,
    #    at least four shifts wide (stage_dp.cu: dp_load)
    def _window(self):
        rlo = rhi = 0.0
        blo, bhi = np.inf, -np.inf
        for k in range(self.Nt):
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
            blo, bhi = min(blo, lo_c, hi_c), max(bhi, lo_c, hi_c)
            rlo += min(0.0, self.shift[k])
            rhi += max(0.0, self.shift[k])
        margin = np.abs(self.shift).max()
        S0, S1 = max(rlo, blo - margin), min(rhi, bhi + margin)
        if not S1 > S0:
            S0, S1 = rlo, rhi
        minw = 4.0 * margin * (1.0 + 16.0 / self.G)
        if S1 - S0 < minw:
            mid = 0.5 * (S0 + S1)
            S0, S1 = mid - 0.5 * minw, mid + 0.5 * minw
        self.S0, self.w = S0, (S1 - S0) / self.G
        self.tailmin = np.concatenate([np.cumsum(np.minimum(self.c, 0.0)[::-1])[::-1], [0.0]])

    def stage_cost(self, k, p, u):
        return self.c[k] * u + float(np.sum(self.q[k] * np.maximum(0.0, self.e * p - self.rhs[k])))

    # -- backward sweep (stage_dp_table_kernel): LB[k][cell] valid for every state of the (slightly widened) cell
    def _table(self):
        Nt, G, w, S0 = self.Nt, self.G, self.w, self.S0
        cells = np.arange(G)
        self.LB = np.zeros((Nt + 1, G), dtype=np.float32)
        cur = np.zeros(G)
        for k in range(Nt - 1, 0, -1):
            lo_edge = self.ak[k] * (S0 + (cells - EDGE_EPS) * w)
            hi_edge = self.ak[k] * (S0 + (cells + 1.0 + EDGE_EPS) * w)
            pen = np.zeros(G)
            for i in range(self.nc):
                edge = lo_edge if self.e[i] >= 0 else hi_edge
                pen += self.q[k, i] * np.maximum(0.0, self.e[i] * edge - self.rhs[k, i])
            out = self.tailmin[k + 1]
            stay = cur                                             # no input: the cell maps onto itself exactly
            r = self.shift[k] / w
            i0 = int(np.floor(r))
            fr = r - i0

            def at(idx):
                v = np.full(G, out)
                ok = (idx >= 0) & (idx < G)
                v[ok] = cur[idx[ok]]
                return v
            move = np.minimum(at(cells + i0), at(cells + i0 + 1))
            if fr < EDGE_EPS:
                move = np.minimum(move, at(cells + i0 - 1))
            if fr > 1.0 - EDGE_EPS:
                move = np.minimum(move, at(cells + i0 + 2))
            best = pen + np.minimum(stay, self.c[k] + move)
            self.LB[k] = _round_down_f32(best)
            cur = self.LB[k].astype(np.float64)

    def bound(self, k, s):
        """lower bound of the cost-to-go from exact state s at stage k (stage_dp_search_kernel's table read)"""
        if k >= self.Nt:
            return 0.0
        cell = np.floor((s - self.S0) / self.w)
        if cell < 0 or cell >= self.G:
            return float(self.tailmin[k])
        return float(self.LB[k][int(cell)])

    # -- exact search (depth-first, best-bound child first); returns (objective, u, nodes)
    def solve(self, max_nodes=2000000):
        Nt = self.Nt
        best, best_u, nodes = np.inf, None, 0
        u = np.zeros(Nt)
        stack = [(0, 0.0, 0.0, -1.0, -np.inf)]
        while stack and nodes < max_nodes:
            k, s, cost, uprev, bd = stack.pop()
            tol = 1e-11 * max(1.0, abs(best)) if np.isfinite(best) else 0.0
            if bd >= best - tol:
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
                if b2 < best - tol:
                    kids.append((b2, k + 1, s2, c2, act))
            for b2, k2, s2, c2, act in sorted(kids, key=lambda t: -t[0]):
                stack.append((k2, s2, c2, act, b2))
        return best, best_u, nodes

    def cost_to_go_exact(self, k, s):
        """true V_k(s) by exhaustive enumeration (small Nt only) -- the yardstick of the validity test"""
        Nt = self.Nt
        if k >= Nt:
            return 0.0
        p = self.ak[k] * s
        return min(self.stage_cost(k, p, act) + self.cost_to_go_exact(k + 1, s + self.shift[k] * act) for act in (0.0, 1.0))


def from_dewh_problem(mats, prob, Nt):
    """(a, b, e, rhs, c, q) of a DEWH agent from its MLD blocks and the oracle's assembled problem"""
    a, b = float(mats["A"][0, 0]), float(mats["B1"][0, 0])
    e = np.asarray(mats["E"], dtype=float)[:, 0]
    cost = prob.c[:3 * Nt].reshape(Nt, 3)
    return a, b, e, prob.rhs.reshape(Nt, 2), cost[:, 0], cost[:, 1:]
