"""DEV TOOL (not imported by the product, not an oracle): numpy twin of the LINEAR-CELL value table of
pyhybridcontrol_b200/csrc/stage_dp.cu (round 2).

Every cell of a stage carries a line (value at its left edge YL, at its right edge YR) that lies below the true
cost-to-go on the whole cell; neighbouring cells need not agree at their common edge (the table is discontinuous, so
the staircase of the nominal regime stays as sharp as with constant cells, and a cell whose line has slope zero IS the
constant-cell bound).  One backward step for cell j:

    g(s) = min( stay line of cell j,  c_k + [line of cell m on [lo, t),  line of cell m+1 on [t, hi)] )   (move)
    line = tightest supporting line of the lower convex hull of g at the hull's minimum vertex
    out  = line + stage penalty where the whole cell violates the row (a cell that straddles the kink gets zero)

tests/test_stage_dp_proto.py checks validity (never above the exhaustive cost-to-go) and exactness of the pruned search.
"""
import numpy as np

from stage_dp_proto import StageDp, EDGE_EPS, from_dewh_problem  # noqa: F401


def hull_line(y_lo, y_t, y_hi, theta):
    """(left, right) values of a line below the three-point piecewise-linear function (lo, y_lo) - (t, y_t) - (hi, y_hi),
    t at fraction theta of the cell; all arguments arrays of one value per cell"""
    chord_t = y_lo + theta * (y_hi - y_lo)
    convex = y_t < chord_t
    itheta, i1theta = 1.0 / theta, 1.0 / (1.0 - theta)
    # kink below the chord: supporting line at the lowest vertex
    flat = y_t <= np.minimum(y_lo, y_hi)
    down = y_lo > y_hi                       # decreasing: extend the right segment backwards
    s2 = (y_hi - y_t) * i1theta
    s1 = (y_t - y_lo) * itheta
    oL = np.where(convex, np.where(flat, y_t, np.where(down, y_hi - s2, y_lo)), y_lo)
    oR = np.where(convex, np.where(flat, y_t, np.where(down, y_hi, y_lo + s1)), y_hi)
    return oL, oR


class StageDpLin(StageDp):
    def _table(self):
        Nt, G, w, S0 = self.Nt, self.G, self.w, self.S0
        cells = np.arange(G)
        self.YL = np.zeros((Nt + 1, G))
        self.YR = np.zeros((Nt + 1, G))
        curL, curR = np.zeros(G), np.zeros(G)
        for k in range(Nt - 1, 0, -1):
            out = self.tailmin[k + 1]

            def at(arr, idx):
                v = np.full(G, out)
                ok = (idx >= 0) & (idx < G)
                v[ok] = arr[idx[ok]]
                return v
            r = self.shift[k] / w
            i0 = int(np.floor(r))
            fr = r - i0
            c = self.c[k]
            if fr < EDGE_EPS or fr > 1.0 - EDGE_EPS:
                # translation lands on a cell boundary: constant bound over every cell that can be touched
                lo_i, hi_i = (i0 - 1, i0 + 1) if fr < EDGE_EPS else (i0, i0 + 2)
                mv = np.full(G, np.inf)
                for d in range(lo_i, hi_i + 1):
                    mv = np.minimum(mv, np.minimum(at(curL, cells + d), at(curR, cells + d)))
                mvL = mvR = c + mv
                y_lo = np.minimum(curL, mvL)
                y_hi = np.minimum(curR, mvR)
                # min of two lines is concave: the chord of the end-point minima is below it
                oL, oR = y_lo, y_hi
            else:
                mL, mR = at(curL, cells + i0), at(curR, cells + i0)
                nL, nR = at(curL, cells + i0 + 1), at(curR, cells + i0 + 1)
                theta = 1.0 - fr
                A_lo = mL + fr * (mR - mL)
                B_hi = nL + fr * (nR - nL)
                y_lo = np.minimum(curL, c + A_lo)
                y_hi = np.minimum(curR, c + B_hi)
                y_t = c + np.minimum(mR, nL)
                oL, oR = hull_line(y_lo, y_t, y_hi, theta)
            # stage penalty: exact line where the whole cell violates the row
            akk = self.ak[k]
            for i in range(self.nc):
                sl = self.e[i] * akk * w
                vL = self.e[i] * akk * (S0 + cells * w) - self.rhs[k, i]
                act = np.minimum(vL, vL + sl) >= 0.0        # straddling cell: the zero line (= the constant-cell bound)
                oL = oL + np.where(act, self.q[k, i] * vL, 0.0)
                oR = oR + np.where(act, self.q[k, i] * (vL + sl), 0.0)
            # floating-point slop of the construction (a few ulp of the values involved)
            slop = 1e-13 * (np.abs(oL) + np.abs(oR))
            self.YL[k], self.YR[k] = oL - slop, oR - slop
            curL, curR = self.YL[k], self.YR[k]

    def bound(self, k, s):
        if k >= self.Nt:
            return 0.0
        x = (s - self.S0) / self.w
        fl = np.floor(x)
        if fl < 0 or fl >= self.G:
            return float(self.tailmin[k])
        j = int(fl)
        fr = x - fl
        L, R = self.YL[k], self.YR[k]
        v = L[j] + fr * (R[j] - L[j])
        if fr < EDGE_EPS:
            v = min(v, R[j - 1] if j > 0 else v)          # s >= S0 is decided exactly (s - S0 is exact near 0)
        if fr > 1.0 - EDGE_EPS:
            v = min(v, L[j + 1] if j + 1 < self.G else v)
        return float(v)


if __name__ == '__main__':
    import sys
    sys.path.insert(0, '/root/repo')
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    S = 32
    wl = syn.dewh_batch(64, N_p, seed=5)
    Nt = wl['Nt']
    rng = np.random.default_rng(123)
    scen = wl["omega"][:, :, None] * rng.uniform(0.5, 1.8, size=(64, Nt, S))
    for b in [62, 45, 17, 13, 52, 63, 0, 1]:
        mats = {k: v[b] for k, v in wl["mats"].items()}
        full, d, vt = omld.complete(mats, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, wl["x0"][b], wl["omega"][b],
                                atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]), omega_scenarios=scen[b])
        args = from_dewh_problem(mats, prob, Nt)
        line = "agent %2d |" % b
        for cls, G in ((StageDp, 8192), (StageDpLin, 8192), (StageDpLin, 2048)):
            dp = cls(*args, cells=G)
            obj, u, nodes = dp.solve(max_nodes=300000)
            line += " %s G=%d: %.6f nodes %d |" % (cls.__name__[7:] or "const", G, obj, nodes)
        print(line, flush=True)
