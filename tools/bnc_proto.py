"""DEV TOOL (not imported by the product, not an oracle): numpy twin of the GPU branch-and-cut kernel
(pyhybridcontrol_b200/csrc/milp_bnc.cu), written with the same control flow so that pivot / node / cut
counts and numerics can be studied on the CPU before spending GPU time.

Algorithm (one problem):  min c'x  s.t.  H x <= rhs,  lo <= x <= hi,  x[bin] in {0,1}

* bounded DUAL simplex on a dense tableau  T = B^-1 N  that only holds the ACTIVE rows (row generation):
  the tableau starts empty, violated rows of H (and of the cut pool) are brought in on demand, expressed
  in the current basis.  All nodes of the search share the one tableau -- a node is only a set of bounds;
  because every binary is boxed the current basis stays dual feasible for any node.
* complemented mixed-integer-rounding (c-MIR) cuts separated from single rows of H (globally valid, so
  they can be separated at any node and kept for all others).
* depth-first branch and bound on the most fractional binary, nearest-integer child first, cutoff by the
  incumbent inside the dual simplex (the dual objective only increases).
"""
import numpy as np

INF = np.inf


class BranchAndCut(object):
    def __init__(self, c, H, rhs, lb, ub, is_bin, big=1e7, ptol=1e-9, itol=1e-6, rmax=96, max_cuts=256,
                 cut_depth=1 << 30, cut_rounds_root=30, cut_rounds_node=2, cuts_per_round=8, need_frac=True,
                 kset=(1, 2, 4, 8)):
        self.m0, self.n = H.shape
        n = self.n
        self.c = np.asarray(c, float)
        self.H = np.asarray(H, float)
        self.rhs = np.asarray(rhs, float)
        self.glo = np.where(np.isfinite(lb), lb, -big)
        self.ghi = np.where(np.isfinite(ub), ub, np.where(self.c < 0, big, INF))
        self.is_bin = np.asarray(is_bin, bool)
        self.bin_idx = np.nonzero(self.is_bin)[0]
        self.poolH = [self.H[i] for i in range(self.m0)]     # pool rows: originals then cuts
        self.poolr = [self.rhs[i] for i in range(self.m0)]
        self.active_of_pool = {}                              # pool id -> tableau row (via basis / nb lookups)
        self.rmax, self.max_cuts = rmax, max_cuts
        self.T = np.zeros((0, n))
        self.bbar = np.zeros(0)
        self.rowpool = []                                     # pool id of each tableau row
        self.basis = np.zeros(0, int)                         # var id basic in row r (>= n: slack of pool row id-n)
        self.nb = np.arange(n)                                # var id in column j
        self.d = self.c.copy()
        self.xN = np.zeros(n)
        self.xB = np.zeros(0)
        self.ptol, self.itol = ptol, itol
        self.stats = dict(pivots=0, nodes=0, cuts=0, rows_added=0, purges=0, lp_solves=0, max_rows=0)
        self.cut_depth, self.cut_rounds_root, self.cut_rounds_node = cut_depth, cut_rounds_root, cut_rounds_node
        self.cuts_per_round = cuts_per_round
        self.need_frac, self.kset = need_frac, kset
        self.stats['sep_evals'] = 0
        self.stats['sep_calls'] = 0

    # ------------------------------------------------------------------ bounds helpers
    def _bounds_of(self, var_ids, lo, hi):
        s = var_ids >= self.n
        l = np.where(s, 0.0, lo[np.minimum(var_ids, self.n - 1)])
        h = np.where(s, INF, hi[np.minimum(var_ids, self.n - 1)])
        return l, h

    def node_setup(self, lo, hi):
        self.lo, self.hi = lo, hi
        l, h = self._bounds_of(self.nb, lo, hi)
        x = np.where(self.d >= 0, l, h)
        x = np.where(np.isfinite(x), x, np.where(np.isfinite(l), l, h))
        self.xN = x
        self.xB = self.bbar - self.T @ self.xN

    def x_struct(self):
        x = np.zeros(self.n)
        sn = self.nb < self.n
        x[self.nb[sn]] = self.xN[sn]
        sb = self.basis < self.n
        x[self.basis[sb]] = self.xB[sb]
        return x

    def objective(self):
        return float(self.c @ self.x_struct())

    # ------------------------------------------------------------------ dual simplex
    def dual_simplex(self, cutoff=INF, maxit=20000):
        T = self.T
        for _ in range(maxit):
            if T.shape[0] == 0:
                return "opt"
            lo_b, hi_b = self._bounds_of(self.basis, self.lo, self.hi)
            viol_lo, viol_hi = lo_b - self.xB, self.xB - hi_b
            viol = np.maximum(viol_lo, viol_hi)
            r = int(np.argmax(viol))
            if viol[r] <= self.ptol:
                return "opt"
            below = viol_lo[r] > viol_hi[r]
            row = T[r]
            l, h = self._bounds_of(self.nb, self.lo, self.hi)
            movable = h > l
            at_lower = self.xN <= l
            at_upper = self.xN >= h
            if below:
                cand = movable & (((row < -1e-9) & at_lower) | ((row > 1e-9) & at_upper))
                target = lo_b[r]
            else:
                cand = movable & (((row > 1e-9) & at_lower) | ((row < -1e-9) & at_upper))
                target = hi_b[r]
            if not cand.any():
                return "inf"
            ratios = np.where(cand, np.abs(self.d) / np.maximum(np.abs(row), 1e-300), INF)
            rmin = ratios.min()
            tie = cand & (ratios <= rmin + 1e-12)
            q = int(np.argmax(np.where(tie, np.abs(row), -1.0)))
            piv = row[q]
            t = (self.xB[r] - target) / piv
            self.xB -= T[:, q] * t
            xq_new = self.xN[q] + t
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
            self.stats["pivots"] += 1
            if cutoff < INF and self.objective() >= cutoff:
                return "cut"
        return "lim"

    # ------------------------------------------------------------------ row generation
    def _row_in_basis(self, g, g0, pool_id):
        """pool row g'x <= g0 in structural space -> tableau row over the current nonbasic columns.
        x_B = bbar - T x_N ; slack s = g0 - g'x."""
        n = self.n
        gB = np.where(self.basis < n, g[np.minimum(self.basis, n - 1)], 0.0)
        gN = np.where(self.nb < n, g[np.minimum(self.nb, n - 1)], 0.0)
        row = gN - gB @ self.T
        rb = g0 - gB @ self.bbar
        return row, rb

    def add_pool_row(self, pid):
        if self.T.shape[0] >= self.rmax:
            self.purge()
            if self.T.shape[0] >= self.rmax:
                return False
        row, rb = self._row_in_basis(self.poolH[pid], self.poolr[pid], pid)
        self.T = np.vstack([self.T, row[None, :]])
        self.bbar = np.concatenate([self.bbar, [rb]])
        self.basis = np.concatenate([self.basis, [self.n + pid]])
        self.rowpool.append(pid)
        self.xB = np.concatenate([self.xB, [rb - row @ self.xN]])
        self.stats["rows_added"] += 1
        self.stats["max_rows"] = max(self.stats["max_rows"], self.T.shape[0])
        return True

    def purge(self):
        """drop tableau rows whose own slack is basic in that row and strictly positive (inactive rows)"""
        keep = []
        for r in range(self.T.shape[0]):
            if self.basis[r] >= self.n and self.xB[r] > 1e-7:
                # slack basic: is it the slack of a pool row?  then the row is redundant right now
                continue
            keep.append(r)
        if len(keep) == self.T.shape[0]:
            return
        self.stats["purges"] += 1
        keep = np.array(keep, int)
        self.T = self.T[keep]
        self.bbar = self.bbar[keep]
        self.basis = self.basis[keep]
        self.xB = self.xB[keep]
        self.rowpool = [self.rowpool[r] for r in keep]
        # NOTE: a purged slack that is basic simply disappears; nonbasic slacks of purged rows cannot occur

    def active_pool_ids(self):
        ids = set(int(v) - self.n for v in self.basis if v >= self.n)
        ids |= set(int(v) - self.n for v in self.nb if v >= self.n)
        return ids

    def scan_rows(self, max_add=16):
        """bring in violated pool rows (originals + cuts).  -> number added"""
        x = self.x_struct()
        act = self.active_pool_ids()
        viol = []
        for pid in range(len(self.poolH)):
            if pid in act:
                continue
            v = self.poolH[pid] @ x - self.poolr[pid]
            if v > 1e-7:
                viol.append((v / max(1.0, np.abs(self.poolH[pid]).max()), pid))
        viol.sort(reverse=True)
        k = 0
        for _, pid in viol[:max_add]:
            if self.add_pool_row(pid):
                k += 1
        return k

    def solve_lp(self, cutoff=INF):
        """dual simplex + row generation until no pool row is violated"""
        self.stats["lp_solves"] += 1
        while True:
            st = self.dual_simplex(cutoff)
            if st != "opt":
                return st
            if self.scan_rows() == 0:
                return "opt"

    # ------------------------------------------------------------------ c-MIR separation on original rows
    def separate(self, x):
        n = self.n
        self.stats['sep_calls'] += 1
        isb = self.is_bin
        found = []
        for i in range(self.m0):
            h = self.H[i]
            r = self.rhs[i]
            nzb = isb & (h != 0)
            if not nzb.any():
                continue
            ib = np.nonzero(nzb)[0]
            ic = np.nonzero(~isb & (h != 0))[0]
            # continuous columns: shift to their global lower bound; positive coefficients are relaxed away
            if any(h[j] > 0 and not np.isfinite(self.glo[j]) for j in ic):
                continue
            if any(h[j] < 0 and not np.isfinite(self.glo[j]) for j in ic):
                continue
            rr = r - sum(h[j] * self.glo[j] for j in ic)
            comp = x[ib] > 0.5
            a = np.where(comp, -h[ib], h[ib])
            bb = rr - h[ib][comp].sum()
            xs = np.where(comp, 1 - x[ib], x[ib])
            sneg = sum(-h[j] * (x[j] - self.glo[j]) for j in ic if h[j] < 0)
            cand = [abs(a[j]) for j in range(ib.size) if 1e-6 < xs[j] < 1 - 1e-6 and abs(a[j]) > 1e-9]
            if self.need_frac and not cand:
                continue
            cand.append(np.abs(a).max())
            best = None
            for d0 in cand:
                for k in self.kset:
                    self.stats['sep_evals'] += 1
                    dl = d0 / k
                    at = a / dl
                    bt = bb / dl
                    f0 = bt - np.floor(bt)
                    if f0 < 0.05 or f0 > 0.95:
                        continue
                    fj = at - np.floor(at)
                    Fa = np.floor(at) + np.maximum(0, fj - f0) / (1 - f0)
                    lhs = Fa @ xs - sneg / (dl * (1 - f0))
                    viol = lhs - np.floor(bt)
                    norm = np.sqrt((Fa ** 2).sum() + 1e-12)
                    if viol > 1e-6 and (best is None or viol / norm > best[0]):
                        best = (viol / norm, dl, f0, Fa.copy(), np.floor(bt))
            if best is None:
                continue
            _, dl, f0, Fa, fb = best
            g = np.zeros(n)
            g0 = fb
            for jj, j in enumerate(ib):
                if comp[jj]:
                    g[j] -= Fa[jj]
                    g0 -= Fa[jj]
                else:
                    g[j] += Fa[jj]
            for j in ic:
                if h[j] < 0:
                    g[j] += h[j] / (dl * (1 - f0))
                    g0 += h[j] * self.glo[j] / (dl * (1 - f0))
            found.append((best[0], g, g0))
        found.sort(key=lambda t: -t[0])
        return found

    def cut_loop(self, rounds, cutoff):
        for _ in range(rounds):
            x = self.x_struct()
            xb = x[self.bin_idx]
            if np.abs(xb - np.round(xb)).max(initial=0) <= self.itol:
                return "opt"
            if len(self.poolH) - self.m0 >= self.max_cuts:
                return "opt"
            cuts = self.separate(x)
            if not cuts:
                return "opt"
            added = 0
            for _, g, g0 in cuts[:self.cuts_per_round]:
                self.poolH.append(g)
                self.poolr.append(g0)
                self.stats["cuts"] += 1
                if self.add_pool_row(len(self.poolH) - 1):
                    added += 1
            if not added:
                return "opt"
            st = self.solve_lp(cutoff)
            if st != "opt":
                return st
        return "opt"

    # ------------------------------------------------------------------ search
    def run(self, max_nodes=200000):
        best, best_x = INF, None
        stack = [(self.glo.copy(), self.ghi.copy(), -INF, 0)]
        while stack and self.stats["nodes"] < max_nodes:
            lo, hi, bound, depth = stack.pop()
            tolb = 1e-9 * max(1.0, abs(best)) if np.isfinite(best) else 0.0
            if bound >= best - tolb:
                continue
            self.stats["nodes"] += 1
            self.node_setup(lo, hi)
            st = self.solve_lp(cutoff=best - tolb)
            if st == "opt" and depth <= self.cut_depth:
                st = self.cut_loop(self.cut_rounds_root if depth == 0 else self.cut_rounds_node, best - tolb)
            if st != "opt":
                continue
            obj = self.objective()
            if obj >= best - tolb:
                continue
            x = self.x_struct()
            xb = x[self.bin_idx]
            frac = np.abs(xb - np.round(xb))
            if frac.max(initial=0) <= self.itol:
                # polish: fix binaries to the rounded values and re-solve for the continuous part
                l2, h2 = lo.copy(), hi.copy()
                l2[self.bin_idx] = h2[self.bin_idx] = np.round(xb)
                self.node_setup(l2, h2)
                if self.solve_lp() == "opt":
                    o2 = self.objective()
                    if o2 < best:
                        best, best_x = o2, self.x_struct()
                continue
            j = self.bin_idx[int(np.argmax(frac))]
            first = 1.0 if x[j] >= 0.5 else 0.0
            for val in (1.0 - first, first):
                l2, h2 = lo.copy(), hi.copy()
                l2[j] = h2[j] = val
                stack.append((l2, h2, obj, depth + 1))
        return best, best_x


if __name__ == "__main__":
    import sys
    import time
    sys.path.insert(0, ".")
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    kw = eval("dict(%s)" % sys.argv[3]) if len(sys.argv) > 3 else {}
    wl = syn.dewh_batch(B, N_p, seed=1)
    Nt = wl["Nt"]
    tot = {}
    ok = 0
    worst = 0
    for b in range(B):
        mats = {k: v[b] for k, v in wl["mats"].items()}
        full, d, vt = omld.complete(mats, nu_l=1)
        evo = oc.condense(full, d, Nt)
        prob = oa.build_problem(evo, d, vt, Nt, wl["x0"][b], wl["omega"][b],
                                atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
        s = BranchAndCut(prob.c, prob.H, prob.rhs, prob.lb, prob.ub, prob.is_bin, **kw)
        t0 = time.perf_counter()
        obj, x = s.run()
        t1 = time.perf_counter()
        st, oref, vref = osv.solve_milp(prob)
        good = abs(obj + prob.c0 - oref) <= 1e-6 * max(1, abs(oref))
        same = x is not None and np.array_equal(np.round(x[prob.is_bin]), np.round(vref[prob.is_bin]))
        ok += good
        worst = max(worst, s.stats["pivots"])
        for k, v in s.stats.items():
            tot[k] = tot.get(k, 0) + v
        if not good or not same or B <= 16:
            print(b, "obj %.9f ref %.9f %s %s" % (obj + prob.c0, oref, "OK" if good else "MISMATCH",
                                                     "same-u" if same else "DIFF-u"), s.stats)
    print("N_p=%d B=%d ok=%d  avg:" % (N_p, B, ok), {k: round(v / B, 1) for k, v in tot.items()}, "worst pivots", worst)
