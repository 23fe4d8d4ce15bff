"""Generates the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE (run in the build container,
where /root/reference is mounted):

    python tests/golden/make_golden.py

* condense_<case>.npz : inputs (MLD matrices, N_tilde) and the 12 condensed matrices produced by the
  reference's own ``MldEvoMatrices`` (controllers/components/mld_evolution_matrices.py) under oracle/ref_shim.py
* lsim_<case>.npz     : inputs and outputs of the reference's own ``MldModel.lsim_k`` (models/mld_model.py:647-699)
* milp_dewh.npz       : synthetic DEWH MPC problems solved by HiGHS 1.12.0 (scipy 1.18.1) through the oracle;
                        the reference cannot pin these (cvxpy/Gurobi are not installable here) -- "parity unpinned".
The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim, mld as omld, condense as oc, assemble as oa, solve as osv, lsim as ol  # noqa: E402
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn  # noqa: E402


def rand_mld(rng, nx, nu, nd, nz, nmu, nw, ny, nc, give_C=True):
    m = {}

    def r(a, b):
        return rng.standard_normal((a, b))
    if nx:
        m["A"] = r(nx, nx) * 0.5
        m["b5"] = r(nx, 1)
        for k, c in (("B1", nu), ("B2", nd), ("B3", nz), ("B4", nw)):
            if c:
                m[k] = r(nx, c)
    if ny:
        if nx and give_C:
            m["C"] = r(ny, nx)
        m["d5"] = r(ny, 1)
        for k, c in (("D1", nu), ("D2", nd), ("D3", nz), ("D4", nw)):
            if c:
                m[k] = r(ny, c)
    if nc:
        for k, c in (("E", nx), ("F1", nu), ("F2", nd), ("F3", nz), ("F4", nw), ("G", ny), ("Psi", nmu)):
            if c:
                m[k] = r(nc, c)
        m["f5"] = r(nc, 1)
    return m


CASES = {
    # name: (mats, Nt, nu_l)
    "appendixA": (dict(A=[[0.99]], B1=[[0.5]], B4=[[-2.0]], b5=[[0.1]], E=[[1], [-1]], F1=[[0], [0]],
                       Psi=[[-1, 0], [0, -1]], f5=[[65], [-50]]), 4, 1),
    "dewh_N24": (ol.dewh_mld(ol.DEWH_PARAMS), 25, 1),
    "dewh_N48": (ol.dewh_mld(ol.DEWH_PARAMS), 49, 1),
    "dewh_N96": (ol.dewh_mld(ol.DEWH_PARAMS), 97, 1),
    "grid_3dev": (ol.grid_mld(ol.GRID_PARAMS, 3), 5, 0),
}


def main():
    rng = np.random.default_rng(20260101)
    cases = dict(CASES)
    cases["rand_small"] = (rand_mld(rng, 3, 2, 1, 1, 2, 2, 2, 4), 7, 1)
    cases["rand_noC"] = (rand_mld(rng, 2, 1, 0, 0, 0, 1, 2, 3, give_C=False), 6, 0)
    cases["rand_mid"] = (rand_mld(rng, 4, 2, 2, 1, 3, 2, 3, 5), 25, 1)
    cases["rand_nomu"] = (rand_mld(rng, 2, 2, 1, 0, 0, 0, 1, 3), 9, 2)
    for name, (mats, Nt, nu_l) in cases.items():
        ref, dims, mld = ref_shim.reference_condense(mats, Nt - 1, Nt, bin_dims=dict(nu_l=nu_l) if nu_l else None)
        out = {"in_" + k: np.asarray(v, dtype=float) for k, v in mats.items()}
        out.update({"out_" + k: v for k, v in ref.items()})
        out["Nt"] = Nt
        out["nu_l"] = nu_l
        out["dims"] = np.array([dims[k] for k in ("nx", "nu", "ndelta", "nz", "nmu", "nomega", "ny", "n_constraints")])
        out["var_type_v"] = np.array([t == "b" for t in mld.mld_info["var_type_v"].ravel()])
        np.savez_compressed(os.path.join(HERE, "condense_%s.npz" % name), **out)
        # lsim_k of the same model on random inputs (all auxiliaries given -> no cvxpy needed)
        full, d, vt = omld.complete(mats, nu_l=nu_l)
        recs = []
        for t in range(4):
            x = rng.standard_normal((d["nx"], 1)) * 3 + 55 * (name.startswith("dewh"))
            u = (rng.random((d["nu"], 1)) > 0.5).astype(float)
            de = (rng.random((d["ndelta"], 1)) > 0.5).astype(float)
            z = rng.standard_normal((d["nz"], 1))
            mu = rng.random((d["nmu"], 1))
            w = rng.random((d["nomega"], 1)) * (0.01 if name.startswith("dewh") else 1.0)
            res = mld.lsim_k(x_k=x if d["nx"] else None, u_k=u if d["nu"] else None,
                             delta_k=de if d["ndelta"] else None, z_k=z if d["nz"] else None,
                             mu_k=mu if d["nmu"] else None, omega_k=w if d["nomega"] else None)
            recs.append(dict(x=x, u=u, delta=de, z=z, mu=mu, w=w, x1=np.asarray(res.x_k1, float),
                             y=np.asarray(res.y, float), cons=np.asarray(res.cons, bool)))
        lo = {"in_" + k: np.asarray(v, dtype=float) for k, v in mats.items()}
        for key in recs[0]:
            lo[key] = np.stack([r[key] for r in recs])
        lo["nu_l"] = nu_l
        np.savez_compressed(os.path.join(HERE, "lsim_%s.npz" % name), **lo)
        print("wrote", name, {k: v.shape for k, v in ref.items() if k in ("H_v", "Gamma_v")})

    # ---- MILP goldens (HiGHS through the oracle): 12 DEWH problems at N_p = 24 and 12 at N_p = 48
    for N_p, B in ((24, 12), (48, 12)):
        wl = syn.dewh_batch(B, N_p, seed=7)
        Nt = wl["Nt"]
        objs, vs = [], []
        for b in range(B):
            full, d, vt = omld.complete({k: v[b] for k, v in wl["mats"].items()}, nu_l=1)
            evo = oc.condense(full, d, Nt)
            prob = oa.build_problem(evo, d, vt, Nt, wl["x0"][b], wl["omega"][b],
                                    atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
            st, obj, v = osv.solve_milp(prob)
            assert st == osv.OPTIMAL
            objs.append(obj)
            vs.append(v)
        np.savez_compressed(os.path.join(HERE, "milp_dewh_N%d.npz" % N_p), seed=7, B=B, N_p=N_p,
                            obj=np.array(objs), v=np.array(vs))
        print("wrote milp_dewh_N%d" % N_p, np.round(objs, 4))


if __name__ == "__main__":
    main()
