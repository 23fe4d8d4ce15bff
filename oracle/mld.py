"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

MLD dimension rules and default blocks, restated from the reference's ``MldInfo``/``MldModel``:

  x(k+1) = A x + B1 u + B2 delta + B3 z + B4 omega + b5
  y(k)   = C x + D1 u + D2 delta + D3 z + D4 omega + d5
  E x + F1 u + F2 delta + F3 z + F4 omega + G y + Psi mu <= f5 ,  mu >= 0

* dimension map              : models/mld_model.py:149-168
* C defaults to eye(A.shape) : models/mld_model.py:515-520
* missing blocks -> zeros of the derived shape (b5/d5 get one column, f5 must be given when nc > 0)
                             : models/mld_model.py:910-950
* variable types: continuous entries first, the LAST n*_l entries binary; delta is all binary, z all
  continuous, v = [u; delta; z; mu]                              : models/mld_model.py:294-345
"""
import numpy as np

STATE_INPUT = ("A", "B1", "B2", "B3", "B4", "b5")
OUTPUT = ("C", "D1", "D2", "D3", "D4", "d5")
CONSTRAINT = ("E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")
ALL_NAMES = STATE_INPUT + OUTPUT + CONSTRAINT
DIM_NAMES = ("nx", "nu", "ndelta", "nz", "nmu", "nomega", "ny", "nc")


def _as2d(a):
    a = np.asarray(a, dtype=float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a[:, None]
    return a


def complete(mats, nu_l=0, nmu_l=0):
    """-> (full dict of the 20 named matrices as float64 2-D arrays, dims dict, var_type_v list)."""
    given = {k: _as2d(v) for k, v in mats.items() if v is not None}
    for k in given:
        if k not in ALL_NAMES:
            raise ValueError("Invalid matrix name: %s" % k)
    shp = {k: (given[k].shape if k in given and 0 not in given[k].shape else (0, 0)) for k in ALL_NAMES}
    if "C" not in given:  # mld_model.py:515-520
        n = shp["A"][0]
        given["C"] = np.eye(n)
        shp["C"] = (n, n) if n else (0, 0)

    def rows(names):
        return max(shp[n][0] for n in names)

    def cols(names):
        return max(shp[n][1] for n in names)

    d = dict(nx=rows(STATE_INPUT), ny=rows(OUTPUT), nc=rows(CONSTRAINT),
             nu=cols(("B1", "D1", "F1")), ndelta=cols(("B2", "D2", "F2")), nz=cols(("B3", "D3", "F3")),
             nomega=cols(("B4", "D4", "F4")), nmu=shp["Psi"][1])
    d["nv"] = d["nu"] + d["ndelta"] + d["nz"] + d["nmu"]
    coldim = dict(A="nx", B1="nu", B2="ndelta", B3="nz", B4="nomega", b5=None,
                  C="nx", D1="nu", D2="ndelta", D3="nz", D4="nomega", d5=None,
                  E="nx", F1="nu", F2="ndelta", F3="nz", F4="nomega", f5=None, G="ny", Psi="nmu")
    rowdim = {**{k: "nx" for k in STATE_INPUT}, **{k: "ny" for k in OUTPUT}, **{k: "nc" for k in CONSTRAINT}}
    full = {}
    for k in ALL_NAMES:
        r = d[rowdim[k]]
        c = 1 if coldim[k] is None else d[coldim[k]]
        if shp[k] == (0, 0):
            if k == "f5" and d["nc"]:
                raise ValueError("Constraint vector 'f5' can only be null if all constraint matrices are null.")
            full[k] = np.zeros((r, c))
        else:
            if given[k].shape != (r, c):
                raise ValueError("Invalid shape for %s: %s, required %s" % (k, given[k].shape, (r, c)))
            full[k] = np.array(given[k], dtype=float)
    d.update(nu_l=int(nu_l), ndelta_l=d["ndelta"], nz_l=0, nmu_l=int(nmu_l))
    vt = (["c"] * (d["nu"] - d["nu_l"]) + ["b"] * d["nu_l"] + ["b"] * d["ndelta"] + ["c"] * d["nz"]
          + ["c"] * (d["nmu"] - d["nmu_l"]) + ["b"] * d["nmu_l"])
    return full, d, vt
