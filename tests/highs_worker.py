"""Worker of the GPU config-size tests: one HiGHS solve of one agent's oracle problem, run in SPAWNED processes (a
fork of a process that has CUDA and thread pools initialised can deadlock in the child).  Imports numpy / scipy and
the oracle only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def highs_job(job):
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    mats, Nt, x0, omega, q_u, q_mu, scen, extra = job
    full, d, vt = omld.complete(mats, nu_l=1)
    prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, x0, omega, atoms=dict(q_u=q_u, q_mu=q_mu),
                            omega_scenarios=scen, extra_constraints=extra)
    st, obj, v = osv.solve_milp(prob, polish=True, time_limit=30.0)
    u = None if v is None else np.round(np.asarray(v)[prob.is_bin])
    return st, obj, u
