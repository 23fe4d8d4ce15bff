"""DEV TOOL (gpurun): randomised cross-check of the two solve kernels -- stage-DP against branch-and-cut -- on many
random MLDs of the scalar-state class (shapes, signs, hard rows, fixed binaries, negative costs)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_stage_dp import random_scalar_mld
from pyhybridcontrol_b200 import cabi
from pyhybridcontrol_b200.batch import BatchMpc

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
total = bad = infeas = undecided = 0
for seed, (nu, ndelta, nc, ny, soft, hard, Nt) in enumerate([(1, 0, 2, 1, True, 0, 14), (1, 0, 2, 1, True, 1, 12), (2, 0, 3, 2, True, 1, 8),
                                                          (1, 1, 2, 1, True, 0, 10), (1, 0, 4, 2, False, 0, 12), (2, 1, 4, 1, True, 2, 6),
                                                          (1, 0, 3, 1, True, 0, 20)]):
    rng = np.random.default_rng(1000 + seed)
    m = random_scalar_mld(rng, B, nu, ndelta, nc, ny, soft, hard)
    if not soft:
        m["f5"] = m["f5"] * 3.0
    nb, nmu = nu + ndelta, (nc if soft else 0)
    x0 = rng.uniform(-1.5, 1.5, size=(B, 1)); om = rng.uniform(-1.0, 1.0, size=(B, Nt))
    cost = np.zeros((B, Nt, nb + nmu))
    cost[:, :, :nb] = rng.uniform(-0.4, 1.0, size=(B, Nt, nb))
    cost[:, :, nb:] = rng.uniform(1.0, 30.0, size=(B, 1, nmu))
    lb = np.tile(np.r_[np.zeros(nb), np.zeros(nmu)], Nt); ub = np.tile(np.r_[np.ones(nb), np.full(nmu, np.inf)], Nt)
    pins = rng.integers(0, Nt * (nb + nmu), size=3)
    for pidx in pins:                                  # pin a few binaries
        if pidx % (nb + nmu) < nb:
            val = float(rng.integers(0, 2)); lb[pidx] = ub[pidx] = val
    res = {}
    for solver in ("stage_dp", "bnc"):
        bm = BatchMpc(m, Nt - 1, Nt, nu_l=nu, device=dev, solver=solver, opts=cabi.default_opts(max_nodes=20000, max_pivots=400000))
        bm.lb_v, bm.ub_v = lb.copy(), ub.copy()
        bm.build()
        r = bm.solve(x0, om, cost_v=cost.reshape(B, -1))
        res[solver] = (r["obj"].cpu().numpy(), r["status"].cpu().numpy(), r["v"].cpu().numpy())
    od, sd, vd = res["stage_dp"]; ob, sb, vb = res["bnc"]
    both = (sd == 0) & (sb == 0)
    rel = np.abs(od - ob) / np.maximum(1.0, np.abs(ob))
    mism = both & (rel > 1e-6)
    inf_mism = ((sd == 1) & (sb == 0)) | ((sd == 0) & (sb == 1))
    total += B; bad += int(mism.sum()) + int(inf_mism.sum()); infeas += int(((sd == 1) & (sb == 1)).sum()); undecided += int((sb >= 2).sum())
    print("shape nu=%d ndelta=%d nc=%d ny=%d soft=%s hard=%d Nt=%d: both optimal %d, both infeasible %d, bnc undecided %d, "
          "objective mismatches %d, feasibility mismatches %d, max rel diff %.2e, dp status %s" % (
              nu, ndelta, nc, ny, soft, hard, Nt, both.sum(), ((sd == 1) & (sb == 1)).sum(), (sb >= 2).sum(), mism.sum(), inf_mism.sum(),
              rel[both].max() if both.any() else 0.0, np.bincount(sd, minlength=6).tolist()))
    for b in np.nonzero(mism | inf_mism)[0][:3]:
        print("   agent", b, "dp", od[b], sd[b], "bnc", ob[b], sb[b])
print("TOTAL %d problems, %d mismatches, %d infeasible in both, %d undecided by bnc" % (total, bad, infeas, undecided))
