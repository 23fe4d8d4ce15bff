"""CPU checks of the mathematics behind the stage-DP kernels, on their numpy twin (tools/stage_dp_proto.py: same
grid, widened cells, FP32 round-down):

* validity  -- the value table never exceeds the true cost-to-go (exhaustive enumeration, short horizons);
* exactness -- the bound-pruned depth-first search returns the HiGHS optimum and decisions (N_p = 24 and 48).
The CUDA kernels themselves are checked on the GPU (tests/test_gpu_stage_dp.py, tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from stage_dp_proto import StageDp, from_dewh_problem  # noqa: E402

from oracle import mld as omld, condense as oc, assemble as oa, solve as osv  # noqa: E402
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn  # noqa: E402


def _problem(wl, b):
    Nt = wl["Nt"]
    mats = {k: v[b] for k, v in wl["mats"].items()}
    full, d, vt = omld.complete(mats, nu_l=1)
    prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, wl["x0"][b], wl["omega"][b],
                            atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
    return mats, prob


@pytest.mark.parametrize("cells", [64, 512])
@pytest.mark.parametrize("x0_shift", [0.0, -12.0])     # nominal start / cold start (slack unavoidable)
def test_value_table_is_a_lower_bound(cells, x0_shift):
    N_p = 9
    wl = syn.dewh_batch(3, N_p, seed=17)
    wl["x0"] = wl["x0"] + x0_shift
    rng = np.random.default_rng(cells)
    for b in range(3):
        mats, prob = _problem(wl, b)
        dp = StageDp(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells)
        # states actually reachable at stage k (all 2^k prefixes) and random states inside the window
        for k in range(1, wl["Nt"]):
            reach = {0.0}
            for j in range(k):
                reach |= {s + dp.shift[j] for s in reach}
            states = list(reach)[:64] + list(dp.S0 + rng.uniform(0, 1, size=8) * dp.G * dp.w)
            for s in states:
                lb, true = dp.bound(k, s), dp.cost_to_go_exact(k, s)
                assert lb <= true + 1e-9 * max(1.0, abs(true)), (b, k, s, lb, true)


@pytest.mark.parametrize("N_p,cells", [(24, 256), (24, 4096), (48, 2048)])
def test_bound_pruned_search_is_exact(N_p, cells):
    wl = syn.dewh_batch(4, N_p, seed=23)
    for b in range(4):
        mats, prob = _problem(wl, b)
        dp = StageDp(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells)
        obj, u, nodes = dp.solve()
        st, oref, vref = osv.solve_milp(prob, polish=True)
        assert st == osv.OPTIMAL and abs(obj + prob.c0 - oref) <= 1e-6 * max(1.0, abs(oref)), (b, obj, oref, nodes)
        assert np.array_equal(u, np.round(vref[prob.is_bin]))
        assert nodes < 200000


# ---- the round-2 candidate: nodal piecewise-linear bound (tools/stage_dp_pwl_proto.py) --------------------------
from stage_dp_pwl_proto import StageDpPwl  # noqa: E402


@pytest.mark.parametrize("cells", [64, 512])
@pytest.mark.parametrize("x0_shift", [0.0, -12.0, 14.0])     # nominal / cold start / above the upper limit
def test_piecewise_linear_table_is_a_lower_bound(cells, x0_shift):
    """the chord-deficiency construction is valid where it matters most: with slack penalties active"""
    wl = syn.dewh_batch(3, 9, seed=17)
    wl["x0"] = wl["x0"] + x0_shift
    rng = np.random.default_rng(cells)
    for b in range(3):
        mats, prob = _problem(wl, b)
        dp = StageDpPwl(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells)
        for k in range(1, wl["Nt"]):
            reach = {0.0}
            for j in range(k):
                reach |= {s + dp.shift[j] for s in reach}
            states = list(reach)[:64] + list(dp.S0 + rng.uniform(0, 1, size=16) * dp.G * dp.w)
            for s in states:
                lb, true = dp.bound(k, s), dp.cost_to_go_exact(k, s)
                assert lb <= true + 1e-9 * max(1.0, abs(true)), (b, k, s, lb, true)


@pytest.mark.parametrize("N_p,cells", [(24, 4096), (48, 2048)])
def test_search_with_the_piecewise_linear_bound_is_exact(N_p, cells):
    wl = syn.dewh_batch(4, N_p, seed=23)
    for b in range(4):
        mats, prob = _problem(wl, b)
        dp = StageDpPwl(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells)
        obj, u, nodes = dp.solve()
        st, oref, vref = osv.solve_milp(prob, polish=True)
        assert st == osv.OPTIMAL and abs(obj + prob.c0 - oref) <= 1e-6 * max(1.0, abs(oref)), (b, obj, oref, nodes)
        assert np.array_equal(u, np.round(vref[prob.is_bin]))


# ---- round 2: moving window, semi-infinite cells, linear cells, threshold passes, split search (tools/stage_dp_mw_proto.py)
from stage_dp_mw_proto import StageDpMw  # noqa: E402


@pytest.mark.parametrize("lin", [False, True])
@pytest.mark.parametrize("cells,x0_shift", [(64, 0.0), (64, -12.0), (512, 14.0)])
def test_moving_window_tables_are_lower_bounds(lin, cells, x0_shift):
    """constant and linear cells over the moving window, and the cells beyond the window: never above the exact
    cost-to-go, for every state a plan can reach and for random states inside and outside the window"""
    N_p = 9
    wl = syn.dewh_batch(3, N_p, seed=17)
    wl["x0"] = wl["x0"] + x0_shift
    rng = np.random.default_rng(cells + int(lin))
    for b in range(3):
        mats, prob = _problem(wl, b)
        dp = StageDpMw(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells, lin=lin)
        for k in range(1, wl["Nt"]):
            reach = {0.0}
            for j in range(k):
                reach |= {s + dp.shift[j] for s in reach}
            lo = dp.o[k] * dp.w
            states = list(reach)[:64] + list(lo + rng.uniform(-0.5, 1.5, size=12) * dp.G * dp.w)
            for s in states:
                lb, true = dp.bound(k, s), dp.cost_to_go_exact(k, s)
                assert lb <= true + 1e-9 * max(1.0, abs(true)), (b, k, s, lb, true)


@pytest.mark.parametrize("N_p,cells,lin", [(24, 256, False), (24, 1024, True), (48, 2048, False)])
def test_threshold_search_and_split_search_are_exact(N_p, cells, lin):
    """iterative deepening on the bound returns the HiGHS optimum and plan; dealing the root's 32 subtrees to eight
    parts by bound rank (the team kernel's split) returns the same optimum, every subtree being searched exactly once"""
    wl = syn.dewh_batch(3, N_p, seed=23)
    for b in range(3):
        mats, prob = _problem(wl, b)
        dp = StageDpMw(*from_dewh_problem(mats, prob, wl["Nt"]), cells=cells, lin=lin)
        obj, u, nodes = dp.solve_ida()
        st, oref, vref = osv.solve_milp(prob, polish=True)
        assert st == osv.OPTIMAL and abs(obj + prob.c0 - oref) <= 1e-6 * max(1.0, abs(oref)), (b, obj, oref, nodes)
        assert np.array_equal(u, np.round(vref[prob.is_bin]))
        for nparts in (1, 3, 8):
            obj2, u2, nodes2 = dp.solve_split(nparts=nparts)
            assert abs(obj2 - obj) <= 1e-11 * max(1.0, abs(obj)), (b, nparts, obj2, obj)
            assert np.array_equal(u2, u), (b, nparts)
            assert nodes2 < 50 * max(nodes, 100)
