"""GPU parity tests (run with -m gpu on the B200 box).  Every check calls the CUDA path through the C ABI
(pyhybridcontrol_b200.cabi -> libhmpc.so) and compares with the committed goldens (generated from the unmodified
reference) and with the oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): condensed matrices
1e-10 relative, objectives 1e-6 relative, binary decisions exact where the optimum is unique."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden

pytestmark = pytest.mark.gpu

EVO = ("Phi_x", "Gamma_v", "Gamma_omega", "Gamma_5", "L_x", "L_v", "L_omega", "L_5", "H_x", "H_v", "H_omega", "H_5")


def _t(a, dev, dtype=None):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype or torch.float64).to(dev)


def _golden_batch(g, dev, B=1):
    from pyhybridcontrol_b200.batch import BatchMpc
    mats = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    return BatchMpc(mats, int(g["Nt"]) - 1, int(g["Nt"]), nu_l=int(g["nu_l"]), B=B, device=dev), mats


@pytest.mark.parametrize("case", golden_cases("condense"))
def test_condense_vs_reference_golden(case, cuda_device):
    g = load_golden("condense", case)
    bm, _ = _golden_batch(g, cuda_device)
    evo = bm.build()
    for name in EVO:
        ref = g["out_" + name]
        got = evo[name][0].cpu().numpy()
        if ref.size == 0:
            assert got.size == 0
            continue
        assert got.shape == ref.shape, name
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10 * max(1.0, np.abs(ref).max()), err_msg=name)


def test_condense_batched_and_broadcast(cuda_device):
    """a batch mixing per-agent and broadcast blocks equals per-agent condensing; B=1 and odd B work."""
    from oracle import mld as omld, condense as oc
    from pyhybridcontrol_b200.batch import BatchMpc
    rng = np.random.default_rng(3)
    B, Nt = 5, 11
    mats = dict(A=rng.standard_normal((B, 2, 2)) * 0.4, B1=rng.standard_normal((B, 2, 1)), B2=rng.standard_normal((1, 2, 1)),
                B4=rng.standard_normal((B, 2, 2)), b5=rng.standard_normal((B, 2, 1)), C=rng.standard_normal((1, 1, 2)),
                D1=rng.standard_normal((B, 1, 1)), d5=rng.standard_normal((B, 1, 1)),
                E=rng.standard_normal((B, 3, 2)), F1=rng.standard_normal((B, 3, 1)), F2=rng.standard_normal((B, 3, 1)),
                F4=rng.standard_normal((1, 3, 2)), G=rng.standard_normal((B, 3, 1)), Psi=-np.ones((1, 3, 1)),
                f5=rng.standard_normal((B, 3, 1)))
    bm = BatchMpc(mats, Nt - 1, Nt, nu_l=1, device=cuda_device)
    evo = bm.build()
    for b in range(B):
        full, d, vt = omld.complete({k: (v[b] if v.shape[0] == B else v[0]) for k, v in mats.items()}, nu_l=1)
        ref = oc.condense(full, d, Nt)
        for name in EVO:
            np.testing.assert_allclose(evo[name][b].cpu().numpy(), ref[name], rtol=0,
                                       atol=1e-10 * max(1.0, np.abs(ref[name]).max()), err_msg="%s[%d]" % (name, b))


def test_condense_linearity_full_size(cuda_device):
    """size-independent property at BASELINE size (B=100, N_p=48): H_5 is affine in f5/b5, Gamma is linear in B1."""
    import torch
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(100, 48, seed=2)
    e1 = BatchMpc(wl["mats"], 48, nu_l=1, device=cuda_device).build()
    m2 = dict(wl["mats"])
    m2["B1"] = 2.0 * m2["B1"]
    e2 = BatchMpc(m2, 48, nu_l=1, device=cuda_device).build()
    u_cols = torch.arange(0, 147, 3, device=cuda_device)
    assert torch.allclose(e2["Gamma_v"][:, :, u_cols], 2.0 * e1["Gamma_v"][:, :, u_cols], rtol=1e-14, atol=0)
    assert torch.equal(e2["Gamma_omega"], e1["Gamma_omega"]) and torch.equal(e2["H_5"], e1["H_5"])
    # x-prediction identity: row i of [Phi | Gamma_v | Gamma_w | Gamma_5] reproduces the recurrence
    x0 = _t(wl["x0"], cuda_device)
    w = _t(wl["omega"], cuda_device)
    v = torch.zeros((100, 147), dtype=torch.float64, device=cuda_device)
    v[:, 0::3] = (torch.arange(49, device=cuda_device) % 5 == 0).double()
    from pyhybridcontrol_b200 import cabi
    xt = cabi.predict(e1["Phi_x"], e1["Gamma_v"], e1["Gamma_omega"], e1["Gamma_5"].reshape(100, -1), x0, v, w).cpu().numpy()
    A, B1, B4, b5 = (wl["mats"][k][:, 0, 0] for k in ("A", "B1", "B4", "b5"))
    x = wl["x0"][:, 0].copy()
    for k in range(49):
        np.testing.assert_allclose(xt[:, k], x, rtol=1e-12)
        x = A * x + B1 * v[:, 3 * k].cpu().numpy() + B4 * wl["omega"][:, k] + b5


def test_constraint_rhs_and_scenarios(cuda_device):
    from oracle import mld as omld, condense as oc, assemble as oa
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B, N_p, S = 7, 24, 32
    wl = syn.dewh_batch(B, N_p, seed=4)
    Nt = wl["Nt"]
    bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=cuda_device)
    evo = bm.build()
    rng = np.random.default_rng(1)
    W = np.abs(rng.standard_normal((B, Nt, S))) * 0.01
    x0, w = _t(wl["x0"], cuda_device), _t(wl["omega"], cuda_device)
    r1 = cabi.constraint_rhs(bm.dims, evo, x0, w).cpu().numpy()
    r2 = cabi.constraint_rhs(bm.dims, evo, x0, None, scenarios=_t(W, cuda_device)).cpu().numpy()
    r3 = cabi.constraint_rhs(bm.dims, evo, x0, w, rows=2 * 8).cpu().numpy()
    for b in range(B):
        full, d, vt = omld.complete({k: v[b] for k, v in wl["mats"].items()}, nu_l=1)
        ref = oc.condense(full, d, Nt)
        _, a1 = oa.evo_rhs(ref, d, wl["x0"][b], wl["omega"][b])
        _, a2 = oa.evo_rhs(ref, d, wl["x0"][b], omega_scenarios=W[b])
        _, a3 = oa.evo_rhs(ref, d, wl["x0"][b], wl["omega"][b], N_tilde=8)
        np.testing.assert_allclose(r1[b], a1, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(r2[b], a2, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(r3[b], a3, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("case", golden_cases("lsim"))
def test_lsim_vs_reference_golden(case, cuda_device):
    from pyhybridcontrol_b200 import cabi
    g = load_golden("lsim", case)
    T = g["x"].shape[0]
    from pyhybridcontrol_b200.batch import BatchMpc
    mats = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    bm = BatchMpc(mats, 0, 1, nu_l=int(g["nu_l"]), B=T, device=cuda_device)
    d = cabi.make_dims(T, 1, nx=bm.dims.nx, nu=bm.dims.nu, ndelta=bm.dims.ndelta, nz=bm.dims.nz, nmu=bm.dims.nmu,
                       nomega=bm.dims.nomega, ny=bm.dims.ny, nc=bm.dims.nc)
    f = lambda k: _t(g[k].reshape(T, -1), cuda_device) if g[k].size else None
    x1, y, cons = cabi.lsim_step(d, bm.mats, f("x"), f("u"), f("delta"), f("z"), f("w"))
    ref_x1, ref_y = g["x1"].reshape(T, -1), g["y"].reshape(T, -1)
    if ref_x1.size:
        np.testing.assert_allclose(x1.cpu().numpy(), ref_x1, rtol=0, atol=1e-12 * max(1.0, np.abs(ref_x1).max()))
    if ref_y.size:
        np.testing.assert_allclose(y.cpu().numpy(), ref_y, rtol=0, atol=1e-12 * max(1.0, np.abs(ref_y).max()))
    assert np.array_equal(cons.cpu().numpy().astype(bool), g["cons"].reshape(T, -1).astype(bool))


def test_dewh_sim_step_and_control_model(cuda_device):
    import torch
    from oracle import lsim as ol
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B = 300
    plist = [syn.dewh_agent_params(i) for i in range(B)]
    params = _t(cabi.pack_dewh_params(plist), cuda_device)
    rng = np.random.default_rng(0)
    T = rng.uniform(10, 70, B)          # includes states below T_w -> clamp path
    u = (rng.random(B) > 0.5).astype(float)
    D = rng.random(B) * 0.02
    T1, model, cons = cabi.dewh_sim_step(params, _t(T, cuda_device), _t(u, cuda_device), _t(D, cuda_device), want_model=True)
    cm = cabi.dewh_control_model(params).cpu().numpy()
    T1 = T1.cpu().numpy()
    for b in range(B):
        x1, xc, c = ol.dewh_sim_step(dict(plist[b]), T[b], u[b], D[b])
        assert abs(T1[b] - x1) <= 1e-11 * max(1.0, abs(x1))
        assert np.array_equal(cons[b].cpu().numpy().astype(bool), c)
        np.testing.assert_allclose(cm[b], ol.dewh_scalars(plist[b]), rtol=1e-13)


def test_aggregate_power(cuda_device):
    import torch
    from pyhybridcontrol_b200 import cabi
    rng = np.random.default_rng(0)
    for B, Nt in ((1, 5), (257, 49), (1000, 49)):
        v = torch.as_tensor(rng.random((B, Nt, 3))).to(cuda_device)
        P = torch.as_tensor(rng.uniform(2700, 3300, B)).to(cuda_device)
        u = v[:, :, 0]                                   # strided view, like the solver output
        got = cabi.aggregate_power(u, P).cpu().numpy()
        ref = (P.cpu().numpy()[:, None] * u.cpu().numpy()).sum(axis=0)
        np.testing.assert_allclose(got, ref, rtol=1e-13)
        assert np.array_equal(got, cabi.aggregate_power(u, P).cpu().numpy())      # deterministic


SOLVERS = ("stage_dp", "bnc")     # both K3/K4 implementations must pass every DEWH solve test


def _dewh_solve(wl, dev, solver="auto", mip_rel_gap=0.0):
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    B, Nt = wl["B"], wl["Nt"]
    bm = BatchMpc(wl["mats"], wl["N_p"], nu_l=1, device=dev, solver=solver,
                  opts=cabi.default_opts(mip_rel_gap=mip_rel_gap), dp_opts=cabi.stage_dp_default_opts(mip_rel_gap=mip_rel_gap))
    bm.build()
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1))
    assert res["solver"] == (solver if solver != "auto" else "stage_dp")
    return bm, {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}, cost.reshape(B, -1)


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("N_p", [24, 48])
def test_milp_vs_highs_golden(N_p, solver, cuda_device):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    g = load_golden("milp", "dewh_N%d" % N_p)
    wl = syn.dewh_batch(int(g["B"]), N_p, seed=int(g["seed"]))
    bm, res, _ = _dewh_solve(wl, cuda_device, solver)
    assert (res["status"] == 0).all()
    np.testing.assert_allclose(res["obj"], g["obj"], rtol=1e-6, atol=1e-9)
    isb = bm.is_bin_v.astype(bool)
    assert np.array_equal(np.round(res["v"][:, isb]), np.round(g["v"][:, isb]))      # binary decisions, exact
    assert np.array_equal(res["v"][:, 0], np.round(g["v"][:, 0]))                    # first applied control


@pytest.mark.parametrize("solver", SOLVERS)
def test_milp_vs_enumeration_small(solver, cuda_device):
    """independent check on problems small enough to enumerate (2^9 assignments)."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(6, 8, seed=9)
    bm, res, _ = _dewh_solve(wl, cuda_device, solver)
    for b in range(6):
        full, d, vt = omld.complete({k: v[b] for k, v in wl["mats"].items()}, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, 9), d, vt, 9, wl["x0"][b], wl["omega"][b],
                                atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
        st, obj, v, second = osv.solve_enumerate(prob)
        assert res["status"][b] == 0 and abs(res["obj"][b] - obj) <= 1e-6 * max(1.0, abs(obj))
        if second - obj > 1e-6:
            assert np.array_equal(np.round(res["v"][b][prob.is_bin]), np.round(v[prob.is_bin]))


@pytest.mark.parametrize("solver", SOLVERS)
def test_milp_full_size_properties(solver, cuda_device):
    """BASELINE config 2 (B=100, N_p=48): every returned point is binary-integral, satisfies H v <= rhs, its
    objective equals c'v, and a sample of agents matches HiGHS."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(100, 48, seed=1)
    bm, res, cost = _dewh_solve(wl, cuda_device, solver)
    assert (res["status"] == 0).all()
    v = res["v"]
    isb = bm.is_bin_v.astype(bool)
    assert np.all((v[:, isb] == 0) | (v[:, isb] == 1))
    assert np.all(v[:, ~isb] >= -1e-9)
    np.testing.assert_allclose((cost * v).sum(axis=1), res["obj"], rtol=1e-9, atol=1e-12)
    H = bm.evo["H_v"].cpu().numpy()
    rhs = cabi.constraint_rhs(bm.dims, bm.evo, _t(wl["x0"], cuda_device), _t(wl["omega"], cuda_device)).cpu().numpy()
    assert np.all(np.einsum("bij,bj->bi", H, v) <= rhs + 1e-6)
    for b in range(0, 100, 9):
        full, d, vt = omld.complete({k: m[b] for k, m in wl["mats"].items()}, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, 49), d, vt, 49, wl["x0"][b], wl["omega"][b],
                                atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
        st, obj, vr = osv.solve_milp(prob)
        assert abs(res["obj"][b] - obj) <= 1e-6 * max(1.0, abs(obj))
        assert np.array_equal(np.round(v[b][isb]), np.round(vr[isb]))


def test_milp_edge_cases(cuda_device):
    import torch
    from pyhybridcontrol_b200 import cabi
    dev = cuda_device
    # (1) pure LP, no binaries: min -x0 - x1  s.t. x0 + x1 <= 1.5, x <= 1
    c = _t([[-1.0, -1.0]], dev)
    H = _t([[[1.0, 1.0]]], dev)
    v, obj, st, stats = cabi.milp_solve(c, H, _t([[1.5]], dev), _t([0.0, 0.0], dev), _t([1.0, 1.0], dev),
                                        _t([0, 0], dev, torch.uint8))
    assert int(st[0]) == 0 and abs(float(obj[0]) + 1.5) < 1e-9
    # (2) same with binaries -> optimum -1 (only one can be on)
    v, obj, st, stats = cabi.milp_solve(c, H, _t([[1.5]], dev), _t([0.0, 0.0], dev), _t([1.0, 1.0], dev),
                                        _t([1, 1], dev, torch.uint8))
    assert int(st[0]) == 0 and abs(float(obj[0]) + 1.0) < 1e-9 and sorted(v[0].cpu().tolist()) == [0.0, 1.0]
    # (3) infeasible: x0 + x1 <= -1 with x >= 0
    v, obj, st, stats = cabi.milp_solve(_t([[1.0, 1.0]], dev), H, _t([[-1.0]], dev), _t([0.0, 0.0], dev),
                                        _t([1.0, 1.0], dev), _t([1, 1], dev, torch.uint8))
    assert int(st[0]) == 1 and not np.isfinite(float(obj[0]))
    # (4) no constraint rows at all (m = 0): every variable goes to the bound its cost prefers
    v, obj, st, stats = cabi.milp_solve(_t([[1.0, -2.0, 0.5]], dev), torch.zeros((1, 0, 3), dtype=torch.float64, device=dev),
                                        torch.zeros((1, 0), dtype=torch.float64, device=dev), _t([0.0, 0.0, 0.0], dev),
                                        _t([1.0, 1.0, 1.0], dev), _t([1, 1, 0], dev, torch.uint8))
    assert int(st[0]) == 0 and v[0].cpu().tolist() == [0.0, 1.0, 0.0] and abs(float(obj[0]) + 2.0) < 1e-12
    # (5) knapsack-cover with ragged batch (B = 3, different rhs): min sum c u  s.t. -sum a u <= -b
    a = np.array([4.3, 4.2, 4.1, 4.0, 3.9, 3.8])
    cc = np.array([1.0, 1.1, 0.9, 1.3, 0.7, 1.2])
    Hk = _t(np.tile(-a, (3, 1, 1)), dev)
    rhs = _t([[-4.0], [-8.1], [-16.5]], dev)
    v, obj, st, stats = cabi.milp_solve(_t(np.tile(cc, (3, 1)), dev), Hk, rhs, _t(np.zeros(6), dev), _t(np.ones(6), dev),
                                        _t(np.ones(6, dtype=np.uint8), dev, torch.uint8))
    import itertools
    for b, need in enumerate((4.0, 8.1, 16.5)):
        best = min(sum(cc[list(s)]) for r in range(7) for s in itertools.combinations(range(6), r) if a[list(s)].sum() >= need)
        assert int(st[b]) == 0 and abs(float(obj[b]) - best) < 1e-9


@pytest.mark.parametrize("solver", SOLVERS)
def test_mip_gap_option(solver, cuda_device):
    """with MIPGap = 1e-2 (what the reference ran) the returned objective is within 1 % of the proven optimum."""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(16, 48, seed=5)
    _, exact, _ = _dewh_solve(wl, cuda_device, solver)
    _, loose, _ = _dewh_solve(wl, cuda_device, solver, mip_rel_gap=1e-2)
    assert (loose["status"] == 0).all()
    assert np.all(loose["obj"] >= exact["obj"] - 1e-9) and np.all(loose["obj"] <= exact["obj"] * 1.01 + 1e-9)


@pytest.mark.parametrize("solver", SOLVERS)
def test_host_front_door_matches_device_path(solver, cuda_device):
    """hmpc_mpc_step_host_f64 (numpy in / numpy out) == device-pointer path, and re-use without recondensing."""
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(10, 24, seed=6)
    bm, res, cost = _dewh_solve(wl, cuda_device, solver)
    plan = cabi.StepPlan(bm.dims, cabi.default_opts(force_general=int(solver == "bnc")))
    hm = dict(wl["mats"])
    hm["C"] = np.ones((1, 1, 1))
    v, obj, st, stats, tm = plan.step(hm, wl["x0"], wl["omega"], cost, bm.lb_v, bm.ub_v, bm.is_bin_v, recondense=True)
    assert (st == 0).all() and plan.last_solver == solver
    np.testing.assert_allclose(obj, res["obj"], rtol=1e-12)
    assert np.array_equal(v, res["v"])
    v2, obj2, st2, _, _ = plan.step(None, wl["x0"], wl["omega"], cost, bm.lb_v, bm.ub_v, bm.is_bin_v, recondense=False)
    assert np.array_equal(obj2, obj)
    h2d, d2h = plan.bytes_per_step(True)
    assert h2d > 0 and d2h == 8 * (10 * 75 + 10) + 4 * (10 + 80)
    plan.close()


def test_smoke_entry(cuda_device):
    import __graft_entry__
    __graft_entry__.smoke()


def test_peer_store_aggregate_exchange_emulated(cuda_device):
    """K6 without a collective: publish / gather kernels of csrc/aggregate.cu with TWO (and three) ranks emulated on
    one GPU -- all windows on this device, the ranks' publishes issued one after the other, so that every flag is up
    before a gather looks at it.  Sums equal numpy's in rank order, step after step (the ring of four slots wraps);
    a missing peer gives NaN and the error word, not a hang."""
    import torch
    from pyhybridcontrol_b200.distributed import PeerExchange
    rng = np.random.default_rng(3)
    Nt = 49
    for world in (1, 2, 3):
        wins = [PeerExchange.new_window(Nt, world, cuda_device) for _ in range(world)]
        ex = [PeerExchange(Nt, cuda_device, world=world, rank=r, windows=wins) for r in range(world)]
        Bs = [37, 100, 5][:world]
        P = [torch.as_tensor(rng.uniform(2700, 3300, b)).to(cuda_device) for b in Bs]
        for step in range(9):
            us = [torch.as_tensor((rng.random((b, Nt)) > 0.5).astype(float)).to(cuda_device) for b in Bs]
            prevs = [torch.empty(Nt, dtype=torch.float64, device=cuda_device) for _ in range(world)]
            for r in range(world):
                ex[r].publish(us[r], P[r], out_prev=prevs[r] if step % 2 else None)
            ref = np.zeros(Nt)
            for r in range(world):          # rank order, chunks of 16 agents in order: the kernel's summation order
                loc = np.zeros(Nt)
                un, pn = us[r].cpu().numpy(), P[r].cpu().numpy()
                for c0 in range(0, Bs[r], 16):
                    acc = np.zeros(Nt)
                    for b in range(c0, min(c0 + 16, Bs[r])):
                        acc = acc + pn[b] * un[b]
                    loc = loc + acc
                ref = ref + loc
            for r in range(world):
                got = ex[r].gather().cpu().numpy()
                assert np.array_equal(got, ref), (world, step, r)
                assert ex[r].error() == 0
                if step % 2 and r == 0:     # (rank 0 publishes first: every peer's previous step was complete by then)
                    assert np.array_equal(prevs[r].cpu().numpy(), last_ref), (world, step)
            last_ref = ref
    # a peer that never publishes: bounded wait, NaN result, error word = the step that was waited for
    wins = [PeerExchange.new_window(Nt, 2, cuda_device) for _ in range(2)]
    ex0 = PeerExchange(Nt, cuda_device, world=2, rank=0, windows=wins)
    ex0.publish(torch.ones((4, Nt), dtype=torch.float64, device=cuda_device), torch.ones(4, dtype=torch.float64, device=cuda_device))
    out = ex0.gather(spin_limit=1000).cpu().numpy()
    assert np.isnan(out).all() and ex0.error() == 1
    # lagged gathers of a longer run (ring of eight slots)
    wins = [PeerExchange.new_window(Nt, 1, cuda_device)]
    ex1 = PeerExchange(Nt, cuda_device, world=1, rank=0, windows=wins)
    hist = []
    one = torch.ones(1, dtype=torch.float64, device=cuda_device)
    for step in range(20):
        u1 = torch.full((1, Nt), float(step), dtype=torch.float64, device=cuda_device)
        back = torch.empty(Nt, dtype=torch.float64, device=cuda_device)
        ex1.publish(u1, one, out_prev=back, lag=4)
        hist.append(back.cpu().numpy()[0])
    assert hist[:4] == [0.0] * 4 and hist[4:] == [float(i) for i in range(16)]
