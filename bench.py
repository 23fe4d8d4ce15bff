#!/usr/bin/env python
"""bench.py -- headline benchmark: DEWH hybrid-MPC MILP solves/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--agents B] [--np N_p]

Workload (config.workload): BASELINE.json configs[1] -- a batch of 100 synthetic DEWH agents per GPU, N_p = 48
(n = 147 variables of which 49 binary, m = 98 rows per agent), independent MILPs, solved to proven optimality
(mip_rel_gap = 0).  One "step" = one pass of the hot path over the batch:
    K1 condense -> K2 constraint rhs -> K3/K4 branch-and-cut solve -> K5 DEWH sim step -> K6 aggregate power
                                                                               (+ NCCL all-reduce when N > 1).
Every step uses a different control instant (new states, demand forecast and prices), so nothing is cached.

* value  : whole-job solves/s with inputs resident in HBM, CUDA-event time of the K steps, max over ranks.
* e2e    : same metric through the host-buffer C-ABI call (hmpc_mpc_step_host_f64): numpy in, numpy out,
           host<->device copies inside the timed region.
* roofline / cpu_baseline objects: see DESIGN.md section 6.
* --impl reference : the reference's CPU path (oracle port: numpy condensing + HiGHS via scipy) on all host
  cores, same workload, same metric.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dewh_hybrid_mpc_milp_solves_per_sec"
UNIT = "solves/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--pool", type=int, default=32, help="distinct control instants the steps cycle through")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one (no CUDA graph)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--agents", type=int, default=100, help="agents per GPU (weak scaling)")
    ap.add_argument("--np", dest="N_p", type=int, default=48)
    ap.add_argument("--cpu-sample", type=int, default=256, help="agent-solves timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=80.0, help="--impl reference: target wall time of the run")
    ap.add_argument("--solver", default="auto", choices=("auto", "bnc", "stage_dp"),
                    help="auto = exact stage-DP kernels for the scalar-state DEWH class, bnc = general branch-and-cut")
    ap.add_argument("--cells", type=int, default=0, help="stage-DP value-table cells per stage (0 = library default)")
    ap.add_argument("--recondense", action="store_true",
                    help="run K1 inside every timed step although the models do not change (the reference condenses "
                         "only when the model's version changed, mld_evolution_matrices.py:79)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the BASELINE configs[2] / configs[3] blocks")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "configs[1]: batch of %d synthetic DEWH agents per GPU, N_p=%d, independent MILPs, gap 0"
                        % (args.agents, args.N_p),
            "agents_per_gpu": args.agents, "N_p": args.N_p, "n_vars": 3 * (args.N_p + 1), "n_binaries": args.N_p + 1,
            "n_rows": 2 * (args.N_p + 1), "parallelism": "agents sharded x%d" % n_gpus,
            "l2_policy": "L2 flushed (256 MiB write) before every timed step",
            "inputs": "a pool of %d distinct control instants, cycled" % min(max(args.warmup, 3) + args.steps, args.pool)}


# ------------------------------------------------------------------------------------------------ CPU reference
def _cpu_solve_one(job):
    """One agent's control step on the CPU: numpy condensing + assembly + HiGHS (the oracle port)."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    mats, Nt, x0, omega, q_u, q_mu = job
    full, d, vt = omld.complete(mats, nu_l=1)
    evo = oc.condense(full, d, Nt)
    prob = oa.build_problem(evo, d, vt, Nt, x0, omega, atoms=dict(q_u=q_u, q_mu=q_mu))
    st, obj, v = osv.solve_milp(prob)
    return obj


def cpu_jobs(args, step, first_agent=0, count=None):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B = args.agents if count is None else count
    wl = syn.dewh_batch(B, args.N_p, seed=1, k0=step, first_agent=first_agent)
    return [({k: v[b] for k, v in wl["mats"].items()}, wl["Nt"], wl["x0"][b], wl["omega"][b], wl["q_u"][b],
             wl["q_mu"][b]) for b in range(B)]


def run_cpu(args, jobs_per_step, steps, warmup, cores):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for s in range(warmup + steps):
            jobs = jobs_per_step(s)
            t0 = time.perf_counter()
            pool.map(_cpu_solve_one, jobs, chunksize=max(1, len(jobs) // (4 * cores)))
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    return times


def main_reference(args):
    """Reference arm: the reference's CPU path (oracle port: numpy condensing + HiGHS) on
    all host cores.  Every step solves a BOUNDED SAMPLE of the step's agents, sized from a probe so that the whole
    run takes about a minute and a half whatever --steps is; throughput is per solve, so the unit is unchanged."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, Wr = args.steps, min(args.warmup, 1)
    probe = run_cpu(args, lambda s: cpu_jobs(args, s, count=min(args.agents, 2 * cores)), 1, 0, cores)
    rate = min(args.agents, 2 * cores) / probe[0]
    n = int(max(min(cores, args.agents), min(args.agents, args.ref_budget_s * rate / (K + Wr))))
    # all K bounded steps are queued at once so that the pool never idles on a straggler: the number is the CPU's
    # best sustained throughput on this workload, not a per-step latency
    import multiprocessing as mp
    jobs = []
    for s_ in range(K):
        jobs += cpu_jobs(args, s_ % args.pool, count=n)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_solve_one, jobs[:cores], chunksize=1)                      # warm-up
        t0 = time.perf_counter()
        pool.map(_cpu_solve_one, jobs, chunksize=max(1, len(jobs) // (16 * cores)))
        total = time.perf_counter() - t0
    times = [total / K] * K
    value = n * K / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.agents / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d of the step's %d agents (bounded sample), numpy condensing + HiGHS "
                                       "1.12 (scipy.optimize.milp, gap 0), multiprocessing pool on all host cores; "
                                       "ms_per_step = time of a full %d-agent step at the measured rate"
                                       % (len(times), n, args.agents, args.agents)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "latency_p50_ms": 1e3 * float(np.median(times)) * args.agents / n, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index, enabled=True):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        self.enabled, self.started = enabled, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def start_once(self):
        if self.enabled and not self.started:
            self.started = True
            self.start()

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                 getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ BASELINE configs[2], [3]
def extra_configs(world, rank, dev):
    """BASELINE.json configs[2] and configs[3], strong-scaled over the ranks of this run, through the public fleet API:
      config3  residential micro-grid: 1,000 DEWHs + PV + residential demand + grid agent, N_p = 48 -- per step every
               rank solves its contiguous shard, the aggregate power is all-reduced, the grid MLD is evaluated;
      config4  10,000 DEWHs, closed loop over 24 h (96 steps of 15 min): per step MLD re-parametrisation, MILP solve,
               simulation step, aggregate exchange.
    Times are CUDA events, max over ranks; every solve is proven optimal unless `not_optimal` says otherwise."""
    import torch
    import torch.distributed as dist
    from pyhybridcontrol_b200 import distributed
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    out = {}
    N_p = 48
    Nt = N_p + 1

    def params(lo, hi):
        base = [syn.dewh_agent_params(a) for a in range(256)]
        return [base[b % 256] for b in range(lo, hi)]

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- config 3
    total = 1000
    lo, hi = distributed.shard_range(total, rank, world)
    B = hi - lo
    fleet = DewhFleet(params(lo, hi), N_p, device=dev)
    rng = np.random.default_rng(0)
    T0_all = rng.integers(55, 65, size=total).astype(float)
    T0 = torch.as_tensor(T0_all[lo:hi]).to(dev)
    demand = torch.as_tensor(np.stack([syn.dhw_demand_profile(Nt, seed=b % 256) for b in range(lo, hi)])).to(dev)
    price = syn.price_profile(Nt, seed=1)
    k = np.arange(Nt)
    p_pv = torch.as_tensor(-3000.0 * total * np.clip(np.sin((k / 96.0) * 2 * np.pi - 0.5 * np.pi), 0, None)).to(dev)
    p_res = torch.as_tensor(1200.0 * total * (1.0 + 0.3 * np.sin(k / 96.0 * 4 * np.pi))).to(dev)
    fleet.build()
    cost = fleet.cost_from_prices(price)

    def microgrid_step():
        res = fleet.control_step(T0.reshape(B, 1), demand, cost)
        p_dev = fleet.aggregate_power(res["u"])            # + NCCL all-reduce when several ranks run
        grid = distributed.grid_evaluate(p_dev, p_pv, p_res)
        return res, grid
    for _ in range(3):
        res, grid = microgrid_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res, grid = microgrid_step()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
    out["config3"] = {"workload": "configs[2]: 1,000 DEWHs + PV + demand + grid agent, N_p=48, agents sharded over %d GPU(s), "
                                  "all-reduce of the aggregate power, grid MLD evaluated every step" % world,
                      "agents": total, "scaling": "strong", "ms_per_step": ms, "solves_per_s": total / ms * 1e3,
                      "not_optimal": int(sum_over_ranks(int((res["status"] != 0).sum()))),
                      "grid_import_cost": float((grid["p_imp"] * torch.as_tensor(price).to(dev)).sum())}
    del fleet
    torch.cuda.empty_cache()
    # ---- config 4
    total, steps = 10000, 96
    lo, hi = distributed.shard_range(total, rank, world)
    B = hi - lo
    fleet = DewhFleet(params(lo, hi), N_p, device=dev)
    T0 = np.random.default_rng(1).integers(55, 65, size=total).astype(float)[lo:hi]
    prof = np.stack([syn.dhw_demand_profile(steps + Nt, seed=b) for b in range(256)])
    demand = prof[np.arange(lo, hi) % 256]
    price = syn.price_profile(steps + Nt, seed=2)
    # warm-up = the whole loop once: the first pass grows the caching allocator's pools (measured: 20-80 ms stalls in a
    # handful of its steps, none in the second pass over the same inputs)
    fleet.closed_loop(T0, demand, price, steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    log = fleet.closed_loop(T0, demand, price, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out["config4"] = {"workload": "configs[3]: 10,000 DEWHs closed loop 24 h (96 x 15 min: re-parametrisation + MILP solve + "
                                  "sim step + aggregate exchange), agents sharded over %d GPU(s)" % world,
                      "agents": total, "sim_steps": steps, "scaling": "strong", "wall_ms": ms,
                      "ms_per_control_step": ms / steps, "solves_per_s": total * steps / ms * 1e3,
                      "not_optimal": int(sum_over_ranks(int((log["status"] != 0).sum()))),
                      "T_min": float(log["T"].min()), "T_max": float(log["T"].max()),
                      "heater_duty": float(log["u"].mean())}
    del fleet, log
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main_ours(args):
    import torch
    import torch.distributed as dist
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N_p, Nt = args.agents, args.N_p, args.N_p + 1
    K, W = args.steps, max(args.warmup, 3)
    first_agent = rank * B

    # ---- synthetic inputs: a pool of distinct control instants (new states, forecasts and prices each), resident
    #      in HBM before timing starts; the timed steps cycle through the pool
    P = min(W + K, args.pool)
    wl0 = syn.dewh_batch(B, N_p, seed=1, k0=0, first_agent=first_agent)
    fleet = DewhFleet(wl0["params"], N_p, device=dev)
    steps_in = []
    for s in range(P):
        wl = wl0 if s == 0 else syn.dewh_batch(B, N_p, seed=1, k0=s, first_agent=first_agent)
        cost = np.zeros((B, Nt, 3))
        cost[:, :, 0] = wl["q_u"]
        cost[:, :, 1:] = wl["q_mu"][:, None, :]
        # one packed block per instant [x0 | omega | cost]: a new instant reaches the step's static buffers in ONE copy
        packed = torch.cat([torch.as_tensor(wl["x0"]).reshape(-1), torch.as_tensor(wl["omega"]).reshape(-1),
                            torch.as_tensor(cost).reshape(-1)]).to(dev)
        steps_in.append(dict(packed=packed, x0=packed[:B].view(B, 1), omega=packed[B:B + B * Nt].view(B, Nt),
                             cost=packed[B + B * Nt:].view(B, -1), host=wl, host_cost=cost.reshape(B, -1)))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    names = ("condense", "rhs", "solve", "sim", "aggregate")
    use_dp = args.solver == "stage_dp" or (args.solver == "auto" and fleet.batch.stage_dp_ok)
    dp_opts = cabi.stage_dp_default_opts(**({"cells": args.cells} if args.cells else {}))
    kernel_ms = {k: 0.0 for k in names}
    solve_stats, statuses = [], []
    fleet.build()                                  # K1 once: the models do not change between the control instants
    sim_stream = torch.cuda.Stream()

    def one_step(inp, timed_events=None):
        ev = timed_events
        if ev:
            ev[0].record()
        if args.recondense or ev:                  # (the per-kernel breakdown pass always runs K1, to time it)
            fleet.build()
        if ev:
            ev[1].record()
        rhs = cabi.constraint_rhs(fleet.batch.dims, fleet.batch.evo, inp["x0"], inp["omega"])
        if ev:
            ev[2].record()
        lb, ub, isb = fleet.batch._bounds_dev()
        if use_dp:
            v, obj, status, stats = cabi.stage_dp_solve(fleet.batch.dims, fleet.batch.mats, rhs, inp["cost"], lb, ub, isb,
                                                        dp_opts)
        else:
            v, obj, status, stats = cabi.milp_solve(inp["cost"], fleet.batch.evo["H_v"], rhs, lb, ub, isb,
                                                    fleet.batch.opts)
        if ev:
            ev[3].record()
        u = v.view(B, Nt, 3)[:, :, 0]
        # K5 and K6 both consume the solution and nothing of each other: K5 runs on a side stream next to K6 (in the
        # breakdown pass, which times them one by one, they stay in line)
        main_s = torch.cuda.current_stream()
        if ev:
            T1, cons = fleet.sim_step(inp["x0"][:, 0].contiguous(), u[:, 0].contiguous(), inp["omega"][:, 0].contiguous())
            ev[4].record()
        else:
            sim_stream.wait_stream(main_s)
            with torch.cuda.stream(sim_stream):
                T1, cons = fleet.sim_step(inp["x0"][:, 0].contiguous(), u[:, 0].contiguous(),
                                          inp["omega"][:, 0].contiguous())
        if peer_ex is not None:
            # K6 with the exchange fused in: the reduction's last pass stores this rank's sums into every rank's window
            # over NVLink; the same launch adds the world's contributions of the step four publishes back (pipelined: a
            # rank may run up to four steps ahead of the slowest one, as with the ring of four NCCL buffers before)
            peer_ex.publish(u, fleet.P_nom, out_prev=p_total_prev, lag=4)
            p_agg = p_total_prev
        else:
            p_agg = cabi.aggregate_power(u, fleet.P_nom)      # this rank's agents; ranks are summed by exchange()
        if ev:
            ev[5].record()
        else:
            main_s.wait_stream(sim_stream)
        return v, obj, status, stats, T1, p_agg

    # K6 across ranks: NCCL all-reduce of the [Nt] aggregate power.  Nothing in the NEXT control step depends on it
    # (the agents are decentralised; the sum goes to the grid agent), so it runs on a side stream, pipelined with the
    # following steps through a small ring of buffers, and is joined before the timed region ends.
    peer_ex, peer_note = None, "single GPU: no exchange"
    p_total_prev = torch.zeros(Nt, dtype=torch.float64, device=dev)
    if world > 1 and not os.environ.get("HMPC_NCCL_EXCHANGE"):
        try:
            from pyhybridcontrol_b200.distributed import PeerExchange
            peer_ex = PeerExchange(Nt, dev)
            peer_note = ("peer-store exchange inside the step's CUDA graph (symmetric memory over NVLink; publish fused "
                         "into the reduction's last pass, gather four steps behind): no host call per step")
        except Exception as exc:
            peer_ex = None
            peer_note = "NCCL all-reduce on a side stream (symmetric memory unavailable: %r)" % (exc,)
    elif world > 1:
        peer_note = "NCCL all-reduce on a side stream, pipelined with the next steps (HMPC_NCCL_EXCHANGE)"
    ok_all = torch.tensor([1.0 if (peer_ex is not None or world == 1) else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)          # every rank must take the same path
        if float(ok_all.item()) < 1.0:
            peer_ex = None
    xstream = torch.cuda.Stream() if world > 1 else None
    ring = [torch.empty(Nt, dtype=torch.float64, device=dev) for _ in range(4)]
    ring_ev = [None] * 4
    ring_evobj = [torch.cuda.Event() for _ in range(4)]

    def exchange(p_agg, s):
        if world == 1 or peer_ex is not None or os.environ.get("HMPC_NO_EXCHANGE"):
            return p_agg                              # (the peer-store exchange is part of the step itself)
        main = torch.cuda.current_stream()
        buf = ring[s % 4]
        if ring_ev[s % 4] is not None:
            main.wait_event(ring_ev[s % 4])          # the exchange that used this buffer 4 steps ago is done
        buf.copy_(p_agg)
        xstream.wait_stream(main)
        with torch.cuda.stream(xstream):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            ring_ev[s % 4] = ring_evobj[s % 4]
            ring_ev[s % 4].record(xstream)
        return buf

    def join_exchange():
        if world > 1:
            torch.cuda.current_stream().wait_stream(xstream)

    def note(msg):
        if os.environ.get("HMPC_BENCH_VERBOSE"):
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank, enabled=(rank == 0))   # one NVML poller per box: queries contend with launches
    # ---- the step as ONE CUDA graph over static input buffers (the per-step inputs are copied in, device to
    #      device, inside the timed region); --no-graph launches the kernels one by one instead
    static_packed = torch.empty_like(steps_in[0]["packed"])
    static = dict(x0=static_packed[:B].view(B, 1), omega=static_packed[B:B + B * Nt].view(B, Nt),
                  cost=static_packed[B + B * Nt:].view(B, -1))

    def load_inputs(inp):
        static_packed.copy_(inp["packed"])

    for s in range(W):
        flush.fill_(float(s))
        load_inputs(steps_in[s % P])
        out = one_step(static)
        exchange(out[5], s)
    join_exchange()
    barrier()
    note("warm-up done")
    # ---- per-kernel breakdown: an extra, untimed pass with events between the launches (no graph)
    Kb = min(K, 20)
    bevs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(Kb)]
    for s in range(Kb):
        flush.fill_(float(s))
        load_inputs(steps_in[s % P])
        one_step(static, bevs[s])
    barrier()
    for e in bevs:
        for i, k in enumerate(names):
            kernel_ms[k] += e[i].elapsed_time(e[i + 1]) * K / Kb
    # One CUDA graph per pool instant: input copies + K1..K6 (local part), so that a timed step costs the host a
    # flush launch, one graph launch and the exchange enqueue -- with 8 ranks on one box the host side is otherwise
    # what limits the step rate.  The graphs share one memory pool (they never run concurrently).
    graph = None
    graphs, g_outs = [], []
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            load_inputs(steps_in[0])
            one_step(static)
        torch.cuda.current_stream().wait_stream(side)
        barrier()
        pool_id = None
        for pidx in range(P):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool_id, capture_error_mode="thread_local"):
                flush.fill_(float(pidx))               # L2 flush: first node of every step's graph
                load_inputs(steps_in[pidx])
                go = one_step(static)
            if pool_id is None:
                pool_id = g.pool()
            graphs.append(g)
            g_outs.append(go)
        graph = graphs[0]
        barrier()
        sampler.start_once()            # clocks are sampled from here on: the same steps run, untimed, right now
        for g in graphs:                # first launch of a graph uploads it to the device: not part of a step
            g.replay()
        barrier()
    note("graph captured")
    sampler.start_once()
    launches0 = cabi.launch_count
    # Timing: ONE pair of events around the K steps (everything on the device between them counts: the steps, the
    # input copies, the pipelined exchanges, launch gaps) minus K L2 flushes.  With graphs the flush is the first node
    # of the step's graph (one host call per step: with 8 ranks on one box the host's launch rate is what limits the
    # step rate otherwise) and its duration is calibrated right before the timed region; without graphs every flush
    # is bracketed by its own events.  Per-step events on every 4th step give the latency percentiles.
    cal = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush.fill_(0.5)
    cal[0].record()
    for i in range(20):
        flush.fill_(float(i))
    cal[1].record()
    barrier()
    flush_one_ms = cal[0].elapsed_time(cal[1]) / 20.0
    evs = {s: [torch.cuda.Event(enable_timing=True) for _ in range(4)] for s in range(K) if graph is None or s % 4 == 0}
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    host_us = [0.0, 0.0]
    ev_start.record()
    for s in range(K):
        e = evs.get(s)
        if graph is not None:
            if e:
                e[0].record()
            out = g_outs[(W + s) % P]
            th0 = time.perf_counter()
            graphs[(W + s) % P].replay()
            host_us[0] += time.perf_counter() - th0
        else:
            e[2].record()
            flush.fill_(float(s))
            e[3].record()
            e[0].record()
            load_inputs(steps_in[(W + s) % P])
            out = one_step(static)
        th0 = time.perf_counter()
        exchange(out[5], s)
        host_us[1] += time.perf_counter() - th0
        if e:
            e[1].record()
        if s % 8 == 0 or s == K - 1:                  # solver statistics of a sample of the steps
            statuses.append(out[2].clone())
            solve_stats.append(out[3].clone())
    t_host_loop = time.perf_counter() - t_wall0
    join_exchange()
    if peer_ex is not None:
        peer_ex.gather(out=p_total_prev, lag=0)          # the last step's sum (the loop gathered the previous ones)
    ev_end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    note("timed region done")
    if peer_ex is not None:
        assert peer_ex.error() == 0, "peer exchange: a rank never published step %d" % peer_ex.error()
        assert bool(torch.isfinite(p_total_prev).all())
    # [K1], K2, K3/K4 (stage-DP: one fused launch up to 296 agents, else three), K5, K6 (two launches)
    launches_per_step = ((1 if args.recondense else 0) + 1 + (cabi.stage_dp_launches(B, dp_opts.fuse_search) if use_dp else 1)
                         + 1 + 2)
    launches = launches_per_step * K if graph is not None else cabi.launch_count - launches0
    if graph is not None:
        step_ms = [e[0].elapsed_time(e[1]) - flush_one_ms for e in evs.values()]
        flush_ms = flush_one_ms * K
    else:
        step_ms = [e[0].elapsed_time(e[1]) for e in evs.values()]
        flush_ms = float(sum(e[2].elapsed_time(e[3]) for e in evs.values()))
    total_ms = float(ev_start.elapsed_time(ev_end)) - flush_ms
    last_obj = out[1].clone()
    # every rank's own SM clock while its GPU is still warm (one NVML query per rank, outside the timed region)
    my_mhz = 0.0
    try:
        import pynvml
        pynvml.nvmlInit()
        my_mhz = float(pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank), pynvml.NVML_CLOCK_SM))
    except Exception:
        pass
    rank_mhz = [my_mhz]
    rank_ms = [total_ms / K]
    if world > 1:
        tm = torch.tensor([my_mhz], dtype=torch.float64, device=dev)
        allm = [torch.empty_like(tm) for _ in range(world)]
        dist.all_gather(allm, tm)
        rank_mhz = [float(x.item()) for x in allm]
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x.item()) / K for x in allt]
        total_ms = max(rank_ms) * K
    outs = [(None, None, st_, ss_) for st_, ss_ in zip(statuses, solve_stats)]
    solve_stats = []
    not_opt = 0
    fma = 0.0
    piv = []
    for o in outs:
        st = o[2].cpu().numpy()
        ss = o[3].cpu().numpy()
        not_opt += int((st != 0).sum())
        fma += float(ss[:, 7].astype(np.float64).sum()) * 1000.0          # FP64-pipe instructions executed (kernel-counted)
        piv.append(ss[:, 1])
        solve_stats.append(ss)
    piv = np.concatenate(piv)
    fma *= K / max(1, len(outs))                       # statistics were sampled on a subset of the steps
    value = world * B * K / (total_ms * 1e-3)

    # ---- e2e: host buffers through hmpc_mpc_step_host_f64 (numpy in, numpy out, copies inside the timed region)
    plan = cabi.StepPlan(fleet.batch.dims, cabi.default_opts(force_general=0 if use_dp else 1))
    hmats = dict(wl0["mats"])
    hmats["C"] = np.ones((1, 1, 1))
    def run_e2e(recondense):
        times, devt = [], []
        for s in range(W + K):
            inp = steps_in[s % P]
            flush.fill_(float(s))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            v_h, obj_h, st_h, stats_h, tm = plan.step(hmats, inp["host"]["x0"], inp["host"]["omega"], inp["host_cost"],
                                                      fleet.batch.lb_v, fleet.batch.ub_v, fleet.batch.is_bin_v,
                                                      recondense=bool(recondense or s == 0))
            u0 = v_h[:, 0].sum()  # the step's result is read on the host
            dt = time.perf_counter() - t0
            if s >= W:
                times.append(dt)
                devt.append(tm)
        total = float(sum(times))
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, devt, obj_h
    e2e_total_rc, e2e_dev_rc, _ = run_e2e(True)
    e2e_total, e2e_dev, obj_h = run_e2e(args.recondense)
    h2d, d2h = plan.bytes_per_step(bool(args.recondense))
    h2d_rc, _ = plan.bytes_per_step(True)
    e2e_value = world * B * K / e2e_total
    # parity spot check inside the bench: e2e path and device path agree on the last step
    assert np.allclose(obj_h, last_obj.cpu().numpy(), rtol=1e-9, atol=1e-12), "host and device paths disagree"
    plan.close()
    extra = {}
    if not args.no_extra_configs:
        try:
            extra = extra_configs(world, rank, dev)
        except Exception as exc:          # (never let the side measurements break the headline line)
            extra = {"extra_configs_error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_peak = cabi.fp64_peak_tflops()
    tab_bytes = 8 if dp_opts.table_fp64 else 4
    solve_ms = kernel_ms["solve"] / K
    cond_ms = kernel_ms["condense"] / K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cond_bytes = 4 * 8 * 0  # placeholder, replaced below
    d = fleet.batch.dims
    # bytes K1 writes per launch in this pipeline (the four constraint matrices) -- algorithmic, see DESIGN.md
    cond_bytes = 8 * B * (d.nc * Nt) * (d.nx + d.nv * Nt + d.nomega * Nt + 1)
    achieved_tf = (2.0 * fma / K) / (solve_ms * 1e-3) / 1e12 if solve_ms > 0 else 0.0
    D_search = 5                                   # depth of one search expansion for one binary per step
    table_bytes = B * ((Nt - 1) // D_search) * int(dp_opts.cells) * tab_bytes if use_dp else None
    allst = np.concatenate(solve_stats, axis=0).astype(np.float64) if solve_stats else np.zeros((1, 8))
    phase_us = [[float(allst[:, c].mean() / 10.0), float(allst[:, c].max() / 10.0)] for c in (1, 2, 4)]
    # FP64-pipe utilisation of the sweep phase alone, on the SMs that have a CTA: sweep instructions of one agent over
    # its own sweep time against one SM's share of the peak
    sweep_frac = None
    if use_dp and phase_us[1][0] > 0 and fp64_peak:
        per_agent_ins = (fma / K) / B
        sweep_frac = (2.0 * per_agent_ins / (phase_us[1][0] * 1e-6) / 1e12) / (fp64_peak / 148.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "run_config": dict(solver="stage_dp" if use_dp else "bnc", exchange=peer_note if peer_ex is not None or world == 1 else
                           (peer_note if "NCCL" in peer_note else "NCCL all-reduce on a side stream (a rank could not map the peers' windows)"),
                       K1="inside every step (--recondense)" if args.recondense else
                          "once, before the timed region: the models do not change (reference: mld_evolution_matrices.py:79); "
                          "ms_per_step_with_recondense adds the measured K1 launch",
                       launch="one CUDA graph per step" if graph is not None else "kernel by kernel",
                       timing="one event pair around the K steps minus K L2 flushes (%.4f ms each, calibrated before the "
                              "timed region; the flush is the first node of each step's graph); the NCCL exchange of "
                              "step s runs on a side stream, pipelined with step s+1, joined before the end"
                              % flush_one_ms,
                       solver_stats="sampled every 8th step"),
        "latency_p50_ms": float(np.median(step_ms)), "latency_max_ms": float(np.max(step_ms)),
        "ms_per_step_by_rank": rank_ms, "sm_mhz_by_rank_after_timed_region": rank_mhz,
        "host_ms_per_step": {"loop": 1e3 * t_host_loop / K, "graph_launch": 1e3 * host_us[0] / K,
                             "exchange_enqueue": 1e3 * host_us[1] / K},
        "ms_per_step_with_recondense": total_ms / K + (0.0 if args.recondense else cond_ms),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_total / K, "api": "hmpc_mpc_step_host_f64 via cabi.StepPlan.step",
                "recondense": bool(args.recondense),
                "with_recondense_every_step": {"value": world * B * K / e2e_total_rc, "ms_per_step": 1e3 * e2e_total_rc / K,
                                               "h2d_bytes_per_step": int(h2d_rc)},
                "device_ms_per_step": dict(zip(("h2d", "kernels", "d2h", "total"),
                                               [float(x) for x in np.mean(np.array(e2e_dev), axis=0)]))},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": dict({k: v / K for k, v in kernel_ms.items()},
                                   note="untimed eager pass with events between the launches (%d steps)" % Kb),
        "roofline": ({"kernel": "stage_dp_table_kernel (+ stage_dp_search_kernel)", "bound": "fp64",
                      "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": achieved_tf / fp64_peak if fp64_peak else None,
                      # DRAM bytes of ONE launch from the round's `ncu --set full` capture of this very configuration
                      # (profiles/r2_final_table_ncu_full_summary.csv: 386,560 B read + 308,992 B written); null elsewhere
                      "traffic": (695552 if (B, N_p, int(dp_opts.cells)) == (100, 48, 4096) else None),
                      "traffic_source": "profiles/r2_final_table_ncu_full_summary.csv (dram__bytes_read.sum + "
                                        "dram__bytes_write.sum, one launch; the table itself stays in L2)",
                      "table_bytes_leaving_sm": table_bytes,
                      "peak_source": "hmpc_fp64_peak_probe (DFMA micro-benchmark, measured in this run; "
                                                              "not in MEASURED_PEAKS.json)",
                      "share_of_step": solve_ms / (total_ms / K),
                      "fp64_pipe_instructions_per_launch": fma / K,
                      "sweep_phase": {"us_mean": phase_us[1][0], "us_max": phase_us[1][1],
                                      "frac_of_peak_while_sweeping": sweep_frac},
                      "phases_us_mean_max": {"setup": phase_us[0], "sweep": phase_us[1], "search": phase_us[2]},
                      "hbm_write_gbs": table_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else None,
                      "note": "achieved = 2 x FP64-pipe instructions the kernels EXECUTE (DFMA, DADD, DMUL, DSETP each "
                              "occupy one issue slot of the FP64 pipe; counted per path by the kernel: 4 per cell where no "
                              "row is violated, 5 + 3 rows elsewhere on the DEWH path, + the search's) / the solve launch's "
                              "CUDA-event duration, i.e. the fraction is FP64-pipe utilisation over the whole launch, "
                              "set-up and search tail included, on the SMs' aggregate peak (100 of 148 SMs have a CTA at "
                              "B = 100); `table_bytes_leaving_sm` = value-table bytes bulk-stored per launch (only the stages "
                              "the search can read: k = D, 2D, ...); they stay in L2, hence the small DRAM `traffic`"} if use_dp else
                     {"kernel": "milp_bnc_kernel", "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak,
                      "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": None,
                      "peak_source": "hmpc_fp64_peak_probe (DFMA micro-benchmark, measured in this run)",
                      "share_of_step": solve_ms / (total_ms / K),
                      "note": "latency-bound tree search: algorithmic FMAs (pivots, row transforms) counted by the "
                              "kernel itself; SURVEY 8(d) names FP64 pipe / latency, not HBM, as the bound"}),
        "roofline_condense": {"kernel": "condense_kernel", "bound": "hbm", "achieved": cond_bytes / (cond_ms * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s",
                              "frac": cond_bytes / (cond_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                              "bytes_per_launch": cond_bytes},
        "solver": ({"kernel": "stage_dp", "cells": int(dp_opts.cells), "table": "fp64" if dp_opts.table_fp64 else "fp32",
                    "not_optimal": not_opt,
                    "nodes_mean": float(np.concatenate([x[:, 0] for x in solve_stats]).mean()),
                    "nodes_max": int(np.concatenate([x[:, 0] for x in solve_stats]).max())} if use_dp else
                   {"kernel": "bnc", "not_optimal": not_opt, "pivots_mean": float(piv.mean()),
                    "pivots_p50": float(np.median(piv)), "pivots_max": int(piv.max())}),
        "wall_s_timed_region": t_wall,
        "clocks": sampler.summary(),
    }
    line.update(extra)
    # K1 at a batch that fills the GPU (the bench batch of 100 agents writes 15 MB: launch/latency-bound)
    try:
        Bl = 8192
        dl = cabi.make_dims(Bl, Nt, nx=1, nu=1, nmu=2, nomega=1, ny=1, nc=2)
        reps_ = (Bl + B - 1) // B
        big = {k: (v.repeat((reps_, 1, 1))[:Bl].contiguous() if v.shape[0] == B else v) for k, v in fleet.batch.mats.items()}
        want = ("H_x", "H_v", "H_omega", "H_5")
        evo_l = cabi.condense(dl, big, want=want)
        best = 1e30
        for _ in range(5):
            flush.fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); cabi.condense(dl, big, want=want, out=evo_l); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        nbytes = sum(evo_l[k].numel() * 8 for k in want)
        line["roofline_condense_large_batch"] = {
            "kernel": "condense_kernel", "bound": "hbm", "agents": Bl, "bytes_per_launch": nbytes, "ms": best,
            "achieved": nbytes / best / 1e6, "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / best / 1e6 / hbm_peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}
        del evo_l, big
    except Exception as exc:      # never let the side measurement break the bench line
        line["roofline_condense_large_batch"] = {"error": str(exc)}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = max(cores, args.cpu_sample)
        times = run_cpu(args, lambda s: cpu_jobs(args, s, count=sample), 1, 0, cores)
        line["cpu_baseline"] = {"value": sample / times[0], "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d agent-solves of the same workload (numpy condensing + HiGHS 1.12 via "
                                          "scipy.optimize.milp, gap 0) on a %d-process pool" % (sample, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
