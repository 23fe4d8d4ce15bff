"""Multi-GPU plumbing: agents (and agent x scenario pairs) are independent, so the batch is sharded in contiguous
blocks, one process per GPU, with NO data-path collective during condense / solve / sim.  The only exchange per
control step is the aggregate power trajectory the grid agent needs
(reference: examples/.../micro_grid_agents.py:625-646 stacks every device's power; the grid model only uses the
sum, micro_grid_models.py:143): an all-reduce of [Nt] FP64 values (392 B at N_p = 48) on the solve stream, or an
all-gather of the per-agent trajectories when the caller wants them.  Works with NCCL (GPU) and gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(num_agents, rank, world_size):
    """Contiguous block [lo, hi) of agents owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(num_agents), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_aggregate(p_local):
    """Sum of the per-rank aggregate power trajectories [Nt]; in place, asynchronous on the current stream."""
    if is_distributed():
        dist.all_reduce(p_local, op=dist.ReduceOp.SUM)
    return p_local


def allreduce_max_int(value, device=None):
    """Largest ``value`` over the ranks (every rank gets it): loop bounds that must agree wherever a collective sits
    inside the loop."""
    if not is_distributed():
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def response_blocks(num_local, groups):
    """Blocks of one best-response pass: ``groups`` (the same on every rank) spans [lo, hi) of the local agents.  A
    rank with fewer agents than blocks gets EMPTY spans -- it still has to take part in the block's all-reduce, so
    the spans are yielded, not skipped."""
    return [shard_range(num_local, g, groups) for g in range(int(groups))]


class PeerExchange(object):
    """The per-step exchange of the [Nt] aggregate power WITHOUT a collective call: every rank owns a window in
    symmetric memory (peer-mapped over NVLink), `publish` is the last pass of the local reduction storing this rank's
    sums into every rank's window, `gather` adds the world's contributions in rank order (csrc/aggregate.cu).  Both are
    ordinary kernels on the current stream, so the whole control step -- exchange included -- is one CUDA graph and the
    host issues nothing per step (the NCCL all-reduce cannot be captured on this stack: tools/gpu_nccl_graph_probe.py).
    `windows` lets a test emulate several ranks on one GPU."""

    def __init__(self, Nt, device, world=None, rank=None, windows=None):
        from . import cabi
        self.Nt = int(Nt)
        self.world = int(world) if world is not None else (dist.get_world_size() if is_distributed() else 1)
        self.rank = int(rank) if rank is not None else (dist.get_rank() if is_distributed() else 0)
        n = cabi.aggregate_window_doubles(self.Nt, self.world)
        if windows is not None:                         # emulation: all windows live on this device
            self.window = windows[self.rank]
            ptrs = [w.data_ptr() for w in windows]
        elif self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.window = symm.empty(n, dtype=torch.float64, device=device)
            self.window.zero_()
            self._hdl = symm.rendezvous(self.window, dist.group.WORLD)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier()
        else:
            self.window = torch.zeros(n, dtype=torch.float64, device=device)
            ptrs = [self.window.data_ptr()]
        self.windows_dev = torch.tensor(ptrs, dtype=torch.int64, device=device)

    @staticmethod
    def new_window(Nt, world, device):
        from . import cabi
        return torch.zeros(cabi.aggregate_window_doubles(Nt, world), dtype=torch.float64, device=device)

    def publish(self, u, P_nom, out_prev=None, lag=1):
        """out_prev [Nt]: the same launch also writes the world's sum of the step `lag` publishes back there; the ranks
        can then drift up to `lag` steps apart without anybody waiting (lag <= 6: the windows hold a ring of 8 steps)"""
        from . import cabi
        self._partial = cabi.aggregate_publish(u, P_nom, self.world, self.rank, self.windows_dev, out_prev=out_prev,
                                               lag=lag)

    def gather(self, out=None, spin_limit=0, lag=0):
        """sum over the ranks of the step published `lag` steps ago (0 = the latest; zeros while nothing that old
        exists).  A pipelined loop uses lag = 1: the peers' contributions of the previous step have long arrived."""
        from . import cabi
        return cabi.aggregate_gather(self.Nt, self.world, self.rank, self.window, out=out, spin_limit=spin_limit, lag=lag)

    def error(self):
        """step number at which a peer never arrived (0 = none)"""
        return int(self.window[:2].view(torch.int64)[1].item())


def allgather_trajectories(traj_local, counts=None):
    """[B_local, Nt] per rank -> [B_total, Nt] on every rank, rank-major (the order the reference stacks devices)."""
    if not is_distributed():
        return traj_local
    world = dist.get_world_size()
    if counts is None:
        out = [torch.empty_like(traj_local) for _ in range(world)]
        dist.all_gather(out, traj_local.contiguous())
        return torch.cat(out, dim=0)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(traj_local.shape[1:]), dtype=traj_local.dtype, device=traj_local.device)
    pad[:traj_local.shape[0]] = traj_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def grid_evaluate(p_devices_sum, p_pv, p_res):
    """Grid agent closed form on the gathered aggregate (reference grid MLD, micro_grid_models.py:145-168):
    y = sum of device powers, delta = [y >= 0], z = delta*y (import), export = y - z."""
    y = p_devices_sum + p_pv + p_res
    delta = (y >= 0).to(y.dtype)
    z = delta * y
    return dict(y=y, delta=delta, p_imp=z, p_exp=y - z)
