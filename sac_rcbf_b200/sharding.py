"""Multi-GPU layout of the safety path: instances shard contiguously by rank, one process per GPU, and there is NO
collective on the step path (no term of the QP / dynamics couples two instances).  The only
communication is an optional, once-per-rollout reduction of a handful of statistics (sum reward, sum cost, #done,
#goal_met, solver counters), which rides on whatever torch.distributed backend is initialised (NCCL over NVLink on the
GPU box, gloo in the CPU tests)."""
import torch

STAT_NAMES = ("instances", "sum_reward", "sum_cost", "n_done", "n_goal_met", "qp_nan", "qp_uncertified", "qp_trivial",
              "qp_pass1_iters", "qp_fallback")


def shard_range(n_total, rank, world_size):
    """Contiguous [lo, hi) slice of `n_total` instances owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_rollout_stats(reward, cost, done, goal_met=None, counters=None):
    """Pack this rank's statistics of one step/rollout into a float64 vector (layout: STAT_NAMES)."""
    dev = reward.device
    c = torch.zeros(8, dtype=torch.float64, device=dev) if counters is None else counters[:8].to(torch.float64)
    gm = torch.zeros((), dtype=torch.float64, device=dev) if goal_met is None else goal_met.sum().to(torch.float64)
    return torch.stack([torch.tensor(float(reward.numel()), dtype=torch.float64, device=dev),
                        reward.sum().to(torch.float64), cost.sum().to(torch.float64), done.sum().to(torch.float64), gm,
                        c[0], c[1], c[3], c[4], c[5]])


def reduce_rollout_stats(local_stats, group=None):
    """Sum the statistics vector over all ranks (all_reduce).  Returns a dict keyed by STAT_NAMES.
    With no process group initialised it is the identity (single GPU)."""
    import torch.distributed as dist

    v = local_stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, v.tolist()))
