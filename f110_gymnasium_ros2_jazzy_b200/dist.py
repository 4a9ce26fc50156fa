"""Multi-GPU: envs are independent, so they shard by index with NO collective on the step path
(SURVEY 8e).  The only exchange is a small all-reduce of episode statistics, issued off the step path.
"""
import torch
import torch.distributed as dist


def shard_range(num_envs, rank, world_size):
    """Contiguous env-index partition: rank g owns [g*N/G, (g+1)*N/G)."""
    lo = (num_envs * rank) // world_size
    hi = (num_envs * (rank + 1)) // world_size
    return lo, hi


class EpisodeStats(object):
    """Sum-reduces the per-rank episode counters (BatchSim.stats) across ranks.

    Works with any initialised process group (nccl on GPUs, gloo on CPU tensors for tests).
    """
    NAMES = ('episodes', 'episode_steps', 'ego_collisions', 'laps_done', 'episode_time', 'r5', 'r6', 'r7')

    def __init__(self, group=None):
        self.group = group

    def reduce(self, local):
        """local: tensor[F110_NUM_STATS] (device tensor under nccl). Returns the global sums as a dict."""
        t = local.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        v = t.detach().cpu().tolist()
        out = dict(zip(self.NAMES, v))
        n = max(out['episodes'], 1.0)
        out['mean_episode_steps'] = out['episode_steps'] / n
        out['mean_episode_time'] = out['episode_time'] / n
        return out
