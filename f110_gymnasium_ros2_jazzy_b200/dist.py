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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index, pci_bus_id=None, sysfs='/sys'):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it
    allocates afterwards (first touch) and the threads that fill them sit next to that GPU's PCIe root.  With one
    process per GPU on a two-socket box this keeps the per-step observation downloads off the socket interconnect.
    Returns the previous affinity (hand it to os.sched_setaffinity to undo) or None when nothing was changed
    (single node, no sysfs entry, affinity already narrower)."""
    import os
    try:
        if pci_bus_id is None:
            p = torch.cuda.get_device_properties(device_index)
            pci_bus_id = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(os.path.join(sysfs, 'bus/pci/devices', pci_bus_id.lower(), 'numa_node')).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(os.path.join(sysfs, 'devices/system/node/node%d/cpulist' % node)).read())
        before = os.sched_getaffinity(0)
        want = cpus & before
        if not want or want == before:
            return None
        os.sched_setaffinity(0, want)
        return before
    except (OSError, ValueError, AttributeError):
        return None


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
