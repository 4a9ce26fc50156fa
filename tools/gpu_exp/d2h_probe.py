"""Concurrent bare device->host copy bandwidth, one process per GPU (torchrun), no kernels: the fabric ceiling the e2e path
of bench.py lives under.  Every rank copies an observation-sized device tensor (4096 x 1088 f32 = 17.8 MB by default) into
pinned host memory over and over; all ranks start together (barrier), each reports its own GB/s, rank 0 prints min / mean
per rank and the aggregate.  Also the host->device direction and both at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/gpu_exp/d2h_probe.py
"""
import json, os, sys, time
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
bind = '--bind' in sys.argv
if bind:
    from f110_gymnasium_ros2_jazzy_b200.dist import bind_host_to_gpu
    bind_host_to_gpu(local)
n = 4096 * 1088
dev = torch.zeros(n, dtype=torch.float32, device='cuda')
host = torch.zeros(n, dtype=torch.float32, pin_memory=True)
host2 = torch.zeros(n, dtype=torch.float32, pin_memory=True)
dev2 = torch.zeros(n, dtype=torch.float32, device='cuda')
s2 = torch.cuda.Stream()


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return iters * n * 4 / (time.perf_counter() - t0) / 1e9


def d2h():
    host.copy_(dev, non_blocking=True)


def h2d():
    dev2.copy_(host2, non_blocking=True)


def both():
    host.copy_(dev, non_blocking=True)
    with torch.cuda.stream(s2):
        dev2.copy_(host2, non_blocking=True)


def sync_each():            # one copy, then wait: the shape of a synchronous env step
    host.copy_(dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()


res = {}
for name, fn in (('d2h', d2h), ('h2d', h2d), ('both_d2h_plus_h2d', both), ('d2h_sync_each_copy', sync_each)):
    g = torch.tensor([timed(fn)], dtype=torch.float64, device='cuda')
    if world > 1:
        allg = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(allg, g)
        v = [float(x) for x in allg]
    else:
        v = [float(g)]
    res[name] = {'per_rank_gbs': [round(x, 1) for x in v], 'min': round(min(v), 1), 'aggregate': round(sum(v) * (2 if 'both' in name else 1), 1)}
if rank == 0:
    print(json.dumps({'ranks': world, 'bytes_per_copy': n * 4, 'bound_to_numa': bind, 'results': res}))
if world > 1:
    dist.destroy_process_group()
