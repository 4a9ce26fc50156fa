"""Which NUMA node each GPU hangs off, and what bind_host_to_gpu would do (diagnostic for the multi-GPU host path)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from f110_gymnasium_ros2_jazzy_b200.dist import bind_host_to_gpu
print("cpus:", len(os.sched_getaffinity(0)), "nodes:", sorted(d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')))
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bid = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        node = open('/sys/bus/pci/devices/%s/numa_node' % bid).read().strip()
    except OSError as e:
        node = repr(e)
    before = os.sched_getaffinity(0)
    prev = bind_host_to_gpu(i)
    print(i, bid, 'numa_node', node, 'bound to', len(os.sched_getaffinity(0)), 'cpus', 'changed' if prev else 'unchanged')
    os.sched_setaffinity(0, before)
