"""the box's device-to-host ceiling: plain cudaMemcpyAsync (torch copy_) from every rank into page-locked host memory,
one copy per buffer; run alone (1 GPU) or under torchrun (N ranks at the same time)"""
import os, sys, time
import torch
rank = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
def numa_info():
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        b = bus[4:] if len(bus.split(':')[0]) == 8 else bus
        node = open('/sys/bus/pci/devices/%s/numa_node' % b).read().strip()
        speed = open('/sys/bus/pci/devices/%s/current_link_speed' % b).read().strip()
        width = open('/sys/bus/pci/devices/%s/current_link_width' % b).read().strip()
        return bus, node, speed, width
    except Exception as e:
        return repr(e)
print('rank', rank, 'pci/numa/link', numa_info(), 'cpus allowed', len(os.sched_getaffinity(0)), flush=True)
for mb in (8, 82, 256):
    n = mb*1000*1000//8
    d = torch.rand(n, dtype=torch.float64, device='cuda')
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    for direction in ('d2h', 'h2d'):
        for _ in range(2):
            (h.copy_(d, non_blocking=True) if direction == 'd2h' else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            (h.copy_(d, non_blocking=True) if direction == 'd2h' else d.copy_(h, non_blocking=True))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/5
        print('rank %d  %s %4d MB: %.3f ms = %.1f GB/s' % (rank, direction, mb, ms, mb/1e3/(ms*1e-3)), flush=True)
