"""Aggregate pinned host->device bandwidth with N ranks copying at once (torchrun): what bounds the e2e metric at 8 GPUs.
Variants: torch pin_memory() buffers vs cudaHostAlloc(WriteCombined) buffers, with / without binding to the GPU's NUMA node."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, lr, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
NB = 542474240
d = torch.empty(NB // 4, dtype=torch.float32, device=dev)

def wc_buffer(nbytes):
    from cuda.bindings import runtime as cudart
    err, ptr = cudart.cudaHostAlloc(nbytes, cudart.cudaHostAllocWriteCombined | cudart.cudaHostAllocPortable)
    assert int(err) == 0, err
    buf = (ctypes.c_byte * nbytes).from_address(int(ptr))
    return torch.frombuffer(buf, dtype=torch.float32)

def measure(h, tag):
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'N={world} {tag}: {float(t):.2f} ms per 542 MB copy per rank = {NB / float(t) / 1e6:.1f} GB/s per GPU, '
              f'{world * NB / float(t) / 1e6:.0f} GB/s aggregate -> e2e ceiling {world * 4 / float(t) * 1e3:.0f} frames/s', flush=True)

measure(torch.empty(NB // 4, dtype=torch.float32).pin_memory(), 'pin_memory, no affinity')
import bench
cpus = bench.pin_to_gpu_numa_node(lr)
measure(torch.empty(NB // 4, dtype=torch.float32).pin_memory(), f'pin_memory, bound to {len(cpus) if cpus else "?"} cores next to the GPU')
try:
    h = wc_buffer(NB); h.zero_()
    measure(h, 'cudaHostAlloc(WriteCombined), bound')
except Exception as exc:
    if rank == 0: print('write-combined buffer failed:', exc)
if rank == 0:
    os.system('nvidia-smi topo -m | head -14')
dist.destroy_process_group()
