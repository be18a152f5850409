"""Probe: does torch symmetric memory (peer pointers, NVLS multicast) work on this box?  torchrun --nproc-per-node 2"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ['RANK']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'world', h.world_size, 'ptrs', [hex(p) for p in h.buffer_ptrs], 'mc', hex(h.multicast_ptr) if h.multicast_ptr else None,
      'has_mc', getattr(h, 'has_multicast_support', None), flush=True)
t.fill_(rank + 1.0)
h.barrier(channel=0)
peer = h.get_buffer((rank + 1) % h.world_size, (1 << 20,), torch.float32)
print(rank, 'peer value', float(peer[0]), flush=True)
h.barrier(channel=0)
dist.destroy_process_group()
