"""Scratch: raw pinned host->device copy bandwidth for the e2e step's payload (542 MB), alone and with a concurrent D2H."""
import torch
n = 542474240 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device='cuda')
o = torch.empty(99680256 // 4, dtype=torch.float32, device='cuda')
ho = torch.empty(99680256 // 4, dtype=torch.float32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s1)
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                ho.copy_(o, non_blocking=True)
    e1.record(s1)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'H2D 542 MB{" + concurrent D2H 100 MB" if both else ""}: {ms:.2f} ms = {542474240 / ms / 1e6:.1f} GB/s '
          f'-> PCIe ceiling of the e2e metric: {4 / ms * 1e3:.0f} frames/s')
