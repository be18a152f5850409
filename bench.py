#!/usr/bin/env python
"""Benchmark of the aggregation hot path (BASELINE.json: aggregation frames/s and % of roofline at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference ...                     # the reference algorithm on the box's host cores

One "step" = one pass of the hot path over one batch of synthetic frames:
    projection table (1 launch) + weight re-layout (3) + tap records (1) + coverage / row lists / unit table (3) + texel lists
    of the quads (1) + image-plane 3xTF32 GEMM over the covered rows (1) + pooling from the lists with bias / ReLU / view-and-
    scale sum (1) + its completion pass (1)   [--flags 32: one fused grid-side kernel instead of the last seven].
Default workload: MultiviewC-shaped (7 views, 1280x720 source, stride-8/16/32 maps 90x160 / 45x80 / 23x40, C = 256,
156x156x5 voxel grid = the shipped config-of-record of "37.5 m x 37.5 m"), B frames per GPU.
Multi-GPU = batch data parallel (frames are independent: no data-path collective, weak scaling).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
API with pinned-host inputs and a device->host read of the result inside the timed region; `roofline` describes the
dominant kernel (pool_list_kernel; the GEMM under `second_kernel`, the whole step under `step`); `cpu_baseline` = the oracle's torch-CPU port of the reference timed on the host cores (N = 1 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = 'aggregation_frames_per_s'
UNIT = 'frames/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--workload', choices=['MultiviewC', 'MultiviewX', 'Wildtrack', 'MultiviewC-37.5m'], default='MultiviewC',
                    help='dataset geometry (config-of-record); MultiviewC-37.5m = the literal 150 x 150 cell reading')
    ap.add_argument('--batch', type=int, default=4, help='frames per GPU per step')
    ap.add_argument('--flags', type=int, default=0, help='vfa_aggregate_fwd flags (1 = force SIMT path)')
    ap.add_argument('--cpu-views', type=int, default=2, help='views of one frame timed for cpu_baseline')
    ap.add_argument('--ref-views', type=int, default=1, help='views of one frame per step of --impl reference')
    ap.add_argument('--mode', choices=['dp', 'slab', 'views', 'train'], default='dp',
                    help='dp: frames sharded over GPUs, no data-path collective (weak scaling, default); slab: BEV row '
                         'slabs over GPUs, features broadcast from rank 0 + output all-gather per step (strong scaling); '
                         'views: cameras sharded over GPUs (features stay on the rank that owns the camera), partial BEV '
                         'maps all-reduced, frame chunks pipelined (strong scaling); '
                         'train: full detector training step, DDP over GPUs (BASELINE config 5; --batch frames per GPU)')
    ap.add_argument('--view-chunk', type=int, default=0,
                    help='--mode views: frames per all-reduce chunk (0 = the whole batch in one call; smaller chunks overlap '
                         'the all-reduce with the next chunk at the price of per-call table / weight / record launches)')
    ap.add_argument('--backward', action='store_true', help='time forward + backward (BASELINE config 4)')
    ap.add_argument('--features', choices=['f32', 'bf16'], default='f32',
                    help='feature-map storage: f32 (parity path, default) or bf16 (half the gather bytes, stated tolerance)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-variants', action='store_true', help='skip the static-camera and bf16-MMA lines of the same workload')
    ap.add_argument('--no-configs', action='store_true', help='skip the MultiviewX / Wildtrack B = 1 lines (BASELINE configs 2-3)')
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the strong-scaling (camera-sharded, fused) block')
    ap.add_argument('--no-config4', action='store_true', help='skip the batch-64 forward+backward block (BASELINE config 4)')
    ap.add_argument('--no-config5', action='store_true', help='skip the whole-detector training-step block (BASELINE config 5)')
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    tflops_burst=p['bf16_tflops'], source='measured (MEASURED_PEAKS.json: copy bandwidth; bf16 sustained for '
                    'the step timed in a long loop, bf16 burst for kernels timed in isolation)')
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source='fallback (B200_PROFILING.md)')


def workload_numbers(geom, batch, esize=4):
    """Algorithmic bytes / flops of one step on one GPU."""
    L, W = geom.grid_shape
    C, nl, V = geom.channels, geom.n_layers, geom.n_views
    px = sum(h * w for h, w in geom.feature_sizes())
    feat_bytes = batch * V * px * C * esize
    out_bytes = batch * C * L * W * 4
    const_bytes = 3 * (C * C * nl + C) * 4 + V * nl * L * W * 16 + L * W * 12 + V * 48
    flops = 2.0 * L * W * (C * nl) * C * 3 * V * batch          # collapse contraction, grid-side (as the reference)
    # feature-side formulation (vfa_fwd_fside.cu): contraction on the image plane, Y = per-layer products in HBM
    fside_flops = 2.0 * batch * V * px * C * (C * nl)
    y_bytes = batch * V * nl * px * C * 4
    rec_bytes = V * len(geom.feature_sizes()) * nl * L * W * 32
    y_frame = V * nl * px * C * 4
    chunk = min(batch, max(1, (6 << 30) // y_frame))            # fside_chunk_frames()
    return dict(bytes=feat_bytes + out_bytes + const_bytes, flops=flops, feat_bytes=feat_bytes, out_bytes=out_bytes,
                fside_flops=fside_flops, y_bytes=y_bytes, rec_bytes=rec_bytes, fside_chunks=-(-batch // chunk))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML, 5 ms period; falls back to
    polling nvidia-smi when the NVML binding is unavailable)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.max_mhz, self._stop_evt = index, [], 0, None, threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                    self.mask |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self._stop_evt.wait(0.005)
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=clocks.sm,clocks.max.sm,'
                                          'clocks_event_reasons.active', '--format=csv,noheader,nounits'],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                    self.sm.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    self.mask |= int(out[2].strip(), 16)
            except Exception:
                self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        reasons = [name for bit, name in self.REASONS.items() if self.mask & bit]
        return {'sm_mhz': statistics.median(self.sm), 'sm_min_mhz': min(self.sm), 'sm_max_mhz': self.max_mhz,
                'reasons': reasons, 'samples': len(self.sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def cpu_port_time(geom, views, threads_note=True):
    """Seconds for `views` views x 3 scales of ONE frame through the oracle's torch-CPU port of the reference
    (fp32, no_grad, boxes re-derived per call like the reference)."""
    from oracle import ref_port
    from vfa_b200 import geometry, synthetic
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    grid = geometry.grid_for(geom)
    calibs = synthetic.ring_calibs(geom)[:views]
    feats = synthetic.features(geom, batch=1, n_views=views, seed=0)
    params = synthetic.collapse_params(geom, seed=0)
    with torch.no_grad():
        t0 = time.perf_counter()
        out = ref_port.aggregate(feats, calibs, grid, params, geom.grid_height, geom.cube_size, geom.name,
                                 geom.image_size, cache_boxes=False)
        dt = time.perf_counter() - t0
    return dt, float(out.mean())


def run_reference(args):
    """`--impl reference`: the reference algorithm (oracle port: same torch operator sequence as reference
    vfa_op.py:61-125 looped as vfanet.py:64-82) on the host cores; rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from vfa_b200 import geometry
    geom = geometry.BENCH_WORKLOADS[args.workload]
    views = max(1, min(args.ref_views, geom.n_views))
    for _ in range(args.warmup):
        cpu_port_time(geom, views)
    times = [cpu_port_time(geom, views)[0] for _ in range(args.steps)]
    per_step = sum(times) / len(times)
    fps = 1.0 / (per_step * geom.n_views / views)
    sample = (f'{views} of {geom.n_views} views x 3 scales of one {args.workload}-shaped frame per step, fp32 torch-CPU '
              f'port of the reference operator sequence; frames/s = 1 / (step_s * {geom.n_views}/{views})')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the same workload as the GPU arm's line (the reference processes the frames of a batch one after the other --
        # vfanet.py:64-82 is batch 1 by construction -- so its frames/s do not depend on the batch size); `sample` says what
        # one timed step covers
        'config': {'workload': f'{args.workload}-shaped aggregation forward', 'batch_per_gpu': args.batch, 'views': geom.n_views,
                   'channels': geom.channels, 'grid': list(geom.grid_shape) + [geom.n_layers],
                   'feature_maps': [list(s_) for s_ in geom.feature_sizes()], 'sample': sample},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample,
                         'host_cpus': os.cpu_count()},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_train(args):
    """`--mode train`: the BASELINE config 5 line on its own (see train_measure)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    line = train_measure(args, world, rank, local_rank, dev, args.batch, args.steps, args.warmup, breakdown=True)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_measure(args, world, rank, local_rank, dev, B, steps, warmup, breakdown):
    """BASELINE config 5: one optimiser step of the whole detector -- GroupNorm ResNet-18 trunk, batched
    laterals, the fused aggregation (forward + backward kernels of this repo), `fuse` and the four heads, a stand-in
    detection loss, SGD -- on synthetic camera images, data parallel over the GPUs of one box (DistributedDataParallel:
    gradients of every parameter, the collapse weights included, all-reduced over NCCL during the backward).  The
    backbone, heads and loss are stock PyTorch / cuDNN (torch default precision settings, as the reference trains);
    they are outside the hand-written path and are here so the aggregation is measured inside its real consumer.
    Needs an initialised process group when world > 1; returns the JSON line as a dict (every rank)."""
    import torch.nn.functional as F
    import vfa_b200
    from types import SimpleNamespace
    from vfa_b200 import geometry, synthetic
    from vfa_b200.network import VFANet
    geom = geometry.BENCH_WORKLOADS[args.workload]
    V = geom.n_views
    H, W = geom.resize_size
    torch.manual_seed(0)                                  # identical initial weights on every rank
    net = VFANet(SimpleNamespace(data=geom.name, image_size=geom.image_size), 'resnet18', geom.grid_height,
                 geom.cube_size, 360, '3D', False, flags=args.flags).to(dev).to(memory_format=torch.channels_last)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank]) if world > 1 else net
    opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)   # reference train.py:256
    grid = geometry.grid_for(geom)[None].to(dev)
    calibs = synthetic.ring_calibs(geom).to(dev)
    L, Wg = grid.shape[1:3]
    gen = torch.Generator().manual_seed(77 + rank)        # every rank trains on its own frames
    host_images = torch.rand(B * V, 3, H, W, generator=gen).pin_memory()
    images = host_images.to(dev)
    tgt = {'heatmap': (torch.rand(B, 1, L, Wg, generator=gen) > 0.98).float().to(dev),
           'loc_offset': torch.rand(B, L, Wg, 2, generator=gen).to(dev),
           'dim_offset': torch.randn(B, L, Wg, 3, generator=gen).to(dev),
           'rotation': torch.randint(0, 360, (B, L, Wg), generator=gen).to(dev)}

    def train_step(x):
        pred = model(x, calibs, grid, batch=B)
        obj = tgt['heatmap'].permute(0, 2, 3, 1)          # regression / orientation terms count on object cells only
        loss = (F.binary_cross_entropy_with_logits(pred['heatmap'], tgt['heatmap'])
                + (obj * (pred['loc_offset'] - tgt['loc_offset']).abs()).mean()
                + (obj * (pred['dim_offset'] - tgt['dim_offset']).abs()).mean()
                + (obj[..., 0] * F.cross_entropy(pred['rotation'].reshape(-1, 360), tgt['rotation'].reshape(-1),
                                                 reduction='none').view(B, L, Wg)).mean())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(3, warmup)):
        loss = train_step(images)
    path = vfa_b200.last_kernel_path()
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms = timed(lambda: train_step(images), steps)
    clocks = sampler.stop()
    value = world * B * steps / (total_ms * 1e-3)

    def e2e_step():
        x = host_images.to(dev, non_blocking=True)        # pinned host -> device, every step
        return float(train_step(x).item())                # the loss read back on the host, every step
    n_e2e = max(3, min(steps, 10))
    e2e_ms = timed(e2e_step, n_e2e)
    final_loss = float(loss.item())
    n_params = sum(p.numel() for p in net.parameters())
    breakdown_ms = None
    if breakdown:
        breakdown_ms = train_breakdown(args, net, images, calibs, grid, B, V, L, Wg, dev, timed, total_ms / steps)

    line = {
        'metric': 'training_frames_per_s', 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
        'warmup': max(3, warmup), 'ms_per_step': total_ms / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32 (aggregation: 3xTF32 tcgen05 + fp32 sums; backbone / heads: cuDNN with torch default settings)',
        'data': 'synthetic',
        'config': {'workload': f'full VFA training step on synthetic {args.workload}-shaped frames (GroupNorm ResNet-18 '
                               f'trunk + laterals + fused aggregation + fuse/heads + stand-in loss + SGD)',
                   'batch_per_gpu': B, 'views': V, 'image': [H, W], 'grid': [L, Wg, geom.n_layers], 'parameters': n_params,
                   'parallelism': f'ddp{world}', 'kernel_path': path,
                   'l2': f'images {host_images.numel() * 4 / 1e6:.0f} MB/step/GPU and every activation exceed the 126 MB L2'},
        'clocks': clocks,
        'e2e': {'value': world * B * n_e2e / (e2e_ms * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': host_images.numel() * 4, 'd2h_bytes_per_step': 4, 'steps': n_e2e,
                'note': 'pinned-host fp32 images copied H2D and the loss read back with .item() every step'},
        'breakdown_ms': breakdown_ms,
        'final_loss': final_loss,
        # this repo's kernels per step: forward 10 (table, weight re-layout, tap records, chunk lists + tile order, row lists,
        # unit table, compacted GEMM, pooling + completion pass) + backward >= 16 (3 transposed weight re-layouts, tap records, tile needs, 5 CSR
        # launches, mask_grad, dy_gather, overflow, dFeature GEMM, dWeight GEMM, layout); cuDNN / ATen launches not counted
        'gpu_launches': 26 * steps,
        'roofline': None, 'cpu_baseline': None,
    }
    return line


def train_breakdown(args, net, images, calibs, grid, B, V, L, Wg, dev, timed, step_ms):
    """Where the training step goes: forward of the whole network, and of the aggregation stage alone, on the same inputs."""
    import vfa_b200
    net.eval()
    with torch.no_grad():
        fwd_ms = timed(lambda: net(images, calibs, grid, batch=B), 5) / 5
        x = (images - net.mean.view(3, 1, 1)) / net.std.view(3, 1, 1)
        f8, f16, f32_ = net.base(x.contiguous(memory_format=torch.channels_last))
        from vfa_b200 import vfanet
        lats = vfanet.lateral_features(net, f8, f16, f32_)
        feats = [t.reshape(B, V, *t.shape[1:]) for t in lats]
        table = vfa_b200.build_table(net.vfa8.geometry((L, Wg)), calibs, grid)
        wts, bs = [m.collapse.weight for m in (net.vfa8, net.vfa16, net.vfa32)], [m.collapse.bias for m in (net.vfa8, net.vfa16, net.vfa32)]
        agg_ms = timed(lambda: vfa_b200.aggregate(feats, table, wts, bs, flags=args.flags), 5) / 5
    net.train()
    with torch.enable_grad():
        fg = [t.detach().requires_grad_(True) for t in feats]
        gout = torch.randn(B, 256, L, Wg, device=dev)

        def agg_fb():
            vfa_b200.aggregate(fg, table, wts, bs, flags=args.flags).backward(gout)
        agg_fb()
        agg_fb_ms = timed(agg_fb, 5) / 5
    return {'train_step': step_ms, 'network_forward_eval': fwd_ms,
            'aggregation_forward': agg_ms, 'aggregation_forward_backward': agg_fb_ms,
            'note': 'aggregation = table + fused forward (+ backward kernels) on the lateral maps of this '
                    'batch; the rest of the step is cuDNN convolutions / GroupNorm / optimiser'}


def time_steps(fn, n, barrier, dev, world):
    """n calls of fn between two barriers, CUDA events on the current stream, max over ranks -> total ms."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pin_to_gpu_numa_node(local_rank):
    """Bind this process to the CPU cores next to its GPU before pinned host buffers are allocated (first touch places the
    pages on that NUMA node): the end-to-end path is bound by host memory / PCIe bandwidth when 8 ranks stream at once."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[local_rank]) if vis and vis.split(',')[local_rank].isdigit() else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.mode == 'train':
        if '--batch' not in sys.argv:
            args.batch = 1
        run_train(args)
        return

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f'--gpus {args.gpus} needs a torchrun launch with {args.gpus} ranks (WORLD_SIZE={world})')
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    import vfa_b200
    from vfa_b200 import distributed as vd
    from vfa_b200 import geometry
    from vfa_b200 import synthetic

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def setup(workload, B, v_range=None, features='f32', seed=1234):
        """Device-resident synthetic problem of one rank: geometry, table inputs, collapse parameters, channels-last
        features [B,V,fH,fW,C] (logical [B,V,C,fH,fW]; the layout the lateral convs emit in torch.channels_last)."""
        g = geometry.BENCH_WORKLOADS[workload]
        zs = list(range(0, g.grid_height, g.cube_size[2]))
        grid = geometry.grid_for(g).to(dev)
        calibs = synthetic.ring_calibs(g).to(dev)
        cg = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
        params = synthetic.collapse_params(g, seed=0)
        ws, bs = [w.to(dev) for w, _ in params], [b.to(dev) for _, b in params]
        gen = torch.Generator(device=dev).manual_seed(seed + rank)      # every GPU owns different frames
        v0, v1 = v_range if v_range is not None else (0, g.n_views)
        feats = [torch.randn(B, v1 - v0, h, w, g.channels, generator=gen, device=dev).relu_() for (h, w) in g.feature_sizes()]
        if features == 'bf16':
            feats = [f.to(torch.bfloat16) for f in feats]
        return dict(g=g, zs=zs, grid=grid, calibs=calibs[v0:v1].contiguous(), cgeom=cg, ws=ws, bs=bs, feats=feats, gen=gen)

    geom = geometry.BENCH_WORKLOADS[args.workload]
    B, V, C = args.batch, geom.n_views, geom.channels
    P = setup(args.workload, B, vd.view_bounds(V, world, rank) if args.mode == 'views' else None, args.features)
    zs, grid, calibs, cgeom, weights, biases, feats_cl, gen = (P[k] for k in ('zs', 'grid', 'calibs', 'cgeom', 'ws', 'bs',
                                                                              'feats', 'gen'))
    v0, v1 = vd.view_bounds(V, world, rank) if args.mode == 'views' else (0, V)
    shape = ws = None
    if args.mode != 'views':
        shape = vfa_b200.make_shape(feats_cl, cgeom.n_layers)
        ws = vfa_b200.workspace_for(cgeom, shape, args.flags, dev)
    out = torch.empty(B, C, grid.shape[0], grid.shape[1], device=dev)
    nums = workload_numbers(geom, B, 2 if args.features == 'bf16' else 4)

    zs_geom = lambda lw: vfa_b200.make_geometry(len(zs), geom.cube_size, zs, lw, geom.name, geom.image_size)  # noqa: E731
    if args.backward:
        for t in feats_cl + weights + biases:
            t.requires_grad_(True)
        gout = torch.randn(out.shape, generator=gen, device=dev)
    slab_compute = None
    if args.mode == 'slab':
        def slab_compute(f, c, grid_slab, w, b_):
            table = vfa_b200.build_table(zs_geom(grid_slab.shape[:2]), c, grid_slab)
            if args.backward:
                return vfa_b200.aggregate(f, table, w, b_, flags=args.flags, channels_last=True)
            return vfa_b200.aggregate_forward_raw(f, table, w, b_, args.flags)

    def views_compute(f, c, g_, w, b_):
        table = vfa_b200.build_table(cgeom, c, g_)
        if args.backward:
            return vfa_b200.aggregate(f, table, w, b_, flags=args.flags, channels_last=True)
        return vfa_b200.aggregate_forward_raw(f, table, w, b_, args.flags)

    def step_general(timed_events=None):
        """slab / views mode and/or backward: through the autograd-capable public entry points."""
        if timed_events is not None:
            timed_events[0].record()
        if args.mode == 'views':
            res = vd.aggregate_views(feats_cl, calibs, grid, weights, biases, views_compute, out_channels=C,
                                     frames_per_chunk=args.view_chunk)
        elif args.mode == 'slab':
            vd.broadcast_features([f.detach() for f in feats_cl], src=0)
            res = vd.aggregate_slab(feats_cl, calibs, grid, weights, biases, slab_compute)
        else:
            table = vfa_b200.build_table(cgeom, calibs, grid)
            res = vfa_b200.aggregate(feats_cl, table, weights, biases, flags=args.flags, channels_last=True)
        if args.backward:
            for t in feats_cl + weights + biases:
                t.grad = None
            res.backward(gout)
            if world > 1 and args.mode == 'dp':
                vd.allreduce_collapse_grads(weights + biases)
        if timed_events is not None:
            timed_events[1].record()

    def step(timed_events=None):
        if args.mode in ('slab', 'views') or args.backward:
            return step_general(timed_events)
        table = vfa_b200.build_table(cgeom, calibs, grid)                           # 1 launch
        vfa_b200.prepare_weights(cgeom, shape, weights, args.flags, workspace=ws)    # 1 launch
        if timed_events is not None:
            timed_events[0].record()
        vfa_b200.aggregate_forward_raw(feats_cl, table, weights, biases, args.flags, out=out, workspace=ws,
                                       prepared=True)
        if timed_events is not None:
            timed_events[1].record()

    for _ in range(max(3, args.warmup)):
        step()
    path = vfa_b200.last_kernel_path()
    tile_pool = path.startswith('fside') and os.environ.get('VFA_POOL_TILE', '1') != '0'
    pool_name = 'pool_tile_kernel' if tile_pool else 'pool_list_kernel'
    # launches of this repo's kernels per step: table_build + prep_weight (all scales) + tap records + {fused grid-side kernel |
    # pooling lists with the coverage bitmap (tile_build_kernel + tile_order_kernel; VFA_POOL_TILE=0: cover_mark_kernel +
    # qlist_build_kernel) + row lists + unit table + per frame chunk: compacted ygemm + pooling kernel + its completion pass
    # (pool_quad_kernel<OVF>)}; memsets not counted
    launches_per_step = 3 + (4 + 3 * nums['fside_chunks'] if path.startswith('fside') else 1)
    if args.mode == 'views':                          # every frame chunk is a complete call (table and weights included)
        launches_per_step *= -(-B // (args.view_chunk or B))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(kev[i])
    ev1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    frames_per_step = B if args.mode in ('slab', 'views') else world * B
    value = frames_per_step * args.steps / (total_ms * 1e-3)
    kern = sum(kernel_ms) / len(kernel_ms)
    plain = args.mode == 'dp' and not args.backward

    # ---- per-kernel durations of the feature-side pair (CUDA events; debug bits switch one kernel off at a time:
    #      64 = GEMM only, 128 = pooling only, 256 = reuse the tap records -> exactly one launch between the events) ----
    per_kernel = None
    if path.startswith('fside') and plain:
        table = vfa_b200.build_table(cgeom, calibs, grid)
        per_kernel = {}
        for kname, bits in ((pool_name, 128 | 256), ('ygemm_kernel', 64 | 256), ('ygemm_kernel_all_tiles', 64)):
            # every loop starts from an idle GPU: back to back, the loop behind 20 launches of the tensor-bound GEMM ran up to
            # 15 % slower on some boxes (clock / power state carried over), while the step itself never moved
            torch.cuda.synchronize()
            time.sleep(0.25)
            os.environ['VFA_UMMA_VARIANT'] = str(bits)
            if kname == 'ygemm_kernel_all_tiles':        # every (tile, layer) multiplied: the GEMM's own efficiency
                os.environ['VFA_FSIDE_NO_SKIP'] = '1'
                table = vfa_b200.build_table(cgeom, calibs, grid)
            vfa_b200.reload_env()                         # the library reads its switches once per process
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for i in range(-2, args.steps):
                if i >= 0:
                    evs[i][0].record()
                vfa_b200.aggregate_forward_raw(feats_cl, table, weights, biases, args.flags, out=out, workspace=ws,
                                               prepared=True)
                if i >= 0:
                    evs[i][1].record()
            torch.cuda.synchronize()
            # median over the launches: a launch that waited for the host (interpreter pause, NVML poll) is not the kernel
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            per_kernel[kname] = ts[len(ts) // 2] / nums['fside_chunks']
        os.environ.pop('VFA_UMMA_VARIANT', None)
        os.environ.pop('VFA_FSIDE_NO_SKIP', None)
        vfa_b200.reload_env()
        vfa_b200.aggregate_forward_raw(feats_cl, table, weights, biases, args.flags, out=out, workspace=ws, prepared=True)
        torch.cuda.synchronize()

    # ---- end to end through the public API: pinned host features -> device -> aggregate -> host, every step ----
    def run_e2e(layout, dtype):
        """layout 'nchw' (the reference's [B,V,C,fH,fW] feature layout; transposed on the device) or 'nhwc'
        ([B,V,fH,fW,C], what a channels-last backbone emits); dtype float32 or bfloat16 (feature storage)."""
        src = [f.float() for f in feats_cl]
        if layout == 'nchw':
            host = [f.permute(0, 1, 4, 2, 3).contiguous().cpu().pin_memory() for f in src]
        else:
            host = [f.to(dtype).cpu().pin_memory() for f in src]
        table0 = vfa_b200.build_table(cgeom, calibs, grid)
        agg = vfa_b200.StreamingAggregator(table0, weights, biases, [tuple(h.shape) for h in host], args.flags, depth=2,
                                           dtype=dtype, channels_last=layout == 'nhwc')
        for _ in range(3):
            agg.submit(host, calibs, grid)
        agg.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 10))
        e0.record(agg.s_h2d)
        last = None
        for _ in range(n_e2e):
            last = agg.submit(host, calibs, grid)
        host_out = agg.result(last)
        e1.record(agg.s_d2h)
        agg.drain()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return {'value': world * B * n_e2e / (float(te.item()) * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': sum(h.numel() * h.element_size() for h in host),
                'd2h_bytes_per_step': host_out.numel() * 4, 'steps': n_e2e,
                'host_layout': '[B,V,C,fH,fW]' if layout == 'nchw' else '[B,V,fH,fW,C]',
                'feature_storage': 'bf16' if dtype == torch.bfloat16 else 'f32'}

    e2e = None
    if not args.no_e2e and plain:
        cpus = pin_to_gpu_numa_node(local_rank)
        if args.features == 'bf16':
            e2e = run_e2e('nhwc', torch.bfloat16)
        else:
            e2e = run_e2e('nchw', torch.float32)
            e2e['variants'] = {'channels_last_f32': run_e2e('nhwc', torch.float32),
                               'channels_last_bf16': run_e2e('nhwc', torch.bfloat16)}
            for v_ in e2e['variants'].values():
                v_.pop('unit', None)
        e2e['cpu_affinity'] = None if cpus is None else f'{len(cpus)} cores next to GPU {local_rank}'
        e2e['note'] = ('vfa_b200.StreamingAggregator: every step copies the pinned-host features H2D (fp32 [B,V,C,fH,fW] = the '
                       'reference layout, transposed to channels-last on the device), rebuilds the table, runs the aggregation '
                       'kernels (weights re-laid once: inference) and copies the full [B,C,L,W] result D2H; copies and compute '
                       'overlap on 3 streams / 2 device slots.  variants: host buffers already channels-last (no transpose), '
                       'fp32 and bf16 feature storage (bf16 halves the H2D bytes; tolerance 4e-3 of the output scale, '
                       'tests/test_gpu_parity.py::test_bf16_feature_storage)')

    # ---- two more lines of the same workload (same run, same inputs): static cameras, and the bf16 tensor-core variant ----
    extra = None
    if path.startswith('fside') and plain and not args.no_variants and args.features == 'f32':
        try:
            extra = {}
            table_s = vfa_b200.build_table(cgeom, calibs, grid)
            vfa_b200.prepare_weights(cgeom, shape, weights, args.flags, workspace=ws)
            vfa_b200.aggregate_forward_raw(feats_cl, table_s, weights, biases, args.flags, out=out, workspace=ws, prepared=True)

            def static_step():          # cameras and weights fixed (inference on a fixed rig): everything derived from them is reused
                vfa_b200.aggregate_forward_raw(feats_cl, table_s, weights, biases, args.flags, out=out, workspace=ws,
                                               prepared=True, table_prepared=True)
            for _ in range(3):
                static_step()
            ms = time_steps(static_step, args.steps, barrier, dev, world) / args.steps
            extra['static_cameras'] = {
                'value': world * B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'dtype': 'f32',
                'note': 'VFA_FLAG_TABLE_PREPARED | VFA_FLAG_WEIGHTS_PREPARED: projection table, tap records, coverage, pooling lists '
                        'and the weight re-layout of the previous call reused (the reference recomputes its boxes every forward, '
                        'vfa_op.py:64-88, although calibrations are per-dataset constants); the headline `value` rebuilds all of it '
                        'every step'}
            bflags = args.flags | vfa_b200.FLAG_BF16_MMA
            ws_b = vfa_b200.workspace_for(cgeom, shape, bflags, dev)

            def bf16_step():
                tb = vfa_b200.build_table(cgeom, calibs, grid)
                vfa_b200.prepare_weights(cgeom, shape, weights, bflags, workspace=ws_b)
                vfa_b200.aggregate_forward_raw(feats_cl, tb, weights, biases, bflags, out=out, workspace=ws_b, prepared=True)
            for _ in range(3):
                bf16_step()
            ms = time_steps(bf16_step, args.steps, barrier, dev, world) / args.steps
            extra['bf16_mma'] = {
                'value': world * B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'kernel_path': vfa_b200.last_kernel_path(),
                'dtype': 'bf16 (one tcgen05 kind::f16 pass, fp32 accumulate; Y stored in bf16; pooling and sums in fp32)',
                'tolerance': 'max |err| <= 4e-3, mean <= 5e-4 of max|out| vs the float64 port (measured 1.6e-3 / 2e-4): '
                             'tests/test_gpu_frame_parity.py::test_bf16_mma_variant',
                'table': 'rebuilt every step'}
            del ws_b
            step()                       # leave `out` / the workspace as the fp32 step left them
        except Exception as exc:      # an optional block must never cost the headline line
            extra = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- BASELINE configs 2-3: MultiviewX / Wildtrack, one frame, one GPU (same run; N = 1 only) ----
    configs = None
    if plain and world == 1 and args.workload == 'MultiviewC' and not args.no_configs:
        try:
            configs = []
            for wl in ('MultiviewX', 'Wildtrack'):
                Q = setup(wl, 1, seed=4321)
                q_shape = vfa_b200.make_shape(Q['feats'], Q['cgeom'].n_layers)
                q_ws = vfa_b200.workspace_for(Q['cgeom'], q_shape, args.flags, dev)
                q_out = torch.empty(1, C, Q['grid'].shape[0], Q['grid'].shape[1], device=dev)

                def q_step():
                    tb = vfa_b200.build_table(Q['cgeom'], Q['calibs'], Q['grid'])
                    vfa_b200.prepare_weights(Q['cgeom'], q_shape, Q['ws'], args.flags, workspace=q_ws)
                    vfa_b200.aggregate_forward_raw(Q['feats'], tb, Q['ws'], Q['bs'], args.flags, out=q_out, workspace=q_ws,
                                                   prepared=True)
                for _ in range(3):
                    q_step()
                ms = time_steps(q_step, args.steps, barrier, dev, world) / args.steps
                qn = workload_numbers(Q['g'], 1)
                configs.append({'workload': f'{wl}-shaped aggregation forward (BASELINE config {2 if wl == "MultiviewX" else 3})',
                                'batch': 1, 'n_gpus': 1, 'views': Q['g'].n_views, 'grid': list(Q['grid'].shape[:2]) + [len(Q['zs'])],
                                'value': 1e3 / ms, 'unit': UNIT, 'ms_per_step': ms, 'kernel_path': vfa_b200.last_kernel_path(),
                                'table': 'rebuilt every step',
                                'algorithmic_bytes_per_frame': qn['bytes'],
                                'hbm_frac_of_step': qn['bytes'] / (ms * 1e-3) / 1e9 / load_peaks()['hbm_gbs']})
                del Q, q_ws, q_out
            torch.cuda.empty_cache()
        except Exception as exc:      # an optional block must never cost the headline line
            configs = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- strong scaling of ONE frame batch over the GPUs: camera sharding, the cross-GPU sum fused into the pooling
    #      kernel (multimem.red over the NVLink multicast address); B frames total, every rank owns V/N cameras ----
    strong = None
    if plain and world > 1 and not args.no_strong:
        try:
            sv0, sv1 = vd.view_bounds(V, world, rank)
            S_ = setup(args.workload, B, (sv0, sv1), seed=777)        # this rank's cameras of the same B frames
            try:
                fused = vd.FusedViewAggregator(cgeom, B, C)
                kind = ('fused reduce-scatter in pool_tile_kernel (red.add.v4 into the band owner\'s replica over NVLink) + multimem.st '
                        'all-gather of the bands (vfa_multicast_copy), 2 barriers per step, no NCCL call'
                        if fused.mode == 'reduce_scatter' else
                        'fused multimem.red.add in pool_tile_kernel on the NVLink multicast address (every replica receives every '
                        'partial tile), 1 barrier per step, no NCCL call')
            except Exception as exc:                                  # no multicast support: NCCL all-reduce of the partial maps
                fused, kind = None, f'NCCL all-reduce of partial maps (fused path unavailable: {exc})'

            def strong_step():
                if sv1 > sv0:
                    tb = vfa_b200.build_table(cgeom, S_['calibs'], grid)
                if fused is not None:
                    return fused(S_['feats'], tb if sv1 > sv0 else None, weights, biases)
                part = (vfa_b200.aggregate_forward_raw(S_['feats'], tb, weights, biases, args.flags) if sv1 > sv0
                        else torch.zeros(B, C, grid.shape[0], grid.shape[1], device=dev))
                dist.all_reduce(part)
                return part
            for _ in range(3):
                strong_step()
            ms = time_steps(strong_step, args.steps, barrier, dev, world) / args.steps
            one_gpu_ms = total_ms / args.steps                        # the dp step above: B frames, all V cameras, one GPU
            map_bytes = B * C * grid.shape[0] * grid.shape[1] * 4
            strong_kind_mode = None if fused is None else fused.mode
            strong = {'mode': 'views (camera sharding)', 'collective': kind, 'collective_mode': strong_kind_mode, 'frames_per_step': B, 'n_gpus': world,
                      'cameras_per_rank': [vd.view_bounds(V, world, r)[1] - vd.view_bounds(V, world, r)[0] for r in range(world)],
                      'value': B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'one_gpu_ms_per_step': one_gpu_ms,
                      'speedup': one_gpu_ms / ms, 'efficiency': one_gpu_ms / ms / world,
                      'max_speedup_from_camera_split': V / max(vd.view_bounds(V, world, r)[1] - vd.view_bounds(V, world, r)[0]
                                                               for r in range(world)),
                      'reduced_bytes_per_rank_per_step': map_bytes,
                      'nvlink_bytes_per_rank_per_step': ({'reduce_scatter_out': map_bytes * (world - 1) // world,
                                                          'all_gather_in': map_bytes * (world - 1) // world}
                                                         if fused is not None and fused.mode == 'reduce_scatter' else
                                                         {'multicast_red_out': map_bytes, 'in': map_bytes * (world - 1)}),
                      'note': 'each rank red.adds (N-1)/N of its partial [B,L,W,C] map into the owners\' bands while the kernel '
                              'still pools the next tiles (overlapped, not separately timeable), then broadcasts its own band '
                              '(1/N of the map out, (N-1)/N in); step time includes the table + weight prep of the rank\'s cameras '
                              'and both barriers'}
            del S_, fused
        except Exception as exc:      # an optional block must never cost the headline line
            strong = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- BASELINE config 4: batch 64 forward + backward, sharded over the GPUs (64 / N frames per rank; dWeight / dBias
    #      all-reduced over NCCL -- the only exchange; camera / slab sharding replicate or serialise work at this size) ----
    config4 = None
    if plain and not args.no_config4 and args.workload == 'MultiviewC':
        try:
            B4 = 64
            per = B4 // world
            Q = setup(args.workload, per, seed=999)
            for t_ in Q['feats'] + Q['ws'] + Q['bs']:
                t_.requires_grad_(True)
            g4 = torch.randn(per, C, grid.shape[0], grid.shape[1], device=dev)

            def c4_step():
                tb = vfa_b200.build_table(cgeom, Q['calibs'], grid)
                res = vfa_b200.aggregate(Q['feats'], tb, Q['ws'], Q['bs'], flags=args.flags, channels_last=True)
                for t_ in Q['feats'] + Q['ws'] + Q['bs']:
                    t_.grad = None
                res.backward(g4)
                if world > 1:
                    vd.allreduce_collapse_grads(Q['ws'] + Q['bs'])
            for _ in range(2):
                c4_step()
            n4 = max(2, min(args.steps, 4))
            ms = time_steps(c4_step, n4, barrier, dev, world) / n4
            config4 = {'workload': 'MultiviewC batch-64 aggregation forward+backward (BASELINE config 4)', 'n_gpus': world,
                       'frames_per_rank': per, 'value': per * world / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': n4,
                       'sharding': 'frames (batch) over GPUs', 'collective': 'NCCL all-reduce of dWeight / dBias '
                       f'({sum(w.numel() + b.numel() for w, b in zip(Q["ws"], Q["bs"])) * 4} bytes per step)',
                       'scaling': 'strong (64 frames in total)'}
            del Q, g4
            torch.cuda.empty_cache()
        except Exception as exc:      # an optional block must never cost the headline line
            config4 = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- BASELINE config 5: one optimiser step of the whole detector, data parallel (DDP over NCCL), 1 frame per GPU ----
    config5 = None
    if plain and not args.no_config5 and args.workload == 'MultiviewC':
        try:
            t5 = train_measure(args, world, rank, local_rank, dev, 1, max(3, min(args.steps, 5)), 3, breakdown=False)
            config5 = {'workload': 'full VFA training step, data parallel (BASELINE config 5): ' + t5['config']['workload'],
                       'n_gpus': world, 'batch_per_gpu': 1, 'value': t5['value'], 'unit': UNIT, 'ms_per_step': t5['ms_per_step'],
                       'steps': t5['steps'], 'e2e': t5['e2e'], 'final_loss': t5['final_loss'],
                       'collective': f"NCCL all-reduce of the gradients of {t5['config']['parameters']} parameters "
                                     f"({t5['config']['parameters'] * 4} bytes per step), overlapped with the backward by DDP",
                       'scaling': 'weak (1 frame per GPU)', 'kernel_path': t5['config']['kernel_path'],
                       'note': 'trunk / heads / loss are stock cuDNN / ATen; `bench.py --mode train` prints the same step '
                               'with its breakdown'}
            del t5
            torch.cuda.empty_cache()
        except Exception as exc:      # an optional block must never cost the headline line
            config5 = {'error': f'{type(exc).__name__}: {exc}'}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        views = max(1, min(args.cpu_views, V))
        cpu_port_time(geom, 1)                                   # warm-up (thread pool, allocator)
        dt, _ = cpu_port_time(geom, views)
        cpu_baseline = {'value': 1.0 / (dt * V / views), 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                        'host_cpus': os.cpu_count(),
                        'sample': f'{views} of {V} views x 3 scales of one {args.workload}-shaped frame ({dt:.2f} s), fp32 '
                                  f'torch-CPU port of the reference operator sequence (oracle/ref_port.py), extrapolated '
                                  f'x{V}/{views} to a frame'}

    if rank == 0:
        peaks = load_peaks()
        work_scale = (3.0 if args.backward else 1.0) / (world if args.mode in ('slab', 'views') else 1)
        tflops = nums['flops'] * work_scale / (kern * 1e-3) / 1e12
        gbs = nums['bytes'] / (kern * 1e-3) / 1e9
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        tj = {}
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)

        def traffic_of(kernel):
            per_frame = tj.get(args.workload, {}).get(kernel)
            return None if per_frame is None else per_frame * B / nums['fside_chunks']

        step_ms = total_ms / args.steps
        # roofline fraction of the step = max(T_bytes, T_flops) / T_measured, both terms printed.
        #   T_bytes: compulsory HBM traffic (features once + BEV map once + weights / boxes / grid once) / measured copy bandwidth
        #   T_flops: the SHIPPED formulation's contraction (feature side: 2 * B*V*sum(fH*fW) * C * C*nl) / measured bf16 tensor
        #            peak (sustained: the step runs inside a long loop); the same / (peak / 6) for what fp32 parity costs on the
        #            tensor cores (3 TF32 passes at half the bf16 rate)
        flops_step = nums['fside_flops'] if path.startswith('fside') else nums['flops']
        t_bytes = nums['bytes'] / (peaks['hbm_gbs'] * 1e9) * 1e3
        t_flops = flops_step / (peaks['tflops'] * 1e12) * 1e3
        t_flops_x3 = flops_step / (peaks['tflops'] / 6.0 * 1e12) * 1e3
        step_block = {
            't_bytes_ms': t_bytes, 't_flops_ms': t_flops, 't_flops_tf32x3_ms': t_flops_x3, 't_measured_ms': step_ms,
            'frac': max(t_bytes, t_flops) / step_ms, 'frac_tf32x3': max(t_bytes, t_flops_x3) / step_ms,
            'algorithmic_bytes': nums['bytes'], 'algorithmic_flops': flops_step,
            'flops_formulation': 'feature-side (image-plane) contraction' if path.startswith('fside') else
                                 'grid-side contraction (as the reference)',
            'hbm_achieved_gbs': nums['bytes'] / (step_ms * 1e-3) / 1e9,
            'hbm_frac': nums['bytes'] / (step_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
            'hbm_peak_gbs': peaks['hbm_gbs'], 'tensor_peak_tflops': peaks['tflops'],
            'peak_source': peaks['source'],
            'note': 'step = table + weight re-layout + tap records + coverage + pooling lists + GEMM + pooling, B frames; '
                    'algorithmic bytes / flops per SURVEY.md section 8(d); frac = max(t_bytes_ms, t_flops_ms) / t_measured_ms'}
        if per_kernel is not None:
            frames = B / nums['fside_chunks']
            alg_bytes_launch = nums['bytes'] / nums['fside_chunks']          # 164.8 MB per MultiviewC frame x frames per launch
            pool_form_bytes = (nums['y_bytes'] + nums['out_bytes']) / nums['fside_chunks'] + nums['rec_bytes']
            t_pool, t_gemm = per_kernel[pool_name], per_kernel['ygemm_kernel']
            t_gemm_all = per_kernel['ygemm_kernel_all_tiles']
            # kernels timed in isolation (one launch between two events, nothing else running): burst bf16 figure
            burst = peaks['tflops_burst']
            gemm_tflops = nums['fside_flops'] / nums['fside_chunks'] / (t_gemm_all * 1e-3) / 1e12
            pool_tr, gemm_tr = traffic_of(pool_name), traffic_of('ygemm_compact_kernel')
            pool_block = {
                'kernel': pool_name, 'bound': 'hbm', 'kernel_ms': t_pool,
                'achieved': alg_bytes_launch / (t_pool * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': alg_bytes_launch / (t_pool * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'algorithmic_bytes_per_launch': alg_bytes_launch, 'formulation_bytes': pool_form_bytes,
                'formulation_gbs': pool_form_bytes / (t_pool * 1e-3) / 1e9,
                'traffic': pool_tr, 'traffic_over_algorithmic': None if pool_tr is None else pool_tr / alg_bytes_launch,
                'kernel_share_of_step': t_pool * nums['fside_chunks'] / step_ms, 'frames_per_launch': frames,
                'note': 'pooling of the per-layer products Y; algorithmic bytes = SURVEY 8(d) per-frame figure x frames per '
                        'launch (what the whole path must move); formulation_bytes = what THIS kernel must move (Y once + '
                        'pooling lists once + output once).  Bound by the shared-memory / LSU data pipe and packed-FMA issue, '
                        'not by HBM (profiles/README.md)'}
            gemm_block = {
                'kernel': 'ygemm_compact_kernel', 'bound': 'tensor', 'kernel_ms': t_gemm, 'kernel_ms_all_rows': t_gemm_all,
                'achieved': gemm_tflops, 'peak': burst, 'unit': 'TFLOP/s', 'frac': gemm_tflops / burst,
                'tf32x3_ceiling_tflops': burst / 6.0, 'frac_of_tf32x3_ceiling': gemm_tflops / (burst / 6.0),
                'rows_skipped_frac_est': max(0.0, 1.0 - t_gemm / t_gemm_all),
                'algorithmic_flops_per_launch': nums['fside_flops'] / nums['fside_chunks'],
                'traffic': gemm_tr, 'kernel_share_of_step': t_gemm * nums['fside_chunks'] / step_ms,
                'peak_source': 'measured bf16 burst (MEASURED_PEAKS.json): kernel timed in isolation',
                'note': 'image-plane contraction 2*B*V*fH*fW*C*(C*nl), quoted on the all-rows run (kernel_ms_all_rows); fp32 '
                        'parity needs 3 TF32 passes at half the bf16 rate: ceiling = peak/6.  The shipped kernel (kernel_ms) '
                        'multiplies only the covered texel rows'}
            dom = pool_block if t_pool >= t_gemm else gemm_block
            roofline = dict(dom)
            roofline['peak_source'] = dom.get('peak_source', peaks['source'])
            roofline['step'] = step_block
            roofline['kernels'] = [pool_block, gemm_block]
        else:
            roofline = {
                'bound': 'tensor', 'achieved': tflops, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                'frac': tflops / peaks['tflops'], 'traffic': traffic_of(path),
                'kernel': 'aggregate_fwd_umma_kernel' if path.startswith('umma') else
                          ('aggregate_fwd_simt_kernel' if path.startswith('simt') else 'whole step (several kernels)'),
                'kernel_ms': kern, 'kernel_share_of_step': kern * args.steps / total_ms,
                'algorithmic_flops_per_launch': nums['flops'], 'algorithmic_bytes_per_launch': nums['bytes'],
                'hbm_achieved_gbs': gbs, 'hbm_peak_gbs': peaks['hbm_gbs'], 'hbm_frac': gbs / peaks['hbm_gbs'],
                'peak_source': peaks['source'],
                'tf32x3_ceiling_tflops': peaks['tflops'] / 6.0,
                'frac_of_tf32x3_ceiling': tflops / (peaks['tflops'] / 6.0),
                'step': step_block,
                'note': 'collapse contraction counted grid-side (2*L*W*K*C per view and scale, as the reference computes '
                        'it); fp32 parity needs 3 TF32 tensor-core passes, so executed tensor flops are 3x the algorithmic '
                        'count and the TF32 rate is half the bf16 figure used as `peak`: the ceiling of this formulation '
                        'is peak/6 (tf32x3_ceiling_tflops)',
            }
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.mode in ('slab', 'views') else 'weak', 'vs_baseline': None,
            'dtype': 'f32' if path == 'simt_fp32' else 'f32 (3xTF32 tcgen05 contraction, fp32 pooling and sums)',
            'data': 'synthetic',
            'config': {'workload': f'{args.workload}-shaped aggregation forward' + ('+backward' if args.backward else ''),
                       'batch_per_gpu': B, 'views': V, 'channels': C, 'grid': list(grid.shape[:2]) + [len(zs)],
                       'feature_maps': [list(s) for s in geom.feature_sizes()], 'layout': 'channels_last',
                       'feature_storage': args.features,
                       'parallelism': f'{args.mode}{world}', 'kernel_path': path, 'pooling': pool_name if path.startswith('fside') else None,
                       'table': 'rebuilt every step',
                       'l2': f'inputs {nums["feat_bytes"] / 1e6:.0f} MB/step/GPU exceed the 126 MB L2'},
            'clocks': clocks,
            'e2e': e2e,
            'gpu_launches': launches_per_step * args.steps,
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'variants': extra,
            'configs': configs,
            'strong': strong,
            'config4': config4,
            'config5': config5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
