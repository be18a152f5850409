"""Multi-GPU execution of the aggregation path on one NVLink/NVSwitch box.

The reference is single-GPU, batch-1 (reference train.py:57-59, :245-246); both modes here are new functionality:

* batch data parallel -- frames are independent, every rank aggregates its own frames with no forward collective;
  `allreduce_collapse_grads` sums dWeight / dBias over ranks after the backward (the only exchange).
* BEV row-slab sharding -- rank g owns rows [r0, r1) of the L x W ground grid for all views and scales: it builds
  the projection table of its slab only, aggregates it, and the slabs are all-gathered (NCCL over NVLink) into
  the full [B, C, L, W] map.  View features must be present on every rank (`broadcast_features`, or replicated
  backbones); in the backward each slab yields a partial dFeature over the whole image, summed by all-reduce.
* camera sharding -- the BEV map is a plain sum over cameras of terms that are already past their ReLU (reference
  vfanet.py:79-82), so rank g aggregates its own cameras over the WHOLE grid and the partial maps are summed with one
  all-reduce (25-44 MB per frame, in-switch reduction on NVSwitch).  Features stay where their backbone ran -- nothing
  of the 136 MB per frame is broadcast -- and no image-plane work is replicated, which is what the feature-side
  forward needs (its GEMM runs per camera); the all-reduce of frame k overlaps the kernels of frame k + 1.

One process per GPU, `torch.distributed` (backend "nccl") for the plumbing.  The functions take the per-rank
compute as a callable so the partition / collective logic is testable on CPU ranks (gloo) with the oracle.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def slab_bounds(n_rows: int, world: int, rank: int):
    """Rows [r0, r1) of rank `rank`: as even as possible, the first n_rows % world ranks hold one extra row."""
    base, extra = divmod(n_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


class _GatherSlabs(torch.autograd.Function):
    """local [B, C, rows_g, W] -> full [B, C, L, W]; backward hands every rank its own rows of the gradient."""

    @staticmethod
    def forward(ctx, local, n_rows, group):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, C, rows, W = local.shape
        r0, r1 = slab_bounds(n_rows, world, rank)
        assert rows == r1 - r0, f'rank {rank} holds {rows} rows, plan says {r1 - r0}'
        max_rows = slab_bounds(n_rows, world, 0)[1]
        # equal-sized NCCL messages: pad the (at most one row) shorter slabs
        send = local.new_zeros(B, C, max_rows, W)
        send[:, :, :rows] = local
        recv = local.new_empty(world * B, C, max_rows, W)          # concatenated along dim 0 (gloo and nccl agree)
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
        recv = recv.view(world, B, C, max_rows, W)
        full = local.new_empty(B, C, n_rows, W)
        for g in range(world):
            a, b = slab_bounds(n_rows, world, g)
            full[:, :, a:b] = recv[g, :, :, :b - a]
        ctx.bounds = (r0, r1)
        return full

    @staticmethod
    def backward(ctx, grad_full):
        r0, r1 = ctx.bounds
        return grad_full[:, :, r0:r1].contiguous(), None, None


class _ReplicatedInput(torch.autograd.Function):
    """Identity on a tensor that is replicated on every rank; backward sums the per-rank partial gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def broadcast_features(feats, src: int = 0, group=None):
    """In-place broadcast of the view features from rank `src` (one NCCL broadcast per scale)."""
    for f in feats:
        dist.broadcast(f, src=src, group=group)
    return feats


def aggregate_slab(feats, calibs, grid, weights, biases, compute, group=None, replicated_grads: bool = True):
    """BEV row-slab sharded aggregation.

    feats / weights / biases  replicated on every rank (see broadcast_features); calibs [V,3,4]; grid [L,W,3]
    compute(feats, calibs, grid_slab, weights, biases) -> [B, C, rows, W]   the single-GPU aggregation of a grid slab
    returns the full [B, C, L, W] on every rank.  With replicated_grads=True gradients w.r.t. feats / weights /
    biases are all-reduced so every rank ends up with the gradient of the full map.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    L = grid.shape[-3]
    r0, r1 = slab_bounds(L, world, rank)
    if replicated_grads:
        feats = [_ReplicatedInput.apply(f, group) if f.requires_grad else f for f in feats]
        weights = [_ReplicatedInput.apply(w, group) if w.requires_grad else w for w in weights]
        biases = [_ReplicatedInput.apply(b, group) if b.requires_grad else b for b in biases]
    grid_slab = grid.reshape(L, grid.shape[-2], 3)[r0:r1].contiguous()
    local = compute(feats, calibs, grid_slab, weights, biases)
    return _GatherSlabs.apply(local, L, group)


class _SumOverRanks(torch.autograd.Function):
    """partial [B, C, L, W] of this rank's cameras -> the sum over ranks, started asynchronously: the caller waits on
    the handle appended to `works` before anything reads the result.  Every rank sees the same gradient of the
    (replicated) sum, so the backward is the identity."""

    @staticmethod
    def forward(ctx, partial, group, works):
        out = partial.contiguous().clone()
        if dist.is_initialized():                    # a single process owns every camera: nothing to exchange
            works.append(dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group, async_op=True))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def row_band(n_rows: int, world: int, rank: int, tile: int = 8):
    """BEV rows [r0, r1) whose reduction rank `rank` owns in the fused reduce-scatter (FusedViewAggregator): bands of whole
    `tile`-row tiles of equal height, the last rank takes what is left (possibly nothing).  Returns (r0, r1, band_rows);
    the kernel routes row cy to rank min(cy // band_rows, world - 1)."""
    band = -(-(-(-n_rows // tile)) // world) * tile
    r0 = min(rank * band, n_rows)
    r1 = n_rows if rank == world - 1 else min((rank + 1) * band, n_rows)
    return r0, r1, band


def view_bounds(n_views: int, world: int, rank: int):
    """Cameras [v0, v1) of rank `rank` (even split; with more ranks than cameras the last ranks hold none)."""
    return slab_bounds(n_views, world, rank)


def aggregate_views(feats_local, calibs_local, grid, weights, biases, compute, out_channels: int, group=None,
                    replicated_grads: bool = True, frames_per_chunk: int | None = None):
    """Camera-sharded aggregation.

    feats_local   this rank's cameras only: S tensors [B, V_g, ...] (V_g may be 0), calibs_local [V_g, 3, 4];
    grid [L, W, 3] and weights / biases replicated on every rank;
    compute(feats, calibs, grid, weights, biases) -> [b, C, L, W]   the single-GPU aggregation of those cameras
    returns the full [B, C, L, W] = sum over all cameras on every rank.  Frames are processed `frames_per_chunk` at a
    time (default: all at once) and each chunk's all-reduce runs while the next chunk is computed.  Gradients: dFeature
    is complete on the owning rank (no exchange); with replicated_grads=True dWeight / dBias are all-reduced.
    """
    L, W = grid.shape[-3], grid.shape[-2]
    grid = grid.reshape(L, W, 3)
    if replicated_grads and dist.is_initialized():
        weights = [_ReplicatedInput.apply(w, group) if w.requires_grad else w for w in weights]
        biases = [_ReplicatedInput.apply(b, group) if b.requires_grad else b for b in biases]
    B, Vg = feats_local[0].shape[0], feats_local[0].shape[1]
    step = B if not frames_per_chunk else max(1, min(B, frames_per_chunk))
    outs, works = [], []
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        if Vg > 0:
            partial = compute([f[b0:b1] for f in feats_local], calibs_local, grid, weights, biases)
        else:
            # a rank without cameras contributes zeros, but stays in the autograd graph of the replicated parameters so
            # that their gradient all-reduce (a collective) is entered on every rank
            tie = sum(w.sum() for w in weights) + sum(b.sum() for b in biases)
            partial = weights[0].new_zeros(b1 - b0, out_channels, L, W) + 0.0 * tie
        outs.append(_SumOverRanks.apply(partial, group, works))
    for w in works:
        w.wait()
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


def shard_frames(n_frames: int, world: int, rank: int):
    """Frames [f0, f1) of rank `rank` for batch data parallelism (same even split as slab_bounds)."""
    return slab_bounds(n_frames, world, rank)


def allreduce_collapse_grads(params, group=None, average: bool = False):
    """Sum (or average) .grad of the collapse parameters over the data-parallel ranks in one flat all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def cuda_compute(geom_of, flags: int = 0):
    """The default `compute` of aggregate_slab: table + fused kernels of this repo for a grid slab.
    geom_of(grid_lw) -> vfa_geometry_t for a slab of that many rows (see vfa_b200.make_geometry)."""
    from . import vfa_op

    def compute(feats, calibs, grid_slab, weights, biases):
        table = vfa_op.build_table(geom_of(grid_slab.shape[:2]), calibs, grid_slab)
        return vfa_op.aggregate(feats, table, weights, biases, flags=flags)
    return compute


class FusedViewAggregator:
    """Camera-sharded aggregation with the cross-GPU sum fused INTO the pooling kernel (inference / latency mode).

    Rank g owns cameras [v0, v1) (`view_bounds`).  Its pooling kernel does not store the partial BEV map: every finished
    8 x 8-cell tile goes straight over NVLink into a symmetric [B, L, W, C] buffer (torch symmetric memory), while the SMs
    are still pooling the next tiles.  No NCCL call, no partial map in HBM.  (Sum over cameras: reference vfanet.py:82;
    every term is already past its ReLU, so partial sums of camera subsets add exactly, up to fp32 summation order.)

    mode='auto' (default): 'multicast_red' up to 4 ranks, 'reduce_scatter' beyond.
    mode='reduce_scatter': the BEV rows are split into one band per rank; a tile is `red.add`-ed into the replica
      of the rank that OWNS its band (VFA_FLAG_OUT_PEERS: peer memory, 16-byte vector reductions) -- a reduce-scatter fused
      into the kernel, (N-1)/N of the map per rank in each direction.  After a barrier each rank broadcasts its finished
      band to every replica with `multimem.st` (vfa_multicast_copy; the NVSwitch replicates the stores): the all-gather
      half, map/N out and one map in per rank.  Two barriers per step.
    mode='multicast_red': `multimem.red.add` of every tile on the multicast address (VFA_FLAG_OUT_MULTICAST) -- one barrier,
      but every replica receives all N partial maps (N x the ingress): fine for 2 ranks, link-bound beyond.

    Two buffer slots alternate; the slot of step i + 1 is zeroed before the (first) barrier of step i, so no extra barrier
    orders "every accumulator is zero".  Needs NVSwitch multicast support (`multicast_ptr`).  Results come back as a
    [B, C, L, W] view with channels-last strides of the local replica (valid until the slot is reused two steps later).
    """

    def __init__(self, geom, batch: int, channels: int = 256, group=None, flags: int = 0, mode: str = 'auto'):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        if mode not in ('auto', 'reduce_scatter', 'multicast_red'):
            raise ValueError(f'unknown mode {mode!r}')
        self.group = group if group is not None else dist.group.WORLD
        if mode == 'auto':
            # multicast_red costs every replica N maps of NVLink ingress, hidden behind the pooling while N x map / 750 GB/s
            # stays below the rank's compute time: measured (MultiviewC, B = 4) +0.03 / +0.11 ms at 2 / 4 ranks against
            # +0.2 ms for the reduce-scatter's second phase; at 8 ranks the ingress (800 MB) is exposed
            mode = 'multicast_red' if dist.get_world_size(self.group) <= 4 else 'reduce_scatter'
        self.geom, self.B, self.C, self.mode = geom, int(batch), int(channels), mode
        self.flags = int(flags) | _lib.FLAG_OUT_NHWC | (_lib.FLAG_OUT_PEERS if mode == 'reduce_scatter'
                                                        else _lib.FLAG_OUT_MULTICAST)
        dev = torch.device('cuda', torch.cuda.current_device())
        L, W = geom.grid_l, geom.grid_w
        self.buf = symm_mem.empty(2, self.B, L, W, self.C, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        if not self.hdl.multicast_ptr:
            raise RuntimeError('symmetric memory has no multicast address on this system (no NVSwitch multicast support)')
        self.slot_bytes = self.buf[0].numel() * 4
        n, rank = self.hdl.world_size, self.hdl.rank
        if n > 16:
            raise ValueError('vfa_peer_outputs_t holds up to 16 ranks')
        # bands of whole 8-row tiles: rank r owns BEV rows [r0, r1)
        self.r0, self.r1, band = row_band(L, n, rank)
        self.band_rows = band
        self.desc = [torch.tensor([n, band] + [int(p_) + s * self.slot_bytes for p_ in self.hdl.buffer_ptrs]
                                  + [0] * (16 - n), dtype=torch.int64, device=dev) for s in range(2)]
        self.buf.zero_()
        self.hdl.barrier(channel=0)
        self._step = 0
        self.workspace = None

    @staticmethod
    def available() -> bool:
        try:
            import torch.distributed._symmetric_memory as symm_mem      # noqa: F401
        except Exception:
            return False
        return dist.is_initialized() and torch.cuda.is_available()

    def __call__(self, feats_local_cl, table_local, weights, biases):
        """feats_local_cl: S channels-last tensors [B, V_g, fH, fW, C] of this rank's cameras (V_g may be 0);
        table_local: ProjectionTable of the same cameras (None when V_g == 0).  Collective: every rank calls it."""
        import ctypes as C
        from . import _lib, vfa_op
        s = self._step & 1
        rs = self.mode == 'reduce_scatter'
        if rs:
            self.buf[1 - s][:, self.r0:self.r1].zero_()          # next step's accumulator: the band this rank owns
        else:
            self.buf[1 - s].zero_()
        if table_local is not None and feats_local_cl[0].shape[1] > 0:
            if self.workspace is None:
                shape = vfa_op.make_shape(feats_local_cl, self.geom.n_layers)
                self.workspace = vfa_op.workspace_for(table_local.geom, shape, self.flags, feats_local_cl[0].device)
            dst = self.desc[s].data_ptr() if rs else self.hdl.multicast_ptr + s * self.slot_bytes
            vfa_op.aggregate_forward_raw(feats_local_cl, table_local, weights, biases, self.flags, workspace=self.workspace,
                                         out_ptr=dst)
        self.hdl.barrier(channel=0)                              # every partial tile has landed in its owner's band
        if rs:
            if self.r1 > self.r0:
                row_bytes = self.geom.grid_w * self.C * 4
                stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for b in range(self.B):
                    off = s * self.slot_bytes + (b * self.geom.grid_l + self.r0) * row_bytes
                    _lib.check(_lib.lib().vfa_multicast_copy(self.buf.data_ptr() + off, self.hdl.multicast_ptr + off,
                                                             (self.r1 - self.r0) * row_bytes, stream))
            self.hdl.barrier(channel=0)                          # every band has been broadcast
        self._step += 1
        return self.buf[s].permute(0, 3, 1, 2)
