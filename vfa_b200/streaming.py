"""Host-to-host streaming of the aggregation: overlap PCIe copies with the aggregation kernels.

The aggregation of a MultiviewC-shaped frame takes ~1 ms on a B200, the host->device copy of its 135.6 MB of fp32
features ~2.6 ms on a Gen5 x16 link -- run back to back on one stream the GPU would idle three quarters of the time.  `StreamingAggregator` keeps `depth` device-side input slots and runs three CUDA streams (H2D, compute, D2H)
chained with events, so the copy of batch i+1 and the read-back of batch i-1 overlap the aggregation of batch i.
Inference only (no autograd); features arrive as pinned-host `[B,V,C,fH,fW]` tensors (the reference's NCHW layout).
"""
from __future__ import annotations

import torch

from . import vfa_op


class StreamingAggregator:
    def __init__(self, table: vfa_op.ProjectionTable, weights, biases, feature_shapes, flags: int = 0, depth: int = 2,
                 dtype=torch.float32, channels_last: bool = False):
        """feature_shapes: list of S shapes (B, V, C, fH, fW) -- or (B, V, fH, fW, C) with channels_last=True, the layout a
        backbone running in torch.channels_last emits: no transpose kernel on the way in; weights/biases: the three
        collapse layers (CUDA).  dtype=torch.bfloat16 (channels_last only): bf16 feature storage, half the PCIe bytes
        (VFA_FLAG_BF16_FEATURES; tolerance: tests/test_gpu_parity.py::test_bf16_feature_storage)."""
        if dtype == torch.bfloat16 and not channels_last:
            raise ValueError('bfloat16 features must be channels-last')
        self.table, self.flags, self.depth = table, int(flags), int(depth)
        self.dtype, self.channels_last = dtype, bool(channels_last)
        self.weights = [w.detach().contiguous() for w in weights]
        self.biases = [b.detach().contiguous() for b in biases]
        dev = self.weights[0].device
        self.device = dev
        geom = table.geom
        B, C = feature_shapes[0][0], feature_shapes[0][4 if channels_last else 2]
        self.in_slots = [[torch.empty(shape, dtype=dtype, device=dev) for shape in feature_shapes]
                         for _ in range(depth)]
        self.out_slots = [torch.empty(B, C, geom.grid_l, geom.grid_w, dtype=torch.float32, device=dev)
                          for _ in range(depth)]
        self.host_out = [torch.empty(B, C, geom.grid_l, geom.grid_w, dtype=torch.float32).pin_memory()
                         for _ in range(depth)]
        self.s_h2d, self.s_compute, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]        # slot's inputs have landed
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]      # slot's inputs have been consumed
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]       # slot's result is on the device
        self.ev_host = [torch.cuda.Event() for _ in range(depth)]      # slot's result is in pinned host memory
        self._n = 0
        # frozen weights: re-lay them once
        cl0 = self.in_slots[0] if channels_last else [vfa_op.to_channels_last(t) for t in self.in_slots[0]]
        self.shape = vfa_op.make_shape(cl0, geom.n_layers)
        self.workspace = vfa_op.workspace_for(geom, self.shape, self.flags, dev)
        vfa_op.prepare_weights(geom, self.shape, self.weights, self.flags, workspace=self.workspace)
        torch.cuda.synchronize(dev)

    def submit(self, host_feats, calibs=None, grid=None) -> int:
        """Enqueue one batch (list of S pinned-host tensors).  Returns a ticket for `result`.
        With calibs / grid given the projection table is rebuilt for this batch (moving cameras)."""
        i = self._n
        k = i % self.depth
        caller = torch.cuda.current_stream(self.device)
        if i >= self.depth:
            self.s_h2d.wait_event(self.ev_free[k])          # the compute that read this slot has finished
            self.s_compute.wait_event(self.ev_host[k])      # and its previous result has left the device
        with torch.cuda.stream(self.s_h2d):
            for dst, src in zip(self.in_slots[k], host_feats):
                dst.copy_(src, non_blocking=True)
            self.ev_in[k].record(self.s_h2d)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(self.ev_in[k])
            # NCHW -> channels-last (own kernel) unless the host buffers already are
            cl = self.in_slots[k] if self.channels_last else [vfa_op.to_channels_last(t) for t in self.in_slots[k]]
            if calibs is not None:
                # moving cameras: the old table may still be read by kernels of this stream, and the new inputs were
                # produced on the caller's stream (cross-stream lifetime / ordering)
                self.s_compute.wait_stream(caller)
                self.table.boxes.record_stream(self.s_compute)
                calibs.record_stream(self.s_compute)
                grid.record_stream(self.s_compute)
                self.table = vfa_op.build_table(self.table.geom, calibs, grid)
            vfa_op.aggregate_forward_raw(cl, self.table, self.weights, self.biases, self.flags, out=self.out_slots[k],
                                         workspace=self.workspace, prepared=True)
            self.ev_free[k].record(self.s_compute)
            self.ev_out[k].record(self.s_compute)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_out[k])
            self.host_out[k].copy_(self.out_slots[k], non_blocking=True)
            self.ev_host[k].record(self.s_d2h)
        self._n += 1
        return i

    def result(self, ticket: int) -> torch.Tensor:
        """Pinned-host [B,C,L,W] result of `ticket` (valid until `depth` more batches have been submitted)."""
        if ticket < self._n - self.depth or ticket >= self._n:
            raise ValueError(f'ticket {ticket} is no longer (or not yet) buffered')
        k = ticket % self.depth
        self.ev_host[k].synchronize()
        return self.host_out[k]

    def drain(self):
        for s in (self.s_h2d, self.s_compute, self.s_d2h):
            s.synchronize()
