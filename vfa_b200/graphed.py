"""Fixed-shape inference through ONE CUDA-graph launch.

A forward call is 12 kernel launches (projection table, tap records, coverage, row lists, texel lists, GEMM, pooling,
...) issued from Python through ctypes; at batch 1 the GPU work of a MultiviewC-shaped frame (~0.9 ms) is no longer much
larger than the host time of issuing it.  `GraphedAggregator` owns static input / output / workspace buffers, captures
the whole sequence once (the library enqueues on the current stream, allocates nothing and never synchronises, so it is
capturable as it stands) and replays it per frame batch: one `cudaGraphLaunch` instead of 10 launches + their Python.
Moving cameras are supported -- the table is rebuilt INSIDE the graph from the static calibration buffer.  Inference
only (frozen, pre-laid weights; no autograd).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, vfa_op


class GraphedAggregator:
    def __init__(self, geom: _lib.Geometry, feature_shapes, weights, biases, flags: int = 0):
        """feature_shapes: S shapes (B, V, fH, fW, C) of the channels-last feature maps; weights / biases: the collapse
        layers of the S scales (CUDA tensors; re-laid once, here)."""
        self.geom, self.flags = geom, int(flags)
        self.weights = [w.detach().contiguous() for w in weights]
        self.biases = [b.detach().contiguous() for b in biases]
        dev = self.weights[0].device
        self.device = dev
        B, V, C_ = feature_shapes[0][0], feature_shapes[0][1], feature_shapes[0][4]
        LW = geom.grid_l * geom.grid_w
        self.feats = [torch.zeros(tuple(s), dtype=torch.float32, device=dev) for s in feature_shapes]
        self.calibs = torch.zeros(V, 3, 4, dtype=torch.float32, device=dev)
        self.grid = torch.zeros(LW, 3, dtype=torch.float32, device=dev)
        self.boxes = torch.empty(V, geom.n_layers, LW, 4, dtype=torch.float32, device=dev)
        self.out = torch.empty(B, C_, geom.grid_l, geom.grid_w, dtype=torch.float32, device=dev)
        self.shape = vfa_op.make_shape(self.feats, geom.n_layers)
        self.workspace = vfa_op.workspace_for(geom, self.shape, self.flags, dev)
        vfa_op.prepare_weights(geom, self.shape, self.weights, self.flags, workspace=self.workspace)
        self.graph = None

    def _enqueue(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vfa_table_build(C.byref(self.geom), self.calibs.shape[0], self.calibs.data_ptr(),
                                                  self.grid.data_ptr(), self.boxes.data_ptr(), vfa_op._stream()))
        vfa_op.aggregate_forward_raw(self.feats, vfa_op.ProjectionTable(self.geom, self.boxes), self.weights, self.biases,
                                     self.flags, out=self.out, workspace=self.workspace, prepared=True)

    def load(self, feats_cl=None, calibs=None, grid=None):
        """Copy new inputs into the static buffers (device-to-device, or from pinned host memory)."""
        if feats_cl is not None:
            for dst, src in zip(self.feats, feats_cl):
                dst.copy_(src, non_blocking=True)
        if calibs is not None:
            self.calibs.copy_(calibs.reshape(-1, 3, 4), non_blocking=True)
        if grid is not None:
            self.grid.copy_(grid.reshape(-1, 3), non_blocking=True)

    def capture(self):
        """Warm up (kernel attributes, lazily cached occupancy queries) and record the graph.  The static calibration /
        grid buffers must hold valid values (`load`) -- the warm-up runs the real kernels."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._enqueue()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()
        return self

    def __call__(self, feats_cl=None, calibs=None, grid=None) -> torch.Tensor:
        """Aggregate one batch; returns the static output buffer [B, C, L, W] (overwritten by the next call)."""
        self.load(feats_cl, calibs, grid)
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.out
