"""Inputs of the network-level golden case, shared by make_golden_net.py (build container, reference present) and
tests/test_gpu_network.py (GPU box, reference absent).  Everything is derived from fixed seeds with numpy's
RandomState, which is stable across versions and machines.

Rig: MultiviewC cuboids (25 x 25 x 32, 5 layers) on every 4th cell of the 156 x 156 grid -> 39 x 39 cells, two ring
cameras with 704 x 704 images (stride-32 map 22 x 22: the smallest square for which every tap of the default clamp
range stays inside the map)."""
import dataclasses

import numpy as np
import torch

from vfa_b200 import geometry, synthetic

IMAGE = (704, 704)
VIEWS = 2
STEP = 4


def case_inputs():
    """-> (geometry, images [V,3,H,W], calibs [V,3,4], grid [1,l,w,3])"""
    g = dataclasses.replace(geometry.MULTIVIEWC, image_size=IMAGE, resize_size=IMAGE)
    images = torch.from_numpy(np.random.RandomState(7).rand(VIEWS, 3, *IMAGE).astype(np.float32))
    calibs = synthetic.ring_calibs(g, n_views=VIEWS)
    grid = geometry.grid_for(g)[::STEP, ::STEP].contiguous()[None]
    return g, images, calibs, grid
