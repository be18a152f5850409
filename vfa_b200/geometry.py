"""Dataset geometries and the BEV ground grid of the aggregation path.

Numbers mirror the reference's configs-of-record (reference vfa/config.py:12-24, :39-52, :67-80) and the
grid->world conversion table of reference vfa/model/vfa_op.py:23-44.  `make_grid` restates reference
vfa/utils.py:16-37 (cell *origins*, z = 0; Wildtrack has its two axes swapped).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch

# grid->world conversion kinds understood by the C-ABI (include/vfa_b200.h: vfa_convert_t)
CONVERT_DIV = 0      # world = p / scale                        (MultiviewC: 1.0, MultiviewX: 40.0)
CONVERT_AFFINE = 1   # world = p * scale - offset  (per axis)   (Wildtrack: 2.5, (300, 900, 0))


@dataclasses.dataclass(frozen=True)
class Geometry:
    name: str
    world_size: tuple          # reference config `world_size`
    cube_size: tuple           # (l, w, h) of one voxel, grid units
    grid_height: int
    image_size: tuple          # (H, W) the boxes are normalised by (reference vfa_op.py:75)
    resize_size: tuple = (720, 1280)   # what the backbone sees -> feature map sizes
    n_views: int = 7
    channels: int = 256
    convert_kind: int = CONVERT_DIV
    convert_scale: float = 1.0
    convert_offset: tuple = (0.0, 0.0, 0.0)

    @property
    def n_layers(self) -> int:
        return len(range(0, self.grid_height, self.cube_size[2]))

    @property
    def grid_shape(self) -> tuple:
        ws = self.world_size[::-1] if self.name == 'Wildtrack' else self.world_size
        length, width = ws
        nx = len(np.arange(0., width, self.cube_size[0]))
        ny = len(np.arange(0., length, self.cube_size[1]))
        return (nx, ny) if self.name == 'Wildtrack' else (ny, nx)

    def feature_sizes(self) -> list:
        """Stride-8/16/32 map sizes of the resized image (ResNet: ceil at stride 32)."""
        h, w = self.resize_size
        out = []
        for s in (8, 16, 32):
            out.append((math.ceil(h / s), math.ceil(w / s)))
        return out


MULTIVIEWC = Geometry('MultiviewC', (3900, 3900), (25, 25, 32), 160, (720, 1280), n_views=7)
MULTIVIEWX = Geometry('MultiviewX', (640, 1000), (4, 4, 8), 64, (1080, 1920), n_views=6,
                      convert_scale=40.0)
WILDTRACK = Geometry('Wildtrack', (480, 1440), (4, 4, 4), 32, (1080, 1920), n_views=7,
                     convert_kind=CONVERT_AFFINE, convert_scale=2.5, convert_offset=(300.0, 900.0, 0.0))

GEOMETRIES = {g.name: g for g in (MULTIVIEWC, MULTIVIEWX, WILDTRACK)}

# BASELINE.json words the MultiviewC ground plane as "37.5 m x 37.5 m" (reference README.md:20); the shipped config is
# 3900 cm -> 156 x 156 cells (the config-of-record the benchmark uses).  The literal reading, 150 x 150 cells:
MULTIVIEWC_37M5 = dataclasses.replace(MULTIVIEWC, world_size=(3750, 3750))
BENCH_WORKLOADS = dict(GEOMETRIES, **{'MultiviewC-37.5m': MULTIVIEWC_37M5})


def convert_descriptor(dataset: str):
    """(kind, scale, offset[3]) for a reference `args.data` string (reference vfa_op.py:37-44)."""
    try:
        g = GEOMETRIES[dataset]
    except KeyError:
        raise ValueError(f"unknown dataset {dataset!r}; expected one of {sorted(GEOMETRIES)}") from None
    return g.convert_kind, g.convert_scale, g.convert_offset


def make_grid(world_size=(3900, 3900), grid_offset=(0, 0, 0), cube_LW=(25, 25), dataset='Wildtrack'):
    """BEV grid of cell origins, fp32 [L, W, 3] (same call signature as reference vfa/utils.py:16)."""
    if dataset == 'Wildtrack':
        length, width = world_size[::-1]
    else:
        length, width = world_size
    xoff, yoff, zoff = grid_offset
    xs = torch.arange(0., width, cube_LW[0]) + xoff
    ys = torch.arange(0., length, cube_LW[1]) + yoff
    if dataset == 'Wildtrack':
        xx, yy = torch.meshgrid(xs, ys, indexing='ij')
    else:
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
    return torch.stack([xx, yy, torch.full_like(xx, float(zoff))], dim=-1)


def grid_for(geom: Geometry) -> torch.Tensor:
    return make_grid(geom.world_size, cube_LW=geom.cube_size[:2], dataset=geom.name)
