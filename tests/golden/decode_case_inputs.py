"""Seeded head maps of the decode-tail golden case (shared by make_golden_decode.py and the tests)."""
import torch

L, W, ANGLES, TOPK, CLS_THRESH = 39, 52, 360, 50, 0.94
GRID_SIZE, WORLD_SIZE = (39.0, 52.0), (975.0, 1300.0)          # world_size / cube_LW of a 25-unit grid
DIM_MEAN = (170.0, 45.0, 55.0)


def case_pred(seed=0, batch=1):
    g = torch.Generator().manual_seed(seed)
    heat = torch.randn(batch, 1, L, W, generator=g) * 1.5 - 1.0
    # a few plateaus of exactly equal logits (NMS keeps every cell of a plateau that is its window's maximum) and peaks
    heat[0, 0, 5:7, 5:8] = 3.2
    heat[0, 0, 20, 30] = 4.0
    heat[0, 0, 0, 0] = 3.0
    heat[0, 0, L - 1, W - 1] = 3.5
    return {'heatmap': heat,
            'loc_offset': torch.randn(batch, L, W, 2, generator=g),
            'dim_offset': torch.randn(batch, L, W, 3, generator=g) * 0.2,
            'rotation': torch.randn(batch, L, W, ANGLES, generator=g)}
