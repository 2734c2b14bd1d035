"""Seeded synthetic inputs of the shape the reference's training graph produces.

params = mean(86) + noise: camera [wh/2, wh/2, wh/2, wh/1.6] * (1 + 0.05 N(0,1)); pose = mean pose (global rotation
zeroed, concat_mean_param.py:14) + 0.2 N(0,1) rad per component (global rotation 0.3 N(0,1)); shape = N(0,1) clipped
to +-3.  numpy only.
"""
from __future__ import annotations

import numpy as np

from . import smpl_io


def make_params(n: int, img_wh: float, seed: int = 0, noise: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    mean = smpl_io.mean_param_vector(img_wh)                      # (1,86) float64
    p = np.repeat(mean, n, axis=0)
    p[:, :4] *= 1.0 + noise * 0.05 * rng.standard_normal((n, 4))
    p[:, 4:7] += noise * 0.3 * rng.standard_normal((n, 3))
    p[:, 7:76] += noise * 0.2 * rng.standard_normal((n, 69))
    p[:, 76:] = np.clip(noise * rng.standard_normal((n, 10)), -3, 3) if noise else p[:, 76:]
    return p.astype(np.float32)
