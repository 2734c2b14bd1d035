#!/usr/bin/env python
"""Error attribution on the GPU box: kernel vs fp32 oracle vs fp64 oracle, per intermediate boundary."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import np_oracle  # noqa: E402

pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
for n, vs in ((8, None), (70, 5)):
    p = synth.make_params(n, 48, seed=7)
    r32 = np_oracle.smpl_layer_call(host, p, return_all=True)
    r64 = np_oracle.smpl_layer_call(host, p.astype(np.float64), return_all=True)
    p32 = np_oracle.orthographic_project([r32["verts"], p], vs)
    p64 = np_oracle.orthographic_project([r64["verts"], p.astype(np.float64)], vs)
    dec = pkg.SmplDecoder(host, 48, vs, device=dev)
    out = dec(torch.as_tensor(p, device=dev), seg=False)
    gv, gj, gp = (out[k].cpu().numpy() for k in ("verts", "joints", "projects"))
    m = lambda a, b: float(np.abs(a - b).max())  # noqa: E731
    print("N=%d vs=%s" % (n, vs))
    print("  verts    kernel-f32 %.2e  kernel-f64 %.2e  f32-f64 %.2e" % (m(gv, r32["verts"]), m(gv, r64["verts"]), m(r32["verts"], r64["verts"])))
    print("  joints   kernel-f32 %.2e  kernel-f64 %.2e  f32-f64 %.2e" % (m(gj, r32["J_transformed"]), m(gj, r64["J_transformed"]), m(r32["J_transformed"], r64["J_transformed"])))
    print("  projects kernel-f32 %.2e  kernel-f64 %.2e  f32-f64 %.2e" % (m(gp, p32), m(gp, p64), m(p32, p64)))
