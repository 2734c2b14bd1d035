#!/usr/bin/env python
"""One silhouette forward+backward at 256x256 from seeded projections: the command ncu wraps (tools/prof_sil.py --batch 296)."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--wh", type=int, default=256)
args = ap.parse_args()
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
dec = pkg.SmplDecoder(host, args.wh, None, device=dev)
x0 = torch.as_tensor(synth.make_params(args.batch, args.wh, seed=0), device=dev)
with torch.no_grad():
    pr = dec(x0, seg=False)["projects"]
g = torch.randn((args.batch, args.wh, args.wh, 2), device=dev)
for _ in range(2):
    x = pr.clone().requires_grad_(True)
    (pkg.projects_to_silhouette(x, args.wh) * g).sum().backward()
torch.cuda.synchronize()
print("ok", float(x.grad.abs().max()))
