#!/usr/bin/env python
"""One forward+backward of the C5 decoder at a small batch: the command ncu wraps (tools/prof_step.py --batch 2048)."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--wh", type=int, default=48)
ap.add_argument("--vs", type=int, default=5)
ap.add_argument("--sil", type=int, default=0, help="also run the silhouette branch at this resolution")
args = ap.parse_args()
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
vs = None if args.vs <= 1 else args.vs
dec = pkg.SmplDecoder(host, args.wh, vs, silhouette_wh=args.sil or None, parts=pkg.smpl_io.golden_part_vertices(vs), device=dev)
x0 = torch.as_tensor(synth.make_params(args.batch, args.wh, seed=0), device=dev)
g = torch.randn((args.batch, args.wh, args.wh, 32), device=dev)
for _ in range(args.steps):
    x = x0.clone().requires_grad_(True)
    out = dec(x)
    loss_terms = [(out["seg"] * g).sum()]
    if args.sil:
        loss_terms.append(out["silhouette"].square().sum())
    sum(loss_terms).backward()
torch.cuda.synchronize()
print("ok", float(x.grad.abs().max()))
