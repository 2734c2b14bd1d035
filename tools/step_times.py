#!/usr/bin/env python
"""C5 step (decode + project + mask + seg, fwd+bwd) at --batch: ms/step and the mean duration of every library kernel
(the library's event profiler).  The quick A/B command between kernel versions; bench.py stays the contract bench."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=40)   # ~0.3 s: the SM clocks ramp from idle
args = ap.parse_args()
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
dec = pkg.SmplDecoder(host, 48, 5, parts=pkg.smpl_io.golden_part_vertices(5), device=dev)
x0 = torch.as_tensor(synth.make_params(args.batch, 48, seed=0), device=dev)
g = torch.randn((args.batch, 48, 48, 32), device=dev)


def step():
    x = x0.clone().requires_grad_(True)
    dec(x)["seg"].backward(g)
    return x.grad


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
pkg.profile_enable(True)
pkg.profile_collect()
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
pkg.profile_enable(False)
print(round(ms, 3), {k: round(t / n, 3) for k, (n, t) in pkg.profile_collect().items()})
