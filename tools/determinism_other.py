#!/usr/bin/env python
"""Repeatability of the paths the C5 checks do not cover: dense decode backward (gradient on all 6890 vertices),
silhouette forward / backward, focal loss.  Same inputs, repeated evaluations; prints the largest difference."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)


def rep(name, fn, reps=5):
    ref = fn()
    worst = 0.0
    for _ in range(reps):
        out = fn()
        for a, b in zip(out, ref):
            worst = max(worst, float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30))
    print("%-46s max relative-to-max difference over %d repeats: %.3e" % (name, reps, worst))


N = 4096
dec = pkg.SmplDecoder(host, 48, None, device=dev)
x0 = torch.as_tensor(synth.make_params(N, 48, seed=0), device=dev)
gv = torch.randn((N, 6890, 3), device=dev)


def dense():
    x = x0.clone().requires_grad_(True)
    out = dec(x, seg=False)
    (out["verts"] * gv).sum().backward()
    return [x.grad, out["verts"].detach()]


rep("decode fwd+bwd, dense vertex gradient, N=4096", dense)
with torch.no_grad():
    pr = dec(x0[:1024], seg=False)["projects"].clone()
for wh in (64, 256):
    gs = torch.randn((pr.shape[0], wh, wh, 2), device=dev)

    def sil():
        p = pr.clone().requires_grad_(True)
        s = pkg.projects_to_silhouette(p, wh)
        s.backward(gs)
        return [s.detach(), p.grad]

    rep("silhouette fwd+bwd %dx%d, N=%d" % (wh, wh, pr.shape[0]), sil, reps=3)
seg = torch.rand((2048, 48, 48, 32), device=dev)
lab = torch.randint(0, 32, (2048, 48, 48), device=dev, dtype=torch.uint8)
loss_fn = pkg.categorical_focal_loss(gamma=2.0, from_logits=True)


def focal():
    s = seg.clone().requires_grad_(True)
    l = loss_fn(lab, s)
    l.sum().backward()
    return [l.detach(), s.grad]


rep("softmax + focal loss fwd+bwd, N=2048", focal)
