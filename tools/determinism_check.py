#!/usr/bin/env python
"""Run-to-run repeatability of the C5 step, stage by stage (same inputs, two evaluations): forward outputs must be
bit-identical; the seg backward may differ in the last bits (rows are handed to the warps on demand)."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
parts = pkg.smpl_io.golden_part_vertices(5)
dec = pkg.SmplDecoder(host, 48, 5, parts=parts, device=dev)
x0 = torch.as_tensor(synth.make_params(args.batch, 48, seed=0), device=dev)
g = torch.randn((args.batch, 48, 48, 32), device=dev)
with torch.no_grad():
    ref = dec(x0)
pr, mk = ref["projects"].clone(), ref["mask"].clone()


def seg_grad():
    p = pr.clone().requires_grad_(True)
    pkg.projects_to_seg([p, mk], 48, 5, parts=parts).backward(g)
    return p.grad


def full_grad():
    x = x0.clone().requires_grad_(True)
    out = dec(x)
    out["seg"].backward(g)
    return x.grad, out


gs0 = seg_grad()
gf0, out0 = full_grad()
gprev = gf0
for r in range(args.reps):
    gs = seg_grad()
    gf, out = full_grad()
    dd = (gf - gf0).abs()
    ii = int(dd.argmax())
    print("   full grad: worst at (sample %d, param %d): %.6e vs first %.6e; vs previous rep max diff %.3e; rows differing: %d"
          % (ii // 86, ii % 86, float(gf.flatten()[ii]), float(gf0.flatten()[ii]), float((gf - gprev).abs().max()),
             int((dd.amax(dim=1) > 1e-4 * gf0.abs().amax(dim=1)).sum())))
    gprev = gf
    d = (gs - gs0).abs().amax(dim=(1, 2))
    worst = int(d.argmax())
    print("rep %d: fwd identical: verts %s projects %s mask %s seg %s | seg_bwd max diff %.3e (sample %d, |g| max %.3e), "
          "samples over 1e-3: %d | full grad max diff %.3e of %.3e"
          % (r, torch.equal(out["verts"], out0["verts"]), torch.equal(out["projects"], out0["projects"]),
             torch.equal(out["mask"], out0["mask"]), torch.equal(out["seg"], out0["seg"]), float(d.max()), worst,
             float(gs0[worst].abs().max()), int((d > 1e-3).sum()), float((gf - gf0).abs().max()), float(gf0.abs().max())))

# decode backward alone: a fixed gradient on the projections
gp = torch.randn_like(pr)


def dec_grad():
    x = x0.clone().requires_grad_(True)
    dec(x, seg=False)["projects"].backward(gp)
    return x.grad


gd0 = dec_grad()
for r in range(args.reps):
    gd = dec_grad()
    d = (gd - gd0).abs()
    i = int(d.argmax())
    print("decode bwd rep %d: max diff %.3e at (sample %d, param %d), value %.3e; samples differing > 1e-4 rel: %d"
          % (r, float(d.max()), i // 86, i % 86, float(gd0.flatten()[i]),
             int(((d.amax(dim=1)) > 1e-4 * gd0.abs().amax(dim=1)).sum())))
