#!/usr/bin/env python
"""Where does the seg gradient's run-to-run difference come from: the forward's saved arg-min bytes, or the backward?"""
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
parts = pkg.smpl_io.golden_part_vertices(5)
dec = pkg.SmplDecoder(host, 48, 5, parts=parts, device=dev)
N = 16384
x0 = torch.as_tensor(synth.make_params(N, 48, seed=0), device=dev)
g = torch.randn((N, 48, 48, 32), device=dev)
with torch.no_grad():
    ref = dec(x0)
pr, mk = ref["projects"].clone(), ref["mask"].clone()
p = pr.clone().requires_grad_(True)
seg = pkg.projects_to_seg([p, mk], 48, 5, parts=parts)
saved0 = seg.grad_fn.saved_tensors[2].clone()
grads = []
for r in range(6):                       # the SAME saved bytes, backward repeated
    p.grad = None
    seg.backward(g, retain_graph=True)
    grads.append(p.grad.clone())
for r in range(1, 6):
    d = (grads[r] - grads[0]).abs().amax(dim=(1, 2))
    print("same saved, backward rep %d: max diff %.3e, samples over 1e-3: %s" % (r, float(d.max()), (d > 1e-3).nonzero().flatten().tolist()[:8]))
for r in range(4):                       # forward repeated: are the saved bytes reproducible?
    p2 = pr.clone().requires_grad_(True)
    s2 = pkg.projects_to_seg([p2, mk], 48, 5, parts=parts)
    sv = s2.grad_fn.saved_tensors[2]
    ne = (sv != saved0)
    print("forward rep %d: seg identical %s, saved bytes differing: %d (samples %s)"
          % (r, torch.equal(s2, seg), int(ne.sum()), (ne.view(N, -1).any(dim=1)).nonzero().flatten().tolist()[:8]))
