#!/usr/bin/env python
"""A/B of seg-kernel builds: tools/seg_ab.py [path/to/alternative.so]  (one library per process).

Prints one JSON line: score / label agreement of projects_to_seg with the NumPy oracle on the oracle's own projections
(48 samples), checksums of the forward output, the saved-state-driven gradient at N = 2048 (to compare builds with each
other), and the mean seg_fwd / seg_bwd durations at N = 16384 from the library's event profiler."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
if len(sys.argv) > 1:
    binding.LIB_PATH = os.path.abspath(sys.argv[1])
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
from oracle import np_oracle  # noqa: E402

dev = torch.device("cuda", 0)
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
parts = pkg.smpl_io.golden_part_vertices(5)
n, wh = 48, 48
p = synth.make_params(n, wh, seed=4242)
pr = np_oracle.orthographic_project([np_oracle.smpl_layer_call(host, p), p], 5)
mk = np_oracle.compute_mask(pr)
ref = np_oracle.projects_to_seg([pr, mk], wh, 5, parts)
got = pkg.projects_to_seg([torch.as_tensor(pr, device=dev), torch.as_tensor(mk, device=dev)], wh, 5, parts=parts).cpu().numpy()
res = {"lib": os.path.basename(binding.LIB_PATH), "max_abs_dscore": float(np.abs(got - ref).max()),
       "label_mismatch_rate": float((got.argmax(-1) != ref.argmax(-1)).mean())}
N = 16384
x = torch.as_tensor(synth.make_params(N, wh, seed=0), device=dev)
dec = pkg.SmplDecoder(host, wh, 5, parts=parts, device=dev, need_verts=False)
with torch.no_grad():
    out = dec(x)
    prj, msk = out["projects"], out["mask"]
g = torch.randn((N, wh, wh, 32), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
xg = prj.clone().requires_grad_(True)
seg = pkg.projects_to_seg([xg, msk], wh, 5, parts=parts)
seg.backward(g)
res["fwd_sum"] = float(seg.double().sum())
res["fwd_sha"] = int(seg.view(torch.int32).to(torch.int64).sum().item() & 0xffffffffffff)
res["grad_abs_sum"] = float(xg.grad.double().abs().sum())
del seg
for _ in range(3):
    xg.grad = None
    pkg.projects_to_seg([xg, msk], wh, 5, parts=parts).backward(g)
pkg.profile_enable(True); pkg.profile_collect()
for _ in range(10):
    xg.grad = None
    pkg.projects_to_seg([xg, msk], wh, 5, parts=parts).backward(g)
torch.cuda.synchronize()
pkg.profile_enable(False)
st = pkg.profile_collect()
res["seg_fwd_ms"] = round(st["seg_fwd"][1] / st["seg_fwd"][0], 4)
res["seg_bwd_ms"] = round(st["seg_bwd"][1] / st["seg_bwd"][0], 4)
print(json.dumps(res))
