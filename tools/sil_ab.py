#!/usr/bin/env python
"""A/B of silhouette-kernel builds: tools/sil_ab.py [alternative.so] [--batch N]  (one library per process).
One JSON line: checksums of projects_to_silhouette's output and gradient at 256x256 from the decoder's full-resolution
projections (to compare builds with each other) and the mean sil_fwd / sil_bwd durations from the event profiler."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if args:
    binding.LIB_PATH = os.path.abspath(args[0])
N = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 2048
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
dev = torch.device("cuda", 0)
wh = 256
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
dec = pkg.SmplDecoder(host, wh, None, device=dev)
x0 = torch.as_tensor(synth.make_params(N, wh, seed=0), device=dev)
with torch.no_grad():
    pr = dec(x0, seg=False)["projects"]
g = torch.randn((N, wh, wh, 2), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
x = pr.clone().requires_grad_(True)
out = pkg.projects_to_silhouette(x, wh)
out.backward(g)
res = {"lib": os.path.basename(binding.LIB_PATH), "batch": N, "fwd_sum": float(out.double().sum()),
       "fwd_bits": int(out.view(torch.int32).to(torch.int64).sum().item() & 0xffffffffffff),
       "grad_abs_sum": float(x.grad.double().abs().sum())}
del out
for _ in range(2):
    x.grad = None
    pkg.projects_to_silhouette(x, wh).backward(g)
pkg.profile_enable(True); pkg.profile_collect()
for _ in range(5):
    x.grad = None
    pkg.projects_to_silhouette(x, wh).backward(g)
torch.cuda.synchronize()
pkg.profile_enable(False)
st = pkg.profile_collect()
res["sil_fwd_ms"] = round(st["sil_fwd"][1] / st["sil_fwd"][0], 4)
res["sil_bwd_ms"] = round(st["sil_bwd"][1] / st["sil_bwd"][0], 4)
print(json.dumps(res))
