#!/usr/bin/env python
"""Times the mesh visualiser (SMPLRenderer, renderer.py:23-115) on the GPU: batches of 224 x 224 lit renders of decoded
meshes, CUDA events, per-kernel times from the library's profiler.  Prints one JSON line."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    pkg = importlib.import_module("indirect_learning_pose-shape_b200")
    synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
    dev = torch.device("cuda", 0)
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    layer = pkg.SMPLLayer(host, device=dev)
    params = torch.as_tensor(synth.make_params(a.batch, 48, seed=3), device=dev)
    verts = layer(params).detach()
    verts = verts * torch.tensor([1.0, -1.0, -1.0], device=dev) + torch.tensor([0.0, 0.0, 2.6], device=dev)   # camera frame
    R = pkg.SMPLRenderer(img_size=a.size, device=dev)
    for _ in range(a.warmup):
        R(verts, as_tensor=True)
    torch.cuda.synchronize()
    pkg.profile_enable(True)
    pkg.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        img = R(verts, as_tensor=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    prof = pkg.profile_collect()
    pkg.profile_enable(False)
    cover = float((img != 255).any(-1).float().mean())
    host_img = R(verts[0])                                                       # one call as the reference makes it
    print(json.dumps({"metric": "mesh visualiser images/s (%dx%d, lit, 13776 faces)" % (a.size, a.size),
                      "value": a.batch / ms * 1e3, "unit": "images/s", "batch": a.batch, "ms_per_batch": ms,
                      "us_per_image": ms / a.batch * 1e3, "coverage": cover,
                      "kernels_ms": {k: v[1] / max(v[0], 1) for k, v in prof.items()},
                      "single_call_shape": list(host_img.shape)}))


if __name__ == "__main__":
    main()
