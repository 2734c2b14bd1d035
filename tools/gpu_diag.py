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
for n, vs in ((8, None), (70, 5), (300, None)):
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

# ---- gradient accuracy of the decode backward (tensor-core path at N >= 64) vs fp64 / fp32 oracles -------------------
from oracle import torch_oracle  # noqa: E402

for n, vs in ((8, 5), (70, 5), (256, 5), (128, None)):
    rng = np.random.default_rng(n)
    p = synth.make_params(n, 48, seed=20 + n)
    Vs = -(-6890 // (vs or 1))
    w_p = rng.standard_normal((n, Vs, 3)).astype(np.float32)
    grads = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        C = torch_oracle.TorchSmplConstants(host, dt)
        x = torch.tensor(p, dtype=dt, requires_grad=True)
        o = torch_oracle.smpl_layer_call(C, x)
        pr = torch_oracle.orthographic_project([o, x], vs)
        (pr * torch.tensor(w_p, dtype=dt)).sum().backward()
        grads[name] = x.grad.numpy().astype(np.float64)
    dec = pkg.SmplDecoder(host, 48, vs, device=dev)
    x = torch.as_tensor(p, device=dev).requires_grad_(True)
    out = dec(x, seg=False)
    (out["projects"] * torch.as_tensor(w_p, device=dev)).sum().backward()
    got = x.grad.cpu().numpy().astype(np.float64)
    scale = np.abs(grads["f64"]).max(axis=0, keepdims=True) + 1e-6
    e_k = np.abs(got - grads["f64"]) / scale
    e_o = np.abs(grads["f32"] - grads["f64"]) / scale
    print("grad N=%d vs=%s: kernel-f64 max %.2e (cam %.2e pose %.2e shape %.2e)   f32oracle-f64 max %.2e"
          % (n, vs, e_k.max(), e_k[:, :4].max(), e_k[:, 4:76].max(), e_k[:, 76:].max(), e_o.max()))
