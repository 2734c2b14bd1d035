#!/usr/bin/env python
"""Which of two nearly equidistant vertices does the seg forward's hot loop pick?  tools/q13_case.py [alternative.so]
Three (A, B, pixel) triples where tf.norm's fl(fl(du^2) + fl(dv^2)) makes A the nearer vertex and the fused form
fma(du, du, fl(dv^2)) makes B nearer (the same triples as tests/test_gpu_parity.py::test_seg_hot_loop_keeps_tf_norm_roundings).
Measured: the round-2 build before the fix (mul.rn.f32x2 + add.rn.f32x2 contracted to FFMA2 by ptxas) sends the pixel's
gradient to B in all three, the current build to A."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
if len(sys.argv) > 1: binding.LIB_PATH = os.path.abspath(sys.argv[1])
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
cases = [(16.4890193939209, 18.45115852355957, 17.89776611328125, 13.83559513092041),
         (14.508923530578613, 23.694658279418945, 11.48696517944336, 15.41930103302002),
         (24.387985229492188, 21.174949645996094, 14.013175010681152, 17.91790199279785)]
gx, gy, wh, Vs = 20, 17, 48, 33
pr = np.zeros((3, Vs, 3), np.float32); pr[:, 2:, 0] = 200.0 + np.arange(Vs - 2); pr[:, 2:, 1] = -150.0
for i, (ua, va, ub, vb) in enumerate(cases): pr[i, 0, :2] = (ua, va); pr[i, 1, :2] = (ub, vb)
mask = np.ones((3, Vs), np.float32); parts = [[0, 1]] + [[2 + k] for k in range(30)]
g = np.zeros((3, wh, wh, 32), np.float32); g[:, wh - 1 - gy, gx, 1] = 1.0
dev = torch.device("cuda", 0)
x = torch.as_tensor(pr, device=dev).requires_grad_(True)
out = pkg.projects_to_seg([x, torch.as_tensor(mask, device=dev)], wh, None, parts=parts)
(out * torch.as_tensor(g, device=dev)).sum().backward()
got = x.grad.cpu().numpy()
print(os.path.basename(binding.LIB_PATH), "gradient on A:", [bool(np.abs(got[i,0]).max() > 0) for i in range(3)], "on B:", [bool(np.abs(got[i,1]).max() > 0) for i in range(3)])
