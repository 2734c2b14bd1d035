#!/usr/bin/env python
"""The decoder's TRAINING step (model.py:108-120 + focal_loss.py): params -> seg -> softmax -> focal loss -> d/d params,
unfused (seg kernel pair + loss kernel pair, the 48x48x32 scores and their gradient round-trip HBM) against fused
(projects_to_seg_focal_loss: the rasteriser evaluates the loss, the backward reads 16 bytes per pixel).  One JSON line."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def kernels(fn, reps=5):
    pkg.profile_enable(True); pkg.profile_collect()
    for _ in range(reps):
        fn()
    pkg.profile_enable(False)
    return {k: round(t / n, 4) for k, (n, t) in pkg.profile_collect().items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, wh, vs = a.batch, 48, 5
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    dec = pkg.SmplDecoder(host, wh, vs, need_verts=False, parts=pkg.smpl_io.golden_part_vertices(vs), device=dev)
    p = torch.as_tensor(synth.make_params(n, wh, seed=0), device=dev)
    lab = torch.randint(0, 32, (n, wh * wh), device=dev, dtype=torch.uint8)
    gl = torch.full((n, wh * wh), 1.0 / (n * wh * wh), device=dev)
    loss_fn = pkg.categorical_focal_loss(2.0, True, from_logits=True)

    def unfused():
        x = p.detach().requires_grad_(True)
        loss_fn(lab, dec(x)["seg"]).backward(gl)
        return x.grad

    def fused():
        x = p.detach().requires_grad_(True)
        dec.focal_loss(x, lab, 2.0, True)["loss"].backward(gl)
        return x.grad

    g_u, g_f = unfused().clone(), fused().clone()
    rel = float((g_u - g_f).abs().max() / g_u.abs().max())
    ms_u, ms_f = timed(unfused, a.steps), timed(fused, a.steps)
    # algorithmic bytes per sample of the fused step: params in, projects + mask out, labels in, loss out | g_loss in, g_params out
    b_f = 344 + 1378 * 12 + 1378 * 4 + wh * wh + wh * wh * 4 + wh * wh * 4 + 344 + 344
    b_u = b_f + 4 * wh * wh * 32 * 4          # + seg written, read by the loss, g_seg written, read by the seg backward
    print(json.dumps({"config": "training step: decode -> project -> mask -> seg -> softmax -> focal loss, fwd+bwd, N=%d, no mesh output" % n,
                      "unfused_ms": ms_u, "fused_ms": ms_f, "speedup": ms_u / ms_f, "grad_rel_diff": rel,
                      "unfused_samples_per_s": n / ms_u * 1e3, "fused_samples_per_s": n / ms_f * 1e3,
                      "fused_alg_bytes_per_sample": b_f, "unfused_tensor_bytes_per_sample": b_u,
                      "fused_frac_of_hbm_peak_on_unfused_bytes": b_u * n / (ms_f * 1e-3) / 1e9 / PEAK,
                      "kernel_ms_unfused": kernels(unfused), "kernel_ms_fused": kernels(fused)}))


if __name__ == "__main__":
    main()
