#!/usr/bin/env python
"""Secondary BASELINE.json configs (bench.py stays the contract bench for C5).  One JSON line per config.

  c2  single-sample latency, decode + projection + mask + 31-part 48x48 seg forward (vertex_sampling=None), CUDA graph,
      p50 / p99 over --reps launches (predict_realtime-shaped)
  c3  LBS + orthographic projection fwd+bwd, batch 4096, fp32 (outputs verts + projects, gradient from projects)
  c4  silhouette 256x256 fwd+bwd from projections (N,6890,3), batch --c4-batch (BASELINE: 8192)
  f1   fused softmax + categorical focal loss fwd+bwd on a (16384,48,48,32) segmentation, uint8 labels
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, steps, warmup):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c3,c4")
    ap.add_argument("--reps", type=int, default=2000)
    ap.add_argument("--c4-batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    for cfg in args.configs.split(","):
        if cfg == "c2":
            dec = pkg.SmplDecoder(host, 48, None, parts=pkg.smpl_io.golden_part_vertices(None), device=dev)
            x = torch.as_tensor(pkg.smpl_io.mean_param_vector(48).astype(np.float32), device=dev)
            with torch.no_grad():
                for _ in range(3):
                    dec(x)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = dec(x)
                lat = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); e1.synchronize()
                    lat.append(e0.elapsed_time(e1) * 1e3)
                eager = timed(lambda: dec(x), 200, 20) * 1e3
                pkg.profile_enable(True); pkg.profile_collect()
                for _ in range(50):
                    dec(x)
                pkg.profile_enable(False)
                kern = {k: round(t / n * 1e3, 2) for k, (n, t) in pkg.profile_collect().items()}
            lat = np.sort(np.asarray(lat))
            print(json.dumps({"config": "C2 single-sample latency (N=1, wh=48, vertex_sampling=None, forward, CUDA graph)",
                              "p50_us": float(lat[len(lat) // 2]), "p99_us": float(lat[int(len(lat) * 0.99)]),
                              "min_us": float(lat[0]), "eager_us_per_call": eager, "reps": args.reps, "kernel_us": kern,
                              "labels_checksum": int(out["seg"].argmax(-1).sum().item())}))
        elif cfg == "c3":
            n = 4096
            dec = pkg.SmplDecoder(host, 48, None, device=dev)
            p = torch.as_tensor(synth.make_params(n, 48, seed=0), device=dev)
            gp = torch.randn((n, 6890, 3), device=dev)

            def step():
                x = p.detach().requires_grad_(True)
                o = dec(x, seg=False)
                o["projects"].backward(gp)
            ms = timed(step, args.steps, 3)
            pkg.profile_enable(True); pkg.profile_collect()
            for _ in range(5):
                step()
            pkg.profile_enable(False)
            kern3 = {k: round(t / n, 3) for k, (n, t) in pkg.profile_collect().items()}
            b = 249072
            print(json.dumps({"config": "C3 LBS + projection fwd+bwd, N=4096, vs=None", "ms_per_step": ms,
                              "samples_per_s": n / ms * 1e3, "alg_bytes_per_sample": b, "kernel_ms": kern3,
                              "frac_of_hbm_peak": b * n / (ms * 1e-3) / 1e9 / PEAK}))
        elif cfg == "c4":
            n, wh = args.c4_batch, 256
            p = synth.make_params(64, wh, seed=0)
            dec = pkg.SmplDecoder(host, wh, None, device=dev)
            with torch.no_grad():
                pr64 = dec(torch.as_tensor(p, device=dev), seg=False)["projects"]
            pr = pr64.repeat((n + 63) // 64, 1, 1)[:n].contiguous()
            gs = torch.randn((n, wh, wh, 2), device=dev)

            def step():
                x = pr.detach().requires_grad_(True)
                pkg.projects_to_silhouette(x, wh).backward(gs)
            ms = timed(step, max(2, args.steps // 3), 2)
            b = 1296616
            print(json.dumps({"config": "C4 silhouette 256x256 fwd+bwd from projections, N=%d" % n, "ms_per_step": ms,
                              "samples_per_s": n / ms * 1e3, "alg_bytes_per_sample": b,
                              "frac_of_hbm_peak": b * n / (ms * 1e-3) / 1e9 / PEAK}))
        elif cfg == "f1":
            n, wh, C = 16384, 48, 32
            seg = torch.rand((n, wh * wh, C), device=dev)
            lab = torch.randint(0, C, (n, wh * wh), device=dev, dtype=torch.uint8)
            loss_fn = pkg.categorical_focal_loss(gamma=2.0, weight_classes=True, from_logits=True)
            gl = torch.full((n, wh * wh), 1.0 / (n * wh * wh), device=dev)

            def step():
                x = seg.detach().requires_grad_(True)
                loss_fn(lab, x).backward(gl)
            ms = timed(step, args.steps, 3)
            pkg.profile_enable(True); pkg.profile_collect()
            for _ in range(5):
                step()
            pkg.profile_enable(False)
            kernf = {k: round(t / n_, 3) for k, (n_, t) in pkg.profile_collect().items()}
            b = wh * wh * (C * 4 + 1 + 4) + wh * wh * (C * 4 + 1 + 4 + C * 4)      # fwd: scores + label + loss; bwd: + g_loss + g_seg
            print(json.dumps({"config": "F1 softmax + focal loss fwd+bwd, N=16384, 48x48x32, uint8 labels", "ms_per_step": ms,
                              "samples_per_s": n / ms * 1e3, "alg_bytes_per_sample": b, "kernel_ms": kernf,
                              "frac_of_hbm_peak": b * n / (ms * 1e-3) / 1e9 / PEAK}))


if __name__ == "__main__":
    main()
