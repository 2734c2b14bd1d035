#!/usr/bin/env python
"""IEF regression module (model.py:63-97) at the decoder's batch: forward+backward time, tensor throughput, and the
step of regressor + decoder (features in HBM -> d loss / d weights).  One JSON line."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("indirect_learning_pose-shape_b200")


def timed(fn, steps=10, warmup=3):
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
    dev = torch.device("cuda", 0)
    n, wh = 16384, 48
    reg = pkg.IEFRegressor(wh, device=dev)
    feat = torch.rand((n, 2048), device=dev)
    g = torch.randn((n, 86), device=dev)

    def fwd():
        with torch.no_grad():
            reg(feat)

    def step():
        for p in reg.parameters():
            p.grad = None
        reg(feat).backward(g)

    ms_f, ms_s = timed(fwd), timed(step)
    macs = n * (2134 * 1024 + 1024 * 1024 + 1024 * 86) * 3            # three iterations
    flop_f, flop_s = 2 * macs, 2 * macs * 3                            # backward: two products per layer
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    dec = pkg.SmplDecoder(host, wh, 5, need_verts=False, parts=pkg.smpl_io.golden_part_vertices(5), device=dev, fused=True)
    gs = torch.randn((n, wh, wh, 32), device=dev)

    def full():
        for p in reg.parameters():
            p.grad = None
        dec(reg(feat))["seg"].backward(gs)
    ms_full = timed(full, 5)
    pkg.profile_enable(True); pkg.profile_collect()
    for _ in range(3):
        step()
    pkg.profile_enable(False)
    kern = {k: round(t / c, 4) for k, (c, t) in pkg.profile_collect().items()}
    print(json.dumps({"config": "IEF regressor (model.py:63-97), N=%d, 3 iterations, shared Dense 2134-1024-1024-86" % n,
                      "fwd_ms": ms_f, "fwd_bwd_ms": ms_s, "useful_tflops_fwd": flop_f / ms_f / 1e9,
                      "useful_tflops_fwd_bwd": flop_s / ms_s / 1e9, "tensor_tflops_fwd_bwd_3xTF32": 3 * flop_s / ms_s / 1e9,
                      "regressor_plus_decoder_fwd_bwd_ms": ms_full, "samples_per_s_full": n / ms_full * 1e3,
                      "dense_scope_ms_per_launch": kern}))


if __name__ == "__main__":
    main()
