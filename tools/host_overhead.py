#!/usr/bin/env python
"""Host cost of one step against its GPU time at the shard sizes of the strong-scaling run (2048 .. 16384 samples per GPU):
eager modular chain, eager fused entry (smpl_b200_full_fwd/_bwd) and CUDA-graph replay.  One JSON line per batch."""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")


def main():
    dev = torch.device("cuda", 0)
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    parts = pkg.smpl_io.golden_part_vertices(5)
    for B in (2048, 4096, 8192, 16384):
        p = torch.as_tensor(synth.make_params(B, 48, seed=0), device=dev)
        g = torch.randn((B, 48, 48, 32), device=dev)
        res = {"batch": B}
        for name, fused in (("modular", False), ("fused", True)):
            dec = pkg.SmplDecoder(host, 48, 5, parts=parts, device=dev, fused=fused)

            def step():
                x = p.detach().requires_grad_(True)
                dec(x)["seg"].backward(g)
            for _ in range(5):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                step()
            t_issue = (time.perf_counter() - t0) / 20
            torch.cuda.synchronize()
            t_total = (time.perf_counter() - t0) / 20
            # host-only cost: issue one step into an idle GPU and stop the clock before waiting for it
            hs = []
            for _ in range(10):
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                step()
                hs.append(time.perf_counter() - t1)
                torch.cuda.synchronize()
            pkg.profile_enable(True); pkg.profile_collect()
            for _ in range(10):
                step()
            pkg.profile_enable(False)
            kern = {k: round(t / n_, 4) for k, (n_, t) in pkg.profile_collect().items()}
            res[name] = {"ms_per_step": t_total * 1e3, "host_issue_ms_back_to_back": t_issue * 1e3,
                         "host_ms_per_step_idle_gpu": sorted(hs)[len(hs) // 2] * 1e3, "kernel_ms": kern}
        dec = pkg.SmplDecoder(host, 48, 5, parts=parts, device=dev, fused=True)
        gs = pkg.GraphedDecoderStep(dec, B, device=dev)
        gs.params.copy_(p); gs.g_seg.copy_(g)
        for _ in range(5):
            gs.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            gs.replay()
        e1.record()
        torch.cuda.synchronize()
        res["graph"] = {"ms_per_step": e0.elapsed_time(e1) / 20, "launches_per_step": gs.launches_per_step}
        print(json.dumps(res), flush=True)
        del gs, dec, p, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
