#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its named config, one JSON line.

Workload (config C5): the full decoder forward+backward -- params (B,86) -> SMPL vertices ->
projection (vertex_sampling=5) -> visibility mask -> 31-part 48x48 segmentation, then d(loss)/d(params) from an
upstream gradient of the segmentation's shape -- on B = 16384 samples per GPU.  A "step" is one such pass.  The
batch shards across ranks with no data-path collective (samples are independent), so per-GPU work is fixed as N grows:
`"scaling": "weak"`; `--scaling strong` shards ONE global batch of 16384 instead (BASELINE config 5's wording).

  value        samples/s over all ranks, inputs resident in HBM, timed on the device with CUDA events, max over ranks
  e2e          the same step driven from pinned HOST params with the gradient read back to the host every step
  roofline     dominant kernel: algorithmic bytes per launch / its mean duration (library's event profiler, live in the
               timed region) against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline the oracle port (torch-CPU twin of the reference's brute-force algorithm) on a bounded sample, rank 0

`--impl reference` times that CPU port alone as the reference arm (the reference itself is Python 2.7 / TF 1.x and
cannot run here).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "indirect_learning_pose-shape_b200"

METRIC = "SMPL decode+project+part-seg samples/s (fwd+bwd, 48x48x32 seg, vertex_sampling=5)"
UNIT = "samples/s"
IMG_WH, VS, V, VS_COUNT, PARTS = 48, 5, 6890, 1378, 31
# algorithmic bytes per sample: params in, every reference-visible output written once, upstream
# gradient read once, param gradient written once; constants amortised to zero
BYTES_FWD = 344 + V * 12 + VS_COUNT * 12 + VS_COUNT * 4 + IMG_WH * IMG_WH * 32 * 4          # 399,984
BYTES_BWD = IMG_WH * IMG_WH * 32 * 4 + 344 + 344                                            # 295,600
BYTES_STEP = BYTES_FWD + BYTES_BWD                                                          # 695,584
SEG_BYTES = IMG_WH * IMG_WH * 32 * 4
LD = (V + 255) // 256 * 768
# per-kernel algorithmic bytes per sample: tensors the kernel must read + write once
KERNEL_BYTES = {
    "pose_fwd": 344 + 224 * 4 + 24 * 12 * 4 + 24 * 3 * 4,
    "blend_fwd": 224 * 4 + LD * 4,
    "lbs_fwd": LD * 4 + 24 * 12 * 4 + V * 12 + VS_COUNT * 12,
    "mask": VS_COUNT * 12 + VS_COUNT * 4,
    "seg_fwd": VS_COUNT * 12 + VS_COUNT * 4 + SEG_BYTES,
    "seg_bwd": SEG_BYTES + VS_COUNT * 12 + VS_COUNT * 4 + VS_COUNT * 12,
    "lbs_bwd_vertex": VS_COUNT * 12 + VS_COUNT * 12 + 24 * 12 * 4 + VS_COUNT * 12,
    "lbs_bwd_joint": VS_COUNT * 12 + VS_COUNT * 12 + 24 * 12 * 4,
    "blend_bwd": VS_COUNT * 12 + 224 * 4,
    "pose_bwd": 344 + 24 * 12 * 4 + 224 * 4 + 344,
}


# DRAM bytes per sample actually moved by each kernel (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full
# capture: the seg kernels from profiles/r1_seg_v32_ncu.txt (tools/prof_step.py --batch 2048), the others from
# profiles/r1_v22_dram_traffic.csv (--batch 4096)); scaled by the batch for `roofline.traffic`
KERNEL_TRAFFIC = {
    "seg_fwd": 45.2e3 + 371.0e3, "seg_bwd": 408.6e3 + 26.1e3, "lbs_fwd": 84.3e3 + 85.8e3, "blend_fwd": 29.1e3 + 72.1e3,
    "lbs_bwd_vertex": 100.5e3 + 28.4e3, "blend_bwd": 35.1e3 + 0.0e3, "mask": 16.5e3 + 0.3e3, "pose_fwd": 0.3e3,
    "pose_bwd": 2.4e3,
}


RAMP_S = 1.0   # untimed clock-ramp period ahead of the warm-up steps (seconds)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"        # /opt/skills/guides/B200_PROFILING.md fallback


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            for prefix in ("nvmlClocksEventReason", "nvmlClocksThrottleReason"):
                if hasattr(nv, prefix + n):
                    names[getattr(nv, prefix + n)] = n
                    break
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    bits = int(get_reasons(self.h))
                    for bit, name in names.items():
                        if bits & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join(1.0)

    def summary(self):
        if not self.samples:
            try:   # fall back to one nvidia-smi query
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [int(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi idle"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (brute force, as the reference evaluates it)
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_step(C, parts, p, g, torch, torch_oracle):
    x = torch.tensor(p, requires_grad=True)
    out = torch_oracle.decode(C, x, IMG_WH, VS, parts)
    (out["seg"] * g).sum().backward()
    return x.grad


def run_cpu_port(sample: int, steps: int, warmup: int):
    import torch
    from oracle import torch_oracle
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    torch.set_num_threads(os.cpu_count() or 1)
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    parts = pkg.smpl_io.golden_part_vertices(VS)
    C = torch_oracle.TorchSmplConstants(host, torch.float32)
    p = synth.make_params(sample, IMG_WH, seed=0)
    g = torch.randn(sample, IMG_WH, IMG_WH, PARTS + 1, generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        cpu_port_step(C, parts, p, g, torch, torch_oracle)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(C, parts, p, g, torch, torch_oracle)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample / dt, dt, torch.get_num_threads()


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    sample = args.cpu_sample
    value, dt, cores = run_cpu_port(sample, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 full decoder fwd+bwd, 31-part 48x48 seg, 5_sampled_part_vertices",
                   "per_step_sample": sample, "note": "reference = CPU port of the reference algorithm (oracle/"
                   "torch_oracle.py, brute-force O(wh^2 V) rasteriser as in projects_to_seg.py:41-56); the TF1/py2 "
                   "reference cannot run in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d samples fwd+bwd per step" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr: stdout carries ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="samples per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=16, help="samples per CPU-port step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seg-only", action="store_true", help="do not materialise the 6890-vertex mesh (590,856 B/sample)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    lib = importlib.import_module(PKG + "._lib")
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    parts = pkg.smpl_io.golden_part_vertices(VS)

    if args.scaling == "weak":
        global_batch = args.batch * world
    else:
        global_batch = args.batch
    lo, hi = pkg.shard_bounds(global_batch, rank, world)
    B = hi - lo
    # every rank derives its slice from the same seeded stream: generate in chunks to bound host memory
    params_np = synth.make_params(global_batch, IMG_WH, seed=0)[lo:hi]
    params_host = torch.from_numpy(np.ascontiguousarray(params_np)).pin_memory()
    params_dev = params_host.to(dev)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    g_seg = torch.randn((B, IMG_WH, IMG_WH, PARTS + 1), device=dev, generator=gen)      # upstream gradient ~ N(0,1)
    grad_host = torch.empty((B, 86), dtype=torch.float32).pin_memory()
    dec = pkg.SmplDecoder(host, IMG_WH, VS, need_verts=not args.seg_only, parts=parts, device=dev)

    def step(x_dev):
        x = x_dev.detach().requires_grad_(True)
        out = dec(x)
        out["seg"].backward(g_seg)
        return x.grad

    def step_e2e():
        x = params_host.to(dev, non_blocking=True)                   # H2D of this step's inputs (pinned)
        g = step(x)
        grad_host.copy_(g, non_blocking=True)                        # D2H of this step's result
        return g

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # SM clocks take a few hundred ms of load to ramp from idle (three 8 ms warm-up steps measured 3 - 6 % slow): run untimed
    # steps for RAMP_S seconds of wall clock first, then the W warm-up steps the caller asked for.
    t_ramp = time.time()
    while time.time() - t_ramp < RAMP_S:
        step(params_dev)
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(params_dev)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs; the library's event profiler rides along --------------
    launches0 = lib.launch_count()
    lib.profile_enable(True)
    lib.profile_collect()
    with ClockSampler(local_rank) as clocks:
        ms_total = timed(lambda: step(params_dev), args.steps)
    lib.profile_enable(False)
    kstats = lib.profile_collect()
    gpu_launches = lib.launch_count() - launches0
    ms_step = ms_total / args.steps
    value = global_batch / (ms_step * 1e-3)

    # ---- e2e: host params in, host gradient out, every step ----------------------------------------------------
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = global_batch / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    peak, peak_src = load_peaks()
    per_kernel = {k: {"launches": n, "ms_per_launch": t / n, "share": t / max(sum(v[1] for v in kstats.values()), 1e-9)}
                  for k, (n, t) in kstats.items() if n}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"]) if per_kernel else None
    roofline = None
    if dom:
        bytes_launch = KERNEL_BYTES.get(dom, 0) * B
        achieved = bytes_launch / (per_kernel[dom]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (KERNEL_TRAFFIC[dom] * B if (dom in KERNEL_TRAFFIC and not args.seg_only) else None),
                    "traffic_source": "ncu dram__bytes_read+write per sample (profiles/r1_seg_v32_ncu.txt, r1_v22_dram_traffic.csv) x batch",
                    "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": per_kernel[dom]["ms_per_launch"]}
    step_bytes = (BYTES_STEP if not args.seg_only else BYTES_STEP - V * 12 - VS_COUNT * 16)
    step_frac = step_bytes * (B / (ms_step * 1e-3)) / 1e9 / peak

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = run_cpu_port(args.cpu_sample, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d samples fwd+bwd per step, 2 steps after 1 warm-up (oracle/torch_oracle.py)" % args.cpu_sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 full decoder fwd+bwd: SMPL decode (6890 verts) -> projection (vertex_sampling=5)"
                                   " -> visibility mask -> 31-part 48x48 seg, 5_sampled_part_vertices",
                       "per_gpu_batch": B, "global_batch": global_batch, "img_wh": IMG_WH, "vertex_sampling": VS,
                       "materialise_verts": not args.seg_only, "smpl_model": "seeded synthetic (real pkl not shipped)",
                       "l2": "per-step working set %.1f GB >> 126 MB L2 (inputs larger than L2)" % (step_bytes * B / 1e9),
                       "clock_ramp_s": RAMP_S,
                       "parallelism": "batch shards, no collective"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * 86 * 4 * world,
                    "d2h_bytes_per_step": B * 86 * 4 * world},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline,
            "step_roofline": {"algorithmic_bytes_per_sample": step_bytes, "frac_of_hbm_peak": step_frac,
                              "achieved_gbs": step_frac * peak},
            "kernels": per_kernel,
            "cpu_baseline": cpu,
        }
        _emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
