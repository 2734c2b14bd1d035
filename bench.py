#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its named config, one JSON line.

Workload (config C5): the full decoder forward+backward -- params (B,86) -> SMPL vertices ->
projection (vertex_sampling=5) -> visibility mask -> 31-part 48x48 segmentation, then d(loss)/d(params) from an
upstream gradient of the segmentation's shape -- on ONE global batch of 16384 samples, sharded in contiguous slices
over the ranks (BASELINE.json config 5: "batch 16384 sharded across 2/4/8 B200"; the reference's analogue splits one
batch over towers, train.py:205-210).  Total work is fixed as N grows: `"scaling": "strong"`.  There is no collective
on the path (samples are independent); at N > 1 an UNTIMED validation leg all-gathers every rank's labels and parameter
gradients over NCCL and rank 0 compares them with the same batch decoded on its one GPU.
`--scaling weak` keeps 16384 samples per GPU instead.

  value        samples/s over all ranks, inputs resident in HBM, timed on the device with CUDA events, max over ranks
  e2e          the same step driven from pinned HOST params with the gradient read back to the host every step
  roofline     dominant kernel: algorithmic bytes per launch / its mean duration (library's event profiler: a CUDA-event
               pair on the launching stream around every launch) against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline the oracle port (torch-CPU twin of the reference's brute-force algorithm) on a bounded sample, rank 0

The step is captured once as a CUDA graph (GraphedDecoderStep: the public SmplDecoder forward + backward, the per-GPU
batch cut into two slices on two streams inside the graph) and the timed region replays it; the per-kernel durations then
come from an eager pass of the same K steps right after, with the library's event profiler (`--graph off`: the eager API
in the timed region itself, profiler riding along).

`--config c2|c3|c4` runs BASELINE.json's secondary configurations with the same contract keys.
`--impl reference` times the CPU port alone as the reference arm (the reference itself is Python 2.7 / TF 1.x and cannot
run here).
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
C3_BYTES, C4_BYTES = 249072, 1296616


def load_kernel_traffic():
    """DRAM bytes per sample actually moved by each kernel: dram__bytes_read.sum + dram__bytes_write.sum of an
    `ncu --set full` capture of this build (tools/prof_step.py), reduced by tools/ncu_traffic.py into
    profiles/kernel_traffic.json ({"source": ..., "batch": ..., "bytes_per_sample": {kernel: bytes}})."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bytes_per_sample", {}), d.get("source", path)
    return {}, None


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


def run_cpu_c1(reps: int = 2):
    """BASELINE config C1: the reference's own CPU-runnable case -- forward only, batch 1, the shipped neutral mean
    parameters, vertex_sampling=None, 48x48 31-part seg -- on the NumPy port (np_oracle.decode, literal per-pixel mask)."""
    import numpy as np
    from oracle import np_oracle
    pkg = importlib.import_module(PKG)
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    parts = pkg.smpl_io.golden_part_vertices(None)
    p = pkg.smpl_io.mean_param_vector(IMG_WH).astype(np.float32)

    def fwd():
        verts = np_oracle.smpl_layer_call(host, p)
        pwd = np_oracle.orthographic_project([verts, p], None)
        mask = np_oracle.compute_mask(pwd, fast=False)               # one pass per pixel of the 64x64 grid, as compute_mask.py:56-60
        return np_oracle.projects_to_seg([pwd, mask], IMG_WH, None, parts)
    fwd()
    t0 = time.perf_counter()
    for _ in range(reps):
        fwd()
    dt = (time.perf_counter() - t0) / reps
    return {"config": "C1 forward, batch 1, neutral mean params, vertex_sampling=None, 48x48 31-part seg",
            "ms_per_sample": dt * 1e3, "value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d forward passes after 1 warm-up (oracle/np_oracle.py)" % reps}


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    sample = args.cpu_sample
    value, dt, cores = run_cpu_port(sample, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 full decoder fwd+bwd, 31-part 48x48 seg, 5_sampled_part_vertices",
                   "per_step_sample": sample, "note": "reference = CPU port of the reference algorithm (oracle/"
                   "torch_oracle.py, brute-force O(wh^2 V) rasteriser as in projects_to_seg.py:41-56, pinned to the "
                   "reference's own source by tests/test_reference_pin.py); the TF1/py2 reference cannot run in this image"},
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


class Harness:
    """Timing plumbing shared by the configs: barrier + synchronize on both sides, CUDA events, max over ranks."""

    def __init__(self, torch, dist, dev, distributed):
        self.torch, self.dist, self.dev, self.distributed = torch, dist, dev, distributed

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """finish: called after the last step and before the closing event (e.g. join side streams into the timed one)"""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.distributed:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def ramp(self, fn, warmup):
        # SM clocks take a few hundred ms of load to ramp from idle (three 8 ms warm-up steps measured 3 - 6 % slow): run
        # untimed steps for RAMP_S seconds of wall clock first, then the W warm-up steps the caller asked for.
        t0 = time.time()
        while time.time() - t0 < RAMP_S:
            fn()
            self.torch.cuda.synchronize()
        for _ in range(warmup):
            fn()
        self.torch.cuda.synchronize()


def kernel_table(kstats):
    tot = max(sum(v[1] for v in kstats.values()), 1e-9)
    return {k: {"launches": n, "ms_per_launch": t / n, "share": t / tot} for k, (n, t) in kstats.items() if n}


def run_c5(args, rank, world, local_rank, torch, dist, dev, h: Harness):
    import numpy as np
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    lib = importlib.import_module(PKG + "._lib")
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    parts = pkg.smpl_io.golden_part_vertices(VS)
    distributed = world > 1

    global_batch = args.batch * world if args.scaling == "weak" else args.batch
    lo, hi = pkg.shard_bounds(global_batch, rank, world)
    B = hi - lo
    # every rank derives its slice from the same seeded streams: the global batch does not depend on how it is sharded
    params_all = synth.make_params(global_batch, IMG_WH, seed=0)
    params_host = torch.from_numpy(np.ascontiguousarray(params_all[lo:hi])).pin_memory()
    params_dev = params_host.to(dev)

    def upstream(lo_, hi_):
        """rows [lo_, hi_) of the global upstream gradient ~ N(0,1): one Philox stream per 1024-sample block, so any slice
        of the batch sees the same values whichever rank generates it"""
        out = torch.empty((hi_ - lo_, IMG_WH, IMG_WH, PARTS + 1), device=dev)
        blk = 1024
        for b0 in range(lo_ // blk * blk, hi_, blk):
            gen = torch.Generator(device=dev).manual_seed(1000 + b0 // blk)
            t = torch.randn((blk, IMG_WH, IMG_WH, PARTS + 1), device=dev, generator=gen)
            s0, s1 = max(b0, lo_), min(b0 + blk, hi_)
            out[s0 - lo_:s1 - lo_] = t[s0 - b0:s1 - b0]
        return out

    g_seg = upstream(lo, hi)
    grad_host = torch.empty((B, 86), dtype=torch.float32).pin_memory()
    dec = pkg.SmplDecoder(host, IMG_WH, VS, need_verts=not args.seg_only, parts=parts, device=dev, fused=not args.no_fused)

    def step(x_dev):
        x = x_dev.detach().requires_grad_(True)
        out = dec(x)
        out["seg"].backward(g_seg)
        return x.grad, out

    use_graph = args.graph in ("on", "auto")
    graphed = None
    if use_graph:
        # two slices on separate streams inside the graph: one slice's partly filled last wave of blocks (and its
        # issue-bound seg kernels) run beside the other slice's kernels.  Measured at 2048 / 4096 / 8192 / 16384 samples
        # per GPU: -9.0 / -4.9 / -4.3 / -1.5 % against one slice; 3 and 4 slices no better
        # (profiles/r2_v10_micro_batches.jsonl)
        mb = args.micro_batches if args.micro_batches > 0 else 2
        graphed = pkg.GraphedDecoderStep(dec, B, device=dev, micro_batches=mb)
        graphed.params.copy_(params_dev)
        graphed.g_seg.copy_(g_seg)
        g_seg = graphed.g_seg                                        # one copy of the 295 KB/sample upstream gradient

        def step_value():
            graphed.replay()

        # e2e: the same graph fed from pinned host memory every step.  Two graphed steps alternate (PipelinedDecoderSteps):
        # step k+1's parameters are uploaded on a copy stream while step k runs, step k's gradient is read back while step
        # k+1 runs; every step does its own H2D and D2H inside the timed region, the closing event waits for the last copies.
        pipe = None if args.no_pipeline else pkg.PipelinedDecoderSteps(dec, B, device=dev, micro_batches=mb, first=graphed)

        def step_e2e():
            if pipe is not None:
                pipe.step(params_host, grad_host)
                return
            graphed.params.copy_(params_host, non_blocking=True)     # H2D of this step's inputs (pinned)
            graphed.replay()
            grad_host.copy_(graphed.g_params, non_blocking=True)     # D2H of this step's result
    else:
        def step_value():
            step(params_dev)

        def step_e2e():
            x = params_host.to(dev, non_blocking=True)               # H2D of this step's inputs (pinned)
            g, _ = step(x)
            grad_host.copy_(g, non_blocking=True)                    # D2H of this step's result

    h.ramp(step_value, args.warmup)

    # ---- timed region: K steps, device-resident inputs -------------------------------------------------------------
    launches0 = lib.launch_count()
    if not use_graph:
        lib.profile_enable(True)
        lib.profile_collect()
    with ClockSampler(local_rank) as clocks:
        ms_total = h.timed(step_value, args.steps)
    ms_step = ms_total / args.steps
    value = global_batch / (ms_step * 1e-3)
    if use_graph:
        gpu_launches = graphed.launches_per_step * args.steps
        # per-kernel durations: an eager pass of the same K steps with the event profiler (graph nodes carry no events)
        for _ in range(2):
            step(params_dev)
        lib.profile_enable(True)
        lib.profile_collect()
        ms_eager = h.timed(lambda: step(params_dev), args.steps) / args.steps
        lib.profile_enable(False)
        kstats = lib.profile_collect()
    else:
        lib.profile_enable(False)
        kstats = lib.profile_collect()
        gpu_launches = lib.launch_count() - launches0
        ms_eager = ms_step

    # ---- e2e: host params in, host gradient out, every step ----------------------------------------------------
    for _ in range(2):
        step_e2e()
    e2e_join = pipe.join if (use_graph and pipe is not None) else None
    ms_e2e = h.timed(step_e2e, args.steps, finish=e2e_join) / args.steps
    if use_graph and pipe is not None:
        graphed.params.copy_(params_dev)                             # the validation / profiler legs below use the device copy
    e2e_value = global_batch / (ms_e2e * 1e-3)

    # ---- validation leg (untimed, N > 1): NCCL all-gather of every rank's labels and parameter gradients -----------------
    validation = None
    if distributed and not args.no_validate:
        g_local, out_local = step(params_dev)
        labels_local = out_local["seg"].argmax(-1).to(torch.uint8)
        t0 = time.time()
        labels_all = pkg.all_gather_outputs(labels_local, global_batch)
        gp_all = pkg.all_gather_outputs(g_local, global_batch)
        proj_all = pkg.all_gather_outputs(out_local["projects"], global_batch)
        torch.cuda.synchronize()
        t_gather = time.time() - t0
        del out_local
        if rank == 0:
            # the same global batch on ONE GPU, in chunks of the per-rank size (bounds memory; chunking is not sharding:
            # rank 0 decodes every sample itself)
            lab_ok, gp_err, gp_scale, proj_equal, n_bad = True, 0.0, 0.0, True, 0
            chunk = max(B, 1)
            for c0 in range(0, global_batch, chunk):
                c1 = min(c0 + chunk, global_batch)
                x = torch.from_numpy(np.ascontiguousarray(params_all[c0:c1])).to(dev).requires_grad_(True)
                o = dec(x)
                o["seg"].backward(upstream(c0, c1))
                lab = o["seg"].argmax(-1).to(torch.uint8)
                n_bad += int((lab != labels_all[c0:c1]).sum().item())
                proj_equal = proj_equal and bool(torch.equal(o["projects"].detach(), proj_all[c0:c1]))
                gp_err = max(gp_err, float((x.grad - gp_all[c0:c1]).abs().max().item()))
                gp_scale = max(gp_scale, float(x.grad.abs().max().item()))
                del o, x, lab
            lab_ok = n_bad == 0
            validation = {"collective": "nccl all_gather of labels (uint8), g_params and projections over %d ranks" % world,
                          "gather_s": t_gather, "labels_equal_to_single_gpu": lab_ok, "label_mismatches": n_bad,
                          "projections_bit_identical": proj_equal,
                          "g_params_max_abs_diff": gp_err, "g_params_max_abs": gp_scale,
                          "g_params_rel": gp_err / max(gp_scale, 1e-30),
                          "note": "gradient sums are formed in on-demand row order inside seg_bwd: last-bit differences only",
                          "ok": bool(lab_ok and gp_err <= 1e-4 * max(gp_scale, 1e-30))}

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    peak, peak_src = load_peaks()
    traffic, traffic_src = load_kernel_traffic()
    per_kernel = kernel_table(kstats)
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"]) if per_kernel else None
    roofline = None
    if dom:
        bytes_launch = KERNEL_BYTES.get(dom, 0) * B
        achieved = bytes_launch / (per_kernel[dom]["ms_per_launch"] * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (traffic[dom] * B if (dom in traffic and not args.seg_only) else None),
                    "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": per_kernel[dom]["ms_per_launch"],
                    "timed_in": "the timed region itself (eager, event profiler live)" if not use_graph else
                                "an eager pass of the same K steps after the graph-replay region (%.3f ms/step eager vs "
                                "%.3f replayed)" % (ms_eager, ms_step)}
    step_bytes = (BYTES_STEP if not args.seg_only else BYTES_STEP - V * 12 - VS_COUNT * 16)
    step_frac = step_bytes * (B / (ms_step * 1e-3)) / 1e9 / peak

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = run_cpu_port(args.cpu_sample, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d samples fwd+bwd per step, 2 steps after 1 warm-up (oracle/torch_oracle.py)" % args.cpu_sample,
               "c1": run_cpu_c1()}

    if rank == 0:
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 full decoder fwd+bwd: SMPL decode (6890 verts) -> projection (vertex_sampling=5)"
                                   " -> visibility mask -> 31-part 48x48 seg, 5_sampled_part_vertices",
                       "per_gpu_batch": B, "global_batch": global_batch, "img_wh": IMG_WH, "vertex_sampling": VS,
                       "materialise_verts": not args.seg_only, "smpl_model": "seeded synthetic (real pkl not shipped)",
                       "upstream_gradient": "g_seg ~ N(0,1), generated on the device before the timed region (a loss "
                                            "lives on the device); params in / g_params out are the host copies of e2e",
                       "l2": "per-step working set %.1f GB >> 126 MB L2 (inputs larger than L2)" % (step_bytes * B / 1e9),
                       "clock_ramp_s": RAMP_S, "api": "SmplDecoder(fused=%s)" % (not args.no_fused),
                       "cuda_graph": bool(use_graph), "micro_batches": graphed.micro_batches if graphed else 1,
                       "parallelism": "one 16384 batch in contiguous shards, no collective on the path" if
                       args.scaling == "strong" else "16384 samples per GPU, no collective on the path"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": global_batch * 86 * 4,
                    "d2h_bytes_per_step": global_batch * 86 * 4,
                    "copies": ("two alternating graphed steps; step k+1's H2D and step k's D2H run on copy streams beside the "
                               "kernels, every step does both inside the timed region") if (use_graph and pipe is not None)
                    else "H2D, step, D2H in order on one stream"},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline,
            "step_roofline": {"algorithmic_bytes_per_sample": step_bytes, "frac_of_hbm_peak": step_frac,
                              "achieved_gbs": step_frac * peak},
            "kernels": per_kernel,
            "validation": validation,
            "cpu_baseline": cpu,
        }
    return None


def run_secondary(args, rank, world, local_rank, torch, dist, dev, h: Harness):
    """BASELINE.json configs C2 / C3 / C4 (single GPU each; with N > 1 every rank runs a replica: no sharding claimed)."""
    import numpy as np
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    lib = importlib.import_module(PKG + "._lib")
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    peak, peak_src = load_peaks()
    cfg = args.config
    base = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "scaling": "weak"}
    if cfg == "c2":
        dec = pkg.SmplDecoder(host, IMG_WH, None, parts=pkg.smpl_io.golden_part_vertices(None), device=dev)
        x_host = torch.from_numpy(pkg.smpl_io.mean_param_vector(IMG_WH).astype(np.float32)).pin_memory()
        x = x_host.to(dev)
        lab_host = torch.empty((1, IMG_WH, IMG_WH), dtype=torch.uint8).pin_memory()
        with torch.no_grad():
            h.ramp(lambda: dec(x), args.warmup)
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g):
                out = dec(x)
                lab = out["seg"].argmax(-1).to(torch.uint8)
            launches = lib.launch_count() - n0
            reps = max(args.steps, 1000)
            lat = []
            with ClockSampler(local_rank) as clocks:
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); e1.synchronize()
                    lat.append(e0.elapsed_time(e1) * 1e3)
            lat = np.sort(np.asarray(lat))
            e2e = []
            for _ in range(200):                                     # host params in, host labels out, per call
                t0 = time.perf_counter()
                x.copy_(x_host, non_blocking=True); g.replay(); lab_host.copy_(lab, non_blocking=True)
                torch.cuda.synchronize()
                e2e.append((time.perf_counter() - t0) * 1e6)
            lib.profile_enable(True); lib.profile_collect()
            for _ in range(50):
                dec(x)
            lib.profile_enable(False)
            per_kernel = kernel_table(lib.profile_collect())
        p50 = float(lat[len(lat) // 2])
        line = dict(base, metric="single-sample decode+project+mask+seg latency (forward, CUDA graph)", value=p50,
                    unit="us", higher_is_better=False, ms_per_step=p50 / 1e3, steps=reps,
                    config={"workload": "C2 single-sample latency: N=1, wh=48, vertex_sampling=None, forward only "
                                        "(predict_realtime-shaped), one CUDA-graph replay per call"},
                    p50_us=p50, p99_us=float(lat[int(len(lat) * 0.99)]), min_us=float(lat[0]),
                    e2e={"value": float(np.median(e2e)), "unit": "us", "h2d_bytes_per_step": 344, "d2h_bytes_per_step": IMG_WH * IMG_WH,
                         "note": "wall clock per call incl. pinned H2D of params and D2H of the label image"},
                    gpu_launches=int(launches * reps), kernels=per_kernel, clocks=clocks.summary(), roofline=None)
        return line if rank == 0 else None
    if cfg == "c3":
        n = args.batch if args.batch != 16384 else 4096
        dec = pkg.SmplDecoder(host, IMG_WH, None, device=dev)
        p_host = torch.from_numpy(synth.make_params(n, IMG_WH, seed=0)).pin_memory()
        p = p_host.to(dev)
        gp = torch.randn((n, V, 3), device=dev)
        g_host = torch.empty((n, 86)).pin_memory()

        def step(xd=p):
            x = xd.detach().requires_grad_(True)
            o = dec(x, seg=False)
            o["projects"].backward(gp)
            return x.grad

        def step_e2e():
            g_host.copy_(step(p_host.to(dev, non_blocking=True)), non_blocking=True)
        alg, name = C3_BYTES, "C3 LBS + orthographic projection fwd+bwd, batch %d, vertex_sampling=None" % n
    elif cfg == "c4":
        n, wh = (args.batch if args.batch != 16384 else 8192), 256
        dec = pkg.SmplDecoder(host, wh, None, device=dev)
        with torch.no_grad():
            pr64 = dec(torch.as_tensor(synth.make_params(64, wh, seed=0), device=dev), seg=False)["projects"]
        pr = pr64.repeat((n + 63) // 64, 1, 1)[:n].contiguous()
        pr_host = pr.cpu().pin_memory()
        gs = torch.randn((n, wh, wh, 2), device=dev)
        g_host = torch.empty((n, V, 3)).pin_memory()

        def step(xd=pr):
            x = xd.detach().requires_grad_(True)
            pkg.projects_to_silhouette(x, wh).backward(gs)
            return x.grad

        def step_e2e():
            g_host.copy_(step(pr_host.to(dev, non_blocking=True)), non_blocking=True)
        alg, name = C4_BYTES, "C4 silhouette 256x256 fwd+bwd from projections (N,6890,3), batch %d" % n
    else:
        raise SystemExit("unknown --config %r" % cfg)
    h.ramp(step, args.warmup)
    n0 = lib.launch_count()
    lib.profile_enable(True); lib.profile_collect()
    with ClockSampler(local_rank) as clocks:
        ms = h.timed(step, args.steps) / args.steps
    lib.profile_enable(False)
    per_kernel = kernel_table(lib.profile_collect())
    launches = lib.launch_count() - n0
    step_e2e()
    ms_e2e = h.timed(step_e2e, max(2, args.steps // 2)) / max(2, args.steps // 2)
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"])
    frac = alg * n / (ms * 1e-3) / 1e9 / peak
    line = dict(base, metric=name.split(",")[0] + " samples/s", value=n * world / (ms * 1e-3), unit=UNIT, higher_is_better=True,
                ms_per_step=ms, config={"workload": name, "per_gpu_batch": n, "replicas": world},
                e2e={"value": n * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                     "h2d_bytes_per_step": int((p_host if cfg == "c3" else pr_host).numel() * 4),
                     "d2h_bytes_per_step": int(g_host.numel() * 4)},
                gpu_launches=int(launches), kernels=per_kernel, clocks=clocks.summary(),
                roofline={"kernel": "whole step (dominant kernel: %s, %.3f ms)" % (dom, per_kernel[dom]["ms_per_launch"]),
                          "bound": "hbm", "achieved": frac * peak, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                          "frac": frac, "traffic": None, "algorithmic_bytes_per_launch": alg * n})
    return line if rank == 0 else None


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c2", "c3", "c4"])
    ap.add_argument("--batch", type=int, default=16384, help="global batch (strong) or samples per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step as a CUDA graph in the timed region (auto = on; off: the eager API)")
    ap.add_argument("--micro-batches", type=int, default=0,
                    help="slices of the per-GPU batch captured on separate streams of the CUDA graph (0 = auto)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="e2e: copy, replay, copy back on one stream instead of the two-graph pipeline with copy streams")
    ap.add_argument("--no-fused", action="store_true", help="chain the modular ops instead of smpl_b200_full_fwd/_bwd")
    ap.add_argument("--no-validate", action="store_true", help="skip the NCCL all-gather validation leg at N > 1")
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
    h = Harness(torch, dist, dev, distributed)
    line = (run_c5 if args.config == "c5" else run_secondary)(args, rank, world, local_rank, torch, dist, dev, h)
    if rank == 0 and line is not None:
        _emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
