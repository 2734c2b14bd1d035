"""world_size-2 gloo test of the multi-GPU host logic: contiguous batch slices, no collective on the decode path,
all-gather only to assemble outputs for validation."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "indirect_learning_pose-shape_b200"


def test_shard_bounds_cover_batch(pkg):
    for n in (0, 1, 7, 16384, 16385):
        for w in (1, 2, 3, 8):
            b = [pkg.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert pkg.shard_bounds(16384, 3, 8) == (6144, 8192)
    with pytest.raises(ValueError):
        pkg.shard_bounds(8, 2, 2)


def _worker(rank, world, port, n_global, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = importlib.import_module(PKG)
        synth = importlib.import_module(PKG + ".synth")
        from oracle import np_oracle
        host = pkg.smpl_io.make_synthetic_smpl(seed=0)
        # every rank generates the same global batch from the same seed and takes its slice
        params = synth.make_params(n_global, 48, seed=0)
        mine = pkg.shard_slice(torch.from_numpy(params), rank, world)
        lo, hi = pkg.shard_bounds(n_global, rank, world)
        assert mine.shape[0] == hi - lo
        # stand-in for the per-rank decode (the CUDA path cannot run in this container): the oracle, per sample
        local = torch.from_numpy(np_oracle.smpl_layer_call(host, mine.numpy())[:, ::100].copy())
        full = pkg.all_gather_outputs(local, n_global)
        if rank == 0:
            ref = np_oracle.smpl_layer_call(host, params)[:, ::100]
            q.put(float(np.abs(full.numpy() - ref).max()))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]   # ragged: 3 + 2 samples
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-6          # per-sample results do not depend on how the batch was split
