"""Full-size runs (BASELINE.json configs C5 and C4) checked through size-independent properties.

The oracle evaluates O(wh^2 V) pairs per sample and cannot run 16384 samples in seconds, so at full size the CUDA
path is checked against ITSELF at a size the parity tests already pin to the oracle, and against identities of the
reference arithmetic:
  * batch invariance  -- every op of the path is per sample, so sample i of the 16384-batch must equal the
    same sample decoded in a batch of 64 (the regime tests/test_gpu_parity.py compares with the oracle);
  * linearity of the backward in the upstream gradient -- doubling g doubles every partial sum exactly;
  * channel identities -- seg[..., 0] = 1 - clip(sum_k seg[..., 1+k], 0, 1) (projects_to_seg.py:61-67),
    sil[..., 0] + sil[..., 1] = 1 (projects_to_silhouette.py:40-41).
"""
import numpy as np
import pytest
import torch

from oracle import np_oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def test_c5_full_batch_properties(pkg, host_model, parts_by_vs, make_params):
    N, wh, vs = 16384, 48, 5
    params = torch.as_tensor(make_params(N, wh, seed=0), device=dev())
    dec = pkg.SmplDecoder(host_model, wh, vs, parts=parts_by_vs[vs], device=dev())
    gen = torch.Generator(device=dev()).manual_seed(3)
    g = torch.randn((N, wh, wh, 32), device=dev(), generator=gen)

    x = params.clone().requires_grad_(True)
    out = dec(x)
    seg = out["seg"]
    seg.backward(g)
    grad1 = x.grad.clone()
    seg = seg.detach()

    # channel identity on every pixel of every sample
    s = seg[..., 1:].sum(-1)
    bg = 1.0 - s.clamp(0.0, 1.0)
    assert float((seg[..., 0] - bg).abs().max()) <= 4e-6
    assert float(seg.min()) >= 0.0 and float(seg.max()) <= 1.0
    assert bool(torch.isfinite(grad1).all())

    # batch invariance on a strided subset (forward outputs bit-identical: each sample is its own block)
    idx = torch.arange(0, N, N // 64, device=dev())[:64]
    xs = params[idx].clone().requires_grad_(True)
    outs = dec(xs)
    assert torch.equal(outs["projects"], out["projects"][idx])
    assert torch.equal(outs["mask"], out["mask"][idx])
    assert torch.equal(outs["seg"], seg[idx])
    assert float((outs["verts"] - out["verts"][idx]).detach().abs().max()) <= 1e-6      # the blend GEMM tiles the batch
    outs["seg"].backward(g[idx])
    scale = float(grad1[idx].abs().max())
    assert float((xs.grad - grad1[idx]).abs().max()) <= 2e-5 * scale

    # linearity: g -> 2 g scales every product and partial sum by exactly 2
    x2 = params.clone().requires_grad_(True)
    dec(x2)["seg"].backward(2.0 * g)
    assert float((x2.grad - 2.0 * grad1).abs().max()) <= 1e-6 * float(grad1.abs().max())

    # ... and DIRECTLY against the oracle: 64 strided samples of the 16384-batch (seconds of CPU), same tolerances as
    # tests/test_gpu_parity.py
    ii = idx.cpu().numpy()
    ref = np_oracle.decode(host_model, params[idx].cpu().numpy(), wh, vs, parts_by_vs[vs])
    assert np.abs(out["verts"][idx].detach().cpu().numpy() - ref["verts"]).max() <= 1e-5
    assert np.abs(out["joints"][idx].detach().cpu().numpy() - ref["J_transformed"]).max() <= 1e-5
    dp = float(np.abs(out["projects"][idx].detach().cpu().numpy() - ref["projects"]).max())
    assert dp <= 2.5e-5, dp
    gm, rm = out["mask"][idx].detach().cpu().numpy(), ref["mask"]
    same = (gm == rm).all(axis=1)
    assert (gm != rm).mean() <= 2e-3 and same.sum() >= 32, ((gm != rm).mean(), int(same.sum()))
    sg, rs = seg[idx].cpu().numpy(), ref["seg"]
    mism = sg.argmax(-1) != rs.argmax(-1)
    top2 = np.sort(rs, axis=-1)[..., -2:]
    near_tie = (top2[..., 1] - top2[..., 0]) <= 2.0 * np.sqrt(2.0) * dp + 4e-6
    assert int((mism & ~near_tie & same[:, None, None]).sum()) == 0          # every label flip is an oracle near-tie
    assert mism[same].mean() <= 1e-3, mism[same].mean()
    dseg = np.abs(sg[same] - rs[same])
    assert (dseg > np.sqrt(2.0) * dp + 3e-6).mean() <= 1e-5                   # visible vertices: exp(-d) is 1-Lipschitz
    assert dseg.max() <= 500.0 * np.sqrt(2.0) * dp + 2e-6                     # occluded ones (w = 500) on a pixel centre
    assert ii.shape == (64,)


def test_c5_seg_backward_is_repeatable(pkg, host_model, parts_by_vs, make_params):
    """The same backward evaluated five times on the same saved state: rows are handed to the warps on demand, so the
    sums may differ in their last bits, but nothing more.  (An L2 bulk prefetch in this kernel once made ~1 pixel per
    16384-sample launch read wrong data: differences of the size of a whole vertex gradient, which this test catches.)"""
    N, wh, vs = 16384, 48, 5
    dec = pkg.SmplDecoder(host_model, wh, vs, parts=parts_by_vs[vs], device=dev())
    with torch.no_grad():
        out = dec(torch.as_tensor(make_params(N, wh, seed=9), device=dev()))
    pr, mk = out["projects"].clone(), out["mask"].clone()
    del out
    gen = torch.Generator(device=dev()).manual_seed(11)
    g = torch.randn((N, wh, wh, 32), device=dev(), generator=gen)
    p = pr.clone().requires_grad_(True)
    seg = pkg.projects_to_seg([p, mk], wh, vs, parts=parts_by_vs[vs])
    grads = []
    for _ in range(5):
        p.grad = None
        seg.backward(g, retain_graph=True)
        grads.append(p.grad.clone())
    scale = float(grads[0].abs().max())
    for k in range(1, 5):
        assert float((grads[k] - grads[0]).abs().max()) <= 1e-5 * scale, k     # measured 6e-7; the bug it guards: 0.5


def test_c4_silhouette_large_batch_properties(pkg, host_model, make_params):
    wh, n_src, N = 256, 32, 8192                     # BASELINE config C4's batch
    dec = pkg.SmplDecoder(host_model, wh, None, device=dev())
    with torch.no_grad():
        pr_src = dec(torch.as_tensor(make_params(n_src, wh, seed=5), device=dev()), seg=False)["projects"]
    pr = pr_src.repeat(N // n_src, 1, 1).contiguous()
    gen = torch.Generator(device=dev()).manual_seed(4)
    g = torch.randn((N, wh, wh, 2), device=dev(), generator=gen)
    x = pr.clone().requires_grad_(True)
    sil = pkg.projects_to_silhouette(x, wh)
    sd = sil.detach()
    assert float((sd[..., 0] + sd[..., 1] - 1.0).abs().max()) <= 1.2e-7
    assert float(sd.min()) >= 0.0 and float(sd.max()) <= 1.0
    # replicas of the same sample are bit-identical (per-sample blocks), whatever their position in the batch
    assert torch.equal(sd[:n_src], sd[N - n_src:])
    sil.backward(g)
    g1 = x.grad.clone()
    sil = sil.detach()
    assert bool(torch.isfinite(g1).all()) and float(g1[..., 2].abs().max()) == 0.0
    # batch invariance against a small launch (the tile split over gridDim.y differs below 296 samples)
    xs = pr[:8].clone().requires_grad_(True)
    sils = pkg.projects_to_silhouette(xs, wh)
    assert torch.equal(sils, sil[:8])
    sils.backward(g[:8])
    assert float((xs.grad - g1[:8]).abs().max()) <= 1e-5 * float(g1[:8].abs().max())
    # linearity in the upstream gradient
    x2 = pr.clone().requires_grad_(True)
    pkg.projects_to_silhouette(x2, wh).backward(2.0 * g)
    assert float((x2.grad - 2.0 * g1).abs().max()) <= 1e-6 * float(g1.abs().max())
