"""The regression module ahead of the decoder (model.py:63-105): Dense layers as 3xTF32 tcgen05 GEMMs,
forward and backward, against the NumPy restatement (fp32 and fp64) and torch-CPU fp64 autograd of the same formulas.

Tolerance: the reference's layers are fp32 Keras Dense (sgemm); a 3xTF32 split product carries ~22 mantissa bits per
operand and accumulates in fp32, so outputs are compared at 2e-5 of the output scale against fp64 -- the distance the
fp32 NumPy evaluation itself keeps from fp64 is measured alongside.
"""
import numpy as np
import pytest
import torch

from oracle import np_oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev())


@pytest.mark.parametrize("M,fin,fout,relu", [(200, 2134, 1024, True), (70, 1024, 86, False), (129, 333, 96, True),
                                             (5, 40, 7, False), (16384, 1024, 86, False)])
def test_dense_forward_backward(pkg, M, fin, fout, relu):
    rng = np.random.default_rng(M + fin)
    x = rng.standard_normal((M, fin)).astype(np.float32)
    w = (rng.standard_normal((fin, fout)) / np.sqrt(fin)).astype(np.float32)
    b = rng.standard_normal(fout).astype(np.float32)
    g = rng.standard_normal((M, fout)).astype(np.float32)
    x64 = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    w64 = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    b64 = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    y64 = x64 @ w64 + b64
    if relu:
        y64 = torch.relu(y64)
    (y64 * torch.tensor(g, dtype=torch.float64)).sum().backward()
    layer = pkg.Dense(fin, fout, "relu" if relu else "linear", device=dev())
    with torch.no_grad():
        layer.kernel.copy_(t(w)); layer.bias.copy_(t(b))
    # a column block of a wider buffer as input: rows are not densely packed
    wide = torch.zeros((M, fin + 5), device=dev())
    wide[:, 2:2 + fin] = t(x)
    xin = wide[:, 2:2 + fin].detach().requires_grad_(True)
    y = layer(xin)
    (y * t(g)).sum().backward()
    ref = y64.detach().numpy()
    scale = max(1.0, np.abs(ref).max())
    f32 = np_oracle.dense(x, w, b, "relu" if relu else "linear")
    assert np.abs(y.detach().cpu().numpy() - ref).max() <= 2e-5 * scale
    assert np.abs(f32 - ref).max() <= 2e-5 * scale                      # the fp32 evaluation's own distance
    for got, want, name in ((xin.grad, x64.grad, "gX"), (layer.kernel.grad, w64.grad, "gW"), (layer.bias.grad, b64.grad, "gb")):
        want = want.numpy()
        s = max(1.0, np.abs(want).max())
        # a ReLU gate may flip where the pre-activation is within rounding of zero: allow a handful of entries
        bad = np.abs(got.cpu().numpy() - want) > 3e-5 * s
        assert bad.mean() <= (1e-4 if relu else 0.0), (name, bad.mean(), np.abs(got.cpu().numpy() - want).max(), s)


def _weights(rng, sizes):
    return [((rng.standard_normal((a, b)) * np.sqrt(2.0 / (a + b))).astype(np.float32),
             (0.1 * rng.standard_normal(b)).astype(np.float32)) for a, b in sizes]


def test_ief_regressor_matches_model_py(pkg):
    """model.py:63-97: three iterations through the shared layers, forward and d/d(weights, features)."""
    rng = np.random.default_rng(7)
    n, wh = 96, 48
    feat = np.abs(rng.standard_normal((n, 2048))).astype(np.float32)           # ResNet features are post-ReLU
    W = _weights(rng, [(2134, 1024), (1024, 1024), (1024, 86)])
    mv = pkg.smpl_io.load_mean_params()
    ref32 = np_oracle.ief_regressor(feat, W, wh, mv)
    ref64 = np_oracle.ief_regressor(feat.astype(np.float64), [(k.astype(np.float64), b.astype(np.float64)) for k, b in W], wh, mv)
    reg = pkg.IEFRegressor(wh, device=dev())
    with torch.no_grad():
        for layer, (k, b) in zip((reg.IEF_layer_1, reg.IEF_layer_2, reg.IEF_layer_3), W):
            layer.kernel.copy_(t(k)); layer.bias.copy_(t(b))
    x = t(feat).requires_grad_(True)
    out = reg(x)
    assert tuple(out.shape) == (n, 86)
    got = out.detach().cpu().numpy()
    assert np.abs(got - ref64).max() <= 1e-5 and np.abs(ref32 - ref64).max() <= 1e-5
    # gradients vs torch fp64 autograd of the same graph
    g = rng.standard_normal((n, 86)).astype(np.float32)
    (out * t(g)).sum().backward()
    f64 = torch.tensor(feat, dtype=torch.float64, requires_grad=True)
    P = [(torch.tensor(k, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True))
         for k, b in W]
    mean = torch.tensor(pkg.smpl_io.mean_param_vector(wh, mv), dtype=torch.float64)
    param = mean.expand(n, -1)
    state = torch.cat([f64, param], 1)
    for _ in range(3):
        d = torch.relu(state @ P[0][0] + P[0][1])
        d = torch.relu(d @ P[1][0] + P[1][1])
        d = d @ P[2][0] + P[2][1]
        param = param + 0.005 * d
        state = torch.cat([f64, param], 1)
    (param * torch.tensor(g, dtype=torch.float64)).sum().backward()
    pairs = [(x.grad, f64.grad)]
    for layer, (k, b) in zip((reg.IEF_layer_1, reg.IEF_layer_2, reg.IEF_layer_3), P):
        pairs += [(layer.kernel.grad, k.grad), (layer.bias.grad, b.grad)]
    for got_g, want_g in pairs:
        want_g = want_g.numpy()
        s = np.abs(want_g).max() + 1e-12
        bad = np.abs(got_g.cpu().numpy() - want_g) > 1e-4 * s
        assert bad.mean() <= 1e-4, (bad.mean(), np.abs(got_g.cpu().numpy() - want_g).max(), s)


def test_plain_regressor_feeds_the_decoder(pkg, host_model, parts_by_vs):
    """model.py:99-118 end to end: features -> params -> decoder -> seg, one backward through everything."""
    rng = np.random.default_rng(11)
    n, wh = 64, 48
    feat = np.abs(rng.standard_normal((n, 2048))).astype(np.float32)
    W = _weights(rng, [(2048, 2048), (2048, 1024), (1024, 86)])
    mv = pkg.smpl_io.load_mean_params()
    ref = np_oracle.plain_regressor(feat.astype(np.float64), [(k.astype(np.float64), b.astype(np.float64)) for k, b in W], wh, mv)
    reg = pkg.PlainRegressor(wh, device=dev())
    with torch.no_grad():
        for layer, (k, b) in zip((reg.dense_1, reg.dense_2, reg.dense_3), W):
            layer.kernel.copy_(t(k)); layer.bias.copy_(t(b))
    params = reg(t(feat))
    assert np.abs(params.detach().cpu().numpy() - ref).max() <= 1e-5
    dec = pkg.SmplDecoder(host_model, wh, 5, need_verts=False, parts=parts_by_vs[5], device=dev(), fused=True)
    seg = dec(params)["seg"]
    seg.square().sum().backward()
    for layer in (reg.dense_1, reg.dense_2, reg.dense_3):
        assert bool(torch.isfinite(layer.kernel.grad).all()) and float(layer.kernel.grad.abs().max()) > 0


def test_axpy_cols_cabi(pkg):
    import ctypes as C
    lib = pkg.load_library()
    a = torch.arange(12, dtype=torch.float32, device=dev()).reshape(3, 4)
    d = torch.ones((3, 6), device=dev())
    out = torch.zeros((3, 5), device=dev())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.smpl_b200_axpy_cols(C.c_void_p(a.data_ptr()), 4, C.c_void_p(d.data_ptr()), 6, 0.5, 3, 4,
                                   C.c_void_p(out.data_ptr()), 5, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(out[:, :4], a + 0.5) and float(out[:, 4].abs().max()) == 0.0
