"""The mesh visualiser (renderer.py:23-115, 146-197, 221-256).

tests/golden/render_vectors.npz was written by oracle/make_render_vectors.py, which runs the reference's OWN, UNMODIFIED
renderer.py (SMPLRenderer.__call__ / rotated -> render_model -> simple_renderer) with `opendr`, `cv2` and `plyfile`
resolved to the stand-ins of oracle/tf_shim/.  That pins the reference's call sites (camera defaults, near / far, lights,
albedo, part colours, background, alpha helpers, uint8 conversion); OpenDR's OpenGL rasteriser itself is restated once,
in oracle/np_oracle.py, and is "parity unpinned" (OpenDR / OpenGL are absent from the image).

CPU tests: the oracle's restatement of the call-site logic against those vectors, closed-form properties of its
rasteriser, and the no-GPU failure mode.  GPU tests: the CUDA renderer through the reference's interface against the
vectors: identical images except at pixels where a sample sits within float32 rounding of a triangle edge or of a
depth tie (bounded below), colours within one 8-bit level elsewhere.
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import np_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "render_vectors.npz")
MESHES = ("tmpl", "posed")


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def fixtures(pkg):
    return pkg.smpl_io.golden_fixtures()


def _cases(ref, name, fixtures):
    v, bg = ref[name + "_verts"], ref["background"]
    return {
        "lit": dict(),
        "seg": dict(render_seg=True),
        "bg_alpha": dict(img=bg, do_alpha=True),
        "alpha": dict(do_alpha=True),
        "cam": dict(cam=[260.0, 50.5, 44.25], img_size=(80, 112), near=1.0, far=float(v[:, 2].mean())),
    }


# ---- CPU: the oracle against the reference's own renderer.py -----------------------------------------------------------
@pytest.mark.parametrize("name", MESHES)
def test_oracle_matches_reference_renderer(ref, fixtures, name):
    v = ref[name + "_verts"]
    for tag, kw in _cases(ref, name, fixtures).items():
        kw = dict(kw)
        kw.setdefault("img_size", None)
        got = np_oracle.render_mesh(v, fixtures["faces"], part_colors=fixtures["ply_rgb"], default_size=96, flength=230.,
                                    **kw)
        assert got.dtype == np.uint8 and np.array_equal(got, ref[name + "_" + tag]), tag
    rot = np_oracle.render_mesh(np_oracle.rotated_verts(v, 60, "y"), fixtures["faces"], do_alpha=True,
                                default_size=96, flength=230.)
    assert np.array_equal(rot, ref[name + "_rot60"])


def test_golden_images_are_not_degenerate(ref, fixtures):
    for name in MESHES:
        lit, seg, al = ref[name + "_lit"], ref[name + "_seg"], ref[name + "_alpha"]
        cover = (lit != 255).any(-1)
        assert 0.04 < cover.mean() < 0.5                                   # a body, not an empty or a full frame
        assert np.array_equal(al[..., 3] == 255, (al[..., :3] != 255).any(-1))     # get_alpha, renderer.py:200-209
        assert len(np.unique(seg[cover].reshape(-1, 3), axis=0)) > 20      # many part colours visible
        bga = ref[name + "_bg_alpha"]
        assert (bga[..., 3] == 255).all()                                  # append_alpha, renderer.py:212-218
        assert np.array_equal(bga[~cover][:, :3], ref["background"][~cover])       # uncovered pixels show the image
        cam = ref[name + "_cam"]
        assert cam.shape == (80, 112, 3) and 0.005 < (cam != 255).any(-1).mean() < 0.5
        unclipped = np_oracle.render_mesh(ref[name + "_verts"], fixtures["faces"], cam=[260.0, 50.5, 44.25], img_size=(80, 112))
        assert (unclipped != 255).any(-1).sum() > (cam != 255).any(-1).sum()        # the far plane at the mean depth cuts the mesh


def test_rasteriser_closed_form():
    """One triangle, hand-checkable: coverage by the fill rule, depth order, clipping, perspective-correct colour."""
    v = np.array([[0.0, 0.0, 2.0], [8.0, 0.0, 2.0], [0.0, 8.0, 2.0],          # z = 2, projects to (0,0) (8,0) (0,8) with f = 2
                  [0.0, 0.0, 1.0], [4.0, 0.0, 1.0], [0.0, 4.0, 1.0]])         # z = 1 in front: (0,0) (8,0) (0,8) too
    col = np.array([[1.0, 0, 0]] * 3 + [[0, 0, 1.0]] * 3)
    f2, c0 = np.array([2.0, 2.0]), np.array([0.0, 0.0])
    back = np_oracle.rasterise(v, np.array([[0, 1, 2]]), col, f2, c0, 10, 10, 0.5, 10.0)
    # samples strictly inside x + y < 8, x > 0, y > 0 are covered; the hypotenuse and the two legs follow the fill rule
    inside = np.array([[(c > 0 and r > 0 and c + r < 8) for c in range(10)] for r in range(10)])
    red = np.all(back == np.array([1.0, 0, 0]), axis=-1)
    assert np.array_equal(red & inside, inside) and not red[9, 9]
    both = np_oracle.rasterise(v, np.array([[0, 1, 2], [3, 4, 5]]), col, f2, c0, 10, 10, 0.5, 10.0)
    assert np.array_equal(np.all(both == np.array([0, 0, 1.0]), axis=-1), red)                 # the nearer face wins everywhere
    clipped = np_oracle.rasterise(v, np.array([[0, 1, 2], [3, 4, 5]]), col, f2, c0, 10, 10, 1.5, 10.0)
    assert np.array_equal(clipped, back)                                                        # near = 1.5 removes z = 1
    # perspective-correct interpolation: a face from z = 1 to z = 3 shaded 0 -> 1 along x
    vv = np.array([[0.0, -4.0, 1.0], [27.0, -12.0, 3.0], [27.0, 36.0, 3.0], [0.0, 12.0, 1.0]])
    cc = np.array([[0.0] * 3, [1.0] * 3, [1.0] * 3, [0.0] * 3])
    im = np_oracle.rasterise(vv, np.array([[0, 1, 2], [0, 2, 3]]), cc, np.array([1.0, 1.0]), c0, 4, 10, 0.5, 10.0)
    for c in range(1, 9):          # screen x = c <-> X/Z = c with Z = 1 + 2 t, X = 27 t  =>  t = c / (27 - 2 c)
        t = c / (27.0 - 2.0 * c)
        assert abs(im[1, c, 0] - np.rint(t * 255) / 255) <= 1.0 / 255 + 1e-12


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree exists only in the build container")
def test_render_vectors_regenerate_identically(tmp_path):
    env = dict(os.environ)
    script = os.path.join(ROOT, "oracle", "make_render_vectors.py")
    src = open(script).read().replace('OUT = os.path.join(ROOT, "tests", "golden", "render_vectors.npz")',
                                      'OUT = %r' % str(tmp_path / "rv.npz"))
    tmp_script = tmp_path / "make_render_vectors.py"
    tmp_script.write_text(src.replace('ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', 'ROOT = %r' % ROOT))
    subprocess.run([sys.executable, str(tmp_script)], check=True, env=env, capture_output=True, cwd=ROOT)
    a, b = np.load(GOLDEN), np.load(str(tmp_path / "rv.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_renderer_has_no_cpu_path(pkg):
    with pytest.raises(pkg.SmplB200Error):
        pkg.SMPLRenderer(img_size=32)


# ---- GPU: the CUDA renderer against the same vectors ---------------------------------------------------------------------
def _compare(got, want, what):
    """Identical coverage except float32 edge hits; one 8-bit level on the colours elsewhere."""
    assert got.shape == want.shape and got.dtype == np.uint8, what
    cov_g, cov_w = (got[..., :3] != 255).any(-1), (want[..., :3] != 255).any(-1)
    flips = float((cov_g != cov_w).mean())
    both = cov_g & cov_w
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    off = float((diff[both].max(-1) > 1).mean()) if both.any() else 0.0
    # a sample within float32 rounding of an edge (coverage flip) or of a depth tie between two faces of different shade
    assert flips <= 2e-3 and off <= 5e-3, (what, flips, off)
    return flips, off


@pytest.mark.gpu
@pytest.mark.parametrize("name", MESHES)
def test_gpu_renderer_matches_reference_vectors(pkg, ref, fixtures, name):
    R = pkg.SMPLRenderer(img_size=96, flength=230.)
    v = ref[name + "_verts"]
    stats = {}
    for tag, kw in _cases(ref, name, fixtures).items():
        stats[tag] = _compare(R(v, color_id=None, **kw), ref[name + "_" + tag], tag)
    stats["rot60"] = _compare(R.rotated(v, 60, color_id=None), ref[name + "_rot60"], "rot60")
    print("render parity (coverage flips, colour > 1 level):", stats)
    # default color_id (0) is light_blue as well (documented deviation: python-2 dict order)
    assert np.array_equal(R(v), R(v, color_id=None))
    assert not np.array_equal(R(v, color_id=1), R(v))


@pytest.mark.gpu
def test_gpu_renderer_batch_and_tensor_io(pkg, ref):
    R = pkg.SMPLRenderer(img_size=64, flength=150.)
    vs = np.stack([ref["tmpl_verts"], ref["posed_verts"], ref["tmpl_verts"]])
    batch = R(torch.as_tensor(vs, device="cuda"), as_tensor=True)
    assert batch.is_cuda and batch.shape == (3, 64, 64, 3) and batch.dtype == torch.uint8
    for i in range(3):
        assert np.array_equal(batch[i].cpu().numpy(), R(vs[i]))
    assert torch.equal(batch[0], batch[2]) and not torch.equal(batch[0], batch[1])
    # run to run: bit-identical (no atomics, fixed face order)
    assert torch.equal(batch, R(torch.as_tensor(vs, device="cuda"), as_tensor=True))


@pytest.mark.gpu
def test_gpu_renderer_full_size_against_oracle(pkg, ref, fixtures):
    """The reference's default 224 x 224 frame (renderer.py:24-25) against the float64 restatement."""
    R = pkg.SMPLRenderer()
    v = ref["posed_verts"]
    want = np_oracle.render_mesh(v, fixtures["faces"])
    _compare(R(v), want, "224")
    wseg = np_oracle.render_mesh(v, fixtures["faces"], render_seg=True, part_colors=fixtures["ply_rgb"])
    _compare(R(v, render_seg=True), wseg, "224 seg")


@pytest.mark.gpu
def test_gpu_renderer_degenerate_inputs(pkg, ref):
    R = pkg.SMPLRenderer(img_size=40)
    v = ref["tmpl_verts"].copy()
    behind = v.copy(); behind[:, 2] -= 20.0                                 # the whole mesh behind the camera
    assert (R(behind) == 255).all()
    tiny = v.copy(); tiny[:, :2] *= 1e-3                                    # every face inside one pixel
    out = R(tiny)
    assert (out != 255).any(-1).sum() <= 4
    far = R(v, near=0.1, far=0.2)                                           # clip planes in front of the mesh
    assert (far == 255).all()
    with pytest.raises(ValueError):
        R(v[:100])
