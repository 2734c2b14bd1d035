"""CPU tests: model / table I/O, the C-ABI surface, host-side logic.  No compute kernels run here."""
import ctypes
import importlib
import os
import pickle
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- smpl_io ---------------------------------------------------------------------------------------------------
def test_synthetic_model_is_smpl_shaped(host_model):
    hm = host_model
    assert hm.v_template.shape == (6890, 3) and hm.shapedirs.shape == (10, 20670) and hm.posedirs.shape == (207, 20670)
    assert np.allclose(hm.lbs_weights.sum(1), 1, atol=1e-6) and ((hm.lbs_weights != 0).sum(1) <= 4).all()
    assert np.allclose(hm.J_regressor.sum(0), 1, atol=1e-5)
    assert list(hm.parents[:6]) == [-1, 0, 0, 0, 1, 2]


def test_pkl_roundtrip_py2_layout(smpl_io, host_model, tmp_path):
    """The loader must digest the HMR release layout: chumpy leaves, scipy CSC regressors, uint32 kintree whose
    first entry is 2**32-1 (batch_smpl.py:39-83), pickled with protocol 2 and latin1 strings."""
    path = str(tmp_path / "neutral_smpl_with_cocoplus_reg.pkl")
    smpl_io.save_smpl_pkl(host_model, path)
    raw = open(path, "rb").read()
    assert b"chumpy" in raw and b"scipy.sparse" in raw
    back = smpl_io.load_smpl_pkl(path)
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "joint_regressor"):
        assert np.array_equal(getattr(back, k), getattr(host_model, k)), k
    assert np.array_equal(back.parents, host_model.parents) and back.parents[0] == -1
    with pytest.raises((IOError, OSError)):
        smpl_io.load_smpl_pkl(str(tmp_path / "missing.pkl"))


def test_part_tables_match_reference_files(smpl_io, parts_by_vs, tmp_path):
    totals = {None: 6879, 2: 3438, 5: 1376}
    for vs, parts in parts_by_vs.items():
        assert len(parts) == 31 and sum(len(p) for p in parts) == totals[vs]
        ptr, idx = smpl_io.sampled_part_table(parts, vs)
        assert ptr[-1] == totals[vs] and idx.max() < -(-6890 // (vs or 1))
    assert smpl_io.part_vertices_filename(None) == "./keras_smpl/part_vertices.pkl"
    assert smpl_io.part_vertices_filename(5) == "./keras_smpl/5_sampled_part_vertices.pkl"
    for proto in (0, 2):                                         # the reference ships both protocols (Q14)
        f = str(tmp_path / ("p%d.pkl" % proto))
        smpl_io.write_part_vertices_pkl(parts_by_vs[5], f, protocol=proto)
        assert smpl_io.load_part_vertices(f) == parts_by_vs[5]
    ref = "/root/reference/keras_smpl/5_sampled_part_vertices.pkl"
    if os.path.exists(ref):                                      # build container only
        assert smpl_io.load_part_vertices(ref) == parts_by_vs[5]


def test_mean_params_h5_reader(smpl_io, tmp_path):
    fx = smpl_io.golden_fixtures()
    f = str(tmp_path / "neutral_smpl_mean_params.h5")
    open(f, "wb").write(bytes(bytearray(fx["h5_bytes"])))
    m = smpl_io.read_mean_params_h5(f)
    assert np.array_equal(m["shape"], fx["mean_shape"]) and np.array_equal(m["pose"], fx["mean_pose"])
    assert abs(m["shape"][0] - 0.20561) < 1e-5 and abs(m["pose"][0] - 0.453144) < 1e-5
    v = smpl_io.mean_param_vector(48, m)
    assert v.shape == (1, 86) and np.all(v[0, 4:7] == 0) and np.allclose(v[0, :4], [24, 24, 24, 30])
    with pytest.raises(ValueError):
        smpl_io.read_mean_params_h5(b"not an hdf5 file" * 10)
    ref = "/root/reference/neutral_smpl_mean_params.h5"
    if os.path.exists(ref):
        assert np.array_equal(smpl_io.read_mean_params_h5(ref)["pose"], m["pose"])


# ---- C ABI -----------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "smpl_b200.h")).read()
    declared = set(re.findall(r"\b(smpl_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    lib = pkg.load_library()
    binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
    assert declared == set(binding.SIGNATURES), declared ^ set(binding.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smpl_b200_abi_version() == 3
    assert isinstance(lib.smpl_b200_launch_count(), int)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg, host_model):
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
    lib = pkg.load_library()
    struct, keep = binding.make_host_model(host_model)
    handle = ctypes.c_void_p()
    rc = lib.smpl_b200_model_create(ctypes.byref(struct), 0, ctypes.byref(handle))
    assert rc == -5 and not handle.value                        # SMPL_B200_ERR_NO_DEVICE
    assert b"no CPU path" in lib.smpl_b200_last_error()
    layer = pkg.SMPLLayer(host_model)
    with pytest.raises(pkg.SmplB200Error):
        layer(torch.zeros(1, 86))
    with pytest.raises(pkg.SmplB200Error):
        pkg.compute_mask(torch.zeros(1, 10, 3))
    with pytest.raises(pkg.SmplB200Error):
        pkg.projects_to_silhouette(torch.zeros(1, 10, 3), 8)


def test_bad_arguments_return_error_codes(pkg):
    lib = pkg.load_library()
    out = ctypes.c_void_p()
    assert lib.smpl_b200_model_create(None, 0, ctypes.byref(out)) == -1
    assert b"null" in lib.smpl_b200_last_error()
    assert lib.smpl_b200_parts_create(0, 0, None, None, 0, ctypes.byref(out)) == -1
    assert lib.smpl_b200_mask_fwd(None, 3, 10, None, None) == -1
    assert lib.smpl_b200_mask_fwd(None, 0, 10, None, None) == 0          # empty batch is a no-op
    assert lib.smpl_b200_workspace_bytes(None, 0, 4, 48, 1) == 0


# ---- host logic ------------------------------------------------------------------------------------------------
def test_layer_interface_matches_reference(pkg, host_model):
    layer = pkg.SMPLLayer("./neutral_smpl_with_cocoplus_reg.pkl", batch_size=4, dtype="float32", joint_type="lsp")
    assert layer.get_config() == {"pkl_path": "./neutral_smpl_with_cocoplus_reg.pkl", "batch_size": 4,
                                  "dtype": "float32"}            # joint_type omitted, as in batch_smpl.py:161-166
    assert layer.compute_output_shape((4, 86)) == (4, 6890, 3)
    assert layer.num_cam == 4 and layer.num_keypoints == 14
    assert pkg.SMPLLayer(host_model, joint_type="cocoplus").num_keypoints == 19
    with pytest.raises(TypeError):
        pkg.SMPLLayer(host_model, dtype="float16")
    assert list(layer.parameters()) == []                        # no trainable weights (batch_smpl.py:92-94)


def test_mean_param_functions_cpu(pkg):
    from oracle import np_oracle
    mean = pkg.smpl_io.load_mean_params()
    f = torch.randn(3, 2048)
    out = pkg.concat_mean_param(f, 48)
    assert out.shape == (3, 2134)
    assert np.array_equal(out.numpy(), np_oracle.concat_mean_param(f.numpy(), 48, mean))
    x = torch.randn(2, 86)
    assert np.array_equal(pkg.set_cam_params(x, 64).numpy(), np_oracle.set_cam_params(x.numpy(), 64))
    assert np.array_equal(pkg.load_mean_set_cam_params(x, 48).numpy(),
                          np_oracle.load_mean_set_cam_params(x.numpy(), 48, mean))


def test_synthetic_params_are_deterministic(make_params):
    a, b = make_params(5, 48, seed=3), make_params(5, 48, seed=3)
    assert a.dtype == np.float32 and a.shape == (5, 86) and np.array_equal(a, b)
    assert not np.array_equal(a, make_params(5, 48, seed=4))
    assert np.abs(a[:, 76:]).max() <= 3.0
