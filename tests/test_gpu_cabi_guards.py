"""C-ABI calls on raw device pointers with guard bands around every output and workspace buffer: any out-of-bounds
global write by a kernel shows up as a damaged sentinel (compute-sanitizer is not available on this pool)."""
import ctypes as C
import importlib

import numpy as np
import pytest
import torch

from oracle import np_oracle

pytestmark = pytest.mark.gpu
GUARD = 4096          # bytes on each side
SENT = 0xA5


class Guarded:
    def __init__(self, nbytes, dev):
        self.n = int(nbytes)
        pad = (-self.n) % 256
        self.buf = torch.full((GUARD + self.n + pad + GUARD,), SENT, dtype=torch.uint8, device=dev)
        self.pad = pad

    @property
    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + GUARD)

    def view(self, dtype, shape):
        return self.buf[GUARD:GUARD + self.n].view(dtype).view(shape)

    def intact(self):
        lo = self.buf[:GUARD]
        hi = self.buf[GUARD + self.n:]
        return bool((lo == SENT).all().item()) and bool((hi == SENT).all().item())


@pytest.mark.parametrize("n,vs,wh", [(3, 5, 48), (65, 5, 48), (2, 1, 64), (1, 2, 33)])
def test_cabi_guard_bands(pkg, host_model, parts_by_vs, make_params, n, vs, wh):
    lib = pkg.load_library()
    binding = importlib.import_module("indirect_learning_pose-shape_b200._lib")
    layers = importlib.import_module("indirect_learning_pose-shape_b200.layers")
    dev = torch.device("cuda", 0)
    dm = layers.get_device_model(host_model, dev)
    V, LD = dm.V, dm.LD
    Vs = -(-V // vs)
    vsk = None if vs == 1 else vs
    table = layers.get_part_table(vsk, Vs, dev, parts=parts_by_vs[vsk])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p_np = make_params(n, wh, seed=7)
    params = Guarded(n * 86 * 4, dev)
    params.view(torch.float32, (n, 86)).copy_(torch.from_numpy(p_np))
    verts, joints = Guarded(n * V * 12, dev), Guarded(n * 24 * 12, dev)
    vposed, proj = Guarded(n * LD * 4, dev), Guarded(n * Vs * 12, dev)
    vps_ld = (Vs * 3 + 3) // 4 * 4                       # SMPL_B200_VPS_LD
    vps = Guarded(n * vps_ld * 4, dev)
    ws_f = Guarded(dm.workspace_bytes(binding.OP_DECODE_FWD, n), dev)
    binding.check(lib.smpl_b200_decode_fwd(dm.handle, params.ptr, n, verts.ptr, joints.ptr, None, 0, vposed.ptr, vps.ptr,
                                           proj.ptr, vs, ws_f.ptr, ws_f.n, stream), "decode_fwd")
    mask = Guarded(n * Vs * 4, dev)
    binding.check(lib.smpl_b200_mask_fwd(proj.ptr, n, Vs, mask.ptr, stream), "mask_fwd")
    seg = Guarded(n * wh * wh * 32 * 4, dev)
    saved = Guarded(lib.smpl_b200_seg_saved_bytes(n, wh), dev)
    binding.check(lib.smpl_b200_seg_fwd(table.handle, proj.ptr, mask.ptr, n, Vs, wh, seg.ptr, saved.ptr, stream), "seg_fwd")
    g_seg = Guarded(n * wh * wh * 32 * 4, dev)
    g_seg.view(torch.float32, (n, wh, wh, 32)).normal_()
    g_proj = Guarded(n * Vs * 12, dev)
    binding.check(lib.smpl_b200_seg_bwd(table.handle, proj.ptr, mask.ptr, g_seg.ptr, saved.ptr, n, Vs, wh, g_proj.ptr, stream),
                  "seg_bwd")
    sil, g_sil, g_proj2 = Guarded(n * wh * wh * 8, dev), Guarded(n * wh * wh * 8, dev), Guarded(n * Vs * 12, dev)
    g_sil.view(torch.float32, (n, wh, wh, 2)).normal_()
    binding.check(lib.smpl_b200_silhouette_fwd(proj.ptr, n, Vs, wh, sil.ptr, None, 0, stream), "sil_fwd")
    binding.check(lib.smpl_b200_silhouette_bwd(proj.ptr, g_sil.ptr, n, Vs, wh, g_proj2.ptr, None, 0, stream), "sil_bwd")
    # the same pair with the arg-min map handed from forward to backward (search-free backward)
    sil_b, g_proj3 = Guarded(n * wh * wh * 8, dev), Guarded(n * Vs * 12, dev)
    sil_ws = Guarded(dm.workspace_bytes(binding.OP_SILHOUETTE_FWD, n, wh, 0), dev)
    assert sil_ws.n >= n * wh * wh * 2
    binding.check(lib.smpl_b200_silhouette_fwd(proj.ptr, n, Vs, wh, sil_b.ptr, sil_ws.ptr, sil_ws.n, stream), "sil_fwd saved")
    binding.check(lib.smpl_b200_silhouette_bwd(proj.ptr, g_sil.ptr, n, Vs, wh, g_proj3.ptr, sil_ws.ptr, sil_ws.n, stream),
                  "sil_bwd saved")
    g_params = Guarded(n * 86 * 4, dev)
    ws_b = Guarded(dm.workspace_bytes(binding.OP_DECODE_BWD, n, 0, vs), dev)
    binding.check(lib.smpl_b200_decode_bwd(dm.handle, params.ptr, n, vposed.ptr, None, None, g_proj.ptr, vs, None,
                                           g_params.ptr, ws_b.ptr, ws_b.n, stream), "decode_bwd")
    # the same backward from the compact sampled copy alone (no full v_posed): same arithmetic, same bits
    g_params_c = Guarded(n * 86 * 4, dev)
    binding.check(lib.smpl_b200_decode_bwd(dm.handle, params.ptr, n, None, vps.ptr, None, g_proj.ptr, vs, None,
                                           g_params_c.ptr, ws_b.ptr, ws_b.n, stream), "decode_bwd compact")
    assert lib.smpl_b200_decode_bwd(dm.handle, params.ptr, n, None, None, None, g_proj.ptr, vs, None, g_params_c.ptr,
                                    ws_b.ptr, ws_b.n, stream) == -1          # neither copy: SMPL_B200_ERR_BAD_ARG
    # dense-gradient variant (g_verts given) needs the full-size workspace
    g_verts = Guarded(n * V * 12, dev)
    g_verts.view(torch.float32, (n, V, 3)).normal_()
    g_params2 = Guarded(n * 86 * 4, dev)
    ws_b2 = Guarded(dm.workspace_bytes(binding.OP_DECODE_BWD, n, 0, 1), dev)
    binding.check(lib.smpl_b200_decode_bwd(dm.handle, params.ptr, n, vposed.ptr, vps.ptr, g_verts.ptr, g_proj.ptr, vs, None,
                                           g_params2.ptr, ws_b2.ptr, ws_b2.n, stream), "decode_bwd dense")
    # the whole path in one call each way (smpl_b200_full_fwd / _bwd): same kernels, so the same bits as the chain above
    f_verts, f_joints, f_proj = Guarded(n * V * 12, dev), Guarded(n * 24 * 12, dev), Guarded(n * Vs * 12, dev)
    f_mask, f_seg, f_gp = Guarded(n * Vs * 4, dev), Guarded(n * wh * wh * 32 * 4, dev), Guarded(n * 86 * 4, dev)
    f_state = Guarded(lib.smpl_b200_full_state_bytes(dm.handle, n, wh, vs), dev)
    ws_ff = Guarded(dm.workspace_bytes(binding.OP_FULL_FWD, n, wh, vs), dev)
    ws_fb = Guarded(dm.workspace_bytes(binding.OP_FULL_BWD, n, wh, vs), dev)
    binding.check(lib.smpl_b200_full_fwd(dm.handle, table.handle, params.ptr, n, wh, vs, f_verts.ptr, f_joints.ptr,
                                         f_proj.ptr, f_mask.ptr, f_seg.ptr, f_state.ptr, ws_ff.ptr, ws_ff.n, stream), "full_fwd")
    binding.check(lib.smpl_b200_full_bwd(dm.handle, table.handle, params.ptr, n, wh, vs, f_proj.ptr, f_mask.ptr, g_seg.ptr,
                                         f_state.ptr, f_gp.ptr, ws_fb.ptr, ws_fb.n, stream), "full_bwd")
    # inference form: no state, verts / joints skipped
    f_seg2, f_proj2, f_mask2 = Guarded(n * wh * wh * 32 * 4, dev), Guarded(n * Vs * 12, dev), Guarded(n * Vs * 4, dev)
    binding.check(lib.smpl_b200_full_fwd(dm.handle, table.handle, params.ptr, n, wh, vs, None, None, f_proj2.ptr,
                                         f_mask2.ptr, f_seg2.ptr, None, ws_ff.ptr, ws_ff.n, stream), "full_fwd inference")
    # stand-alone projection
    proj_s, g_verts_s, g_params_s = Guarded(n * Vs * 12, dev), Guarded(n * V * 12, dev), Guarded(n * 86 * 4, dev)
    binding.check(lib.smpl_b200_project_fwd(verts.ptr, params.ptr, n, V, vs, proj_s.ptr, stream), "project_fwd")
    binding.check(lib.smpl_b200_project_bwd(verts.ptr, params.ptr, g_proj.ptr, n, V, vs, g_verts_s.ptr, g_params_s.ptr, stream),
                  "project_bwd")
    torch.cuda.synchronize()
    for name, g in dict(params=params, verts=verts, joints=joints, vposed=vposed, proj=proj, ws_f=ws_f, mask=mask, seg=seg,
                        saved=saved, g_seg=g_seg, g_proj=g_proj, sil=sil, g_sil=g_sil, g_proj2=g_proj2, g_params=g_params,
                        ws_b=ws_b, g_verts=g_verts, g_params2=g_params2, ws_b2=ws_b2, proj_s=proj_s, g_verts_s=g_verts_s,
                        g_params_s=g_params_s, vps=vps, g_params_c=g_params_c, f_verts=f_verts, f_joints=f_joints,
                        f_proj=f_proj, f_mask=f_mask, f_seg=f_seg, f_gp=f_gp, f_state=f_state, ws_ff=ws_ff, ws_fb=ws_fb,
                        f_seg2=f_seg2, f_proj2=f_proj2, f_mask2=f_mask2, sil_b=sil_b, g_proj3=g_proj3, sil_ws=sil_ws).items():
        assert g.intact(), "guard band damaged around %s" % name
    # and the results are the real thing
    ref = np_oracle.smpl_layer_call(host_model, p_np)
    assert np.abs(verts.view(torch.float32, (n, V, 3)).cpu().numpy() - ref).max() <= 1e-5
    assert np.array_equal(proj.view(torch.float32, (n, Vs, 3)).cpu().numpy(), proj_s.view(torch.float32, (n, Vs, 3)).cpu().numpy())
    assert torch.isfinite(g_params.view(torch.float32, (n, 86))).all() and torch.isfinite(g_params2.view(torch.float32, (n, 86))).all()
    assert float(seg.view(torch.float32, (n, wh, wh, 32)).sum(-1).min()) > 0.99
    f32 = lambda g, shape: g.view(torch.float32, shape)      # noqa: E731
    gc, gm = f32(g_params_c, (n, 86)), f32(g_params, (n, 86))
    assert float((gc - gm).abs().max()) <= 1e-6 * float(gm.abs().max())       # same arithmetic; the small-batch blend backward sums with atomics
    assert torch.equal(f32(f_verts, (n, V, 3)), f32(verts, (n, V, 3)))
    assert torch.equal(f32(f_proj, (n, Vs, 3)), f32(proj, (n, Vs, 3))) and torch.equal(f32(f_mask, (n, Vs)), f32(mask, (n, Vs)))
    assert torch.equal(f32(f_seg, (n, wh, wh, 32)), f32(seg, (n, wh, wh, 32)))
    assert torch.equal(f32(f_seg2, (n, wh, wh, 32)), f32(seg, (n, wh, wh, 32)))
    assert torch.equal(f32(sil_b, (n, wh, wh, 2)), f32(sil, (n, wh, wh, 2)))
    g3, g2 = f32(g_proj3, (n, Vs, 3)), f32(g_proj2, (n, Vs, 3))
    assert float((g3 - g2).abs().max()) <= 1e-5 * float(g2.abs().max())      # same terms, summed in another order
    ga, gb = f32(f_gp, (n, 86)), f32(g_params, (n, 86))
    assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max())     # the seg backward sums rows in on-demand order
