"""The product's arithmetic headers (csrc/g2s_math.cuh, g2s_raster.cuh), compiled for the host by tests/emu, must
reproduce the oracle bit for bit: projected vertices -> face-index map, z, recon_depth, warp grids."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import emu_cam, emu_lib, golden, oracle_renderer, vp
from oracle import nr_port
from g2s_b200 import synthetic


def _emu_view(emu, cam, S, depth_img, R, t):
    ndc = torch.empty(S * S, 3)
    emu.emu_project(ctypes.byref(cam), vp(depth_img), vp(R), vp(t), vp(ndc))
    fi = torch.empty(2 * S, 2 * S, dtype=torch.int32)
    rec = torch.empty(S, S)
    emu.emu_raster(ctypes.byref(cam), vp(ndc), vp(fi), vp(rec), None)
    return fi, rec


@pytest.mark.parametrize("S,P,rot,seed", [(16, 4, 60.0, 1), (32, 4, 120.0, 2), (64, 3, 60.0, 3), (128, 2, 60.0, 4),
                                          (33, 3, 90.0, 5)])
def test_emu_bit_exact_vs_oracle(S, P, rot, seed):
    emu = emu_lib()
    case = synthetic.make_case(S, P, seed=seed, rot_deg=rot)
    orc = oracle_renderer(S)
    orc.set_transform_matrices(case["view"])
    rd = orc.warp_canon_depth(case["depth"].expand(P, S, S))
    fo = nr_port.LAST["face_index_map"].flip(1)
    gi = orc.get_inv_warped_2d_grid(rd)
    gf = orc.get_warped_2d_grid(case["depth"].expand(P, S, S))
    cam = emu_cam(orc, S)
    dimg = case["depth"][0].contiguous()
    for b in range(P):
        R = orc.rot_mat[b].contiguous()
        t = orc.trans_xyz[b].reshape(3).contiguous()
        fi, rec = _emu_view(emu, cam, S, dimg, R, t)
        assert torch.equal(fi, fo[b])
        assert torch.equal(rec, rd[b])
        g = torch.empty(S, S, 2)
        emu.emu_warp_grid(ctypes.byref(cam), vp(rd[b].contiguous()), vp(R), vp(t), S, S, 1, vp(g))
        assert torch.equal(g, gi[b])
        emu.emu_warp_grid(ctypes.byref(cam), vp(dimg), vp(R), vp(t), S, S, 0, vp(g))
        assert torch.equal(g, gf[b])


@pytest.mark.parametrize("name", ["s16_p3", "s32_p2", "s32_p2_wide"])
def test_emu_vs_golden(name):
    emu = emu_lib()
    g = golden(name)
    S, P = g["depth"].shape[-1], g["view"].shape[0]
    orc = oracle_renderer(S)
    cam = emu_cam(orc, S)
    dimg = torch.tensor(g["depth"][0]).contiguous()
    for b in range(P):
        R = torch.tensor(g["rot_mat"][b]).contiguous()
        t = torch.tensor(g["trans_xyz"][b]).reshape(3).contiguous()
        fi, rec = _emu_view(emu, cam, S, dimg, R, t)
        assert np.array_equal(fi.numpy(), g["face_idx"][b])
        assert np.array_equal(rec.numpy(), g["recon_depth"][b])


def test_face_vertices_closed_form_matches_get_face_idx():
    from oracle import renderer_oracle as ro
    emu = emu_lib()
    S = 7
    faces = ro.get_face_idx(1, S, S)[0]
    full = torch.cat([faces, faces[:, [2, 1, 0]]], 0)
    v = (ctypes.c_int * 3)()
    for f in range(full.shape[0]):
        emu.emu_face_vertices(f, S, v)
        assert list(v) == full[f].tolist()
