"""Known answers for NON-TRIVIAL geometry that pin the restated rasteriser (oracle/nr_raster.c) -- and, with the same
scenes and the same assertions, the CUDA path -- against an independent float64 3-D ray caster (tests/raycast_ref.py) and
against closed forms:

  * a tilted plane under yaw + pitch: the depth at every covered sub-pixel is the ray-plane intersection; the face map is
    what the ray hits first;
  * a tilted plane sampled on the pixel grid: closed-form face map (the identity-view layout);
  * occlusion: a stepped surface under yaw (the near sheet hides the far one and the wall between them) -- the winner is the
    first hit along the ray; fill_back: walls seen from behind carry the index of the reversed copy (f + 2(S-1)^2);
  * near / far rejection: sheets beyond `far` (or before `near`) leave background, sheets behind them win;
  * rgb: perspective-correct barycentric blend of the vertex colours of a colour field that is linear on a plane (with the
    reference's own assignment of colours to vertices, utils.py:98-109);
  * gradients: backward_depth_map against the formula of SURVEY.md App. A.6 evaluated independently in float64, and that
    formula against a float64 finite difference of the ray caster (it is the exact derivative of the interpolated depth).

The reference holds no golden vector for the external rasteriser, so this is the strongest available
substitute: the oracle is not compared with itself or with a closed form for flat geometry only, but with a different
formulation of the same image-formation model.  The CPU tests pin the oracle; the `gpu`-marked ones run the identical
comparison on the product (through the C ABI).
"""
import math

import numpy as np
import pytest
import torch

import raycast_ref as rc
from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, oracle_renderer
from oracle import nr_port, renderer_oracle as ro

NEAR, FAR = 0.1, 100.0


# ------------------------------------------------------------------------------------------------ scenes
def _warped(S, depth, view):
    """camera-space vertices of the reference's mesh for (depth, view): renderer.py:90-95 through the oracle"""
    orc = oracle_renderer(S)
    orc.set_transform_matrices(view)
    return orc.get_warped_3d_grid(depth).reshape(1, -1, 3).contiguous(), orc


def scene_tilted_plane(S):
    """flat canonical depth seen under yaw + pitch + roll and a translation: a plane tilted against the camera"""
    view = torch.tensor([[0.31, -0.47, 0.12, 0.02, -0.015, 0.03]])
    return _warped(S, torch.full((1, S, S), 0.97), view)


def scene_step(S):
    """two sheets (0.92 left, 1.06 right) joined by a 1-px wall, yawed so that the near sheet hides part of the far one
    from one side and the step opens up (the wall stretches) from the other"""
    d = torch.full((1, S, S), 0.92)
    d[:, :, S // 2:] = 1.06
    out = []
    for yaw in (0.55, -0.55):
        out.append(_warped(S, d, torch.tensor([[0.05, yaw, 0.0, 0.0, 0.0, 0.0]])))
    return out


def scene_fold(S):
    """a steep ramp at the mesh boundary (depth 1.1 -> 0.9 over the six leftmost columns) under a 0.9 rad yaw: the ramp turns
    its back to the camera with nothing in front of it, so the reversed (fill_back) copies of its faces win"""
    d = torch.full((1, S, S), 0.9)
    d[:, :, :6] = torch.linspace(1.1, 0.9, 6)
    return _warped(S, d, torch.tensor([[0.0, 0.9, 0.0, 0.0, 0.0, 0.0]]))


def scene_bumpy(S, seed=3):
    """the synthetic ellipsoid + noise + border walls of the bench workload under a strong yaw / pitch"""
    from g2s_b200 import synthetic
    gen = torch.Generator().manual_seed(seed)
    d = synthetic.make_depth(S, gen)
    return _warped(S, d, torch.tensor([[-0.4, 0.7, 0.2, 0.03, 0.02, -0.02]]))


def scene_far(S):
    """the step scene scaled so that the far sheet lies beyond `far` and part of the near sheet before `near`: rejected hits
    leave background or let what is behind them win"""
    (v, orc), _ = scene_step(S)
    return v * 80.0, orc           # z in ~[70, 90] x ... the far sheet's z*80 > 100 after the yaw for part of it


# ------------------------------------------------------------------------------------------------ back ends
def _oracle_maps(orc, verts, S, near=NEAR, far=FAR):
    """face map / depth map (image orientation, is x is) of the ORACLE for explicit camera-space vertices"""
    faces = orc.renderer._fill_back_faces(ro.get_face_idx(1, S, S))
    v = nr_port.projection(verts, orc.renderer.K, orc.renderer.R, orc.renderer.t, orc.renderer.dist_coeffs, S)
    maps = nr_port.forward_face_index_map(nr_port.vertices_to_faces(v, faces), 2 * S, near, far, mode="brute")
    return maps["face_index_map"].flip(1)[0].numpy().astype(np.int64), maps["depth_map"].flip(1)[0].numpy().astype(np.float64)


def _cuda_maps(orc, verts, S, near=NEAR, far=FAR):
    """the same from the PRODUCT: g2s_render_depth_fwd through the C ABI; it returns the face map and the 2x2-pooled depth,
    so the per-sub-pixel depth is taken from the pooled map only where the whole 2x2 block is unambiguous (see _check)"""
    import ctypes
    import g2s_b200
    from g2s_b200 import _lib
    from g2s_b200.functional import ZBuffer, _p, _stream
    lib = _lib.load()
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH)
    cam = _lib.Camera()
    ctypes.memmove(ctypes.byref(cam), ctypes.byref(ren._camera(depth_pass=True)), ctypes.sizeof(cam))
    cam.near_z, cam.far_z = near, far
    v = verts.cuda().contiguous()
    zb = ZBuffer().get(1, S, far, v.device)
    out = torch.empty(1, S, S, device="cuda")
    fidx = torch.empty(1, 2 * S, 2 * S, device="cuda", dtype=torch.int32)
    _lib.check(lib.g2s_render_depth_fwd(ctypes.byref(cam), _p(v), 1, _p(zb), _p(out), _p(fidx), _stream()), "render_depth")
    return fidx[0].cpu().numpy().astype(np.int64), out[0].cpu().numpy().astype(np.float64)


def _check(name, verts, orc, S, maps_fn, near=NEAR, far=FAR, pooled=False, min_unambiguous=0.9, expect=None):
    face, depth = maps_fn(orc, verts, S, near, far)
    ref = rc.raycast(verts[0].numpy(), S, orc.K[0].numpy(), near, far)
    ok = rc.unambiguous(ref)
    frac = float(ok.mean())
    assert frac > min_unambiguous, (name, frac)
    bad = ok & (face != ref["face"])
    assert int(bad.sum()) == 0, (name, int(bad.sum()), np.argwhere(bad)[:5], face[bad][:5], ref["face"][bad][:5])
    # depth: fp32 evaluation of the weights carries ~1e-3 x (z-range of the face) of noise on long faces (cancellation
    # in the 3x3 inverse at large sub-pixel coordinates); allow 3e-6 relative + that term
    cov = ok & (ref["tri"] >= 0)
    F = rc.grid_faces(S)
    zv = verts[0].numpy().astype(np.float64)[:, 2]
    zr = zv[F].max(1) - zv[F].min(1)
    if pooled:
        blk = ok.reshape(S, 2, S, 2).all(axis=(1, 3)) & (ref["tri"] >= 0).reshape(S, 2, S, 2).all(axis=(1, 3))
        want = ref["z"].reshape(S, 2, S, 2).mean(axis=(1, 3))
        tolmap = (3e-6 * want + 2e-3 * zr[np.maximum(ref["tri"], 0)].reshape(S, 2, S, 2).max(axis=(1, 3)))
        err = np.abs(depth - want)
        assert blk.sum() > 0.5 * (ref["tri"] >= 0).sum() / 4
        assert np.all(err[blk] <= tolmap[blk]), (name, float((err[blk] / want[blk]).max()))
    else:
        tol = 3e-6 * ref["z"] + 2e-3 * zr[np.maximum(ref["tri"], 0)]
        err = np.abs(depth - ref["z"])
        assert np.all(err[cov] <= tol[cov]), (name, float((err[cov] / ref["z"][cov]).max()))
        assert np.all(depth[ok & (ref["tri"] < 0)] == np.float32(far))
    if expect:
        expect(face, ref, ok)
    return face, ref, ok


def _run_all(maps_fn, pooled, S=24):
    Q2 = 2 * (S - 1) ** 2
    v, orc = scene_tilted_plane(S)
    face, ref, ok = _check("tilted_plane", v, orc, S, maps_fn, pooled=pooled)
    assert np.all(ref["nhit"][ref["tri"] >= 0] == 1)                  # a plane: nothing hides anything
    assert np.all(face[ok] < Q2)                                      # and every face is seen from the front
    # analytic depth: the plane through three of the vertices, intersected with every ray (independent of the caster)
    P = v[0].numpy().astype(np.float64)
    n = np.cross(P[1] - P[0], P[S] - P[0])
    c = n @ P[0]
    K = orc.K[0].numpy().astype(np.float64)
    cols = (np.arange(2 * S) + 0.5) / 2
    vv, uu = np.meshgrid(cols, cols, indexing="ij")
    D = np.stack([(uu - K[0, 2]) / K[0, 0], (vv - K[1, 2]) / K[1, 1], np.ones_like(uu)], -1)
    z_plane = c / (D @ n)
    cov = ok & (ref["tri"] >= 0)
    assert np.abs(ref["z"][cov] / z_plane[cov] - 1).max() < 1e-6

    seen_back = occluded = 0
    for k, (v, orc) in enumerate(scene_step(S)):                      # one yaw opens the step, the other folds it over
        face, ref, ok = _check("step%d" % k, v, orc, S, maps_fn, pooled=pooled, min_unambiguous=0.85)
        occluded += int((ref["nhit"][ok] > 1).sum())
        seen_back += int((face[ok] >= Q2).sum())
    assert occluded > 20                                              # real occlusion: rays that cross two sheets

    v, orc = scene_fold(S)
    face, ref, ok = _check("fold", v, orc, S, maps_fn, pooled=pooled, min_unambiguous=0.85)
    seen_back += int((face[ok] >= Q2).sum())
    assert seen_back > 200                                            # surfaces seen from behind: the fill_back copies win

    v, orc = scene_bumpy(S)
    _check("bumpy", v, orc, S, maps_fn, pooled=pooled, min_unambiguous=0.85)

    v, orc = scene_far(S)
    zmin, zmax = float(v[0, :, 2].min()), float(v[0, :, 2].max())
    assert zmin < 80.0 < zmax
    for near, far in ((0.1, 80.0), (78.0, 100.0), (76.0, 81.0)):
        face, ref, ok = _check("range[%g,%g]" % (near, far), v, orc, S, maps_fn, near=near, far=far, pooled=pooled,
                               min_unambiguous=0.85)
        assert (face[ok] >= 0).sum() > 50 and (face[ok] < 0).sum() > 50


def test_oracle_vs_raycaster_all_scenes():
    _run_all(_oracle_maps, pooled=False)


@pytest.mark.gpu
def test_cuda_vs_raycaster_all_scenes():
    _run_all(_cuda_maps, pooled=True)


# ------------------------------------------------------------------------------------------------ closed-form face map
def _grid_plane_vertices(S, a, b, c):
    """vertices ON the pixel rays (x, y) of the S x S grid, on the plane Z = a X + b Y + c: they project back to the
    integer pixel grid, so the face map has the identity-view layout whatever the tilt"""
    orc = oracle_renderer(S)
    K = orc.K[0].double()
    xs = torch.arange(S, dtype=torch.float64)
    yy, xx = torch.meshgrid(xs, xs, indexing="ij")
    rx, ry = (xx - K[0, 2]) / K[0, 0], (yy - K[1, 2]) / K[1, 1]
    s = c / (1.0 - a * rx - b * ry)
    return torch.stack([rx * s, ry * s, s], -1).reshape(1, -1, 3).float().contiguous(), orc


def _closed_form_face_map_checks(face, S):
    Q = (S - 1) ** 2
    qy, qx = np.meshgrid(np.arange(S - 1), np.arange(S - 1), indexing="ij")
    f1 = qy * (S - 1) + qx
    n = 2 * (S - 1)
    assert np.array_equal(face[0:n:2, 0:n:2], f1)                     # sub-pixel (0,0) of a quad: faces1
    assert np.array_equal(face[1:n:2, 1:n:2], f1 + Q)                 # (1,1): faces2
    for off in (face[0:n:2, 1:n:2], face[1:n:2, 0:n:2]):              # on the diagonal: either, decided by rounding
        assert np.all((off == f1) | (off == f1 + Q))
    assert np.all(face[n:, :] == -1) and np.all(face[:, n:] == -1)


def test_oracle_grid_plane_closed_form():
    S = 16
    v, orc = _grid_plane_vertices(S, 1.3, -0.8, 1.0)
    face, depth = _oracle_maps(orc, v, S)
    _closed_form_face_map_checks(face, S)
    K = orc.K[0].numpy().astype(np.float64)
    cols = (np.arange(2 * S) + 0.5) / 2
    vv, uu = np.meshgrid(cols, cols, indexing="ij")
    z = 1.0 / (1.0 - 1.3 * (uu - K[0, 2]) / K[0, 0] + 0.8 * (vv - K[1, 2]) / K[1, 1])
    cov = face >= 0
    assert np.abs(depth[cov] / z[cov] - 1).max() < 1e-6               # ray-plane intersection at every covered sub-pixel


@pytest.mark.gpu
def test_cuda_grid_plane_closed_form():
    S = 16
    v, orc = _grid_plane_vertices(S, 1.3, -0.8, 1.0)
    face, pooled = _cuda_maps(orc, v, S)
    _closed_form_face_map_checks(face, S)
    K = orc.K[0].numpy().astype(np.float64)
    cols = (np.arange(2 * S) + 0.5) / 2
    vv, uu = np.meshgrid(cols, cols, indexing="ij")
    z = 1.0 / (1.0 - 1.3 * (uu - K[0, 2]) / K[0, 0] + 0.8 * (vv - K[1, 2]) / K[1, 1])
    want = z.reshape(S, 2, S, 2).mean(axis=(1, 3))
    assert np.abs(pooled[:S - 1, :S - 1] / want[:S - 1, :S - 1] - 1).max() < 1e-6


# ------------------------------------------------------------------------------------------------ rgb
def _linear_colour_case(S):
    v, orc = scene_tilted_plane(S)
    P = v[0].double()
    # a colour field linear in the 3-D position: linear on the plane, so perspective-correct interpolation reproduces it
    coef = torch.tensor([[3.0, -2.0, 1.5], [-1.0, 4.0, 0.5], [2.0, 2.0, -3.0]], dtype=torch.float64)
    off = torch.tensor([0.1, -0.2, 0.3], dtype=torch.float64)
    col = ((P - P.mean(0)) @ coef.T + off).float()                    # [V,3], |col| < 1 for this scene
    im = col.T.reshape(1, 3, S, S).contiguous()
    return v, orc, im, col


def _expected_rgb(v, orc, S, col, near, far, bg=1.0):
    """perspective-correct barycentric blend of the vertex colours at the ray's hit point, with the REFERENCE's assignment of
    colours to vertices: get_textures_from_im (utils.py:98-109) stacks the colours of (y,x), (y,x+1), (y+1,x) for a faces1
    triangle whose vertices get_face_idx (utils.py:76-80) lists as (y,x), (y+1,x), (y,x+1) -- the two off-diagonal vertices
    carry each other's colour (in faces2 the first two vertices do).  That is what the reference renders; it is reproduced, not corrected."""
    ref = rc.raycast(v[0].numpy(), S, orc.K[0].numpy(), near, far)
    F = rc.grid_faces(S)
    Q = (S - 1) ** 2
    assign = F.copy()                             # vertex k of a face carries the colour of vertex assign[k]
    assign[:Q] = F[:Q][:, [0, 2, 1]]              # faces1 (y,x),(y+1,x),(y,x+1)   <- colours of (y,x),(y,x+1),(y+1,x)
    assign[Q:] = F[Q:][:, [1, 0, 2]]              # faces2 (y,x+1),(y+1,x),(y+1,x+1) <- colours of (y+1,x),(y,x+1),(y+1,x+1)
    tri = np.maximum(ref["tri"], 0)
    c = col.numpy().astype(np.float64)            # [V,3]
    colour = (ref["bary"][..., None] * c[assign[tri]]).sum(-2)
    colour[ref["tri"] < 0] = bg
    ok = rc.unambiguous(ref)
    blk = ok.reshape(S, 2, S, 2).all(axis=(1, 3))
    return colour.reshape(S, 2, S, 2, 3).mean(axis=(1, 3)).transpose(2, 0, 1), blk


def test_oracle_rgb_linear_colour_plane():
    S = 20
    v, orc, im, col = _linear_colour_case(S)
    with torch.no_grad():
        out = orc._mesh_view(im, v, 1, S, S)[0].numpy().astype(np.float64)     # nr.render_rgb(...).clamp(-1, 1)
    want, blk = _expected_rgb(v, orc, S, col, orc.renderer_min_depth, orc.renderer_max_depth)
    assert blk.mean() > 0.7
    # texture sampling clamps the cube coordinate to 1 - eps (eps = 1e-3): a deviation of <= 1e-3 x the colour step between
    # neighbouring vertices, only next to a vertex
    assert np.abs(out - np.clip(want, -1, 1))[:, blk].max() < 2e-5


@pytest.mark.gpu
def test_cuda_rgb_linear_colour_plane():
    import g2s_b200
    S = 20
    v, orc, im, col = _linear_colour_case(S)
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH)
    out = ren._render_rgb(v.cuda(), im.cuda())[0].cpu().numpy().astype(np.float64)
    want, blk = _expected_rgb(v, orc, S, col, ren.renderer_min_depth, ren.renderer_max_depth)
    assert np.abs(out - np.clip(want, -1, 1))[:, blk].max() < 2e-5


# ------------------------------------------------------------------------------------------------ gradients
def _gradient_case(S=12):
    v, orc = scene_bumpy(S, seed=5)
    faces = orc.renderer._fill_back_faces(ro.get_face_idx(1, S, S))
    vn = nr_port.projection(v, orc.renderer.K, orc.renderer.R, orc.renderer.t, orc.renderer.dist_coeffs, S)
    return v, orc, faces, vn


def test_oracle_backward_depth_map_matches_formula_f64():
    """oracle/nr_raster.c nr_backward_depth_map == SURVEY.md App. A.6 evaluated independently in float64 (from the same fp32
    projected vertices and the same face map); the x / y terms are cancelling sums, so fp32 carries ~1e-4 of noise there"""
    S = 12
    v, orc, faces, vn = _gradient_case(S)
    f9 = nr_port.vertices_to_faces(vn, faces).clone().requires_grad_(True)
    rgb, alpha, depth = nr_port.Rasterize.apply(f9, None, 2 * S, NEAR, FAR, 1e-4, [0, 0, 0], False, False, True)
    g = torch.randn(depth.shape, generator=torch.Generator().manual_seed(1))
    (depth * g).sum().backward()
    fmap = nr_port.LAST["face_index_map"][0].numpy()
    want = rc.depth_gradient_f64(vn[0].numpy(), fmap, g[0].numpy().astype(np.float64), S)
    got = f9.grad[0].numpy().astype(np.float64)
    scale = np.abs(want).max(axis=(0, 1), keepdims=True)
    assert np.abs(got - want)[..., 2].max() < 1e-5 * scale[..., 2].max()          # z: well conditioned
    assert np.abs(got - want)[..., :2].max() < 2e-3 * scale[..., :2].max()        # x, y: cancelling sums in fp32


def test_formula_is_the_exact_derivative_of_the_interpolated_depth():
    """App. A.6's x / y term is not an ad-hoc approximation for interior sub-pixels: a float64 finite difference of the
    perspective-correct depth with respect to a projected vertex coordinate reproduces it (unclamped weights)"""
    S = 12
    v, orc, faces, vn = _gradient_case(S)
    Vn = vn[0].numpy().astype(np.float64)
    F = rc.grid_faces(S)
    is_ = 2 * S

    def zp_of(Vn_, f, xi, yi):
        tri = Vn_[F[f]]
        p = 0.5 * (tri[:, :2] * is_ + is_ - 1)
        M = np.stack([p[:, 0], p[:, 1], np.ones(3)], 0)
        w = np.linalg.inv(M) @ np.array([xi, yi, 1.0])
        return 1.0 / (w / tri[:, 2]).sum(), w

    rng = np.random.default_rng(0)
    checked = 0
    for f in rng.permutation(F.shape[0])[:200]:
        tri = Vn[F[f]]
        p = 0.5 * (tri[:, :2] * is_ + is_ - 1)
        c = p.mean(0)
        xi, yi = int(round(c[0])), int(round(c[1]))
        if not (0 <= xi < is_ and 0 <= yi < is_):
            continue
        zp, w = zp_of(Vn, f, xi, yi)
        if w.min() < 0.05:
            continue
        fmap = np.full((is_, is_), -1)
        fmap[yi, xi] = f
        gmap = np.zeros((is_, is_))
        gmap[yi, xi] = 1.0
        ana = rc.depth_gradient_f64(Vn, fmap, gmap, S)[f]
        for k in range(3):
            for l in range(3):
                h = 1e-7
                Vp, Vm = Vn.copy(), Vn.copy()
                Vp[F[f][k], l] += h
                Vm[F[f][k], l] -= h
                fd = (zp_of(Vp, f, xi, yi)[0] - zp_of(Vm, f, xi, yi)[0]) / (2 * h)
                assert abs(fd - ana[k, l]) <= 1e-5 * max(1.0, abs(ana).max()), (f, k, l, fd, ana[k, l])
        checked += 1
    assert checked > 50


@pytest.mark.gpu
def test_cuda_backward_depth_map_matches_formula_f64():
    """the product's render_depth backward (vertex gradient through the C ABI) against the float64 formula, chained through
    the projection in float64 autograd"""
    import g2s_b200
    import g2s_b200.nr_compat as nrc
    S = 12
    v, orc, faces, vn = _gradient_case(S)
    r = nrc.Renderer(camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0., K=orc.K,
                     R=torch.eye(3)[None], t=torch.zeros(1, 3), near=0.1, far=10., image_size=S, orig_size=S,
                     fill_back=True, background_color=[1, 1, 1])
    vc = v.cuda().requires_grad_(True)
    out = r.render_depth(vc, ro.get_face_idx(1, S, S).cuda())
    g = torch.randn(1, S, S, generator=torch.Generator().manual_seed(2))
    (out * g.cuda()).sum().backward()
    # float64 reference: formula on the is x is map (every sub-pixel gets g / 4 of its output pixel, rows flipped back to
    # nr's native order), then the projection's Jacobian by float64 autograd
    orc_ren = orc
    _ = orc_ren.renderer.render_depth(v, ro.get_face_idx(1, S, S))
    fmap = nr_port.LAST["face_index_map"][0].numpy()
    g_sub = np.repeat(np.repeat(g[0].numpy().astype(np.float64) / 4, 2, 0), 2, 1)[::-1]
    gf = rc.depth_gradient_f64(vn[0].numpy(), fmap, g_sub, S)
    F4 = np.concatenate([rc.grid_faces(S), rc.grid_faces(S)[:, ::-1]], 0)
    gv_ndc = np.zeros((S * S, 3))
    np.add.at(gv_ndc, F4.reshape(-1), gf.reshape(-1, 3))
    vd = v[0].double().clone().requires_grad_(True)
    K = orc.K[0].double()
    x_, y_ = vd[:, 0] / (vd[:, 2] + 1e-9), vd[:, 1] / (vd[:, 2] + 1e-9)
    u = K[0, 0] * x_ + K[0, 1] * y_ + K[0, 2]
    w = K[1, 0] * x_ + K[1, 1] * y_ + K[1, 2]
    ndc = torch.stack([2 * (u - S / 2) / S, 2 * ((S - w) - S / 2) / S, vd[:, 2]], -1)
    (ndc * torch.tensor(gv_ndc)).sum().backward()
    want, got = vd.grad.numpy(), vc.grad[0].cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() < 2e-3 * np.abs(want).max()
