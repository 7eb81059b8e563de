"""Analytic known-answer tests that pin the oracle's restatement of the external rasteriser:
the reference holds no test or golden vector for it, so these closed forms are what anchors it."""
import torch

from helpers import oracle_renderer
from oracle import nr_port, renderer_oracle as ro


def _identity_view(P=1):
    return torch.zeros(P, 6)


def test_identity_flat_depth_known_answer():
    S, d0 = 8, 0.97
    orc = oracle_renderer(S)
    orc.set_transform_matrices(_identity_view())
    nr_port.MODE["raster"] = "brute"
    try:
        rd = orc.warp_canon_depth(torch.full((1, S, S), d0))
    finally:
        nr_port.MODE["raster"] = "culled"
    # half-pixel shift: rows/cols 0..S-2 see the surface, the last row and column are background -> clamp 1.2
    assert torch.allclose(rd[0, :S - 1, :S - 1], torch.full((S - 1, S - 1), d0), atol=2e-7)
    assert torch.all(rd[0, S - 1, :] == 1.2 - 0.0) or torch.allclose(rd[0, S - 1, :], torch.full((S,), 1.2))
    assert torch.allclose(rd[0, :, S - 1], torch.full((S,), 1.2))
    # closed-form face-index map (image orientation): quad (qy,qx) covers sub-pixels (2qy..2qy+1, 2qx..2qx+1);
    # (0,0),(0,1),(1,0) -> faces1 index, (1,1) -> faces2 index; the two diagonal sub-pixels are exact ties won by
    # the lower index
    fim = nr_port.LAST["face_index_map"].flip(1)[0]
    Q = (S - 1) ** 2
    for qy in range(S - 1):
        for qx in range(S - 1):
            f1 = qy * (S - 1) + qx
            assert fim[2 * qy, 2 * qx] == f1 and fim[2 * qy, 2 * qx + 1] == f1 and fim[2 * qy + 1, 2 * qx] == f1
            assert fim[2 * qy + 1, 2 * qx + 1] == Q + f1
    assert torch.all(fim[2 * (S - 1):, :] == -1) and torch.all(fim[:, 2 * (S - 1):] == -1)


def test_tie_rule_and_bbox_candidates():
    S = 8
    orc = oracle_renderer(S)
    orc.set_transform_matrices(_identity_view())
    g = orc.get_warped_3d_grid(torch.full((1, S, S), 1.0)).reshape(1, -1, 3)
    faces = orc.renderer._fill_back_faces(ro.get_face_idx(1, S, S))
    v = nr_port.projection(g, orc.renderer.K, orc.renderer.R, orc.renderer.t, orc.renderer.dist_coeffs, S)
    maps = nr_port.forward_face_index_map(nr_port.vertices_to_faces(v, faces), 2 * S, 0.1, 100.0, mode="brute",
                                          want_stats=True)
    covered = int((maps["face_index_map"] >= 0).sum())
    assert covered == 4 * (S - 1) ** 2
    # the two diagonal sub-pixels of every quad lie exactly on the shared edge: both faces pass the inside test with
    # (nearly) equal zp; exact-bit ties are a subset of those
    assert 0 < maps["stats"]["ties"] <= covered // 2
    assert maps["stats"]["outside_bbox"] == 0


def test_culled_equals_brute():
    from g2s_b200 import synthetic
    S, P = 24, 3
    case = synthetic.make_case(S, P, seed=7, rot_deg=120.0)
    orc = oracle_renderer(S)
    orc.set_transform_matrices(case["view"])
    g = orc.get_warped_3d_grid(case["depth"].expand(P, S, S)).reshape(P, -1, 3)
    faces = orc.renderer._fill_back_faces(ro.get_face_idx(P, S, S))
    v = nr_port.projection(g, orc.renderer.K, orc.renderer.R, orc.renderer.t, orc.renderer.dist_coeffs, S)
    fv = nr_port.vertices_to_faces(v, faces)
    a = nr_port.forward_face_index_map(fv, 2 * S, 0.1, 100.0, mode="brute", want_stats=True)
    b = nr_port.forward_face_index_map(fv, 2 * S, 0.1, 100.0, mode="culled")
    for k in ("face_index_map", "weight_map", "depth_map", "face_inv_map"):
        assert torch.equal(a[k], b[k]), k
    assert a["stats"]["outside_bbox"] == 0


def test_inv_grid_identity_is_linspace():
    S = 16
    orc = oracle_renderer(S)
    orc.set_transform_matrices(_identity_view())
    g = orc.get_inv_warped_2d_grid(torch.full((1, S, S), 1.03))
    lin = torch.linspace(-1, 1, S)
    assert torch.allclose(g[0, :, :, 0], lin.view(1, S).expand(S, S), atol=1e-5)
    assert torch.allclose(g[0, :, :, 1], lin.view(S, 1).expand(S, S), atol=1e-5)


def test_plane_normals_constant():
    S = 16
    orc = oracle_renderer(S)
    n = orc.get_normal_from_depth(torch.full((1, S, S), 1.0))
    # |tu x tv| = (2 d / f)^2 ~ 1e-3..1e-5, so the +1e-7 in the denominator is visible: n_z = len / (len + 1e-7)
    inner = n[0, 1:-1, 1:-1]
    assert inner[..., :2].abs().max() < 1e-6
    assert torch.allclose(inner[..., 2], inner[0, 0, 2].expand(S - 2, S - 2), atol=1e-6) and 0.9 < inner[0, 0, 2] <= 1.0
    assert torch.allclose(n[0, 0, :, :], torch.tensor([0., 0., 1.]).expand(S, 3), atol=1e-6)


def test_zp_gradient_wrt_vertex_z_matches_finite_difference():
    """d(zp)/d(z_k) of [nr] backward_depth_map is exact (the x/y gradient is approximate by design)."""
    torch.manual_seed(0)
    S = 8
    orc = oracle_renderer(S)
    orc.set_transform_matrices(_identity_view())
    depth = (1.0 + 0.02 * torch.rand(1, S, S)).double().float()
    g3 = orc.get_warped_3d_grid(depth).reshape(1, -1, 3)
    faces = orc.renderer._fill_back_faces(ro.get_face_idx(1, S, S))
    v = nr_port.projection(g3, orc.renderer.K, orc.renderer.R, orc.renderer.t, orc.renderer.dist_coeffs, S)
    fv = nr_port.vertices_to_faces(v, faces).clone().requires_grad_(True)
    _, _, dm = nr_port.Rasterize.apply(fv, None, 2 * S, 0.1, 100.0, 1e-4, [0, 0, 0], False, False, True)
    pix = (3, 5)
    dm[0, pix[0], pix[1]].backward()
    fn = int(nr_port.LAST["face_index_map"][0, pix[0], pix[1]])
    assert fn >= 0
    for k in range(3):
        h = 1e-3
        fp = fv.detach().clone(); fp[0, fn, k, 2] += h
        fm = fv.detach().clone(); fm[0, fn, k, 2] -= h
        zp = nr_port.forward_face_index_map(fp, 2 * S, 0.1, 100.0)["depth_map"][0, pix[0], pix[1]]
        zm = nr_port.forward_face_index_map(fm, 2 * S, 0.1, 100.0)["depth_map"][0, pix[0], pix[1]]
        fd = float(zp - zm) / (2 * h)
        assert abs(fd - float(fv.grad[0, fn, k, 2])) < 2e-3 * max(1.0, abs(fd))
