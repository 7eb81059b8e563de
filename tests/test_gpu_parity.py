"""Parity of the CUDA path (through the C ABI, via the Renderer drop-in) against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): face-index maps bit-exact (exact-depth-tie pixels exempt -- none are needed: the
packed-key atomicMin reproduces the reference's tie rule, so the tests ask for ZERO mismatches when R, t are the same
bits on both sides); depth, images and gradients within 1e-5 relative (fp32; `rel_err` = max|a-b| / max|b|)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, close_except_few, golden, log_stats, oracle_renderer, rel_err
from oracle import nr_port, renderer_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cuda_renderer(S, align_corners=False):
    import g2s_b200
    return g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, align_corners=align_corners)


def _case(S, P, seed, rot):
    from g2s_b200 import synthetic
    return synthetic.make_case(S, P, seed=seed, rot_deg=rot)


def _inject(ren, orc):
    """same R, t bits on both sides (sin/cos differ by an ulp between the CPU and CUDA libms)"""
    ren.rot_mat = orc.rot_mat.detach().cuda()
    ren.trans_xyz = orc.trans_xyz.detach().cuda()


@pytest.mark.parametrize("S,P,rot,seed", [(16, 4, 60.0, 1), (32, 5, 120.0, 2), (33, 3, 90.0, 5), (64, 4, 60.0, 3),
                                          (128, 4, 60.0, 4), (128, 2, 180.0, 6)])
def test_warp_canon_depth_forward_bit_exact(S, P, rot, seed):
    case = _case(S, P, seed, rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    orc.set_transform_matrices(case["view"])
    _inject(ren, orc)
    rd_o = orc.warp_canon_depth(case["depth"].expand(P, S, S))
    f_o = nr_port.LAST["face_index_map"].flip(1)
    rd, fidx = ren.warp_canon_depth(case["depth"].cuda().expand(P, S, S), return_face_idx=True)
    assert int((fidx.cpu() != f_o).sum()) == 0
    assert torch.equal(rd.cpu(), rd_o)
    # materialised (non-expanded) depth gives the same answer, and the z-buffer was left clean
    rd2, fidx2 = ren.warp_canon_depth(case["depth"].cuda().repeat(P, 1, 1), return_face_idx=True)
    assert torch.equal(rd2, rd) and torch.equal(fidx2, fidx)


def test_warp_canon_depth_256_bit_exact_and_backward():
    """BASELINE.json face config size (256^2, 512^2 sub-pixels): one view against the oracle, forward bit-exact,
    backward to 1e-5"""
    S, P = 256, 1
    case = _case(S, P, 256, 60.0)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    d_o = case["depth"].clone().requires_grad_(True)
    R_o = ro.get_transform_matrices(case["view"])[0].clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    rd_o = orc.warp_canon_depth(d_o.expand(P, S, S))
    f_o = nr_port.LAST["face_index_map"].flip(1)
    cot = torch.randn(P, S, S)
    (rd_o * cot).sum().backward()
    d = case["depth"].cuda().requires_grad_(True)
    ren.rot_mat = R_o.detach().cuda().requires_grad_(True)
    ren.trans_xyz = t_o.detach().cuda().requires_grad_(True)
    rd, fidx = ren.warp_canon_depth(d.expand(P, S, S), return_face_idx=True)
    assert int((fidx.cpu() != f_o).sum()) == 0
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    (rd * cot.cuda()).sum().backward()
    assert rel_err(d.grad.cpu(), d_o.grad) < TOL
    assert rel_err(ren.rot_mat.grad.cpu(), R_o.grad) < TOL


@pytest.mark.parametrize("name", ["s16_p3", "s32_p2", "s32_p2_wide"])
def test_golden_forward(name):
    g = golden(name)
    S, P = g["depth"].shape[-1], g["view"].shape[0]
    ren = _cuda_renderer(S)
    ren.rot_mat = torch.tensor(g["rot_mat"]).cuda()
    ren.trans_xyz = torch.tensor(g["trans_xyz"]).cuda()
    depth = torch.tensor(g["depth"]).cuda()
    rd, fidx = ren.warp_canon_depth(depth.expand(P, S, S), return_face_idx=True)
    assert np.array_equal(fidx.cpu().numpy(), g["face_idx"])
    assert np.array_equal(rd.cpu().numpy(), g["recon_depth"])
    assert rel_err(ren.get_inv_warped_2d_grid(rd).cpu(), g["inv_grid"]) < 1e-6
    assert rel_err(ren.get_warped_2d_grid(depth.expand(P, S, S)).cpu(), g["fwd_grid"]) < 1e-6
    assert rel_err(ren.get_normal_from_depth(depth).cpu(), g["normal"]) < 1e-6


@pytest.mark.parametrize("name", ["s16_p3", "s32_p2", "s32_p2_wide"])
def test_golden_fused_chain_forward_backward(name):
    """whole path from (depth, albedo, view, light) against the vectors the REFERENCE's own code produced
    (tests/golden/make_golden.py).  The rasteriser sees the golden R bits (CUDA's sincosf differs from the CPU libm by an ulp,
    checked separately: <= 3e-7), while the backward still runs through k_view_bwd to `view`: faces and depth bit-exact,
    image and every gradient to 1e-5 -- no allowance for flipped pixels."""
    import g2s_b200
    g = golden(name)
    S, P = g["depth"].shape[-1], g["view"].shape[0]
    ren = _cuda_renderer(S)
    depth = torch.tensor(g["depth"]).cuda().requires_grad_(True)
    albedo = torch.tensor(g["albedo"]).cuda().requires_grad_(True)
    view = torch.tensor(g["view"]).cuda().requires_grad_(True)
    light = torch.tensor(g["light"]).cuda().requires_grad_(True)
    R_gpu, t_gpu = g2s_b200.functional.ViewToRtFn.apply(view)
    R_gold = torch.tensor(g["rot_mat"]).cuda()
    assert (R_gpu.detach() - R_gold).abs().max().item() <= 3e-7
    assert torch.equal(t_gpu.detach().reshape(-1), torch.tensor(g["trans_xyz"]).cuda().reshape(-1))
    R = R_gpu + (R_gold - R_gpu).detach()          # forward value: the golden bits; backward: through k_view_bwd
    assert torch.equal(R.detach(), R_gold)
    light5 = g2s_b200.functional.LightFn.apply(light)
    recon_im, recon_depth, fidx = g2s_b200.functional.RenderChainFn.apply(depth, albedo, R, t_gpu, light5, ren, P, False)
    assert np.array_equal(fidx.cpu().numpy(), g["face_idx"])
    assert np.array_equal(recon_depth.detach().cpu().numpy(), g["recon_depth"])
    assert rel_err(recon_im.detach().cpu(), g["recon_im"]) < TOL
    (recon_im * torch.tensor(g["cotangent"]).cuda()).sum().backward()
    errs = dict(gd=rel_err(depth.grad.cpu(), g["grad_depth"]), ga=rel_err(albedo.grad.cpu(), g["grad_albedo"]),
                gv=rel_err(view.grad.cpu(), g["grad_view"]), gl=rel_err(light.grad.cpu(), g["grad_light"]))
    log_stats("golden_fused_chain", name=name, **errs)
    for k, v in errs.items():
        assert v < TOL, (k, v)


@pytest.mark.parametrize("S,P,rot,seed", [(32, 3, 60.0, 11), (64, 2, 120.0, 12)])
@pytest.mark.parametrize("inverse", [False, True])
def test_warp_grid_forward_backward(S, P, rot, seed, inverse):
    case = _case(S, P, seed, rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    d_o = (case["depth"].expand(P, S, S) + 0.01 * torch.rand(P, S, S)).requires_grad_(True)
    R_o = ro.get_transform_matrices(case["view"])[0].clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    g_o = orc.get_inv_warped_2d_grid(d_o) if inverse else orc.get_warped_2d_grid(d_o)
    cot = torch.randn(P, S, S, 2)
    (g_o * cot).sum().backward()
    d = d_o.detach().cuda().requires_grad_(True)
    ren.rot_mat = R_o.detach().cuda().requires_grad_(True)
    ren.trans_xyz = t_o.detach().cuda().requires_grad_(True)
    g = ren.get_inv_warped_2d_grid(d) if inverse else ren.get_warped_2d_grid(d)
    assert torch.equal(g.detach().cpu(), g_o.detach())
    (g * cot.cuda()).sum().backward()
    assert rel_err(d.grad.cpu(), d_o.grad) < TOL
    assert rel_err(ren.rot_mat.grad.cpu(), R_o.grad) < TOL
    assert rel_err(ren.trans_xyz.grad.cpu(), t_o.grad) < TOL


@pytest.mark.parametrize("S,B", [(16, 2), (64, 3), (37, 1)])
def test_normal_forward_backward(S, B):
    from g2s_b200 import synthetic
    gen = torch.Generator().manual_seed(3)
    depth = synthetic.make_depth(S, gen, B)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    d_o = depth.clone().requires_grad_(True)
    n_o = orc.get_normal_from_depth(d_o)
    cot = torch.randn(B, S, S, 3, generator=gen)
    (n_o * cot).sum().backward()
    d = depth.cuda().requires_grad_(True)
    n = ren.get_normal_from_depth(d)
    assert rel_err(n.detach().cpu(), n_o.detach()) < TOL
    (n * cot.cuda()).sum().backward()
    assert rel_err(d.grad.cpu(), d_o.grad) < TOL


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("align", [False, True])
def test_grid_sample_forward_backward(mode, align):
    import g2s_b200
    gen = torch.Generator().manual_seed(9)
    B, C, H, W, Ho, Wo = 3, 3, 24, 20, 17, 29
    inp = torch.randn(B, C, H, W, generator=gen)
    grid = torch.rand(B, Ho, Wo, 2, generator=gen) * 2.6 - 1.3      # includes out-of-range samples
    grid[0, 0, 1] = torch.tensor([1e30, -1e30])
    cot = torch.randn(B, C, Ho, Wo, generator=gen)
    i_o, g_o = inp.clone().requires_grad_(True), grid.clone().requires_grad_(True)
    out_o = F.grid_sample(i_o, g_o, mode=mode, padding_mode="zeros", align_corners=align)
    (out_o * cot).sum().backward()
    i_c, g_c = inp.cuda().requires_grad_(True), grid.cuda().requires_grad_(True)
    out = g2s_b200.functional.grid_sample(i_c, g_c, mode, align)
    assert torch.allclose(out.detach().cpu(), out_o.detach(), rtol=1e-5, atol=1e-5, equal_nan=False)
    (out * cot.cuda()).sum().backward()
    assert torch.allclose(i_c.grad.cpu(), i_o.grad, rtol=1e-5, atol=1e-5)
    ok = torch.isfinite(g_o.grad).all(-1)
    assert torch.allclose(g_c.grad.cpu()[ok], g_o.grad[ok], rtol=1e-4, atol=1e-4)
    # broadcast (batch-stride-0) input, as render_given_view passes an expanded image
    one = inp[:1].cuda()
    o2 = g2s_b200.functional.grid_sample(one.expand(B, C, H, W), g_c.detach(), mode, align)
    o3 = g2s_b200.functional.grid_sample(one.repeat(B, 1, 1, 1), g_c.detach(), mode, align)
    assert torch.equal(o2, o3)


@pytest.mark.parametrize("S,P,rot,seed", [(16, 3, 60.0, 21), (32, 4, 90.0, 22), (64, 2, 60.0, 23)])
def test_warp_canon_depth_backward(S, P, rot, seed):
    case = _case(S, P, seed, rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    d_o = case["depth"].clone().requires_grad_(True)
    R_o = ro.get_transform_matrices(case["view"])[0].clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    rd_o = orc.warp_canon_depth(d_o.expand(P, S, S))
    cot = torch.randn(P, S, S)
    (rd_o * cot).sum().backward()
    d = case["depth"].cuda().requires_grad_(True)
    ren.rot_mat = R_o.detach().cuda().requires_grad_(True)
    ren.trans_xyz = t_o.detach().cuda().requires_grad_(True)
    rd = ren.warp_canon_depth(d.expand(P, S, S))
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    (rd * cot.cuda()).sum().backward()
    assert rel_err(d.grad.cpu(), d_o.grad) < TOL
    assert rel_err(ren.rot_mat.grad.cpu(), R_o.grad) < TOL
    assert rel_err(ren.trans_xyz.grad.cpu(), t_o.grad) < TOL


def _oracle_chain(orc, case, R_o, t_o, align):
    S = orc.image_size
    P = R_o.shape[0]
    depth = case["depth"].clone().requires_grad_(True)
    albedo = case["albedo"].clone().requires_grad_(True)
    light5 = torch.cat(ro.get_lighting_directions(case["light"]), 1).clone().requires_grad_(True)
    normal = orc.get_normal_from_depth(depth)
    _, texture = ro.get_shading(normal, light5[:, 0:1], light5[:, 1:2], light5[:, 2:5], albedo)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    rd = orc.warp_canon_depth(depth.expand(P, S, S))
    f = nr_port.LAST["face_index_map"].flip(1).clone()
    grid = orc.get_inv_warped_2d_grid(rd)
    im = F.grid_sample(texture, grid, mode="bilinear", align_corners=align).clamp(min=-1, max=1)
    return depth, albedo, light5, rd, f, im


@pytest.mark.parametrize("S,P,rot,seed,align", [(16, 3, 60.0, 31, False), (32, 4, 90.0, 32, False),
                                                (32, 2, 60.0, 33, True), (64, 3, 60.0, 34, False)])
def test_fused_chain_vs_oracle(S, P, rot, seed, align):
    import g2s_b200
    case = _case(S, P, seed, rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S, align_corners=align)
    R_o = ro.get_transform_matrices(case["view"])[0].clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
    depth_o, albedo_o, light_o, rd_o, f_o, im_o = _oracle_chain(orc, case, R_o, t_o, align)
    cot_im, cot_d = case["cotangent"], torch.randn(P, S, S) / (S * S)
    ((im_o * cot_im).sum() + (rd_o * cot_d).sum()).backward()

    depth = case["depth"].cuda().requires_grad_(True)
    albedo = case["albedo"].cuda().requires_grad_(True)
    R = R_o.detach().cuda().requires_grad_(True)
    t = t_o.detach().cuda().requires_grad_(True)
    light5 = light_o.detach().cuda().requires_grad_(True)
    im, rd, fidx = g2s_b200.functional.RenderChainFn.apply(depth, albedo, R, t, light5, ren, P, align)
    assert int((fidx.cpu() != f_o).sum()) == 0
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    assert rel_err(im.detach().cpu(), im_o.detach()) < TOL
    ((im * cot_im.cuda()).sum() + (rd * cot_d.cuda()).sum()).backward()
    assert rel_err(depth.grad.cpu(), depth_o.grad) < TOL
    assert rel_err(albedo.grad.cpu(), albedo_o.grad) < TOL
    assert rel_err(R.grad.cpu(), R_o.grad) < TOL
    assert rel_err(t.grad.cpu(), t_o.grad) < TOL
    assert rel_err(light5.grad.cpu(), light_o.grad) < TOL


def test_fused_equals_composition_multi_image():
    """n_images > 1: the fused launch equals the composition of the standalone operators image by image"""
    import g2s_b200
    from g2s_b200 import synthetic
    S, N, P = 32, 3, 4
    case = synthetic.make_case(S, P, seed=41, n_images=N)
    ren = _cuda_renderer(S)
    depth, albedo = case["depth"].cuda().requires_grad_(True), case["albedo"].cuda().requires_grad_(True)
    view, light = case["view"].cuda(), case["light"].cuda()
    cot = case["cotangent"].cuda()
    im, rd, fidx = ren.render_chain(depth, albedo, view, light, views_per_image=P)
    (im * cot).sum().backward()
    gd, ga = depth.grad.clone(), albedo.grad.clone()
    depth.grad = albedo.grad = None
    ims = []
    for i in range(N):
        sl = slice(i * P, (i + 1) * P)
        normal = ren.get_normal_from_depth(depth[i:i + 1])
        a, b, d = g2s_b200.get_lighting_directions(light[sl])
        _, tex = g2s_b200.get_shading(normal, a, b, d, albedo[i:i + 1])
        ren.set_transform_matrices(view[sl])
        rdi, fi = ren.warp_canon_depth(depth[i:i + 1].expand(P, S, S), return_face_idx=True)
        assert torch.equal(fi, fidx[sl]) and torch.equal(rdi, rd[sl])
        grid = ren.get_inv_warped_2d_grid(rdi)
        ims.append(ren.grid_sample(tex, grid).clamp(-1, 1))
    im2 = torch.cat(ims, 0)
    assert rel_err(im2.detach(), im.detach()) < TOL
    (im2 * cot).sum().backward()
    assert rel_err(depth.grad, gd) < TOL
    assert rel_err(albedo.grad, ga) < TOL


def test_render_pseudo_views_vs_oracle():
    """sample_pseudo_imgs path (model.py:291-328): relit texture warped by render_given_view(grid_sample=True) plus the
    nearest-neighbour mask warp, as one fused forward"""
    S, P = 32, 5
    case = _case(S, P, 71, 90.0)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    gen = torch.Generator().manual_seed(5)
    a = torch.rand(P, 1, generator=gen) * 0.5 + 0.3
    b = torch.rand(P, 1, generator=gen) * 0.5 + 0.3
    d = torch.cat([torch.rand(P, 2, generator=gen) - 0.5, torch.ones(P, 1)], 1)
    d = d / ((d ** 2).sum(1, keepdim=True)) ** 0.5
    mask = (torch.rand(1, 1, S, S, generator=gen) > 0.3).float()
    R = ro.get_transform_matrices(case["view"])[0]
    t = case["view"][:, 3:].reshape(P, 1, 3)
    with torch.no_grad():
        normal = orc.get_normal_from_depth(case["depth"])
        _, tex = ro.get_shading(normal, a, b, d, case["albedo"])
        orc.rot_mat, orc.trans_xyz = R, t
        rd = orc.warp_canon_depth(case["depth"].expand(P, S, S))
        grid = orc.get_inv_warped_2d_grid(rd)
        im_o = F.grid_sample(tex, grid, mode="bilinear", align_corners=False).clamp(-1, 1)
        m_o = F.grid_sample(mask.expand(P, 1, S, S), grid, mode="nearest", align_corners=False)
    import g2s_b200
    light5 = torch.cat([a, b, d], 1).cuda()
    with torch.no_grad():
        im, rd_c, fidx, m = g2s_b200.functional.RenderChainFn.apply(case["depth"].cuda(), case["albedo"].cuda(), R.cuda(),
                                                                   t.cuda(), light5, ren, P, False, mask[:, 0].cuda(), True)
        _, m1 = ren.render_pseudo_views(case["depth"].cuda(), case["albedo"].cuda(), case["view"].cuda(), a.cuda(),
                                        b.cuda(), d.cuda())
    assert torch.equal(rd_c.cpu(), rd)
    assert rel_err(im.cpu(), im_o) < TOL
    assert float((m.cpu() != m_o).float().mean()) < 2e-3      # nearest: a tap exactly between two texels may round apart
    assert m1.shape == (P, 1, S, S) and float(m1.min()) >= 0.0 and float(m1.max()) == 1.0


@pytest.mark.parametrize("name", ["s16_p3", "s32_p2"])
def test_render_yaw_mesh_branch_golden(name):
    g = golden(name)
    S = g["depth"].shape[-1]
    ren = _cuda_renderer(S)
    yaw = ren.render_yaw(torch.tensor(g["albedo"]).cuda(), torch.tensor(g["depth"]).cuda(), maxr=40, nsample=3)
    assert yaw.shape == g["render_yaw"].shape
    assert close_except_few(yaw.cpu(), g["render_yaw"], tol=1e-4)


@pytest.mark.parametrize("S,seed,yaw", [(32, 61, 0.5), (64, 62, -0.9), (48, 63, 0.0)])
def test_render_rgb_identical_vertices(S, seed, yaw):
    """mesh-texture render with the SAME 3-D vertices on both sides: face indices bit-exact, colours to 1e-5"""
    case = _case(S, 1, seed, 60.0)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    im, depth = case["albedo"], case["depth"]
    R, _ = ro.get_transform_matrices(torch.tensor([[0.1, yaw, -0.05]]))
    verts = orc.rotate_pts(orc.depth_to_3d_grid(depth).reshape(1, -1, 3), R)
    with torch.no_grad():
        out_o = orc._mesh_view(im, verts, 1, S, S)
    f_o = nr_port.LAST["face_index_map"].flip(1)
    out, fidx = ren._render_rgb(verts.cuda().contiguous(), im.cuda(), return_face_idx=True)
    assert int((fidx.cpu() != f_o).sum()) == 0
    assert rel_err(out.cpu(), out_o) < TOL


@pytest.mark.parametrize("S,seed", [(32, 51), (64, 52)])
def test_sweeps_and_given_view_vs_oracle(S, seed):
    case = _case(S, 2, seed, 60.0)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    im, depth = case["albedo"], case["depth"]
    vb = case["view"][:1] * 0.3
    va = case["view"][1:2] * 0.3
    with torch.no_grad():
        for kw in (dict(maxr=60, nsample=4), dict(maxr=30, nsample=2, v_before=vb, v_after=va),
                   dict(maxr=30, nsample=2, crop_mesh=(2, 1, 3, 0)), dict(maxr=30, nsample=3, grid_sample=True)):
            y_o = orc.render_yaw(im, depth, **kw)
            kc = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
            y = ren.render_yaw(im.cuda(), depth.cuda(), **kc)
            assert close_except_few(y.cpu(), y_o, tol=1e-4), kw
        v_o = orc.render_view(im, depth, maxr=[10, 40], nsample=[2, 3])
        v = ren.render_view(im.cuda(), depth.cuda(), maxr=[10, 40], nsample=[2, 3])
        assert close_except_few(v.cpu(), v_o, tol=1e-4)
        P = 2
        d2, im2, mask = depth.expand(P, S, S), im.expand(P, 3, S, S), torch.ones(P, 1, S, S)
        for gs in (True, False):
            a_o, m_o = orc.render_given_view(im2, d2, case["view"], mask=mask, grid_sample=gs)
            a, m = ren.render_given_view(im2.cuda(), d2.cuda(), case["view"].cuda(), mask=mask.cuda(), grid_sample=gs)
            assert close_except_few(a.cpu(), a_o, tol=1e-4)
            assert close_except_few(m.cpu(), m_o, tol=1e-4)


@pytest.mark.parametrize("S,seed", [(32, 61), (64, 62)])
def test_render_rgb_texture_gradient_vs_oracle(S, seed):
    """mesh-texture branch (render_yaw / render_given_view with grid_sample=False): gradient with respect to the image =
    [nr] backward_textures through get_textures_from_im, the fill_back permutation, the 2x2 mean and the clamp.  The
    geometry gets no gradient on either side (the oracle raises for it, the product returns None)."""
    case = _case(S, 2, seed, 60.0)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    depth = case["depth"]
    gen = torch.Generator().manual_seed(seed)
    im = case["albedo"] * 1.3                                       # some pixels beyond [-1,1]: the clamp masks their gradient
    kw = dict(maxr=50, nsample=3)
    im_o = im.clone().requires_grad_(True)
    y_o = orc.render_yaw(im_o, depth, **kw)
    cot = torch.randn(y_o.shape, generator=gen)
    (y_o * cot).sum().backward()
    im_c = im.cuda().requires_grad_(True)
    y = ren.render_yaw(im_c, depth.cuda(), **kw)
    (y * cot.cuda()).sum().backward()
    assert close_except_few(y.detach().cpu(), y_o.detach(), tol=1e-4)
    assert close_except_few(im_c.grad.cpu(), im_o.grad, tol=1e-4, frac=5e-3)
    # render_given_view, two views of an `expand`ed image (stride-0 batch) and a materialised copy give the same gradient
    P = 2
    d2 = depth.expand(P, S, S)
    cot2 = torch.randn(P, 3, S, S, generator=gen)
    im_o2 = im.clone().requires_grad_(True)
    a_o = orc.render_given_view(im_o2.expand(P, 3, S, S), d2, case["view"], grid_sample=False)
    (a_o * cot2).sum().backward()
    grads = []
    for materialise in (False, True):
        im_c2 = im.cuda().requires_grad_(True)
        src = im_c2.expand(P, 3, S, S)
        a = ren.render_given_view(src.contiguous() if materialise else src, d2.cuda(), case["view"].cuda(), grid_sample=False)
        (a * cot2.cuda()).sum().backward()
        grads.append(im_c2.grad.cpu())
    assert rel_err(grads[0], grads[1]) < TOL
    assert close_except_few(grads[0], im_o2.grad, tol=1e-4, frac=5e-3)


# ---- size-independent properties at BASELINE.json's full sizes ----------------------------------------------------
@pytest.mark.parametrize("S", [128, 256])
def test_identity_flat_depth_known_answer_full_size(S):
    ren = _cuda_renderer(S)
    ren.set_transform_matrices(torch.zeros(2, 6, device="cuda"))
    d0 = 0.97
    rd, fidx = ren.warp_canon_depth(torch.full((1, S, S), d0, device="cuda").expand(2, S, S), return_face_idx=True)
    assert torch.allclose(rd[:, :S - 1, :S - 1], torch.full((2, S - 1, S - 1), d0, device="cuda"), atol=3e-7)
    assert torch.all(rd[:, S - 1, :] == rd[0, S - 1, 0]) and abs(float(rd[0, S - 1, 0]) - 1.2) < 1e-6
    Q = (S - 1) ** 2
    qy, qx = torch.meshgrid(torch.arange(S - 1), torch.arange(S - 1), indexing="ij")
    f1 = (qy * (S - 1) + qx).int().cuda()
    for b in range(2):
        assert torch.equal(fidx[b, 0:2 * (S - 1):2, 0:2 * (S - 1):2], f1)
        assert torch.equal(fidx[b, 1:2 * (S - 1):2, 1:2 * (S - 1):2], f1 + Q)
        off = fidx[b, 0:2 * (S - 1):2, 1:2 * (S - 1):2]
        assert torch.all((off == f1) | (off == f1 + Q))     # diagonal sub-pixels: near-ties between the two faces
    assert torch.all(fidx[:, 2 * (S - 1):, :] == -1) and torch.all(fidx[:, :, 2 * (S - 1):] == -1)


@pytest.mark.parametrize("S,P", [(128, 16), (256, 8)])
def test_full_size_shard_invariance_and_determinism(S, P):
    """rendering views in two shards (as two ranks would) equals rendering them in one launch, bit for bit; forward is
    deterministic; gradients of the shards sum to the gradient of the whole (allclose: float atomics)."""
    case = _case(S, P, 77, 60.0)
    ren = _cuda_renderer(S)
    depth, albedo = case["depth"].cuda(), case["albedo"].cuda()
    view, light, cot = case["view"].cuda(), case["light"].cuda(), case["cotangent"].cuda()

    def run(sl):
        d, a = depth.clone().requires_grad_(True), albedo.clone().requires_grad_(True)
        im, rd, f = ren.render_chain(d, a, view[sl], light[sl])
        (im * cot[sl]).sum().backward()
        return im.detach(), rd.detach(), f, d.grad, a.grad

    whole = run(slice(0, P))
    again = run(slice(0, P))
    for x, y in zip(whole[:3], again[:3]):
        assert torch.equal(x, y)
    h0, h1 = run(slice(0, P // 2)), run(slice(P // 2, P))
    for k in range(3):
        assert torch.equal(torch.cat([h0[k], h1[k]], 0), whole[k])
    assert rel_err(h0[3] + h1[3], whole[3]) < TOL
    assert rel_err(h0[4] + h1[4], whole[4]) < TOL
    assert float((whole[2] >= 0).float().mean()) > 0.3
    assert torch.isfinite(whole[3]).all() and torch.isfinite(whole[4]).all()


@pytest.mark.parametrize("w", [3, 5, 6])
def test_view_and_light_kernels(w):
    """set_transform_matrices (utils.py:33-73) and get_lighting_directions (model.py:347-353) as single kernels"""
    import g2s_b200
    torch.manual_seed(w)
    view = (torch.rand(7, w) - 0.5) * 1.2
    light = torch.rand(7, 4) * 2 - 1
    v_o, l_o = view.clone().requires_grad_(True), light.clone().requires_grad_(True)
    R_o, t_o = ro.get_transform_matrices(v_o)
    l5_o = torch.cat(ro.get_lighting_directions(l_o), 1)
    cR, ct, cl = torch.randn(7, 3, 3), torch.randn(7, 1, 3), torch.randn(7, 5)
    ((R_o * cR).sum() + (t_o * ct).sum() + (l5_o * cl).sum()).backward()
    v, l = view.cuda().requires_grad_(True), light.cuda().requires_grad_(True)
    R, t = g2s_b200.functional.ViewToRtFn.apply(v)
    l5 = g2s_b200.functional.LightFn.apply(l)
    assert t.shape == (7, 1, 3)
    assert torch.allclose(R.detach().cpu(), R_o.detach(), atol=3e-7) and torch.equal(t.detach().cpu(), t_o.detach())
    assert torch.allclose(l5.detach().cpu(), l5_o.detach(), atol=3e-7)
    ((R * cR.cuda()).sum() + (t * ct.cuda()).sum() + (l5 * cl.cuda()).sum()).backward()
    assert rel_err(v.grad.cpu(), v_o.grad) < TOL
    assert rel_err(l.grad.cpu(), l_o.grad) < TOL


def test_shared_reciprocal_division_is_ieee_exact():
    """the kernels' division (one reciprocal shared by several quotients) must equal __fdiv_rn bit for bit"""
    import ctypes
    from g2s_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2, 3):
        _lib.check(lib.g2s_selftest_division(1 << 31, seed, ctypes.c_void_p(bad.data_ptr()), None), "selftest")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_face_index_division_exact_for_every_face():
    """face index -> vertex indices on the device (reciprocal estimate + fix-up) == integer arithmetic, every face, sizes up
    to the largest supported"""
    import ctypes
    from g2s_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for S in (2, 3, 17, 33, 128, 255, 256, 1000, 2047, 2048):
        _lib.check(lib.g2s_selftest_face_vertices(S, ctypes.c_void_p(bad.data_ptr()), None), "selftest_face_vertices")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_raster_fast_path_equals_ieee_path():
    """the per-face / per-hit fast arithmetic (shared reciprocals, structural operand-range guards) must equal the plain
    IEEE formulation bit for bit, including on degenerate and extreme-magnitude triangles that take the fall-backs"""
    import ctypes
    from g2s_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed, S in ((1, 128), (2, 256), (3, 33)):
        _lib.check(lib.g2s_selftest_raster(1 << 26, seed, S, ctypes.c_void_p(bad.data_ptr()), None), "selftest_raster")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


@pytest.mark.parametrize("S,P,N", [(32, 4, 1), (128, 16, 8), (64, 16, 132)])
def test_cuda_graph_step_matches_eager(S, P, N):
    """forward + backward captured once into a CUDA graph (small, launch-bound batches) == eager; the last case has
    more views than one forward chunk, so the capture includes the two-lane fork / join (events, internal stream)"""
    import g2s_b200
    from g2s_b200 import synthetic
    if N > 100:
        assert N * P > g2s_b200._lib.load().g2s_chunk_views(S)
    case = synthetic.make_case(S, P, seed=81, n_images=N)
    ren = _cuda_renderer(S)
    g = g2s_b200.graphs.GraphedRenderStep(ren, N, P)
    dev = {k: v.cuda() for k, v in case.items()}
    for _ in range(2):     # replay twice: the z-buffer must come back clean each time
        im, rd, fidx, grads = g.step(dev["depth"], dev["albedo"], dev["view"], dev["light"], dev["cotangent"])
    d, a = dev["depth"].clone().requires_grad_(True), dev["albedo"].clone().requires_grad_(True)
    v, l = dev["view"].clone().requires_grad_(True), dev["light"].clone().requires_grad_(True)
    im_e, rd_e, f_e = ren.render_chain(d, a, v, l, views_per_image=P)
    (im_e * dev["cotangent"]).sum().backward()
    assert torch.equal(fidx, f_e) and torch.equal(rd, rd_e) and torch.equal(im, im_e)
    for gg, ref in zip(grads, (d.grad, a.grad, v.grad, l.grad)):
        assert rel_err(gg, ref) < TOL


def test_host_render_step_matches_direct_call():
    """hostio.HostRenderStep (pinned host buffers in and out, three overlapped streams, two slots) returns what the
    direct device call returns, for every step of a stream of different batches"""
    import g2s_b200
    S, P, N = 32, 4, 3
    ren = _cuda_renderer(S)
    cases = [g2s_b200.synthetic.make_case(S, P, seed=40 + i, n_images=N) for i in range(5)]
    cot = cases[0]["cotangent"].cuda()
    step = g2s_b200.hostio.HostRenderStep(ren, N, P, cot, outputs=("recon_im",))
    got = []
    for c in cases:
        batch = {k: c[k].pin_memory() for k in ("depth", "albedo", "view", "light")}
        r = step.submit(batch)
        if r is not None:
            got.append({k: v.clone() for k, v in r.items()})
    got += [{k: v.clone() for k, v in r.items()} for r in step.drain()]
    assert len(got) == len(cases)
    for c, r in zip(cases, got):
        d, a, v, l = (c[k].cuda().requires_grad_(True) for k in ("depth", "albedo", "view", "light"))
        im = ren.render_chain(d, a, v, l, views_per_image=P)[0]
        gd, ga, gv, gl = torch.autograd.grad([im], [d, a, v, l], grad_outputs=[cot])
        assert torch.equal(r["recon_im"], im.detach().cpu())
        for name, want in (("grad_depth", gd), ("grad_albedo", ga), ("grad_view", gv), ("grad_light", gl)):
            assert rel_err(r[name], want.cpu()) < TOL, name


def test_error_behaviour():
    import g2s_b200
    ren = _cuda_renderer(32)
    with pytest.raises(Exception):
        ren.set_transform_matrices(torch.zeros(2, 4, device="cuda"))       # utils.py:70-71
    ren.set_transform_matrices(torch.zeros(2, 6, device="cuda"))
    with pytest.raises(RuntimeError):
        ren.warp_canon_depth(torch.ones(2, 16, 16, device="cuda"))          # wrong side
    with pytest.raises(RuntimeError):
        ren.warp_canon_depth(torch.ones(2, 32, 32))                         # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        g2s_b200.functional.grid_sample(torch.ones(1, 1, 4, 4, device="cuda"), torch.zeros(1, 2, 2, 2, device="cuda"),
                                        "bicubic", False)
    # 3- and 5-wide views (utils.py:60-69)
    for w in (3, 5):
        ren.set_transform_matrices(torch.zeros(2, w, device="cuda"))
        assert ren.trans_xyz.shape == (2, 1, 3)
        ren.warp_canon_depth(torch.ones(2, 32, 32, device="cuda"))
