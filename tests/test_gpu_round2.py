"""Round-2 additions, CUDA (through the C ABI) against the oracle:
  * the four public Renderer methods the reference class has and round 1 lacked (renderer.py:74-102): depth_to_3d_grid,
    grid_3d_to_2d, get_warped_3d_grid, get_inv_warped_3d_grid -- forward bit-exact, backward to 1e-5;
  * the grid operators and the fused chain after downscale_K (the rasteriser keeps the K captured at construction, the
    grid operators follow the rescaled K: renderer.py:56-59 vs 47-50);
  * the geometry gradient of an rgb render: neural_renderer's backward_pixel_map (renderer.py:196, 230, 248, 272, 275) on
    identical vertices;
  * workspace sizes, the caller-owned context, a second device.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, el_err, log_stats, oracle_renderer, rel_err
from oracle import nr_port, renderer_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cuda_renderer(S, **kw):
    import g2s_b200
    return g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, **kw)


def _views(P, seed):
    from g2s_b200 import synthetic
    return synthetic.make_views(P, torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("S,P,seed", [(24, 3, 1), (64, 2, 2)])
def test_public_3d_grid_methods_forward_backward(S, P, seed):
    from g2s_b200 import synthetic
    gen = torch.Generator().manual_seed(seed)
    depth = synthetic.make_depth(S, gen, P)
    view = _views(P, seed)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    for name in ("depth_to_3d_grid", "get_warped_3d_grid", "get_inv_warped_3d_grid"):
        d_o = depth.clone().requires_grad_(True)
        R_o = ro.get_transform_matrices(view)[0].clone().requires_grad_(True)
        t_o = view[:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
        orc.rot_mat, orc.trans_xyz = R_o, t_o
        out_o = getattr(orc, name)(d_o)
        cot = torch.randn(out_o.shape, generator=gen)
        (out_o * cot).sum().backward()
        d = depth.cuda().requires_grad_(True)
        ren.rot_mat = R_o.detach().cuda().requires_grad_(True)
        ren.trans_xyz = t_o.detach().cuda().requires_grad_(True)
        out = getattr(ren, name)(d)
        assert out.shape == out_o.shape
        assert torch.equal(out.detach().cpu(), out_o.detach()), name
        (out * cot.cuda()).sum().backward()
        assert rel_err(d.grad.cpu(), d_o.grad) < TOL, name
        if name != "depth_to_3d_grid":
            assert rel_err(ren.rot_mat.grad.cpu(), R_o.grad) < TOL, name
            assert rel_err(ren.trans_xyz.grad.cpu(), t_o.grad) < TOL, name
        # the expanded (batch-stride-0) depth the reference's callers pass gives the same values
        out2 = getattr(ren, name)(depth[:1].cuda().expand(P, S, S))
        out3 = getattr(ren, name)(depth[:1].cuda().repeat(P, 1, 1))
        assert torch.equal(out2, out3)
    # grid_3d_to_2d of an arbitrary 3-D grid, and the composition the reference writes (renderer.py:104-114)
    g3_o = (orc.get_warped_3d_grid(depth) + 0.01 * torch.randn(P, S, S, 3, generator=gen)).detach().requires_grad_(True)
    g2_o = orc.grid_3d_to_2d(g3_o)
    cot = torch.randn(g2_o.shape, generator=gen)
    (g2_o * cot).sum().backward()
    g3 = g3_o.detach().cuda().requires_grad_(True)
    g2 = ren.grid_3d_to_2d(g3)
    assert torch.equal(g2.detach().cpu(), g2_o.detach())
    (g2 * cot.cuda()).sum().backward()
    assert rel_err(g3.grad.cpu(), g3_o.grad) < TOL
    d = depth.cuda()
    assert torch.equal(ren.grid_3d_to_2d(ren.get_warped_3d_grid(d)), ren.get_warped_2d_grid(d))
    assert torch.equal(ren.grid_3d_to_2d(ren.get_inv_warped_3d_grid(d)), ren.get_inv_warped_2d_grid(d))


def test_after_downscale_K_grid_operators_follow_K_and_rasteriser_does_not():
    """renderer.py:56-59 rescales K / inv_K, but neural_renderer captured K at construction (renderer.py:47-50): the grid
    operators follow the new K, the rasteriser projects with the old one, and the fused chain equals that composition"""
    from g2s_b200 import synthetic
    S, P = 32, 3
    case = synthetic.make_case(S, P, seed=17)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    orc.downscale_K(2)
    ren.downscale_K(2)
    R = ro.get_transform_matrices(case["view"])[0]
    t = case["view"][:, 3:].reshape(P, 1, 3)
    orc.rot_mat, orc.trans_xyz = R, t
    ren.rot_mat, ren.trans_xyz = R.cuda(), t.cuda()
    d = case["depth"].expand(P, S, S)
    with torch.no_grad():
        rd_o = orc.warp_canon_depth(d)
        f_o = nr_port.LAST["face_index_map"].flip(1)
        grid_o = orc.get_inv_warped_2d_grid(rd_o)
        fwd_o = orc.get_warped_2d_grid(d)
        rd, fidx = ren.warp_canon_depth(d.cuda(), return_face_idx=True)
        assert int((fidx.cpu() != f_o).sum()) == 0 and torch.equal(rd.cpu(), rd_o)
        assert torch.equal(ren.get_inv_warped_2d_grid(rd).cpu(), grid_o)
        assert torch.equal(ren.get_warped_2d_grid(d.cuda()).cpu(), fwd_o)
        # fused chain == composition of the standalone operators with the SAME camera state
        import g2s_b200
        light5 = torch.cat(ro.get_lighting_directions(case["light"]), 1)
        normal_o = orc.get_normal_from_depth(case["depth"])
        _, tex_o = ro.get_shading(normal_o, light5[:, 0:1], light5[:, 1:2], light5[:, 2:5], case["albedo"])
        im_o = F.grid_sample(tex_o, grid_o, mode="bilinear", align_corners=False).clamp(-1, 1)
        im, rd2, f2 = g2s_b200.functional.RenderChainFn.apply(case["depth"].cuda(), case["albedo"].cuda(), R.cuda(), t.cuda(),
                                                              light5.cuda(), ren, P, False)
        assert torch.equal(f2, fidx) and torch.equal(rd2, rd)
        assert rel_err(im.cpu(), im_o) < TOL


@pytest.mark.parametrize("S,seed,yaw", [(24, 5, 0.5), (48, 6, -0.8)])
def test_render_rgb_geometry_gradient_vs_oracle(S, seed, yaw):
    """neural_renderer's backward_pixel_map through vertices_to_faces and the projection, on identical 3-D vertices.  One
    thread owns a face on both sides and the walks are the same, so the values agree to fp32 rounding; the sums over the
    faces around a vertex are accumulated with float atomics on the device (order tolerance: 1e-5 of the largest entry)."""
    from g2s_b200 import synthetic
    case = synthetic.make_case(S, 1, seed=seed)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    im, depth = case["albedo"], case["depth"]
    R, _ = ro.get_transform_matrices(torch.tensor([[0.1, yaw, -0.05]]))
    verts = orc.rotate_pts(orc.depth_to_3d_grid(depth).reshape(1, -1, 3), R).detach()
    gen = torch.Generator().manual_seed(seed)
    v_o, im_o = verts.clone().requires_grad_(True), im.clone().requires_grad_(True)
    out_o = orc._mesh_view(im_o, v_o, 1, S, S)
    cot = torch.randn(out_o.shape, generator=gen)
    (out_o * cot).sum().backward()
    v_c, im_c = verts.cuda().requires_grad_(True), im.cuda().requires_grad_(True)
    out = ren._render_rgb(v_c, im_c)
    (out * cot.cuda()).sum().backward()
    e_v, e_im = rel_err(v_c.grad.cpu(), v_o.grad), rel_err(im_c.grad.cpu(), im_o.grad)
    # the gradient is dominated by a few silhouette vertices: also compare the bulk
    big = v_o.grad.abs() > 1e-3 * v_o.grad.abs().max()
    e_bulk = ((v_c.grad.cpu() - v_o.grad).abs()[big] / v_o.grad.abs()[big]).max().item()
    log_stats("render_rgb_geometry_gradient", S=S, yaw=yaw, grad_v=e_v, grad_v_elementwise=e_bulk, grad_im=e_im,
              nonzero=float((v_o.grad != 0).float().mean()))
    assert rel_err(out.detach().cpu(), out_o.detach()) < TOL
    assert float((v_o.grad != 0).float().mean()) > 0.3
    assert e_v < TOL and e_im < TOL
    assert e_bulk < 1e-3
    # through the drop-in neural_renderer surface as well (nr.Renderer.render_rgb with cubes from get_textures_from_im)
    import g2s_b200
    import g2s_b200.nr_compat as nrc
    r = nrc.Renderer(camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0., K=orc.K,
                     R=torch.eye(3)[None], t=torch.zeros(1, 3), near=orc.renderer_min_depth, far=orc.renderer_max_depth,
                     image_size=S, orig_size=S, fill_back=True, background_color=[1, 1, 1])
    v2 = verts.cuda().requires_grad_(True)
    tex = g2s_b200.get_textures_from_im(im.cuda(), 2)
    out2 = r.render_rgb(v2, g2s_b200.get_face_idx(1, S, S).cuda(), tex).clamp(-1, 1)
    (out2 * cot.cuda()).sum().backward()
    assert rel_err(v2.grad.cpu(), v_o.grad) < TOL


def test_workspace_sizes_and_context():
    from g2s_b200 import _lib
    lib = _lib.load()
    S, n = 32, 5
    assert lib.g2s_workspace_bytes(_lib.WS_ZBUFFER, n, S) == lib.g2s_zbuffer_bytes(n, S)
    assert lib.g2s_workspace_bytes(_lib.WS_RASTER_BWD, n, S) == n * 9 * S * S * 4
    assert lib.g2s_workspace_bytes(_lib.WS_TEX_BWD, n, S) == n * 4 * S * S * 4
    assert lib.g2s_workspace_bytes(_lib.WS_TEXELS, n, S) == n * 8 * S * S * 4
    assert lib.g2s_workspace_bytes(_lib.WS_GRAD_NORMAL, n, S) == n * 3 * S * S * 4
    assert lib.g2s_workspace_bytes(_lib.WS_RGB_MAP, n, S) == n * 20 * S * S * 4
    assert lib.g2s_workspace_bytes(99, n, S) == 0 and lib.g2s_workspace_bytes(_lib.WS_TEXELS, 0, S) == 0
    a, b = _lib.Context(), _lib.Context()          # caller-owned: any number of them, independent
    assert a.handle.value and b.handle.value and a.handle.value != b.handle.value
    del a, b
    # a NULL context is accepted: the fused forward then runs single-lane
    import g2s_b200
    from g2s_b200 import synthetic
    S = 64
    n_img = lib.g2s_chunk_views(S) // 4 + 8                          # more views than one chunk: would use two lanes
    case = synthetic.make_case(S, 4, seed=3, n_images=n_img)
    ren = _cuda_renderer(S)
    dev = {k: v.cuda() for k, v in case.items()}
    with torch.no_grad():
        im, rd, f = ren.render_chain(dev["depth"], dev["albedo"], dev["view"], dev["light"], views_per_image=4)
        ren._context = lambda device: type("NullCtx", (), {"handle": ctypes.c_void_p(None)})()
        im2, rd2, f2 = ren.render_chain(dev["depth"], dev["albedo"], dev["view"], dev["light"], views_per_image=4)
    assert torch.equal(im, im2) and torch.equal(rd, rd2) and torch.equal(f, f2)


def test_second_device_if_present():
    """the > 48 KB shared-memory opt-in of the rasteriser's second stage is per DEVICE (round 1 did it once per process)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    from g2s_b200 import synthetic
    S, P = 32, 3
    case = synthetic.make_case(S, P, seed=9, rot_deg=150.0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            ren = _cuda_renderer(S, device=dev)
            ren.set_transform_matrices(case["view"].to(dev))
            rd, f = ren.warp_canon_depth(case["depth"].to(dev).expand(P, S, S), return_face_idx=True)
            outs.append((rd.cpu(), f.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_mesh_export_matches_the_rasterised_mesh():
    """mesh_export.depth_to_mesh: the vertices are the 3-D grid the rasteriser renders (renderer.py:74-80 / 90-95)"""
    import numpy as np
    from g2s_b200 import mesh_export as me, synthetic
    S = 16
    case = synthetic.make_case(S, 2, seed=4)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    meshes = me.depth_to_mesh(ren, case["depth"].cuda().expand(2, S, S), image=case["albedo"].cuda(), view=case["view"].cuda())
    orc.set_transform_matrices(case["view"])
    want = orc.get_warped_3d_grid(case["depth"].expand(2, S, S)).reshape(2, -1, 3).numpy()
    assert len(meshes) == 2
    for b in range(2):
        assert np.abs(meshes[b]["vertices"] - want[b]).max() < 1e-6        # R from the CUDA sincos: an ulp off the CPU's
        assert meshes[b]["faces"].shape == (2 * (S - 1) ** 2, 3) and meshes[b]["colors"].shape == (S * S, 3)
    plain = me.depth_to_mesh(ren, case["depth"].cuda())
    assert np.array_equal(plain[0]["vertices"], orc.depth_to_3d_grid(case["depth"]).reshape(-1, 3).numpy())


# ---- fused render + masked photometric loss (model.py:243-274 in one pass) ------------------------------------------
@pytest.mark.parametrize("S,N,P,use_mask,extra_cot", [(32, 1, 4, True, True), (64, 2, 3, True, False), (21, 1, 2, False, True),
                                                      (128, 1, 8, True, True)])
def test_render_chain_loss_vs_oracle(S, N, P, use_mask, extra_cot):
    """loss = PhotometricLoss(recon_im, target, mask=(recon_depth < max_depth + margin) * masks) taken inside the render
    (k_resolve epilogue + k_render_bwd_pixel<LOSS>) against the oracle's render chain + the oracle's loss: value and all
    five gradients, with and without a further cotangent on recon_im / recon_depth (the perceptual loss's path); and
    against the unfused CUDA composition render_chain -> PhotometricLoss."""
    import g2s_b200
    from g2s_b200 import synthetic
    from oracle import callers_oracle as co
    case = synthetic.make_case(S, P, seed=500 + S + P, n_images=N)
    B = N * P
    gen = torch.Generator().manual_seed(77 + S)
    target = torch.rand(B, 3, S, S, generator=gen) * 2 - 1
    masks = (torch.rand(B, 1, S, S, generator=gen) > 0.2).float() if use_mask else None
    cot_im = torch.randn(B, 3, S, S, generator=gen) / (3 * S * S * B)
    cot_d = torch.randn(B, S, S, generator=gen) / (S * S * B)
    thresh = MAX_DEPTH + (MAX_DEPTH - MIN_DEPTH) / 2

    # oracle: per image (its render chain takes one depth map), injected R, t so that both sides see the same views
    orc = oracle_renderer(S)
    R_all = ro.get_transform_matrices(case["view"])[0]
    d_o = case["depth"].clone().requires_grad_(True)
    a_o = case["albedo"].clone().requires_grad_(True)
    R_o = R_all.clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(B, 1, 3).clone().requires_grad_(True)
    l_o = case["light"].clone().requires_grad_(True)
    ims, rds = [], []
    for i in range(N):
        sl = slice(i * P, (i + 1) * P)
        normal = orc.get_normal_from_depth(d_o[i:i + 1])
        la, lb, ld = ro.get_lighting_directions(l_o[sl])
        _, tex = ro.get_shading(normal, la, lb, ld, a_o[i:i + 1])
        orc.rot_mat, orc.trans_xyz = R_o[sl], t_o[sl]
        rd = orc.warp_canon_depth(d_o[i:i + 1].expand(P, S, S))
        grid = orc.get_inv_warped_2d_grid(rd)
        ims.append(F.grid_sample(tex, grid, mode="bilinear", align_corners=False).clamp(min=-1, max=1))
        rds.append(rd)
    im_o, rd_o = torch.cat(ims, 0), torch.cat(rds, 0)
    loss_o = co.photometric_loss(im_o, target, co.recon_im_mask(rd_o, MIN_DEPTH, MAX_DEPTH, masks))
    total_o = loss_o * 1.7
    if extra_cot:
        total_o = total_o + (im_o * cot_im).sum() + (rd_o * cot_d).sum()
    total_o.backward()

    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")
    d = case["depth"].cuda().requires_grad_(True)
    a = case["albedo"].cuda().requires_grad_(True)
    R = R_all.cuda().requires_grad_(True)
    t = t_o.detach().cuda().requires_grad_(True)
    light5 = g2s_b200.functional.LightFn.apply(case["light"].cuda()).detach().requires_grad_(True)
    tg, mk = target.cuda(), masks.cuda() if use_mask else None
    im, rd, fidx, loss = g2s_b200.functional.RenderChainLossFn.apply(d, a, R, t, light5, tg, mk, ren, P, False, thresh)
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    assert rel_err(im.detach().cpu(), im_o.detach()) < TOL
    assert rel_err(loss.detach().cpu(), loss_o.detach()) < TOL
    total = loss * 1.7
    if extra_cot:
        total = total + (im * cot_im.cuda()).sum() + (rd * cot_d.cuda()).sum()
    total.backward()
    errs = dict(gd=rel_err(d.grad.cpu(), d_o.grad), ga=rel_err(a.grad.cpu(), a_o.grad), gR=rel_err(R.grad.cpu(), R_o.grad),
                gt=rel_err(t.grad.cpu(), t_o.grad))
    log_stats("render_chain_loss", S=S, N=N, P=P, loss=rel_err(loss.detach().cpu(), loss_o.detach()), **errs)
    assert max(errs.values()) < TOL, errs
    # light: compare through the raw light (LightFn's chain is tested elsewhere)
    lraw = case["light"].cuda().requires_grad_(True)
    g2s_b200.functional.LightFn.apply(lraw).backward(light5.grad)
    assert rel_err(lraw.grad.cpu(), l_o.grad) < TOL

    # the unfused CUDA composition gives the same loss and gradients
    d2 = case["depth"].cuda().requires_grad_(True)
    a2 = case["albedo"].cuda().requires_grad_(True)
    im2, rd2, _ = g2s_b200.functional.RenderChainFn.apply(d2, a2, R.detach(), t.detach(), light5.detach(), ren, P, False)
    assert torch.equal(im2, im.detach()) and torch.equal(rd2, rd.detach())
    kw = g2s_b200.recon_im_mask(rd2.detach(), MIN_DEPTH, MAX_DEPTH)
    loss2 = g2s_b200.PhotometricLoss()(im2, tg, mask=mk, **kw)
    assert rel_err(loss2.detach(), loss.detach()) < 1e-6
    tot2 = loss2 * 1.7
    if extra_cot:
        tot2 = tot2 + (im2 * cot_im.cuda()).sum() + (rd2 * cot_d.cuda()).sum()
    tot2.backward()
    assert rel_err(d2.grad, d.grad) < TOL and rel_err(a2.grad, a.grad) < TOL      # fp32 atomics: order noise only


def test_render_chain_loss_api_and_errors():
    import g2s_b200
    from g2s_b200 import synthetic
    S, P = 32, 4
    case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=3).items()}
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")
    target = torch.zeros(P, 3, S, S, device="cuda")
    d = case["depth"].clone().requires_grad_(True)
    loss, im, rd, fidx = ren.render_chain_loss(d, case["albedo"], case["view"], case["light"], target, views_per_image=P)
    # against the plain mean over the valid pixels of |recon_im - 0|
    m = (rd < MAX_DEPTH + (MAX_DEPTH - MIN_DEPTH) / 2).float().unsqueeze(1).expand_as(im)
    assert abs(loss.item() - ((im.abs() * m).sum() / m.sum()).item()) < 1e-6
    loss.backward()
    assert d.grad is not None and torch.isfinite(d.grad).all() and d.grad.abs().sum() > 0
    with pytest.raises(RuntimeError):
        ren.render_chain_loss(d, case["albedo"], case["view"], case["light"], target[:, :2], views_per_image=P)
    lib = g2s_b200._lib.load()
    assert lib.g2s_workspace_bytes(g2s_b200._lib.WS_LOSS, P, S) == (P * 1 * (S // 4) + 512) * 16


def test_projection_handoff_equals_recompute():
    """the forward hands its projected vertices to the backward (proj_ws); the gradients equal the ones of a backward that
    projects the mesh again (share_projection = False), bit for bit up to the order of the float atomics"""
    import g2s_b200
    from g2s_b200 import synthetic
    for S, N, P in ((33, 2, 3), (128, 1, 8)):           # 33: the mesh's last row / column sits in a tile's 17th slot
        case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=9, n_images=N).items()}
        grads = []
        for share in (True, False):
            ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")
            ren.share_projection = share
            d = case["depth"].clone().requires_grad_(True)
            v = case["view"].clone().requires_grad_(True)
            im, rd, _ = ren.render_chain(d, case["albedo"], v, case["light"], views_per_image=P)
            ((im * case["cotangent"]).sum() + rd.sum() * 1e-3).backward()
            grads.append((d.grad.clone(), v.grad.clone()))
        assert rel_err(grads[0][0], grads[1][0]) < TOL and rel_err(grads[0][1], grads[1][1]) < TOL


def test_render_chain_loss_multi_chunk_two_lanes():
    """640 views at 128^2: two forward chunks on two lanes, the two-lane backward and the projection hand-off all active;
    the fused loss and its gradients equal the unfused CUDA composition (which is held to the oracle at smaller sizes)"""
    import g2s_b200
    from g2s_b200 import synthetic
    S, N, P = 128, 40, 16
    case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=21, n_images=N).items()}
    B = N * P
    assert B > g2s_b200._lib.load().g2s_chunk_views(S)
    gen = torch.Generator().manual_seed(3)
    target = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).cuda()
    masks = (torch.rand(B, 1, S, S, generator=gen) > 0.3).float().cuda()
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")
    outs = []
    for fused in (True, False):
        d = case["depth"].clone().requires_grad_(True)
        a = case["albedo"].clone().requires_grad_(True)
        v = case["view"].clone().requires_grad_(True)
        l = case["light"].clone().requires_grad_(True)
        if fused:
            loss, im, rd, _ = ren.render_chain_loss(d, a, v, l, target, masks, views_per_image=P)
        else:
            im, rd, _ = ren.render_chain(d, a, v, l, views_per_image=P)
            loss = g2s_b200.PhotometricLoss()(im, target, mask=masks, **g2s_b200.recon_im_mask(rd.detach(), MIN_DEPTH, MAX_DEPTH))
        loss.backward()
        outs.append((loss.detach(), im.detach(), d.grad, a.grad, v.grad, l.grad))
    assert torch.equal(outs[0][1], outs[1][1])
    assert rel_err(outs[0][0], outs[1][0]) < 1e-6
    for k in range(2, 6):
        assert rel_err(outs[0][k], outs[1][k]) < TOL, k


def test_projection_handoff_short_last_chunk_keeps_scratch_clean():
    """More views than one backward chunk with a SHORT last chunk: its masked quarter gradient must not land in the
    zero-at-rest front of the raster scratch (it did when the layout followed the chunk's own size), or the next call
    reads stale values as vertex gradients.  Two calls in a row against the recomputing backward."""
    import g2s_b200
    from g2s_b200 import synthetic
    S = 64
    chunk = g2s_b200._lib.load().g2s_chunk_views_bwd(S)
    P = chunk + 300
    case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=13, n_images=1).items()}
    ref = None
    for share, reps in ((False, 1), (True, 2)):
        ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")
        ren.share_projection = share
        for _ in range(reps):
            d = case["depth"].clone().requires_grad_(True)
            v = case["view"].clone().requires_grad_(True)
            im, rd, _ = ren.render_chain(d, case["albedo"], v, case["light"], views_per_image=P)
            (im * case["cotangent"]).sum().backward()
            if ref is None:
                ref = (d.grad.clone(), v.grad.clone())
            else:
                # (10 000 views' fp32 atomics into one image: the summation order alone is worth a few 1e-6)
                assert rel_err(d.grad, ref[0]) < 2e-5 and rel_err(v.grad, ref[1]) < TOL
        if share:      # and the front of the kept scratch is zero at rest
            (buf, _), = ren._raster_scratch.buf.values()
            assert float(buf[: chunk * 4 * S * S].abs().max()) == 0.0


def test_index_math_exact():
    """the kernels' cheap index arithmetic against integer division: v / S, v % S (float estimate + fix-up) for every vertex at
    sizes up to the largest supported; view / views_per_image (multiply-high by ceil(2^32 / vpi), or the division when the
    constant is not exact for the call's view count) for every view, incl. non-power-of-two and large view counts"""
    from g2s_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for S in (2, 3, 17, 33, 128, 255, 256, 1000, 2047, 2048):
        _lib.check(lib.g2s_selftest_index_math(S, 16, 4096, ctypes.c_void_p(bad.data_ptr()), None), "selftest_index_math")
    for vpi, n in ((1, 1000), (2, 5000), (3, 100000), (7, 3000000), (16, 1 << 20), (1000, 4000000), (1024, 1 << 22),
                   (12345, 1 << 24), (3, 1 << 30), (65536, 1 << 30)):
        _lib.check(lib.g2s_selftest_index_math(128, vpi, n, ctypes.c_void_p(bad.data_ptr()), None), "selftest_index_math")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_more_views_than_the_grid_limit_at_a_tiny_size():
    """S = 8 with 33 000 views of one image: more views than one launch may carry (32 768: gridDim limits), forward and
    backward; equals the same views rendered in two halves"""
    import g2s_b200
    from g2s_b200 import synthetic
    S, P = 8, 33000
    case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=17, n_images=1).items()}
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device="cuda")

    def run(sl):
        d = case["depth"].clone().requires_grad_(True)
        a = case["albedo"].clone().requires_grad_(True)
        v = case["view"][sl].clone().requires_grad_(True)
        n = v.shape[0]
        im, rd, f = ren.render_chain(d, a, v, case["light"][sl], views_per_image=n)
        (im * case["cotangent"][sl]).sum().backward()
        return im.detach(), rd.detach(), f, d.grad, a.grad, v.grad

    whole = run(slice(0, P))
    h1, h2 = run(slice(0, P // 2)), run(slice(P // 2, P))
    for k in range(3):
        assert torch.equal(whole[k], torch.cat([h1[k], h2[k]], 0))
    # grad_depth / grad_albedo: 33 000 views' fp32 atomics into ONE 8 x 8 image -- the summation order alone moves the result by
    # ~sqrt(n) ulp (5e-6 was seen); the per-view gradient has no such accumulation
    assert rel_err(whole[3], h1[3] + h2[3]) < 1e-4 and rel_err(whole[4], h1[4] + h2[4]) < 1e-4
    assert rel_err(whole[5], torch.cat([h1[5], h2[5]], 0)) < TOL
