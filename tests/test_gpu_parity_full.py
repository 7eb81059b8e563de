"""Parity of the CUDA path against the oracle AT THE BENCHED CONFIGURATIONS (BASELINE.json configs[1] and a slice of
configs[3]): the fused projected-view render (model.py:243-270, `RenderChainFn`) forward + backward, and
`warp_canon_depth` backward at 128^2.

Bars (BASELINE.json north_star): face-index maps bit-exact -- ZERO mismatches when both sides rasterise from the same R, t
bits; depth, images and every gradient within 1e-5.  Two measures are checked for every tensor:
  * rel_err  = max|a-b| / max|b|                      (the round-1 measure)
  * el_err   = max_i |a_i-b_i| / (|b_i| + floor)      (element-wise, floor = 1e-2 max|b|; helpers.el_err)
R, t reach both sides with the same bits (the oracle's get_transform_matrices, or the CUDA kernel's own output fed to the
oracle: `test_view_path_exact_given_cuda_R`), so nothing here needs a "few flipped pixels" allowance.

The measured errors are appended to gpurun_out/parity_stats.jsonl.
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, el_err, log_stats, oracle_renderer, rel_err
from oracle import nr_port, renderer_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-5        # north star: 1e-5 relative
TOL_EL = 1e-4     # element-wise: every element within 1e-4 of its OWN magnitude (floored at 1 % of the tensor's maximum)


def _cuda_renderer(S, align_corners=False):
    import g2s_b200
    return g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, align_corners=align_corners)


def _oracle_chain(orc, depth, albedo, light5, R_o, t_o, P, align):
    """model.py:243-270 for the P views of ONE image on the oracle; returns (recon_depth, face_idx, recon_im)"""
    S = orc.image_size
    normal = orc.get_normal_from_depth(depth)
    _, texture = ro.get_shading(normal, light5[:, 0:1], light5[:, 1:2], light5[:, 2:5], albedo)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    rd = orc.warp_canon_depth(depth.expand(P, S, S))
    f = nr_port.LAST["face_index_map"].flip(1).clone()
    grid = orc.get_inv_warped_2d_grid(rd)
    im = F.grid_sample(texture, grid, mode="bilinear", align_corners=align).clamp(min=-1, max=1)
    return rd, f, im


@pytest.mark.parametrize("S,P,N,rot,seed", [(128, 16, 1, 60.0, 1234), (128, 16, 3, 60.0, 99), (256, 8, 1, 60.0, 1234),
                                            (128, 32, 1, 120.0, 7)])
def test_benched_config_fused_chain_vs_oracle(S, P, N, rot, seed):
    """BASELINE.json configs[1] (cat: 128^2 x 16 views, also 3 images of it), a slice of configs[3] (face: 256^2 x 8) and the
    bulk shape's 32 views per image: forward bit-exact (faces, depth), image and all five gradients to 1e-5."""
    import g2s_b200
    from g2s_b200 import synthetic
    case = synthetic.make_case(S, P, seed=seed, n_images=N, rot_deg=rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    B = N * P
    R_all = ro.get_transform_matrices(case["view"])[0]
    t_all = case["view"][:, 3:].reshape(B, 1, 3)
    l5_all = torch.cat(ro.get_lighting_directions(case["light"]), 1)
    gen = torch.Generator().manual_seed(seed + 1)
    cot_im, cot_d = case["cotangent"], torch.randn(B, S, S, generator=gen) / (S * S)

    o = dict(rd=[], f=[], im=[], gd=[], ga=[], gR=[], gt=[], gl=[])
    for i in range(N):
        sl = slice(i * P, (i + 1) * P)
        d_o = case["depth"][i:i + 1].clone().requires_grad_(True)
        a_o = case["albedo"][i:i + 1].clone().requires_grad_(True)
        R_o, t_o = R_all[sl].clone().requires_grad_(True), t_all[sl].clone().requires_grad_(True)
        l_o = l5_all[sl].clone().requires_grad_(True)
        rd_o, f_o, im_o = _oracle_chain(orc, d_o, a_o, l_o, R_o, t_o, P, False)
        ((im_o * cot_im[sl]).sum() + (rd_o * cot_d[sl]).sum()).backward()
        for k, v in zip(("rd", "f", "im", "gd", "ga", "gR", "gt", "gl"),
                        (rd_o.detach(), f_o, im_o.detach(), d_o.grad, a_o.grad, R_o.grad, t_o.grad, l_o.grad)):
            o[k].append(v)
    o = {k: torch.cat(v, 0) for k, v in o.items()}

    depth = case["depth"].cuda().requires_grad_(True)
    albedo = case["albedo"].cuda().requires_grad_(True)
    R, t = R_all.cuda().requires_grad_(True), t_all.cuda().requires_grad_(True)
    l5 = l5_all.cuda().requires_grad_(True)
    im, rd, fidx = g2s_b200.functional.RenderChainFn.apply(depth, albedo, R, t, l5, ren, P, False)
    n_face = int((fidx.cpu() != o["f"]).sum())
    ((im * cot_im.cuda()).sum() + (rd * cot_d.cuda()).sum()).backward()
    got = dict(im=im.detach(), gd=depth.grad, ga=albedo.grad, gR=R.grad, gt=t.grad, gl=l5.grad)
    errs = {k: (rel_err(v.cpu(), o[k]), el_err(v.cpu(), o[k])) for k, v in got.items()}
    log_stats("benched_fused_chain", S=S, P=P, N=N, rot=rot, face_mismatches=n_face,
              depth_equal=bool(torch.equal(rd.detach().cpu(), o["rd"])), covered=float((o["f"] >= 0).float().mean()),
              **{k: v for k, v in errs.items()})
    assert n_face == 0
    assert torch.equal(rd.detach().cpu(), o["rd"])
    assert errs["im"][0] < TOL and errs["im"][1] < TOL_EL
    for k in ("gd", "ga", "gl", "gR", "gt"):
        assert errs[k][0] < TOL, (k, errs[k])
        assert errs[k][1] < TOL_EL, (k, errs[k])


@pytest.mark.parametrize("S,P,rot,seed", [(128, 8, 60.0, 5), (128, 4, 150.0, 6)])
def test_warp_canon_depth_backward_128(S, P, rot, seed):
    from g2s_b200 import synthetic
    case = synthetic.make_case(S, P, seed=seed, rot_deg=rot)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    d_o = case["depth"].clone().requires_grad_(True)
    R_o = ro.get_transform_matrices(case["view"])[0].clone().requires_grad_(True)
    t_o = case["view"][:, 3:].reshape(P, 1, 3).clone().requires_grad_(True)
    orc.rot_mat, orc.trans_xyz = R_o, t_o
    rd_o = orc.warp_canon_depth(d_o.expand(P, S, S))
    f_o = nr_port.LAST["face_index_map"].flip(1)
    cot = torch.randn(P, S, S, generator=torch.Generator().manual_seed(seed))
    (rd_o * cot).sum().backward()
    d = case["depth"].cuda().requires_grad_(True)
    ren.rot_mat = R_o.detach().cuda().requires_grad_(True)
    ren.trans_xyz = t_o.detach().cuda().requires_grad_(True)
    rd, fidx = ren.warp_canon_depth(d.expand(P, S, S), return_face_idx=True)
    assert int((fidx.cpu() != f_o).sum()) == 0
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    (rd * cot.cuda()).sum().backward()
    errs = dict(gd=(rel_err(d.grad.cpu(), d_o.grad), el_err(d.grad.cpu(), d_o.grad)),
                gR=(rel_err(ren.rot_mat.grad.cpu(), R_o.grad), el_err(ren.rot_mat.grad.cpu(), R_o.grad)),
                gt=(rel_err(ren.trans_xyz.grad.cpu(), t_o.grad), el_err(ren.trans_xyz.grad.cpu(), t_o.grad)))
    log_stats("warp_canon_depth_backward_128", S=S, P=P, rot=rot, **errs)
    for k, v in errs.items():
        assert v[0] < TOL and v[1] < TOL_EL, (k, v)


@pytest.mark.parametrize("S,P,seed", [(64, 6, 3), (128, 16, 1234)])
def test_view_path_exact_given_cuda_R(S, P, seed):
    """The path a user calls: view -> k_view_fwd -> render.  CUDA's sincosf differs from the CPU libm by an ulp, so instead of
    a blanket tolerance (a) R, t from the kernel are compared with the oracle's (<= 3e-7 absolute), (b) the oracle is run on the
    kernel's OWN R, t bits: zero face mismatches, depth bit-equal, image / gradients to 1e-5, and (c) the number of
    sub-pixels that flip when the oracle uses the CPU libm's R instead is counted and reported (they are rounding-decided)."""
    import g2s_b200
    from g2s_b200 import synthetic
    case = synthetic.make_case(S, P, seed=seed)
    orc, ren = oracle_renderer(S), _cuda_renderer(S)
    depth = case["depth"].cuda().requires_grad_(True)
    albedo = case["albedo"].cuda().requires_grad_(True)
    view = case["view"].cuda().requires_grad_(True)
    light = case["light"].cuda().requires_grad_(True)
    im, rd, fidx = ren.render_chain(depth, albedo, view, light)
    (im * case["cotangent"].cuda()).sum().backward()
    R_cuda, t_cuda = ren.rot_mat.detach().cpu(), ren.trans_xyz.detach().cpu()
    # (a)
    v_o = case["view"].clone().requires_grad_(True)
    R_cpu, t_cpu = ro.get_transform_matrices(v_o)
    assert (R_cuda - R_cpu.detach()).abs().max().item() <= 3e-7 and torch.equal(t_cuda, t_cpu.detach())
    # (b) oracle on the kernel's R bits; the gradient reaches `view` through the oracle's own get_transform_matrices
    d_o = case["depth"].clone().requires_grad_(True)
    a_o = case["albedo"].clone().requires_grad_(True)
    l_o = case["light"].clone().requires_grad_(True)
    l5_o = torch.cat(ro.get_lighting_directions(l_o), 1)
    R_in = R_cpu + (R_cuda - R_cpu).detach()
    assert torch.equal(R_in.detach(), R_cuda)
    rd_o, f_o, im_o = _oracle_chain(orc, d_o, a_o, l5_o, R_in, t_cpu, P, False)
    (im_o * case["cotangent"]).sum().backward()
    n_face = int((fidx.cpu() != f_o).sum())
    errs = dict(im=rel_err(im.detach().cpu(), im_o.detach()), gd=rel_err(depth.grad.cpu(), d_o.grad),
                ga=rel_err(albedo.grad.cpu(), a_o.grad), gv=rel_err(view.grad.cpu(), v_o.grad),
                gl=rel_err(light.grad.cpu(), l_o.grad))
    # (c) how many sub-pixels are decided by the last bit of R
    with torch.no_grad():
        orc.rot_mat, orc.trans_xyz = R_cpu.detach(), t_cpu.detach()
        orc.warp_canon_depth(case["depth"].expand(P, S, S))
        flips = int((nr_port.LAST["face_index_map"].flip(1) != f_o).sum())
    log_stats("view_path_exact_given_cuda_R", S=S, P=P, face_mismatches=n_face, flips_cpu_libm_R=flips,
              subpixels=int(f_o.numel()), **errs)
    assert n_face == 0
    assert torch.equal(rd.detach().cpu(), rd_o.detach())
    for k, v in errs.items():
        assert v < TOL, (k, v)
    assert flips <= 1e-4 * f_o.numel()
