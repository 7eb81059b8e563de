"""Exercises every kernel family once at small sizes (for compute-sanitizer runs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import g2s_b200
from g2s_b200 import synthetic
S, P, N = 33, 3, 2
cfg = {"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}
ren = g2s_b200.Renderer(cfg, S, 0.9, 1.1)
case = {k: v.cuda() for k, v in synthetic.make_case(S, P, seed=5, n_images=N).items()}
d, a = case["depth"].requires_grad_(True), case["albedo"].requires_grad_(True)
v, l = case["view"].requires_grad_(True), case["light"].requires_grad_(True)
im, rd, fidx = ren.render_chain(d, a, v, l, views_per_image=P)
pl, sl = g2s_b200.PhotometricLoss(), g2s_b200.SmoothLoss()
loss = pl(im, torch.rand_like(im), **g2s_b200.recon_im_mask(rd, 0.9, 1.1)) + sl(rd) + (rd * 0.01).sum()
loss.backward()
raw = torch.randn(N, S, S, device="cuda", requires_grad=True)
dep = g2s_b200.get_clamped_depth(raw, S, S, 0.9, 1.1, per_image=True)
nrm = ren.get_normal_from_depth(dep)
la, lb, ld = g2s_b200.get_lighting_directions(case["light"][:N])
dif, tex = g2s_b200.get_shading(nrm, la, lb, ld, case["albedo"])
(sl(dif) + tex.sum() * 1e-3).backward()
ren.set_transform_matrices(case["view"][:P])
rd2 = ren.warp_canon_depth(case["depth"][:1].detach().expand(P, S, S).clone().requires_grad_(True))
g = ren.get_inv_warped_2d_grid(rd2)
ren.grid_sample(case["albedo"][:1].expand(P, 3, S, S), g).sum().backward()
im2 = case["albedo"][:1].detach().clone().requires_grad_(True)
y = ren.render_yaw(im2, case["depth"][:1].detach(), maxr=40, nsample=3)
y.sum().backward()
ren.render_view(case["albedo"][:1].detach(), case["depth"][:1].detach(), maxr=[10, 30], nsample=[2, 2])
ren.render_pseudo_views(case["depth"].detach(), case["albedo"].detach(), case["view"].detach(), torch.rand(N * P, 1, device="cuda"),
                        torch.rand(N * P, 1, device="cuda"), torch.nn.functional.normalize(torch.rand(N * P, 3, device="cuda"), dim=1))
nr = g2s_b200.nr_compat.Renderer(camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0., K=ren.K[0].cpu(),
                                 near=0.1, far=10.0, image_size=S, orig_size=S, fill_back=True, background_color=[1, 1, 1])
verts = ren._grid3d(case["depth"][:1].detach().expand(P, S, S), None, None, None, case["view"][:P].detach()).requires_grad_(True)
nr.render_depth(verts, g2s_b200.get_face_idx(P, S, S).cuda()).sum().backward()
torch.cuda.synchronize()
print("tiny_all ok")
