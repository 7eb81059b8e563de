"""Generates tests/golden/*.npz by running the reference's OWN GAN2Shape/renderer code (unmodified, on torch-CPU,
through oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The rasteriser behind the reference's `nr.Renderer` calls is oracle/nr_port.py + oracle/nr_raster.c (neural_renderer is
not vendored by the reference; see oracle/nr_raster.c header).  The two caller-side formulas that live in
GAN2Shape/model.py (lighting directions :347-353, shading :355-360) are taken from oracle/renderer_oracle.py because
model.py cannot be imported without the StyleGAN2/LPIPS stack.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import g2s_b200  # noqa: E402
from g2s_b200 import synthetic  # noqa: E402
from oracle import ref_shim, nr_port, renderer_oracle as ro  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("s16_p3", 16, 3, 1234, 60.0), ("s32_p2", 32, 2, 4321, 90.0), ("s32_p2_wide", 32, 2, 99, 240.0)]


def run_case(S, P, seed, rot_deg, align_corners=False):
    case = synthetic.make_case(S, P, seed=seed, rot_deg=rot_deg)
    ref = ref_shim.make_renderer(S)
    depth = case["depth"].clone().requires_grad_(True)
    albedo = case["albedo"].clone().requires_grad_(True)
    view = case["view"].clone().requires_grad_(True)
    light = case["light"].clone().requires_grad_(True)
    out = {k: v.numpy() for k, v in case.items()}
    normal = ref.get_normal_from_depth(depth)
    a, b, d = ro.get_lighting_directions(light)
    _, texture = ro.get_shading(normal, a, b, d, albedo)
    ref.set_transform_matrices(view)
    out["rot_mat"] = ref.rot_mat.detach().numpy()
    out["trans_xyz"] = ref.trans_xyz.detach().numpy()
    recon_depth = ref.warp_canon_depth(depth.expand(P, S, S))
    out["face_idx"] = nr_port.LAST["face_index_map"].flip(1).numpy().astype(np.int32)
    grid = ref.get_inv_warped_2d_grid(recon_depth)
    recon_im = F.grid_sample(texture, grid, mode="bilinear", align_corners=align_corners).clamp(min=-1, max=1)
    fwd_grid = ref.get_warped_2d_grid(depth.expand(P, S, S))
    (recon_im * case["cotangent"]).sum().backward()
    out.update(normal=normal.detach().numpy(), recon_depth=recon_depth.detach().numpy(),
               inv_grid=grid.detach().numpy(), fwd_grid=fwd_grid.detach().numpy(), recon_im=recon_im.detach().numpy(),
               grad_depth=depth.grad.numpy(), grad_albedo=albedo.grad.numpy(), grad_view=view.grad.numpy(),
               grad_light=light.grad.numpy())
    # mesh-texture branch: render_yaw with 3 yaw angles (renderer.py:141-198, grid_sample=False)
    with torch.no_grad():
        yaw = ref.render_yaw(case["albedo"].clone(), case["depth"].clone(), maxr=40, nsample=3)
    out["render_yaw"] = yaw.numpy()
    return out


def main():
    nr_port.MODE["raster"] = "brute"   # the faithful loop
    for name, S, P, seed, rot in CASES:
        out = run_case(S, P, seed, rot)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
