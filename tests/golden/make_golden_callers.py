"""Generates tests/golden/callers_*.npz by running the reference's OWN GAN2Shape/model.py and GAN2Shape/losses.py
(unmodified, on torch-CPU, through oracle/ref_model_shim.py) on seeded inputs: get_clamped_depth, get_shading, the
validity mask + PhotometricLoss, SmoothLoss, with their gradients.  Build container only (needs /root/reference):

    python tests/golden/make_golden_callers.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_model_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MIN_D, MAX_D = 0.9, 1.1
CASES = [("callers_s16_b3", 16, 3, 7), ("callers_s32_b2", 32, 2, 11)]


def run_case(S, B, seed):
    model, losses = ref_model_shim.load()
    G, ns = ref_model_shim.model_self(MIN_D, MAX_D)
    g = torch.Generator().manual_seed(seed)
    out = {}
    # ---- depth prologue: model.py:337-345 (one image, as the reference runs it)
    raw = (torch.randn(1, S, S, generator=g) * 0.7).requires_grad_(True)
    cot_d = torch.randn(1, S, S, generator=g)
    depth = G.get_clamped_depth(ns, raw, S, S)
    (depth * cot_d).sum().backward()
    out.update(depth_raw=raw.detach().numpy(), cot_depth=cot_d.numpy(), depth=depth.detach().numpy(),
               grad_depth_raw=raw.grad.numpy())
    # ---- shading: model.py:347-360 (one normal / albedo map broadcast over B lights, as step 3 does)
    normal = torch.nn.functional.normalize(torch.randn(1, S, S, 3, generator=g) + torch.tensor([0., 0., 1.5]), dim=3)
    normal = normal.requires_grad_(True)
    albedo = torch.tanh(torch.randn(1, 3, S, S, generator=g)).requires_grad_(True)
    light = (torch.rand(B, 4, generator=g) * 2 - 1).requires_grad_(True)
    a, b, d = G.get_lighting_directions(ns, light)
    diffuse, texture = G.get_shading(ns, normal, a, b, d, albedo)
    cot_dif = torch.randn(B, 1, S, S, generator=g)
    cot_tex = torch.randn(B, 3, S, S, generator=g)
    ((diffuse * cot_dif).sum() + (texture * cot_tex).sum()).backward()
    out.update(normal=normal.detach().numpy(), albedo=albedo.detach().numpy(), light=light.detach().numpy(),
               diffuse=diffuse.detach().numpy(), texture=texture.detach().numpy(), cot_diffuse=cot_dif.numpy(),
               cot_texture=cot_tex.numpy(), grad_normal=normal.grad.numpy(), grad_albedo=albedo.grad.numpy(),
               grad_light=light.grad.numpy())
    # ---- validity mask + photometric loss: model.py:265-274, losses.py:39-51
    recon_im = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).requires_grad_(True)
    target = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).requires_grad_(True)
    recon_depth = 0.8 + 0.4 * torch.rand(B, S, S, generator=g)
    recon_depth[recon_depth > 1.15] = 1.2                                    # background pixels sit exactly on the clamp
    masks = (torch.rand(B, 1, S, S, generator=g) > 0.2).float()
    margin = (MAX_D - MIN_D) / 2
    m = (recon_depth < MAX_D + margin).float().unsqueeze(1).detach() * masks  # model.py:268-269
    loss = losses.PhotometricLoss()(recon_im, target, mask=m)
    loss.backward()
    out.update(recon_im=recon_im.detach().numpy(), target=target.detach().numpy(), recon_depth=recon_depth.numpy(),
               masks=masks.numpy(), photo_loss=loss.detach().numpy(), grad_recon_im=recon_im.grad.numpy(),
               grad_target=target.grad.numpy(),
               photo_loss_nomask=losses.PhotometricLoss()(recon_im, target).detach().numpy())
    # ---- smooth loss: losses.py:54-79 on a depth-like [1,S,S] map and a shading-like [B,1,S,S] map
    sm = losses.SmoothLoss()
    d_map = depth.detach().clone().requires_grad_(True)
    s_map = diffuse.detach().clone().requires_grad_(True)
    l_d, l_s = sm(d_map), sm(s_map)
    (l_d * 1.5 + l_s * 0.5).backward()
    out.update(smooth_depth=l_d.detach().numpy(), smooth_shading=l_s.detach().numpy(), grad_smooth_depth=d_map.grad.numpy(),
               grad_smooth_shading=s_map.grad.numpy())
    return out


def main():
    for name, S, B, seed in CASES:
        out = run_case(S, B, seed)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
