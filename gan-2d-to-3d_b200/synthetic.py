"""Synthetic inputs for parity tests and bench.py.

All tensors are built on the CPU with a seeded generator and copied by the caller, so the oracle and
the CUDA path see bit-identical inputs.  Shapes and value ranges follow the reference's own callers:
ellipsoid depth prior (GAN2Shape/priors.py:74-97), depth clamp + border rule (model.py:337-345),
view scaling (model.py:330-335), lighting mapping (model.py:347-353).
"""
import math

import torch
import torch.nn.functional as F

MIN_DEPTH, MAX_DEPTH = 0.9, 1.1           # model.py:49-50
BORDER_DEPTH = 0.7 * MAX_DEPTH + 0.3 * MIN_DEPTH  # model.py:51


def make_depth(S, gen, n_images=1):
    """[n,S,S] ellipsoid (near .91, far 1.02, radius .4, r_pixel .35 S) + smooth noise, clipped, bordered."""
    near, far, radius = 0.91, 1.02, 0.4
    r_pixel = 0.35 * S
    c = (S - 1) / 2
    i, j = torch.meshgrid(torch.linspace(0, S - 1, S), torch.linspace(0, S - 1, S), indexing="ij")
    temp = math.sqrt(radius ** 2 - (radius - (far - near)) ** 2)
    dist = torch.sqrt((i - c) ** 2 + (j - c) ** 2)
    area = dist <= r_pixel
    dist_rescale = dist / r_pixel * temp
    ell = radius - torch.sqrt(torch.abs(radius ** 2 - dist_rescale ** 2)) + near
    base = torch.full((S, S), far)
    base[area] = ell[area]
    lo = max(S // 8, 2)
    noise = torch.randn(n_images, 1, lo, lo, generator=gen)
    noise = F.interpolate(noise, size=(S, S), mode="bicubic", align_corners=False)[:, 0]
    depth = (base.unsqueeze(0) + 0.01 * noise).clamp(MIN_DEPTH, MAX_DEPTH)
    if S > 4:
        border = torch.zeros(1, S, S - 4)
        border = F.pad(border, (2, 2), mode="constant", value=1.02)   # model.py:341-343 (sic: 1.02)
        depth = depth * (1 - border) + border * BORDER_DEPTH
    return depth.float().contiguous()


def make_views(P, gen, rot_deg=60.0, xy=0.1, z=0.1):
    """[P,6] = U(-.5,.5) scaled as model.py:330-335 (=> +-30 degrees by default)."""
    u = torch.rand(P, 6, generator=gen) - 0.5
    return torch.cat([u[:, :3] * math.pi / 180 * rot_deg, u[:, 3:5] * xy, u[:, 5:] * z], 1).float().contiguous()


def make_albedo(S, gen, n_images=1):
    lo = max(S // 4, 2)
    a = torch.randn(n_images, 3, lo, lo, generator=gen)
    return torch.tanh(F.interpolate(a, size=(S, S), mode="bicubic", align_corners=False)).float().contiguous()


def make_light(P, gen):
    """Raw lighting vector [P,4] = U(-1,1), mapped to (ambient, diffuse, direction) by the renderer."""
    return (torch.rand(P, 4, generator=gen) * 2 - 1).float().contiguous()


def make_cotangent(P, S, gen):
    return (torch.randn(P, 3, S, S, generator=gen) / (3 * S * S)).float().contiguous()


def make_case(S, P, seed=1234, n_images=1, rot_deg=60.0):
    """One workload: n_images images, P views each (view-major per image)."""
    gen = torch.Generator().manual_seed(seed)
    return dict(
        depth=make_depth(S, gen, n_images),
        albedo=make_albedo(S, gen, n_images),
        view=make_views(n_images * P, gen, rot_deg=rot_deg),
        light=make_light(n_images * P, gen),
        cotangent=make_cotangent(n_images * P, S, gen),
    )
