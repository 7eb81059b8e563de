"""oracle/callers_oracle.py (the CPU restatement of the callers either side of the path: SURVEY.md 8f rows 1 and 3) pinned
(a) against the reference's own GAN2Shape/model.py + GAN2Shape/losses.py imported unmodified (build container only) and
(b) against the committed golden vectors those files produced (tests/golden/callers_*.npz, make_golden_callers.py)."""
import pytest
import torch

from helpers import golden, rel_err
from oracle import callers_oracle as co, ref_model_shim, renderer_oracle as ro

MIN_D, MAX_D = 0.9, 1.1
needs_ref = pytest.mark.skipif(not ref_model_shim.available(), reason="/root/reference not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("S,B,seed", [(16, 2, 0), (24, 3, 1)])
def test_oracle_matches_reference_code_bitwise(S, B, seed):
    model, losses = ref_model_shim.load()
    G, ns = ref_model_shim.model_self(MIN_D, MAX_D)
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(1, S, S, generator=g)
    assert torch.equal(co.get_clamped_depth(raw, S, S, MIN_D, MAX_D), G.get_clamped_depth(ns, raw, S, S))
    assert torch.equal(co.get_clamped_depth(raw, S, S, MIN_D, MAX_D, clamp_border=False),
                       G.get_clamped_depth(ns, raw, S, S, clamp_border=False))
    assert torch.equal(co.rescale_depth(raw.tanh(), MIN_D, MAX_D), G.rescale_depth(ns, raw.tanh()))
    normal = torch.nn.functional.normalize(torch.randn(1, S, S, 3, generator=g), dim=3)
    albedo = torch.tanh(torch.randn(1, 3, S, S, generator=g))
    light = torch.rand(B, 4, generator=g) * 2 - 1
    for x, y in zip(ro.get_lighting_directions(light), G.get_lighting_directions(ns, light)):
        assert torch.equal(x, y)
    a, b, d = ro.get_lighting_directions(light)
    for x, y in zip(ro.get_shading(normal, a, b, d, albedo), G.get_shading(ns, normal, a, b, d, albedo)):
        assert torch.equal(x, y)
    im1, im2 = torch.rand(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g)
    rd = 0.8 + 0.45 * torch.rand(B, S, S, generator=g)
    masks = (torch.rand(B, 1, S, S, generator=g) > 0.3).float()
    m_ref = (rd < MAX_D + (MAX_D - MIN_D) / 2).float().unsqueeze(1).detach() * masks      # model.py:265-269
    assert torch.equal(co.recon_im_mask(rd, MIN_D, MAX_D, masks), m_ref)
    assert torch.equal(co.photometric_loss(im1, im2, m_ref), losses.PhotometricLoss()(im1, im2, mask=m_ref))
    assert torch.equal(co.photometric_loss(im1, im2), losses.PhotometricLoss()(im1, im2))
    for sig in (0.05 + torch.rand(B, 1, S, S, generator=g), 0.05 + torch.rand(B, 3, S, S, generator=g)):
        assert torch.equal(co.photometric_loss(im1, im2, m_ref, sig),
                           losses.PhotometricLoss()(im1, im2, mask=m_ref, conf_sigma=sig))
        assert torch.equal(co.photometric_loss(im1, im2, None, sig), losses.PhotometricLoss()(im1, im2, conf_sigma=sig))
    for m in (raw, im1[:, :1], [raw, raw[:, ::2, ::2]]):
        assert torch.equal(co.smooth_loss(m), losses.SmoothLoss()(m))


@pytest.mark.parametrize("name", ["callers_s16_b3", "callers_s32_b2"])
def test_oracle_matches_golden(name):
    z = {k: torch.from_numpy(v) for k, v in golden(name).items()}
    S = z["depth_raw"].shape[-1]
    raw = z["depth_raw"].clone().requires_grad_(True)
    depth = co.get_clamped_depth(raw, S, S, MIN_D, MAX_D)
    (depth * z["cot_depth"]).sum().backward()
    assert torch.equal(depth.detach(), z["depth"]) and torch.equal(raw.grad, z["grad_depth_raw"])
    normal, albedo, light = (z[k].clone().requires_grad_(True) for k in ("normal", "albedo", "light"))
    a, b, d = ro.get_lighting_directions(light)
    diffuse, texture = ro.get_shading(normal, a, b, d, albedo)
    ((diffuse * z["cot_diffuse"]).sum() + (texture * z["cot_texture"]).sum()).backward()
    assert torch.equal(diffuse.detach(), z["diffuse"]) and torch.equal(texture.detach(), z["texture"])
    assert torch.equal(normal.grad, z["grad_normal"]) and torch.equal(albedo.grad, z["grad_albedo"])
    assert torch.equal(light.grad, z["grad_light"])
    im, tg = z["recon_im"].clone().requires_grad_(True), z["target"].clone().requires_grad_(True)
    loss = co.photometric_loss(im, tg, co.recon_im_mask(z["recon_depth"], MIN_D, MAX_D, z["masks"]))
    loss.backward()
    assert torch.equal(loss.detach(), z["photo_loss"]) and torch.equal(im.grad, z["grad_recon_im"])
    assert torch.equal(tg.grad, z["grad_target"])
    assert torch.equal(co.photometric_loss(z["recon_im"], z["target"]), z["photo_loss_nomask"])
    dm, sm = z["depth"].clone().requires_grad_(True), z["diffuse"].clone().requires_grad_(True)
    l_d, l_s = co.smooth_loss(dm), co.smooth_loss(sm)
    (l_d * 1.5 + l_s * 0.5).backward()
    assert torch.equal(l_d.detach(), z["smooth_depth"]) and torch.equal(l_s.detach(), z["smooth_shading"])
    # the branches of the smooth loss reach the map through several autograd paths whose accumulation order is not fixed
    assert rel_err(dm.grad, z["grad_smooth_depth"]) < 1e-6 and rel_err(sm.grad, z["grad_smooth_shading"]) < 1e-6


def test_known_answers():
    # an integer-valued ramp has zero second differences; identical images have zero photometric loss
    S = 12
    ramp = (torch.arange(S).float()[None, :, None] * 3 + torch.arange(S).float()[None, None, :] * 2)
    assert co.smooth_loss(ramp).item() == 0.0
    im = torch.rand(2, 3, S, S)
    assert co.photometric_loss(im, im, torch.ones(2, 1, S, S)).item() == 0.0
    # border rule: the 2 outer columns on each side are -0.02 * depth + 1.02 * border_depth (the literal pad value 1.02)
    raw = torch.zeros(1, S, S)
    d = co.get_clamped_depth(raw, S, S, MIN_D, MAX_D)
    mid, border = (MIN_D + MAX_D) / 2, 0.7 * MAX_D + 0.3 * MIN_D
    assert torch.allclose(d[0, :, 2:-2], torch.full((S, S - 4), mid))
    assert torch.allclose(d[0, :, :2], torch.full((S, 2), -0.02 * mid + 1.02 * border))
