"""Parity of the CUDA callers either side of the path (csrc/g2s_callers.cuh through the C ABI)
against the oracle (oracle/callers_oracle.py, pinned to the reference's own model.py / losses.py) and against the golden
vectors the reference's code produced.  Floating point: 1e-5 relative (max|a-b| / max|b|), the north star's bar; the
losses are sums of up to 10^8 terms whose order differs from torch-CPU's, so scalars are held to 1e-5 relative as well."""
import numpy as np
import pytest
import torch

from helpers import golden, rel_err
from oracle import callers_oracle as co, renderer_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-5
MIN_D, MAX_D = 0.9, 1.1
THRESH = MAX_D + (MAX_D - MIN_D) / 2


def _g():
    import g2s_b200
    return g2s_b200


@pytest.mark.parametrize("name", ["callers_s16_b3", "callers_s32_b2"])
def test_callers_against_golden(name):
    g2s = _g()
    z = {k: torch.from_numpy(v).cuda() for k, v in golden(name).items()}
    S = z["depth_raw"].shape[-1]
    raw = z["depth_raw"].clone().requires_grad_(True)
    depth = g2s.get_clamped_depth(raw, S, S, MIN_D, MAX_D)
    (depth * z["cot_depth"]).sum().backward()
    assert rel_err(depth.detach().cpu(), z["depth"].cpu()) < TOL
    assert rel_err(raw.grad.cpu(), z["grad_depth_raw"].cpu()) < TOL
    normal, albedo, light = (z[k].clone().requires_grad_(True) for k in ("normal", "albedo", "light"))
    a, b, d = g2s.get_lighting_directions(light)
    diffuse, texture = g2s.get_shading(normal, a, b, d, albedo)
    ((diffuse * z["cot_diffuse"]).sum() + (texture * z["cot_texture"]).sum()).backward()
    for got, want in ((diffuse, "diffuse"), (texture, "texture"), (normal.grad, "grad_normal"), (albedo.grad, "grad_albedo"),
                      (light.grad, "grad_light")):
        assert rel_err(got.detach().cpu(), z[want].cpu()) < TOL, want
    im, tg = z["recon_im"].clone().requires_grad_(True), z["target"].clone().requires_grad_(True)
    loss = g2s.PhotometricLoss()(im, tg, mask=z["masks"], **g2s.recon_im_mask(z["recon_depth"], MIN_D, MAX_D))
    loss.backward()
    assert rel_err(loss.detach().cpu(), z["photo_loss"].cpu()) < TOL
    assert rel_err(im.grad.cpu(), z["grad_recon_im"].cpu()) < TOL and rel_err(tg.grad.cpu(), z["grad_target"].cpu()) < TOL
    assert rel_err(g2s.PhotometricLoss()(z["recon_im"], z["target"]).cpu(), z["photo_loss_nomask"].cpu()) < TOL
    dm, sm = z["depth"].clone().requires_grad_(True), z["diffuse"].clone().requires_grad_(True)
    l_d, l_s = g2s.SmoothLoss()(dm), g2s.SmoothLoss()(sm)
    (l_d * 1.5 + l_s * 0.5).backward()
    assert rel_err(l_d.detach().cpu(), z["smooth_depth"].cpu()) < TOL
    assert rel_err(l_s.detach().cpu(), z["smooth_shading"].cpu()) < TOL
    assert rel_err(dm.grad.cpu(), z["grad_smooth_depth"].cpu()) < TOL
    assert rel_err(sm.grad.cpu(), z["grad_smooth_shading"].cpu()) < TOL


@pytest.mark.parametrize("N,S,per_image,clamp_border", [(1, 128, False, True), (3, 33, False, True), (5, 64, True, True),
                                                         (2, 17, True, False)])
def test_clamped_depth_vs_oracle(N, S, per_image, clamp_border):
    g2s = _g()
    gen = torch.Generator().manual_seed(N * 100 + S)
    raw = torch.randn(N, S, S, generator=gen)
    cot = torch.randn(N, S, S, generator=gen)
    r_o = raw.clone().requires_grad_(True)
    d_o = co.get_clamped_depth(r_o, S, S, MIN_D, MAX_D, clamp_border=clamp_border, per_image=per_image)
    (d_o * cot).sum().backward()
    r = raw.cuda().requires_grad_(True)
    d = g2s.get_clamped_depth(r, S, S, MIN_D, MAX_D, clamp_border=clamp_border, per_image=per_image)
    (d * cot.cuda()).sum().backward()
    assert rel_err(d.detach().cpu(), d_o.detach()) < TOL and rel_err(r.grad.cpu(), r_o.grad) < TOL


@pytest.mark.parametrize("B,S,shared", [(4, 32, True), (3, 21, False), (1, 64, True)])
def test_shading_vs_oracle(B, S, shared):
    g2s = _g()
    gen = torch.Generator().manual_seed(B * 10 + S)
    nb = 1 if shared else B
    normal = torch.nn.functional.normalize(torch.randn(nb, S, S, 3, generator=gen) + torch.tensor([0., 0., 1.]), dim=3)
    albedo = torch.tanh(torch.randn(nb, 3, S, S, generator=gen))
    light = torch.rand(B, 4, generator=gen) * 2 - 1
    cd, ct = torch.randn(B, 1, S, S, generator=gen), torch.randn(B, 3, S, S, generator=gen)
    outs = []
    for dev in ("cpu", "cuda"):
        n, a, l = (t.to(dev).clone().requires_grad_(True) for t in (normal, albedo, light))
        if dev == "cpu":
            dif, tex = ro.get_shading(n, *ro.get_lighting_directions(l), a)
        else:
            dif, tex = g2s.get_shading(n, *g2s.get_lighting_directions(l), a)
        ((dif * cd.to(dev)).sum() + (tex * ct.to(dev)).sum()).backward()
        outs.append([t.detach().cpu() for t in (dif, tex, n.grad, a.grad, l.grad)])
    for want, got, what in zip(outs[0], outs[1], ("diffuse", "texture", "grad_normal", "grad_albedo", "grad_light")):
        assert rel_err(got, want) < TOL, what
    # diffuse-only cotangent (SmoothLoss(diffuse_shading), model.py:162)
    n = normal.cuda().requires_grad_(True)
    dif, _ = g2s.get_shading(n, *g2s.get_lighting_directions(light.cuda()), albedo.cuda())
    (dif * cd.cuda()).sum().backward()
    n_o = normal.clone().requires_grad_(True)
    dif_o, _ = ro.get_shading(n_o, *ro.get_lighting_directions(light), albedo)
    (dif_o * cd).sum().backward()
    assert rel_err(n.grad.cpu(), n_o.grad) < TOL


@pytest.mark.parametrize("B,C,S,use_depth,use_mask,bcast", [(4, 3, 32, True, True, False), (3, 3, 21, True, False, False),
                                                            (2, 1, 17, False, True, False), (5, 3, 64, True, True, True),
                                                            (2, 3, 16, False, False, False)])
def test_photometric_vs_oracle(B, C, S, use_depth, use_mask, bcast):
    g2s = _g()
    gen = torch.Generator().manual_seed(B + C + S)
    im1 = torch.rand(B, C, S, S, generator=gen) * 2 - 1
    im2 = torch.rand(1 if bcast else B, C, S, S, generator=gen) * 2 - 1
    im2[..., :2, :] = im1[:1 if bcast else B, :, :2, :]           # exact zeros: sign(0) = 0 in the gradient
    rd = 0.8 + 0.4 * torch.rand(B, S, S, generator=gen)
    rd[rd > 1.12] = THRESH                                         # background sits exactly on the clamp: not < thresh
    masks = (torch.rand(B, 1, S, S, generator=gen) > 0.25).float()
    x_o, y_o = im1.clone().requires_grad_(True), im2.clone().requires_grad_(not bcast)
    m_o = None
    if use_depth:
        m_o = co.recon_im_mask(rd, MIN_D, MAX_D, masks if use_mask else None)
    elif use_mask:
        m_o = masks
    l_o = co.photometric_loss(x_o, y_o.expand(B, C, S, S), m_o)
    (l_o * 0.7).backward()
    x, y = im1.cuda().requires_grad_(True), im2.cuda().requires_grad_(not bcast)
    kw = g2s.recon_im_mask(rd.cuda(), MIN_D, MAX_D) if use_depth else {}
    loss = g2s.PhotometricLoss()(x, y, mask=masks.cuda() if use_mask else None, **kw)
    (loss * 0.7).backward()
    assert rel_err(loss.detach().cpu(), l_o.detach()) < TOL
    assert rel_err(x.grad.cpu(), x_o.grad) < TOL
    if not bcast:
        assert rel_err(y.grad.cpu(), y_o.grad) < TOL


@pytest.mark.parametrize("B,C,S,sigma_c,use_mask", [(4, 3, 32, 1, True), (3, 3, 21, 3, True), (2, 1, 17, 1, False),
                                                    (2, 3, 64, 1, False)])
def test_photometric_conf_sigma_vs_oracle(B, C, S, sigma_c, use_mask):
    """losses.py:44-45: the confidence-weighted form, value and the gradients to both images and to sigma."""
    g2s = _g()
    gen = torch.Generator().manual_seed(11 * B + C + S)
    im1 = torch.rand(B, C, S, S, generator=gen) * 2 - 1
    im2 = torch.rand(B, C, S, S, generator=gen) * 2 - 1
    im2[..., :2, :] = im1[..., :2, :]
    sig = 0.05 + torch.rand(B, sigma_c, S, S, generator=gen)
    rd = 0.8 + 0.4 * torch.rand(B, S, S, generator=gen)
    masks = (torch.rand(B, 1, S, S, generator=gen) > 0.25).float()
    x_o, y_o, s_o = (t.clone().requires_grad_(True) for t in (im1, im2, sig))
    m_o = co.recon_im_mask(rd, MIN_D, MAX_D, masks) if use_mask else None
    l_o = co.photometric_loss(x_o, y_o, m_o, s_o)
    (l_o * 1.3).backward()
    x, y, sg = (t.cuda().requires_grad_(True) for t in (im1, im2, sig))
    kw = dict(mask=masks.cuda(), **g2s.recon_im_mask(rd.cuda(), MIN_D, MAX_D)) if use_mask else {}
    loss = g2s.PhotometricLoss()(x, y, conf_sigma=sg, **kw)
    (loss * 1.3).backward()
    assert rel_err(loss.detach().cpu(), l_o.detach()) < TOL
    assert rel_err(x.grad.cpu(), x_o.grad) < TOL
    assert rel_err(y.grad.cpu(), y_o.grad) < TOL
    assert rel_err(sg.grad.cpu(), s_o.grad) < TOL
    # sigma = 2**0.5 - EPS (to rounding) and the log term vanishing: equal to the plain loss when log(sigma) is subtracted
    one = torch.full((B, 1, S, S), 2 ** 0.5, device="cuda")
    plain = g2s.PhotometricLoss()(x.detach(), y.detach(), **kw)
    withs = g2s.PhotometricLoss()(x.detach(), y.detach(), conf_sigma=one, **kw)
    assert abs(withs.item() - plain.item() - 0.5 * np.log(2.0)) < 1e-5


@pytest.mark.parametrize("shape", [(1, 128, 128), (3, 1, 33, 20), (2, 3, 3), (4, 1, 64, 64)])
def test_smooth_vs_oracle(shape):
    g2s = _g()
    gen = torch.Generator().manual_seed(sum(shape))
    m = torch.randn(*shape, generator=gen)
    m[..., :2, :] = 1.0                                            # flat patch: exact-zero second differences
    m_o = m.clone().requires_grad_(True)
    l_o = co.smooth_loss(m_o)
    (l_o * 1.3).backward()
    mc = m.cuda().requires_grad_(True)
    loss = g2s.SmoothLoss()(mc)
    (loss * 1.3).backward()
    assert rel_err(loss.detach().cpu(), l_o.detach()) < TOL and rel_err(mc.grad.cpu(), m_o.grad) < TOL
    # list form: weights 1, 1/2.3 (losses.py:61-71)
    if shape[-1] < 6:
        return
    pair = [m.cuda(), m.cuda()[..., ::2, ::2].contiguous()]
    assert rel_err(g2s.SmoothLoss()(pair).cpu(), co.smooth_loss([m, m[..., ::2, ::2]])) < TOL


def test_full_size_properties():
    """BASELINE bulk sizes (256 images x 16 views at 128^2 = 4096 views): size-independent properties."""
    g2s = _g()
    B, S = 4096, 128
    gen = torch.Generator(device="cuda").manual_seed(5)
    im = torch.rand(B, 3, S, S, device="cuda", generator=gen)
    rd = torch.full((B, S, S), 1.0, device="cuda")
    rd[:, :, S // 2:] = THRESH                                     # right half invalid
    pl = g2s.PhotometricLoss()
    assert pl(im, im, **g2s.recon_im_mask(rd, MIN_D, MAX_D)).item() == 0.0
    # |im - (im + c)| = c on the valid half, whatever the invalid half holds
    other = im + 0.25
    other[..., S // 2:] = 7.0
    assert abs(pl(im, other, **g2s.recon_im_mask(rd, MIN_D, MAX_D)).item() - 0.25) < 1e-6
    # linearity in the mask: all-ones mask == plain mean
    ones = torch.ones(B, 1, S, S, device="cuda")
    assert rel_err(pl(im, other, mask=ones).cpu(), pl(im, other).cpu()) < 1e-6
    # smooth loss: integer ramps are exactly flat to second order; scaling the map scales the loss
    ramp = (torch.arange(S, device="cuda").float()[None, :, None] * 3 + torch.arange(S, device="cuda").float()[None, None, :] * 2)
    assert g2s.SmoothLoss()(ramp.expand(64, S, S).contiguous()).item() == 0.0
    m = im[:64, 0]
    assert rel_err(g2s.SmoothLoss()(m * 4).cpu(), g2s.SmoothLoss()(m).cpu() * 4) < 1e-6


def test_errors():
    g2s = _g()
    with pytest.raises(RuntimeError):
        g2s.get_clamped_depth(torch.zeros(1, 8, 8), 8, 8, MIN_D, MAX_D)            # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        g2s.get_clamped_depth(torch.zeros(1, 8, 9, device="cuda"), 8, 8, MIN_D, MAX_D)
    with pytest.raises(RuntimeError):
        g2s.PhotometricLoss()(torch.zeros(1, 3, 4, 4, device="cuda"), torch.zeros(1, 3, 4, 4, device="cuda"),
                              conf_sigma=torch.ones(1, 2, 4, 4, device="cuda"))                  # 1 or C channels
    with pytest.raises(RuntimeError):
        g2s.SmoothLoss()(torch.zeros(1, 2, 2, device="cuda"))                        # H, W >= 3
