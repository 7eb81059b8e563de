"""The LOWER boundary of the path: g2s_b200.nr_compat.Renderer (drop-in for the part of `neural_renderer` that
GAN2Shape/renderer/renderer.py:6, 47-54, 120, 196... uses) against the oracle's restatement of neural_renderer
(oracle/nr_port.py) on the same vertices / faces / textures: depth bit-exact, rgb and gradients to 1e-5."""
import pytest
import torch

from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, close_except_few, oracle_renderer, rel_err
from oracle import nr_port, renderer_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(S, seed, rot=60.0, P=2):
    from g2s_b200 import synthetic
    case = synthetic.make_case(S, P, seed=seed, rot_deg=rot)
    orc = oracle_renderer(S)
    orc.set_transform_matrices(case["view"])
    verts = orc.get_warped_3d_grid(case["depth"].expand(P, S, S)).reshape(P, -1, 3).detach()   # renderer.py:117-118
    faces = ro.get_face_idx(P, S, S)
    kw = dict(camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0., K=orc.K,
              R=torch.eye(3)[None], t=torch.zeros(1, 3), near=0.1, far=10.0, image_size=S, orig_size=S, fill_back=True,
              background_color=[1, 1, 1])
    return case, verts, faces, kw


@pytest.mark.parametrize("S,seed", [(32, 71), (64, 72)])
def test_render_depth_forward_backward(S, seed):
    import g2s_b200
    case, verts, faces, kw = _setup(S, seed)
    nr_o, nr_c = nr_port.Renderer(**kw), g2s_b200.nr_compat.Renderer(**kw)
    v_o = verts.clone().requires_grad_(True)
    d_o = nr_o.render_depth(v_o, faces)
    cot = torch.randn(d_o.shape, generator=torch.Generator().manual_seed(seed)) * (d_o < 50).float()
    (d_o * cot).sum().backward()
    v_c = verts.cuda().requires_grad_(True)
    d_c = nr_c.render_depth(v_c, faces.cuda())
    (d_c * cot.cuda()).sum().backward()
    assert torch.equal(d_c.detach().cpu(), d_o.detach())                  # same vertex bits in, same depth bits out
    assert rel_err(v_c.grad.cpu(), v_o.grad) < TOL


@pytest.mark.parametrize("S,seed", [(32, 73)])
def test_render_rgb_forward_and_texture_gradient(S, seed):
    import g2s_b200
    case, verts, faces, kw = _setup(S, seed)
    P = verts.shape[0]
    nr_o, nr_c = nr_port.Renderer(**kw), g2s_b200.nr_compat.Renderer(**kw)
    im_o = case["albedo"].expand(P, 3, S, S).clone().requires_grad_(True)
    rgb_o = nr_o.render_rgb(verts, faces, ro.get_textures_from_im(im_o, tx_size=2))
    cot = torch.randn(rgb_o.shape, generator=torch.Generator().manual_seed(seed))
    (rgb_o * cot).sum().backward()
    im_c = case["albedo"].expand(P, 3, S, S).clone().cuda().requires_grad_(True)
    rgb_c = nr_c.render_rgb(verts.cuda(), faces.cuda(), g2s_b200.get_textures_from_im(im_c, tx_size=2))
    (rgb_c * cot.cuda()).sum().backward()
    assert rel_err(rgb_c.detach().cpu(), rgb_o.detach()) < TOL
    assert rel_err(im_c.grad.cpu(), im_o.grad) < TOL


def test_unsupported_configurations_raise():
    import g2s_b200
    case, verts, faces, kw = _setup(16, 74)
    for bad in (dict(camera_mode='look_at'), dict(light_intensity_directional=0.5), dict(fill_back=False),
                dict(orig_size=32), dict(R=torch.eye(3)[None] * 2)):
        with pytest.raises(NotImplementedError):
            g2s_b200.nr_compat.Renderer(**{**kw, **bad})
    r = g2s_b200.nr_compat.Renderer(**kw)
    with pytest.raises(NotImplementedError):
        r.render_depth(verts.cuda(), faces.cuda().flip(1))                 # not the grid topology
    with pytest.raises(RuntimeError):
        r.render_depth(verts, faces)                                        # CPU tensors: no fallback
