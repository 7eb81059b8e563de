"""The callers either side of the renderer path, as torch.autograd.Functions over the C ABI.

Mirrors, with the reference's names and argument meaning:
  * `get_clamped_depth` / `rescale_depth`   GAN2Shape/model.py:337-345, 85-86   (depth prologue: centre, tanh, rescale, border)
  * `get_shading`                           GAN2Shape/model.py:355-360          (diffuse shading + shaded texture)
  * `recon_im_mask` + `PhotometricLoss`     GAN2Shape/model.py:146-150, 265-269; GAN2Shape/losses.py:39-51 (one fused pass)
  * `SmoothLoss`                            GAN2Shape/losses.py:54-79
The bodies are CUDA kernels (csrc/g2s_callers.cuh); there is no CPU / PyTorch fallback.
"""
import torch

from . import _lib
from .functional import _f32c, _p, _require_cuda, _stream


def _reduce_ws(device):
    return torch.empty(_lib.load().g2s_reduce_ws_bytes() // 8, dtype=torch.float64, device=device)


# ----------------------------------------------------------------------------------------------------------
class ClampedDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth_raw, n_groups, min_depth, max_depth, border_depth, clamp_border):
        _require_cuda(depth_raw)
        lib = _lib.load()
        raw = _f32c(depth_raw)
        N, H, W = raw.shape
        if N % n_groups:
            raise RuntimeError("get_clamped_depth: %d maps do not split into %d groups" % (N, n_groups))
        ws = _reduce_ws(raw.device)
        mean = torch.empty(n_groups, device=raw.device, dtype=torch.float32)
        depth = torch.empty_like(raw)
        _lib.check(lib.g2s_clamped_depth_fwd(_p(raw), n_groups, N // n_groups, H, W, min_depth, max_depth, border_depth,
                                             int(clamp_border), _p(ws), _p(mean), _p(depth), _stream()),
                   "g2s_clamped_depth_fwd")
        ctx.save_for_backward(raw, mean)
        ctx.meta = (n_groups, min_depth, max_depth, int(clamp_border))
        return depth

    @staticmethod
    def backward(ctx, g_depth):
        lib = _lib.load()
        raw, mean = ctx.saved_tensors
        n_groups, lo, hi, cb = ctx.meta
        N, H, W = raw.shape
        g = _f32c(g_depth)
        ws = _reduce_ws(raw.device)
        g_raw = torch.empty_like(raw)
        _lib.check(lib.g2s_clamped_depth_bwd(_p(raw), _p(mean), _p(g), n_groups, N // n_groups, H, W, lo, hi, cb, _p(ws),
                                             _p(g_raw), _stream()), "g2s_clamped_depth_bwd")
        return g_raw, None, None, None, None, None


def get_clamped_depth(depth_raw, h, w, min_depth, max_depth, border_depth=None, clamp_border=True, per_image=False):
    """model.py:337-345 (with rescale_depth, model.py:85-86, folded in).  depth_raw [N,h,w].

    The reference subtracts `depth_raw.view(1,-1).mean(1)`: ONE mean over the whole tensor (it only ever runs with one
    image); `per_image=True` centres every map on its own mean instead (the batched form).  `border_depth` defaults to the
    model's 0.7*max + 0.3*min (model.py:51)."""
    if depth_raw.dim() != 3 or depth_raw.shape[1] != h or depth_raw.shape[2] != w:
        raise RuntimeError("get_clamped_depth: depth_raw must be [N,%d,%d]" % (h, w))
    if border_depth is None:
        border_depth = 0.7 * max_depth + 0.3 * min_depth
    n_groups = depth_raw.shape[0] if per_image else 1
    return ClampedDepthFn.apply(depth_raw, n_groups, float(min_depth), float(max_depth), float(border_depth), clamp_border)


# ----------------------------------------------------------------------------------------------------------
def _view_strided(t, B, inner):
    """[1 or B, ...] tensor -> (contiguous storage, floats between views)."""
    t = _f32c(t)
    if t.shape[0] == B and B > 1:
        return t, inner
    if t.shape[0] == 1:
        return t, (0 if B > 1 else inner)
    raise RuntimeError("get_shading: batch %d does not broadcast to %d views" % (t.shape[0], B))


class ShadingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normal, light5, albedo):
        _require_cuda(normal, light5, albedo)
        lib = _lib.load()
        L = _f32c(light5)
        B = L.shape[0]
        H, W = normal.shape[1], normal.shape[2]
        HW = H * W
        n, ns = _view_strided(normal, B, 3 * HW)
        a, as_ = _view_strided(albedo, B, 3 * HW)
        diffuse = torch.empty(B, 1, H, W, device=L.device, dtype=torch.float32)
        texture = torch.empty(B, 3, H, W, device=L.device, dtype=torch.float32)
        _lib.check(lib.g2s_shading_fwd(_p(n), ns, _p(L), _p(a), as_, B, HW, _p(diffuse), _p(texture), _stream()),
                   "g2s_shading_fwd")
        ctx.save_for_backward(n, L, a)
        ctx.meta = (ns, as_, B, HW, normal.shape, albedo.shape)
        ctx.set_materialize_grads(False)     # an unused output arrives as None, not as a zero-filled tensor
        return diffuse, texture

    @staticmethod
    def backward(ctx, g_diffuse, g_texture):
        lib = _lib.load()
        n, L, a = ctx.saved_tensors
        ns, as_, B, HW, nshape, ashape = ctx.meta
        if g_diffuse is None and g_texture is None:
            return None, None, None
        gd = _f32c(g_diffuse) if g_diffuse is not None else None
        gt = _f32c(g_texture) if g_texture is not None else None
        g_n = torch.zeros(nshape, device=L.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        g_L = torch.zeros(B, 5, device=L.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        g_a = torch.zeros(ashape, device=L.device, dtype=torch.float32) if ctx.needs_input_grad[2] else None
        _lib.check(lib.g2s_shading_bwd(_p(n), ns, _p(L), _p(a), as_, B, HW, _p(gd), _p(gt), _p(g_n), ns, _p(g_L), _p(g_a),
                                       as_, _stream()), "g2s_shading_bwd")
        return g_n, g_L, g_a


def get_shading(normal, lighting_a, lighting_b, lighting_d, albedo):
    """model.py:355-360: normal [1|B,H,W,3], a/b [B,1], d [B,3], albedo [1|B,3,H,W] -> (diffuse [B,1,H,W], texture [B,3,H,W])."""
    light5 = torch.cat([lighting_a.reshape(-1, 1), lighting_b.reshape(-1, 1), lighting_d.reshape(-1, 3)], 1)
    return ShadingFn.apply(normal, light5, albedo)


# ----------------------------------------------------------------------------------------------------------
class PhotometricFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im1, im2, recon_depth, depth_thresh, mask, conf_sigma=None):
        _require_cuda(im1, im2, recon_depth, mask, conf_sigma)
        lib = _lib.load()
        x = _f32c(im1)
        B, C, H, W = x.shape
        HW = H * W
        if im2.shape[0] == 1 and B > 1:
            y, ys = _f32c(im2), 0
        else:
            y, ys = _f32c(im2.expand(B, C, H, W)), C * HW
        rd = _f32c(recon_depth.reshape(B, HW)) if recon_depth is not None else None
        mk = None
        if mask is not None:
            if mask.dim() == 4 and mask.shape[1] != 1:
                raise RuntimeError("PhotometricLoss: the mask must be [B,1,H,W] (it is expanded over the channels)")
            mk = _f32c(mask.expand(B, 1, H, W).reshape(B, HW))
        sg, sc = None, 1
        if conf_sigma is not None:      # losses.py:44-45; [B,1,H,W] (broadcast over the channels) or [B,C,H,W]
            if conf_sigma.dim() != 4 or conf_sigma.shape[1] not in (1, C):
                raise RuntimeError("PhotometricLoss: conf_sigma must be [B,1,H,W] or [B,C,H,W]")
            sc = conf_sigma.shape[1]
            sg = _f32c(conf_sigma.expand(B, sc, H, W))
        ws = _reduce_ws(x.device)
        out = torch.empty(3, device=x.device, dtype=torch.float32)
        _lib.check(lib.g2s_photometric_fwd(_p(x), _p(y), ys, _p(rd), depth_thresh, _p(mk), _p(sg), sc, B, C, HW, _p(ws),
                                           _p(out), _stream()), "g2s_photometric_fwd")
        ctx.save_for_backward(x, y, rd, mk, out, sg)
        ctx.meta = (ys, depth_thresh, im2.shape, sc, None if conf_sigma is None else conf_sigma.shape)
        return out[0]

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        x, y, rd, mk, out, sg = ctx.saved_tensors
        ys, thresh, yshape, sc, sshape = ctx.meta
        B, C, H, W = x.shape
        g = _f32c(g_loss).reshape(1)
        need2 = ctx.needs_input_grad[1]
        needs = sg is not None and len(ctx.needs_input_grad) > 5 and ctx.needs_input_grad[5]
        if need2 and ys == 0:
            raise RuntimeError("PhotometricLoss: no gradient to a broadcast (batch 1) second image; expand it first")
        g1 = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(x) if need2 else None
        gs = torch.empty_like(sg) if needs else None
        if g1 is None and g2 is None and gs is None:
            return None, None, None, None, None, None
        _lib.check(lib.g2s_photometric_bwd(_p(x), _p(y), ys, _p(rd), thresh, _p(mk), _p(sg), sc, B, C, H * W, _p(out), _p(g),
                                           _p(g1), _p(g2), _p(gs), _stream()), "g2s_photometric_bwd")
        if g2 is not None:
            g2 = g2.sum_to_size(yshape)
        if gs is not None:
            gs = gs.sum_to_size(sshape)
        return g1, g2, None, None, None, gs


class PhotometricLoss:
    """losses.py:39-51.  `mask` is the reference's [B,1,H,W] mask.  The validity mask the callers build first
    (model.py:146-150, 265-269: `recon_depth < max_depth + margin`, times the pseudo-view masks) can be left to the kernel:
    pass `recon_depth` and `depth_thresh` and only the extra masks (or None) as `mask`."""
    EPS = 1e-7

    def __call__(self, image1, image2, mask=None, conf_sigma=None, recon_depth=None, depth_thresh=None):
        if (recon_depth is None) != (depth_thresh is None):
            raise RuntimeError("PhotometricLoss: recon_depth and depth_thresh go together")
        return PhotometricFn.apply(image1, image2, recon_depth, float(depth_thresh) if depth_thresh is not None else 0.0,
                                   mask, conf_sigma)


def recon_im_mask(recon_depth, min_depth, max_depth):
    """model.py:146-150: threshold of the validity mask, `max_depth + (max_depth - min_depth) / 2`; returns
    (recon_depth, depth_thresh) keyword values for PhotometricLoss."""
    return dict(recon_depth=recon_depth, depth_thresh=max_depth + (max_depth - min_depth) / 2)


# ----------------------------------------------------------------------------------------------------------
class SmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred):
        _require_cuda(pred)
        lib = _lib.load()
        p = _f32c(pred)
        m = p.reshape(-1, p.shape[-2], p.shape[-1])       # losses.py:75-76
        M, H, W = m.shape
        ws = _reduce_ws(p.device)
        out = torch.empty(5, device=p.device, dtype=torch.float32)
        _lib.check(lib.g2s_smooth_fwd(_p(m), M, H, W, _p(ws), _p(out), _stream()), "g2s_smooth_fwd")
        ctx.save_for_backward(m)
        ctx.shape = pred.shape
        return out[0]

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        (m,) = ctx.saved_tensors
        M, H, W = m.shape
        g = _f32c(g_loss).reshape(1)
        g_map = torch.empty_like(m)
        _lib.check(lib.g2s_smooth_bwd(_p(m), M, H, W, _p(g), _p(g_map), _stream()), "g2s_smooth_bwd")
        return g_map.reshape(ctx.shape)


class SmoothLoss:
    """losses.py:54-79: second-order smoothness of one map or of a list of maps (weights 1, 1/2.3, 1/2.3^2, ...)."""

    def __call__(self, pred_map):
        if type(pred_map) not in [tuple, list]:
            pred_map = [pred_map]
        loss = 0
        weight = 1
        for scaled_map in pred_map:
            loss = loss + SmoothFn.apply(scaled_map) * weight
            weight /= 2.3
        return loss
