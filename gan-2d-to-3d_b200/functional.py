"""torch.autograd.Functions over the C ABI of csrc/libg2s_b200.so.

Each Function is the drop-in for one operator of the reference's Renderer path; the bodies are hand-written
sm_100a kernels (csrc/g2s_kernels.cu), reached through ctypes with raw device pointers and the current CUDA
stream.  torch is used for device memory, streams and autograd bookkeeping only.  There is no CPU or
PyTorch fallback: non-CUDA tensors raise.
"""
import ctypes
import os

import torch

from . import _lib


FWD_LANES = int(os.environ.get("G2S_FWD_LANES", "2"))


try:        # the raw handle of the current stream without building a torch.cuda.Stream object (19 us -> 0.3 us per call; the
    _raw_current_stream = torch._C._cuda_getCurrentRawStream      # single-image step is host-bound and asks six times)
    _current_device = torch._C._cuda_getDevice
except AttributeError:      # pragma: no cover
    def _raw_current_stream(index):
        return torch.cuda.current_stream(index).cuda_stream

    _current_device = torch.cuda.current_device


def _raw_stream(device=None):
    index = device.index if device is not None and device.index is not None else _current_device()
    return _raw_current_stream(index)


def _stream():
    return ctypes.c_void_p(_raw_stream())


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(*tensors):
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("gan-2d-to-3d_b200: this operator runs only on CUDA tensors (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # the C ABI launches on the current device (one process per GPU); fail loudly instead of touching the
            # wrong device's memory
            raise RuntimeError("gan-2d-to-3d_b200: tensor on cuda:%d but the current device is cuda:%d "
                               "(torch.cuda.set_device first)" % (t.device.index, cur))


def _f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _batched_image(t):
    """[B,...] tensor that may be an `expand`ed (batch-stride-0) view, as model.py:260-262 passes.
    Returns (storage tensor, batch stride in elements) without materialising the expansion."""
    if t.dtype != torch.float32:
        t = t.float()
    inner = t[0]
    if t.shape[0] > 1 and t.stride(0) == 0 and inner.is_contiguous():
        return inner, 0
    t = t.contiguous()
    return t, t[0].numel()


def _Rt(R, t, B):
    R = _f32c(R.expand(B, 3, 3))
    t = _f32c(t.reshape(-1, 3).expand(B, 3))
    return R, t


class ZBuffer:
    """Packed-key z-buffer workspaces owned by a Renderer: one per (far value, device, CUDA stream), grown on demand.
    Every word is the EMPTY key at rest (g2s_zbuffer_init) and every forward leaves it so; a buffer is therefore only
    filled when it is (re)allocated.  Separate streams get separate buffers (two streams rasterising into one buffer would
    race on its keys), a buffer that is replaced by a larger one is handed to the caching allocator with `record_stream`,
    and `invalidate()` -- called when a launch reports an error -- drops every buffer so that the next call re-initialises."""

    def __init__(self):
        self.buf = {}

    def get(self, n_views, S, far, device):
        key = (float(far), device, _raw_stream(device))
        lib = _lib.load()
        need = (lib.g2s_workspace_bytes(_lib.WS_ZBUFFER, n_views, S) + 7) // 8     # keys + work list + counters
        cur = self.buf.get(key)
        if cur is None or cur.numel() < need:
            if cur is not None:
                cur.record_stream(torch.cuda.current_stream(device))
            cur = torch.empty(need, dtype=torch.int64, device=device)
            # laid out per view by every call, so initialise it as `n_views` views of side S: uniformly EMPTY
            _lib.check(lib.g2s_zbuffer_init(_p(cur), n_views, S, far, _stream()), "g2s_zbuffer_init")
            self.buf[key] = cur
        return cur

    def invalidate(self):
        self.buf.clear()


class RasterScratch:
    """The raster backward's scratch when the forward handed over its projected vertices: its vertex-gradient part must be
    zero on entry and k_vertex_bwd leaves it zero, so it is cleared ONCE, when allocated, and kept by the Renderer (one per
    device and stream, grown on demand, dropped after a failed launch)."""

    def __init__(self):
        self.buf = {}

    def get(self, n_views, S, device):
        key = (device, _raw_stream(device))
        need = _lib.ws_floats(_lib.WS_RASTER_BWD, n_views, S)
        cur = self.buf.get(key)
        if cur is None or cur[1] != (n_views, S):
            if cur is not None:
                cur[0].record_stream(torch.cuda.current_stream(device))
            cur = (torch.zeros(need, device=device, dtype=torch.float32), (n_views, S))
            self.buf[key] = cur
        return cur[0]

    def invalidate(self):
        self.buf.clear()


def _checked(renderer, code, what):
    """_lib.check that also drops the renderer's persistent workspaces on failure (a failed launch may leave stale keys /
    gradients behind)"""
    if code != 0:
        renderer._zbuf.invalidate()
        renderer._raster_scratch.invalidate()
    _lib.check(code, what)


def _proj_buffer(renderer, n_views, S, device, needs_grad):
    """[n_views, S*S, 4] buffer in which the forward hands its projected vertices to the backward (which then skips
    k_project_verts): 16 S^2 bytes per view held until the backward; Renderer.share_projection = False turns it off."""
    if not needs_grad or not getattr(renderer, "share_projection", True):
        return None
    return torch.empty(n_views, S * S, 4, device=device, dtype=torch.float32)


def _ws(kind, n, S, device):
    """caller-owned workspace sized by g2s_workspace_bytes"""
    return torch.empty(_lib.ws_floats(kind, n, S), device=device, dtype=torch.float32)


# ----------------------------------------------------------------------------------------------------------
class WarpCanonDepthFn(torch.autograd.Function):
    """renderer.py:116-125 warp_canon_depth (+ nr.Renderer.render_depth).  Returns (recon_depth, face_idx)."""

    @staticmethod
    def forward(ctx, depth, R, t, renderer):
        _require_cuda(depth, R, t)
        lib = _lib.load()
        B, H, W = depth.shape
        S = renderer.image_size
        if H != S or W != S:
            raise RuntimeError("warp_canon_depth: depth must be [B,%d,%d] (image_size of the Renderer)" % (S, S))
        dstore, dstride = _batched_image(depth)
        Rc, tc = _Rt(R, t, B)
        cam = renderer._camera(depth_pass=True)
        zbuf = renderer._zbuf.get(B, S, cam.far_z, depth.device)
        recon = torch.empty(B, S, S, device=depth.device, dtype=torch.float32)
        fidx = torch.empty(B, 2 * S, 2 * S, device=depth.device, dtype=torch.int32)
        _checked(renderer, lib.g2s_warp_depth_fwd(ctypes.byref(cam), _p(dstore), dstride, _p(Rc), _p(tc), B, _p(zbuf),
                                                  _p(recon), _p(fidx), _stream()), "g2s_warp_depth_fwd")
        ctx.save_for_backward(dstore, Rc, tc, fidx, recon)
        ctx.dstride = dstride
        ctx.renderer = renderer
        ctx.shapes = (depth.shape, R.shape, t.shape)
        ctx.mark_non_differentiable(fidx)
        ctx.set_materialize_grads(False)     # no zero-filled int32 [B,2S,2S] "gradient" for the face-index map
        return recon, fidx

    @staticmethod
    def backward(ctx, g_recon, _g_fidx):
        if g_recon is None:
            return None, None, None, None
        lib = _lib.load()
        dstore, Rc, tc, fidx, recon = ctx.saved_tensors
        (B, S, _), Rshape, tshape = ctx.shapes
        cam = ctx.renderer._camera(depth_pass=True)
        g = _f32c(g_recon)
        ws = _ws(_lib.WS_RASTER_BWD, B, S, g.device)             # projected verts | vertex grads (uvz-) | g_sub
        g_depth = torch.zeros(B, S, S, device=g.device, dtype=torch.float32)
        need_view = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gR = torch.zeros(B, 3, 3, device=g.device, dtype=torch.float32) if need_view else None
        gt = torch.zeros(B, 3, device=g.device, dtype=torch.float32) if need_view else None
        _lib.check(lib.g2s_warp_depth_bwd(ctypes.byref(cam), _p(dstore), ctx.dstride, _p(Rc), _p(tc), B, _p(fidx),
                                          _p(recon), _p(g), _p(ws), _p(g_depth), S * S, _p(gR), _p(gt), _stream()),
                   "g2s_warp_depth_bwd")
        if need_view:
            gR = gR.sum_to_size(Rshape)
            gt = gt.reshape(B, *([1] * (len(tshape) - 2)), 3).sum_to_size(tshape)
        return g_depth, gR, gt, None


class WarpGridFn(torch.autograd.Function):
    """renderer.py:104-114 get_warped_2d_grid (inverse=False) / get_inv_warped_2d_grid (inverse=True)."""

    @staticmethod
    def forward(ctx, depth, R, t, renderer, inverse):
        _require_cuda(depth, R, t)
        lib = _lib.load()
        B, H, W = depth.shape
        dstore, dstride = _batched_image(depth)
        Rc, tc = _Rt(R, t, B)
        cam = renderer._camera()
        grid = torch.empty(B, H, W, 2, device=depth.device, dtype=torch.float32)
        _lib.check(lib.g2s_warp_grid_fwd(ctypes.byref(cam), _p(dstore), dstride, _p(Rc), _p(tc), B, H, W,
                                         int(inverse), _p(grid), _stream()), "g2s_warp_grid_fwd")
        ctx.save_for_backward(dstore, Rc, tc)
        ctx.meta = (dstride, renderer, int(inverse), depth.shape, R.shape, t.shape)
        return grid

    @staticmethod
    def backward(ctx, g_grid):
        lib = _lib.load()
        dstore, Rc, tc = ctx.saved_tensors
        dstride, renderer, inverse, (B, H, W), Rshape, tshape = ctx.meta
        cam = renderer._camera()
        g = _f32c(g_grid)
        g_depth = torch.empty(B, H, W, device=g.device, dtype=torch.float32)
        need_view = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gR = torch.zeros(B, 3, 3, device=g.device, dtype=torch.float32) if need_view else None
        gt = torch.zeros(B, 3, device=g.device, dtype=torch.float32) if need_view else None
        _lib.check(lib.g2s_warp_grid_bwd(ctypes.byref(cam), _p(dstore), dstride, _p(Rc), _p(tc), B, H, W, inverse,
                                         _p(g), _p(g_depth), _p(gR), _p(gt), _stream()), "g2s_warp_grid_bwd")
        if need_view:
            gR = gR.sum_to_size(Rshape)
            gt = gt.reshape(B, *([1] * (len(tshape) - 2)), 3).sum_to_size(tshape)
        return g_depth, gR, gt, None, None


class NormalFromDepthFn(torch.autograd.Function):
    """renderer.py:127-139 get_normal_from_depth."""

    @staticmethod
    def forward(ctx, depth, renderer):
        _require_cuda(depth)
        lib = _lib.load()
        d = _f32c(depth)
        B, H, W = d.shape
        cam = renderer._camera()
        normal = torch.empty(B, H, W, 3, device=d.device, dtype=torch.float32)
        _lib.check(lib.g2s_normal_fwd(ctypes.byref(cam), _p(d), B, H, W, _p(normal), _stream()), "g2s_normal_fwd")
        ctx.save_for_backward(d)
        ctx.renderer = renderer
        return normal

    @staticmethod
    def backward(ctx, g_normal):
        lib = _lib.load()
        (d,) = ctx.saved_tensors
        B, H, W = d.shape
        cam = ctx.renderer._camera()
        g = _f32c(g_normal)
        g_depth = torch.empty(B, H, W, device=d.device, dtype=torch.float32)
        _lib.check(lib.g2s_normal_bwd(ctypes.byref(cam), _p(d), B, H, W, _p(g), _p(g_depth), 0, _stream()),
                   "g2s_normal_bwd")
        return g_depth, None


class ViewToRtFn(torch.autograd.Function):
    """utils.py:52-73 get_transform_matrices (+ :33-49 get_rotation_matrix): view [B,3|5|6] -> (R [B,3,3], t [B,1,3])."""

    @staticmethod
    def forward(ctx, view):
        _require_cuda(view)
        lib = _lib.load()
        v = _f32c(view)
        B, w = v.shape
        if w not in (3, 5, 6):
            raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")   # utils.py:70-71
        R = torch.empty(B, 3, 3, device=v.device, dtype=torch.float32)
        t = torch.empty(B, 1, 3, device=v.device, dtype=torch.float32)
        _lib.check(lib.g2s_view_fwd(_p(v), w, B, _p(R), _p(t), _stream()), "g2s_view_fwd")
        ctx.save_for_backward(v)
        return R, t

    @staticmethod
    def backward(ctx, gR, gt):
        lib = _lib.load()
        (v,) = ctx.saved_tensors
        B, w = v.shape
        gR = _f32c(gR) if gR is not None else None
        gt = _f32c(gt) if gt is not None else None
        gv = torch.empty_like(v)
        _lib.check(lib.g2s_view_bwd(_p(v), w, B, _p(gR), _p(gt), _p(gv), _stream()), "g2s_view_bwd")
        return gv


class LightFn(torch.autograd.Function):
    """model.py:347-353 get_lighting_directions, packed: raw light [B,4] -> [B,5] = (a, b, dx, dy, dz)."""

    @staticmethod
    def forward(ctx, light):
        _require_cuda(light)
        lib = _lib.load()
        l = _f32c(light)
        B = l.shape[0]
        if l.shape[1] != 4:
            raise RuntimeError("light must be [B,4]")
        out = torch.empty(B, 5, device=l.device, dtype=torch.float32)
        _lib.check(lib.g2s_light_fwd(_p(l), B, _p(out), _stream()), "g2s_light_fwd")
        ctx.save_for_backward(l)
        return out

    @staticmethod
    def backward(ctx, g5):
        lib = _lib.load()
        (l,) = ctx.saved_tensors
        g = _f32c(g5)
        gl = torch.empty_like(l)
        _lib.check(lib.g2s_light_bwd(_p(l), l.shape[0], _p(g), _p(gl), _stream()), "g2s_light_bwd")
        return gl


_MODES = {"bilinear": 0, "nearest": 1}


class GridSampleFn(torch.autograd.Function):
    """F.grid_sample(input, grid, mode, padding_mode='zeros', align_corners) as the reference calls it
    (model.py:151, 270; renderer.py:179, 223, 241, 261, 263)."""

    @staticmethod
    def forward(ctx, inp, grid, mode, align_corners):
        _require_cuda(inp, grid)
        if mode not in _MODES:
            raise RuntimeError("grid_sample: mode must be 'bilinear' or 'nearest'")
        lib = _lib.load()
        B, C, H, W = inp.shape
        gB, Ho, Wo, two = grid.shape
        if gB != B or two != 2:
            raise RuntimeError("grid_sample: grid must be [B,Ho,Wo,2] with the batch size of input")
        istore, istride = _batched_image(inp)
        g = _f32c(grid)
        out = torch.empty(B, C, Ho, Wo, device=inp.device, dtype=torch.float32)
        _lib.check(lib.g2s_sample_fwd(_p(istore), istride, _p(g), B, C, H, W, Ho, Wo, _MODES[mode],
                                      int(bool(align_corners)), _p(out), _stream()), "g2s_sample_fwd")
        ctx.save_for_backward(istore, g)
        ctx.meta = (istride, (B, C, H, W), (Ho, Wo), _MODES[mode], int(bool(align_corners)))
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        istore, g = ctx.saved_tensors
        istride, (B, C, H, W), (Ho, Wo), mode, align = ctx.meta
        go = _f32c(g_out)
        g_in = torch.zeros(B, C, H, W, device=go.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        g_grid = torch.empty(B, Ho, Wo, 2, device=go.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _lib.check(lib.g2s_sample_bwd(_p(istore), istride, _p(g), _p(go), B, C, H, W, Ho, Wo, mode, align, _p(g_in),
                                      C * H * W, _p(g_grid), _stream()), "g2s_sample_bwd")
        return g_in, g_grid, None, None


def grid_sample(inp, grid, mode="bilinear", align_corners=False):
    return GridSampleFn.apply(inp, grid, mode, align_corners)


class RenderChainFn(torch.autograd.Function):
    """The fused projected-view render (model.py:243-270): normal -> shading -> warp_canon_depth ->
    get_inv_warped_2d_grid -> grid_sample(...).clamp(-1,1), for n_images images x views_per_image views.
    Inputs: depth [N,S,S], albedo [N,3,S,S], R [B,3,3], t [B,3], light [B,5] = (a, b, dx, dy, dz).
    Returns (recon_im [B,3,S,S], recon_depth [B,S,S], face_idx int32 [B,2S,2S])."""

    @staticmethod
    def forward(ctx, depth, albedo, R, t, light, renderer, views_per_image, align_corners, mask=None,
                want_mask=False, _also_save=()):
        _require_cuda(depth, albedo, R, t, light, mask)
        lib = _lib.load()
        N, S, _ = depth.shape
        B = N * views_per_image
        if S != renderer.image_size or depth.shape[2] != S or tuple(albedo.shape) != (N, 3, S, S):
            raise RuntimeError("render_chain: depth must be [N,S,S] and albedo [N,3,S,S] with S = image_size")
        if R.shape[0] != B or light.shape != (B, 5):
            raise RuntimeError("render_chain: R/t/light must have n_images * views_per_image rows")
        d, a = _f32c(depth), _f32c(albedo)
        Rc, tc = _Rt(R, t, B)
        L = _f32c(light)
        cam = renderer._camera(depth_pass=True)
        dev = d.device
        # z-buffer for FWD_LANES chunks: the forward rotates its chunks over that many lanes (streams)
        ws_views = min(B, FWD_LANES * lib.g2s_chunk_views(S))
        zbuf = renderer._zbuf.get(ws_views, S, cam.far_z, dev)
        normal = _ws(_lib.WS_TEXELS, N, S, dev)               # packed texels [N,S,S,8]: normal xyz, albedo rgb, pad
        recon_im = torch.empty(B, 3, S, S, device=dev, dtype=torch.float32)
        recon_depth = torch.empty(B, S, S, device=dev, dtype=torch.float32)
        fidx = torch.empty(B, 2 * S, 2 * S, device=dev, dtype=torch.int32)
        mask_in = _f32c(mask.reshape(N, S, S)) if mask is not None else None
        mask_out = torch.empty(B, 1, S, S, device=dev, dtype=torch.float32) if want_mask else None
        proj = _proj_buffer(renderer, B, S, dev, any(ctx.needs_input_grad[:5]))
        _checked(renderer, lib.g2s_render_fused_fwd(renderer._context(dev).handle, ctypes.byref(cam), _p(d), _p(a), _p(Rc),
                                                    _p(tc), _p(L), N, views_per_image, int(bool(align_corners)), _p(zbuf),
                                                    ws_views, _p(normal), _p(recon_im), _p(recon_depth), _p(fidx),
                                                    _p(mask_in), _p(mask_out), _p(proj), _stream()),
                 "g2s_render_fused_fwd")
        ctx.save_for_backward(d, a, Rc, tc, L, normal, recon_depth, fidx, proj, *_also_save)
        ctx.meta = (renderer, views_per_image, int(bool(align_corners)), R.shape, t.shape)
        ctx.mark_non_differentiable(fidx)
        # unused outputs reach backward as None instead of zero-filled tensors (autograd would otherwise fill a
        # [B,S,S] float and a [B,2S,2S] int32 tensor per step just to say "no gradient")
        ctx.set_materialize_grads(False)
        if want_mask:
            ctx.mark_non_differentiable(mask_out)
            return recon_im, recon_depth, fidx, mask_out
        return recon_im, recon_depth, fidx

    @staticmethod
    def backward(ctx, g_im, g_depth_out, _g_fidx, _g_mask=None):
        lib = _lib.load()
        d, a, Rc, tc, L, normal, recon_depth, fidx, proj = ctx.saved_tensors[:9]
        renderer, vpi, align, Rshape, tshape = ctx.meta
        N, S, _ = d.shape
        B = N * vpi
        dev = d.device
        cam = renderer._camera(depth_pass=True)
        if g_im is None and g_depth_out is None:
            return (None,) * 10
        gi = _f32c(g_im) if g_im is not None else torch.zeros(B, 3, S, S, device=dev)
        gd_out = _f32c(g_depth_out) if g_depth_out is not None else None
        ws_views = min(B, lib.g2s_chunk_views_bwd(S))
        # projected verts | vertex grads (uvz-) | g_sub; with the forward's projection: vertex grads (zero at rest) | - | g_sub
        ws_sub = renderer._raster_scratch.get(ws_views, S, dev) if proj is not None else _ws(_lib.WS_RASTER_BWD, ws_views, S, dev)
        ws_tex = _ws(_lib.WS_TEX_BWD, ws_views, S, dev)
        ws_nrm = _ws(_lib.WS_GRAD_NORMAL, N, S, dev)
        g_depth = torch.empty(N, S, S, device=dev, dtype=torch.float32)
        g_albedo = torch.empty(N, 3, S, S, device=dev, dtype=torch.float32)
        gR = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
        gt = torch.empty(B, 3, device=dev, dtype=torch.float32)
        gL = torch.empty(B, 5, device=dev, dtype=torch.float32)
        _checked(renderer, lib.g2s_render_fused_bwd(renderer._context(dev).handle, ctypes.byref(cam), _p(d), _p(a), _p(Rc), _p(tc), _p(L), N, vpi, align,
                                            _p(normal), _p(recon_depth), _p(fidx), _p(gi), _p(gd_out), _p(proj), ws_views, _p(ws_sub),
                                            _p(ws_tex), _p(ws_nrm), _p(g_depth), _p(g_albedo), _p(gR), _p(gt), _p(gL),
                                            _stream()), "g2s_render_fused_bwd")
        gR = gR.sum_to_size(Rshape)
        gt = gt.reshape(B, *([1] * (len(tshape) - 2)), 3).sum_to_size(tshape)
        return g_depth, g_albedo, gR, gt, gL, None, None, None, None, None


class RenderChainViewFn(torch.autograd.Function):
    """RenderChainFn from the RAW view [B,3|5|6] (utils.py:52-73) and light [B,4] (model.py:347-353) as ONE autograd node:
    what Renderer.render_chain runs.  The small batches of the reference (one image x 16 views per step) are host-bound, and
    three nodes (view -> R, t; light -> directions; the chain) cost a quarter of the step more than one.  Leaves the
    renderer's rot_mat / trans_xyz set as set_transform_matrices would (without autograd history: the gradient to `view`
    flows through this node).  Returns (recon_im, recon_depth, face_idx)."""

    @staticmethod
    def forward(ctx, depth, albedo, view, light, renderer, views_per_image, align_corners):
        _require_cuda(depth, albedo, view, light)
        lib = _lib.load()
        v, l = _f32c(view), _f32c(light)
        B, w = v.shape
        if w not in (3, 5, 6):
            raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")   # utils.py:70-71
        if tuple(l.shape) != (B, 4):
            raise RuntimeError("light must be [B,4]")
        R = torch.empty(B, 3, 3, device=v.device, dtype=torch.float32)
        t = torch.empty(B, 1, 3, device=v.device, dtype=torch.float32)
        L = torch.empty(B, 5, device=v.device, dtype=torch.float32)
        _lib.check(lib.g2s_view_light_fwd(_p(v), w, _p(l), B, _p(R), _p(t), _p(L), _stream()), "g2s_view_light_fwd")
        renderer.rot_mat, renderer.trans_xyz = R, t
        # needs_input_grad[:5] of this node = depth, albedo, view, light, renderer: the same "any gradient wanted" test
        return RenderChainFn.forward(ctx, depth, albedo, R, t, L, renderer, views_per_image, align_corners, None, False,
                                     (v, l))

    @staticmethod
    def backward(ctx, g_im, g_depth_out, _g_fidx):
        if g_im is None and g_depth_out is None:
            return (None,) * 7
        lib = _lib.load()
        g_depth, g_albedo, gR, gt, gL = RenderChainFn.backward(ctx, g_im, g_depth_out, None)[:5]
        v, l = ctx.saved_tensors[9:11]
        B, w = v.shape
        gv, gl = torch.empty_like(v), torch.empty_like(l)
        _lib.check(lib.g2s_view_light_bwd(_p(v), w, _p(l), B, _p(_f32c(gR)), _p(_f32c(gt)), _p(gL), _p(gv), _p(gl), _stream()),
                   "g2s_view_light_bwd")
        return g_depth, g_albedo, gv, gl, None, None, None


class RenderChainLossFn(torch.autograd.Function):
    """RenderChainFn + the step-3 photometric loss of the reference taken INSIDE the render (model.py:243-274):
    loss = PhotometricLoss(recon_im, target, mask=(recon_depth < depth_thresh) * view_mask) (losses.py:39-51).  The
    forward's resolve kernel accumulates the masked sums, the backward's pixel stage forms the loss's cotangent, so the loss
    costs no pass of its own over recon_im / target / recon_depth and no cotangent image is written or read.
    Returns (recon_im, recon_depth, face_idx, loss); cotangents on recon_im (e.g. from a perceptual loss) and recon_depth
    are added to the loss's."""

    @staticmethod
    def forward(ctx, depth, albedo, R, t, light, target, view_mask, renderer, views_per_image, align_corners, depth_thresh,
                _also_save=()):
        _require_cuda(depth, albedo, R, t, light, target, view_mask)
        lib = _lib.load()
        N, S, _ = depth.shape
        B = N * views_per_image
        if S != renderer.image_size or depth.shape[2] != S or tuple(albedo.shape) != (N, 3, S, S):
            raise RuntimeError("render_chain_loss: depth must be [N,S,S] and albedo [N,3,S,S] with S = image_size")
        if R.shape[0] != B or light.shape != (B, 5):
            raise RuntimeError("render_chain_loss: R/t/light must have n_images * views_per_image rows")
        if tuple(target.shape) != (B, 3, S, S):
            raise RuntimeError("render_chain_loss: target must be [n_views,3,S,S]")
        if view_mask is not None and view_mask.numel() != B * S * S:
            raise RuntimeError("render_chain_loss: view_mask must be [n_views,1,S,S]")
        d, a = _f32c(depth), _f32c(albedo)
        Rc, tc = _Rt(R, t, B)
        L = _f32c(light)
        tg = _f32c(target.detach())
        vm = _f32c(view_mask.detach().reshape(B, S, S)) if view_mask is not None else None
        cam = renderer._camera(depth_pass=True)
        dev = d.device
        ws_views = min(B, FWD_LANES * lib.g2s_chunk_views(S))
        zbuf = renderer._zbuf.get(ws_views, S, cam.far_z, dev)
        normal = _ws(_lib.WS_TEXELS, N, S, dev)
        loss_ws = _ws(_lib.WS_LOSS, B, S, dev)
        recon_im = torch.empty(B, 3, S, S, device=dev, dtype=torch.float32)
        recon_depth = torch.empty(B, S, S, device=dev, dtype=torch.float32)
        fidx = torch.empty(B, 2 * S, 2 * S, device=dev, dtype=torch.int32)
        out3 = torch.empty(3, device=dev, dtype=torch.float32)
        la = _lib.PhotoLoss(_p(tg), _p(vm), float(depth_thresh))
        proj = _proj_buffer(renderer, B, S, dev, any(ctx.needs_input_grad[:5]))
        _checked(renderer, lib.g2s_render_fused_loss_fwd(renderer._context(dev).handle, ctypes.byref(cam), _p(d), _p(a),
                                                         _p(Rc), _p(tc), _p(L), N, views_per_image,
                                                         int(bool(align_corners)), _p(zbuf), ws_views, _p(normal),
                                                         _p(recon_im), _p(recon_depth), _p(fidx), ctypes.byref(la),
                                                         _p(loss_ws), _p(out3), _p(proj), _stream()),
                 "g2s_render_fused_loss_fwd")
        ctx.save_for_backward(d, a, Rc, tc, L, normal, recon_depth, fidx, tg, vm, out3, proj, *_also_save)
        ctx.meta = (renderer, views_per_image, int(bool(align_corners)), R.shape, t.shape, float(depth_thresh))
        ctx.mark_non_differentiable(fidx)
        ctx.set_materialize_grads(False)
        return recon_im, recon_depth, fidx, out3[0]

    @staticmethod
    def backward(ctx, g_im, g_depth_out, _g_fidx, g_loss):
        lib = _lib.load()
        d, a, Rc, tc, L, normal, recon_depth, fidx, tg, vm, out3, proj = ctx.saved_tensors[:12]
        renderer, vpi, align, Rshape, tshape, thresh = ctx.meta
        N, S, _ = d.shape
        B = N * vpi
        dev = d.device
        cam = renderer._camera(depth_pass=True)
        if g_im is None and g_depth_out is None and g_loss is None:
            return (None,) * 11
        gi = _f32c(g_im) if g_im is not None else None
        gd_out = _f32c(g_depth_out) if g_depth_out is not None else None
        gl = _f32c(g_loss).reshape(1) if g_loss is not None else torch.zeros(1, device=dev)
        ws_views = min(B, lib.g2s_chunk_views_bwd(S))
        ws_sub = renderer._raster_scratch.get(ws_views, S, dev) if proj is not None else _ws(_lib.WS_RASTER_BWD, ws_views, S, dev)
        ws_tex = _ws(_lib.WS_TEX_BWD, ws_views, S, dev)
        ws_nrm = _ws(_lib.WS_GRAD_NORMAL, N, S, dev)
        g_depth = torch.empty(N, S, S, device=dev, dtype=torch.float32)
        g_albedo = torch.empty(N, 3, S, S, device=dev, dtype=torch.float32)
        gR = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
        gt = torch.empty(B, 3, device=dev, dtype=torch.float32)
        gL = torch.empty(B, 5, device=dev, dtype=torch.float32)
        la = _lib.PhotoLoss(_p(tg), _p(vm), thresh)
        _checked(renderer, lib.g2s_render_fused_loss_bwd(renderer._context(dev).handle, ctypes.byref(cam), _p(d), _p(a), _p(Rc),
                                                 _p(tc), _p(L), N, vpi, align, _p(normal), _p(recon_depth), _p(fidx),
                                                 _p(gi), _p(gd_out), ctypes.byref(la), _p(out3), _p(gl), _p(proj), ws_views,
                                                 _p(ws_sub), _p(ws_tex), _p(ws_nrm), _p(g_depth), _p(g_albedo), _p(gR),
                                                 _p(gt), _p(gL), _stream()), "g2s_render_fused_loss_bwd")
        gR = gR.sum_to_size(Rshape)
        gt = gt.reshape(B, *([1] * (len(tshape) - 2)), 3).sum_to_size(tshape)
        return g_depth, g_albedo, gR, gt, gL, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------------
class RenderRgbFn(torch.autograd.Function):
    """nr.Renderer.render_rgb(vertices, get_face_idx, get_textures_from_im(im, 2)) (+ clamp) as renderer.py:194-196,
    230, 248, 272, 275 call it.  Differentiable with respect to `im` (the per-vertex colours: neural_renderer's
    backward_textures through get_textures_from_im) and with respect to the vertices (neural_renderer's approximate
    backward_pixel_map edge gradient chained through the projection; 3-channel images)."""

    @staticmethod
    def forward(ctx, vertices3d, im, renderer, clamp):
        _require_cuda(vertices3d, im)
        lib = _lib.load()
        B = vertices3d.shape[0]
        Bi, C, H, W = im.shape
        S = renderer.image_size
        if H != S or W != S or Bi not in (1, B):
            raise RuntimeError("render_rgb: image must be [B or 1,C,%d,%d]" % (S, S))
        if C > 4:
            raise RuntimeError("render_rgb: at most 4 channels")
        if Bi == 1 and B > 1:                       # one image for all views (a sweep of one input)
            istore, istride = _f32c(im[0]), 0
        else:
            istore, istride = _batched_image(im)
        verts = _f32c(vertices3d)
        cam = renderer._camera(rgb_pass=True)
        zbuf = renderer._zbuf.get(B, S, cam.far_z, im.device)
        out = torch.empty(B, C, S, S, device=im.device, dtype=torch.float32)
        fidx = torch.empty(B, 2 * S, 2 * S, device=im.device, dtype=torch.int32)
        bg = (ctypes.c_float * 4)(*([renderer.background_color[i % 3] for i in range(4)]))
        _checked(renderer, lib.g2s_render_rgb_fwd(ctypes.byref(cam), _p(verts), _p(istore), istride, B, C,
                                                  renderer.tex_cube_size, bg, int(clamp), _p(zbuf), _p(out), _p(fidx),
                                                  _stream()), "g2s_render_rgb_fwd")
        ctx.save_for_backward(verts, istore, fidx)
        ctx.meta = (renderer, int(clamp), istride, im.shape)
        ctx.mark_non_differentiable(fidx)
        ctx.set_materialize_grads(False)
        return out, fidx

    @staticmethod
    def backward(ctx, g_out, _g_fidx):
        need_v, need_im = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_out is None or not (need_v or need_im):
            return None, None, None, None
        lib = _lib.load()
        verts, istore, fidx = ctx.saved_tensors
        renderer, clamp, istride, imshape = ctx.meta
        C, S = imshape[1], imshape[2]
        g = _f32c(g_out)
        B = g.shape[0]
        cam = renderer._camera(rgb_pass=True)
        if need_v and C != 3:
            raise RuntimeError("render_rgb: the geometry gradient (backward_pixel_map) needs a 3-channel image")
        g_im = torch.zeros(istore.shape, device=g.device, dtype=torch.float32) if need_im else None    # [C,S,S] when shared
        g_v = torch.empty_like(verts) if need_v else None
        rgb_ws = _ws(_lib.WS_RGB_MAP, B, S, g.device) if need_v else None
        ras_ws = _ws(_lib.WS_RASTER_BWD, B, S, g.device) if need_v else None
        bg = (ctypes.c_float * 4)(*([renderer.background_color[i % 3] for i in range(4)]))
        _lib.check(lib.g2s_render_rgb_bwd(ctypes.byref(cam), _p(verts), _p(istore), istride, B, C, renderer.tex_cube_size,
                                          bg, clamp, _p(fidx), _p(g), _p(g_im), istride, _p(rgb_ws), _p(ras_ws), _p(g_v),
                                          _stream()), "g2s_render_rgb_bwd")
        if need_im and istride == 0:
            if imshape[0] != 1:
                # an `expand`ed [B,C,S,S] view of one image: autograd will sum B identical slices back onto it
                g_im = (g_im / imshape[0]).unsqueeze(0).expand(imshape)
            else:
                g_im = g_im.unsqueeze(0)
        return g_v, g_im, None, None


class Grid3dFn(torch.autograd.Function):
    """renderer.py:74-80 depth_to_3d_grid (mode 0), :90-95 get_warped_3d_grid (1), :97-102 get_inv_warped_3d_grid (2):
    depth [B,H,W] (+ R [B,3,3], t [B,1,3]) -> [B,H,W,3]."""

    @staticmethod
    def forward(ctx, depth, R, t, renderer, mode):
        _require_cuda(depth, R, t)
        lib = _lib.load()
        B, H, W = depth.shape
        dstore, dstride = _batched_image(depth)
        Rc = tc = None
        if mode != 0:
            Rc, tc = _Rt(R, t, B)
        cam = renderer._camera()
        out = torch.empty(B, H, W, 3, device=depth.device, dtype=torch.float32)
        R0, t0, R2, t2 = (Rc, tc, None, None) if mode == 2 else (None, None, Rc, tc)
        _lib.check(lib.g2s_grid3d_fwd(ctypes.byref(cam), _p(dstore), dstride, B, H, W, None, _p(R0), _p(t0), None, _p(R2),
                                      _p(t2), _p(out), _stream()), "g2s_grid3d_fwd")
        ctx.save_for_backward(dstore, Rc, tc)
        ctx.meta = (dstride, renderer, mode, depth.shape, None if R is None else R.shape, None if t is None else t.shape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        dstore, Rc, tc = ctx.saved_tensors
        dstride, renderer, mode, (B, H, W), Rshape, tshape = ctx.meta
        cam = renderer._camera()
        g = _f32c(g_out)
        g_depth = torch.empty(B, H, W, device=g.device, dtype=torch.float32)
        need_view = mode != 0 and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        gR = torch.zeros(B, 3, 3, device=g.device, dtype=torch.float32) if need_view else None
        gt = torch.zeros(B, 3, device=g.device, dtype=torch.float32) if need_view else None
        _lib.check(lib.g2s_grid3d_bwd(ctypes.byref(cam), _p(dstore), dstride, B, H, W, mode, _p(Rc), _p(tc), _p(g),
                                      _p(g_depth), _p(gR), _p(gt), _stream()), "g2s_grid3d_bwd")
        if need_view:
            gR = gR.sum_to_size(Rshape)
            gt = gt.reshape(B, *([1] * (len(tshape) - 2)), 3).sum_to_size(tshape)
        return g_depth, gR, gt, None, None


class Grid3dTo2dFn(torch.autograd.Function):
    """renderer.py:82-88 grid_3d_to_2d: [B,H,W,3] -> [B,H,W,2] in [-1,1]."""

    @staticmethod
    def forward(ctx, grid_3d, renderer):
        _require_cuda(grid_3d)
        lib = _lib.load()
        g3 = _f32c(grid_3d)
        B, H, W, _ = g3.shape
        cam = renderer._camera()
        out = torch.empty(B, H, W, 2, device=g3.device, dtype=torch.float32)
        _lib.check(lib.g2s_grid_3d_to_2d_fwd(ctypes.byref(cam), _p(g3), B, H, W, _p(out), _stream()), "g2s_grid_3d_to_2d_fwd")
        ctx.save_for_backward(g3)
        ctx.renderer = renderer
        return out

    @staticmethod
    def backward(ctx, g_grid):
        lib = _lib.load()
        (g3,) = ctx.saved_tensors
        B, H, W, _ = g3.shape
        cam = ctx.renderer._camera()
        out = torch.empty_like(g3)
        _lib.check(lib.g2s_grid_3d_to_2d_bwd(ctypes.byref(cam), _p(g3), B, H, W, _p(_f32c(g_grid)), _p(out), _stream()),
                   "g2s_grid_3d_to_2d_bwd")
        return out, None


class RenderChainLossViewFn(torch.autograd.Function):
    """RenderChainLossFn from the raw view / light as ONE autograd node (see RenderChainViewFn): what
    Renderer.render_chain_loss runs.  Returns (recon_im, recon_depth, face_idx, loss)."""

    @staticmethod
    def forward(ctx, depth, albedo, view, light, target, view_mask, renderer, views_per_image, align_corners, depth_thresh):
        _require_cuda(depth, albedo, view, light, target, view_mask)
        lib = _lib.load()
        v, l = _f32c(view), _f32c(light)
        B, w = v.shape
        if w not in (3, 5, 6):
            raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")   # utils.py:70-71
        if tuple(l.shape) != (B, 4):
            raise RuntimeError("light must be [B,4]")
        R = torch.empty(B, 3, 3, device=v.device, dtype=torch.float32)
        t = torch.empty(B, 1, 3, device=v.device, dtype=torch.float32)
        L = torch.empty(B, 5, device=v.device, dtype=torch.float32)
        _lib.check(lib.g2s_view_light_fwd(_p(v), w, _p(l), B, _p(R), _p(t), _p(L), _stream()), "g2s_view_light_fwd")
        renderer.rot_mat, renderer.trans_xyz = R, t
        return RenderChainLossFn.forward(ctx, depth, albedo, R, t, L, target, view_mask, renderer, views_per_image,
                                         align_corners, depth_thresh, (v, l))

    @staticmethod
    def backward(ctx, g_im, g_depth_out, _g_fidx, g_loss):
        if g_im is None and g_depth_out is None and g_loss is None:
            return (None,) * 10
        lib = _lib.load()
        g_depth, g_albedo, gR, gt, gL = RenderChainLossFn.backward(ctx, g_im, g_depth_out, None, g_loss)[:5]
        v, l = ctx.saved_tensors[12:14]
        B, w = v.shape
        gv, gl = torch.empty_like(v), torch.empty_like(l)
        _lib.check(lib.g2s_view_light_bwd(_p(v), w, _p(l), B, _p(_f32c(gR)), _p(_f32c(gt)), _p(gL), _p(gv), _p(gl), _stream()),
                   "g2s_view_light_bwd")
        return g_depth, g_albedo, gv, gl, None, None, None, None, None, None
