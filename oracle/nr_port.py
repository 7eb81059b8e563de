"""oracle/nr_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the Python layer of the third-party `neural_renderer` package (daniilidis-group
fork, unpinned HEAD in the reference: /root/reference/README.md:32-37), limited to what the reference
calls: `nr.Renderer(camera_mode='projection', ...)` (GAN2Shape/renderer/renderer.py:47-54),
`.render_depth(vertices, faces)` (renderer.py:120) and `.render_rgb(vertices, faces, textures)`
(renderer.py:196, 230, 248, 272, 275).  neural_renderer is NOT in /root/reference; this follows its
published algorithm.  PARITY UNPINNED at this boundary (no reference test or
golden vector exists); see oracle/nr_raster.c for what pins it instead.

The CUDA kernels of neural_renderer are restated in oracle/nr_raster.c and reached through ctypes.
All tensors are torch CPU fp32; every 3-wide matmul goes through Mm3 (oracle/fma_mm.c), the explicit
fma chain that torch-CPU matmul evaluates in the build container, so values do not depend on the host
BLAS of the machine the oracle runs on.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import build as _build

DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100.0
DEFAULT_EPS = 1e-4

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        ci, cf = ctypes.c_int, ctypes.c_float
        L.nr_forward_face_index_map.argtypes = [f32p, ci, ci, ci, cf, cf, i32p, f32p, f32p, f32p, i64p]
        L.nr_forward_face_index_map.restype = None
        L.nr_forward_face_index_map_culled.argtypes = [f32p, ci, ci, ci, cf, cf, i32p, f32p, f32p, f32p]
        L.nr_forward_face_index_map_culled.restype = None
        L.nr_backward_depth_map.argtypes = [f32p, f32p, i32p, f32p, f32p, f32p, f32p, ci, ci, ci]
        L.nr_backward_depth_map.restype = None
        L.nr_forward_texture_sampling.argtypes = [f32p, f32p, i32p, f32p, f32p, f32p, i32p, f32p, f32p,
                                                  f32p, ci, ci, ci, ci, cf]
        L.nr_forward_texture_sampling.restype = None
        L.nr_backward_textures.argtypes = [i32p, f32p, i32p, f32p, f32p, ci, ci, ci, ci]
        L.nr_backward_textures.restype = None
        L.nr_backward_pixel_map.argtypes = [f32p, i32p, f32p, f32p, f32p, f32p, f32p, ci, ci, ci, cf, ci, ci]
        L.nr_backward_pixel_map.restype = None
        cl = ctypes.c_long
        L.fma_mm3_nt.argtypes = [f32p, f32p, f32p, cl, cl, ci]
        L.fma_mm3_nt.restype = None
        L.fma_mm_k3.argtypes = [f32p, f32p, f32p, cl, ci, ci, ci, ci]
        L.fma_mm_k3.restype = None
        _lib = L
    return _lib


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int32))


class Mm3(torch.autograd.Function):
    """v[B,N,3] @ M[b,3,3]^T as the explicit chain fma(v2,m2, fma(v1,m1, v0*m0)) (oracle/fma_mm.c) --
    bit-for-bit what torch.matmul does for the reference on the build container's CPU.
    Backward uses plain matmuls (gradients are compared to tolerance, not bit-for-bit)."""

    @staticmethod
    def forward(ctx, v, M):
        vc = v.detach().contiguous().float()
        Mc = M.detach().contiguous().float()
        B, N = vc.shape[0], vc.shape[1]
        out = torch.empty_like(vc)
        lib().fma_mm3_nt(_fp(vc), _fp(Mc), _fp(out), B, N, 1 if Mc.shape[0] == B and B > 1 else 0)
        ctx.save_for_backward(vc, Mc)
        return out

    @staticmethod
    def backward(ctx, g):
        v, M = ctx.saved_tensors
        gv = g.matmul(M)
        gM = g.transpose(1, 2).matmul(v)
        if M.shape[0] == 1:
            gM = gM.sum(0, keepdim=True)
        return gv, gM


def mm3(v, M):
    """v[B,...,3] @ M[1|B,3,3]^T with the pinned fma chain."""
    shp = v.shape
    assert M.shape[0] in (1, shp[0])
    return Mm3.apply(v.reshape(shp[0], -1, 3), M).reshape(shp)


class _MmK3(torch.autograd.Function):
    """A[na,R,3] @ B[nb,3,C] (na, nb equal or 1) with the pinned fma chain; the backward uses plain matmuls (gradients are
    compared to tolerance, not bit-for-bit)."""

    @staticmethod
    def forward(ctx, A, Bm):
        R, C = A.shape[-2], Bm.shape[-1]
        na, nb = A.shape[0], Bm.shape[0]
        n = max(na, nb)
        out = torch.empty(n, R, C)
        lib().fma_mm_k3(_fp(A), _fp(Bm), _fp(out), n, R, C, 1 if na > 1 else 0, 1 if nb > 1 else 0)
        ctx.save_for_backward(A, Bm)
        return out

    @staticmethod
    def backward(ctx, g):
        A, Bm = ctx.saved_tensors
        gA = g.matmul(Bm.transpose(1, 2))
        gB = A.transpose(1, 2).matmul(g)
        if A.shape[0] == 1 and gA.shape[0] != 1:
            gA = gA.sum(0, keepdim=True)
        if Bm.shape[0] == 1 and gB.shape[0] != 1:
            gB = gB.sum(0, keepdim=True)
        return gA, gB


def mm_k3(A, Bm):
    """A[...,R,3] @ B[...,3,C] with the pinned fma chain; batch dims equal or 1."""
    A3 = A.contiguous().float().reshape(-1, A.shape[-2], 3)
    B3 = Bm.contiguous().float().reshape(-1, 3, Bm.shape[-1])
    return _MmK3.apply(A3, B3)


# Rasteriser mode for the whole oracle: "brute" = the faithful O(is^2 * nf) loop of the reference,
# "culled" = bounding-box culled, bit-identical outputs (checked in tests/test_oracle_raster.py).
MODE = {"raster": "culled"}
# Side channel: the maps of the most recent rasterisation (the nr API hides the face-index map).
LAST = {}


def forward_face_index_map(faces, image_size, near, far, mode=None, want_stats=False):
    """[nr] rasterize_cuda.forward_face_index_map. faces [B,NF,3,3] -> dict of maps (nr-native rows)."""
    mode = mode or MODE["raster"]
    faces = faces.detach().contiguous().float()
    B, NF = faces.shape[:2]
    s = image_size
    fim = torch.empty(B, s, s, dtype=torch.int32)
    wm = torch.empty(B, s, s, 3)
    dm = torch.empty(B, s, s)
    fvm = torch.empty(B, s, s, 3, 3)
    stats = np.zeros(2, dtype=np.int64)
    if mode == "brute":
        lib().nr_forward_face_index_map(_fp(faces), B, NF, s, near, far, _ip(fim), _fp(wm), _fp(dm), _fp(fvm),
                                        stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
                                        if want_stats else None)
    else:
        lib().nr_forward_face_index_map_culled(_fp(faces), B, NF, s, near, far, _ip(fim), _fp(wm), _fp(dm),
                                               _fp(fvm))
    out = dict(face_index_map=fim, weight_map=wm, depth_map=dm, face_inv_map=fvm)
    if want_stats:
        out["stats"] = dict(ties=int(stats[0]), outside_bbox=int(stats[1]))
    return out


class Rasterize(torch.autograd.Function):
    """[nr] rasterize.RasterizeFunction (forward maps, approximate-gradient backward).

    backward_pixel_map (the silhouette / colour edge gradient, only run when rgb or alpha is requested) is restated in
    oracle/nr_raster.c nr_backward_pixel_map; the reference itself never differentiates render_rgb (SURVEY.md 8a')."""

    @staticmethod
    def forward(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb,
                return_alpha, return_depth):
        B, NF = faces.shape[:2]
        s = image_size
        maps = forward_face_index_map(faces, s, near, far)
        ctx.maps = maps
        ctx.dims = (B, NF, s)
        ctx.eps = eps
        ctx.flags = (return_rgb, return_alpha, return_depth)
        ctx.has_tex = textures is not None
        facesc = faces.detach().contiguous().float()
        rgb = torch.zeros(B, s, s, 3)
        alpha = torch.zeros(B, s, s)
        if return_rgb:
            tex = textures.detach().contiguous().float()
            ts = tex.shape[2]
            sidx = torch.zeros(B, s, s, 8, dtype=torch.int32)
            swt = torch.zeros(B, s, s, 8)
            bg = torch.tensor(list(background_color), dtype=torch.float32)
            lib().nr_forward_texture_sampling(_fp(facesc), _fp(tex), _ip(maps["face_index_map"]),
                                              _fp(maps["weight_map"]), _fp(maps["depth_map"]), _fp(rgb),
                                              _ip(sidx), _fp(swt), _fp(alpha), _fp(bg), B, NF, s, ts, eps)
            ctx.sampling = (sidx, swt, ts)
            ctx.rgb_map = rgb.clone()
        elif return_alpha:
            alpha = (maps["face_index_map"] >= 0).float()
        ctx.save_for_backward(facesc)
        LAST.clear()
        LAST.update(maps)
        LAST["rgb_map"] = rgb
        return rgb.clone(), alpha.clone(), maps["depth_map"].clone()

    @staticmethod
    def backward(ctx, grad_rgb, grad_alpha, grad_depth):
        (faces,) = ctx.saved_tensors
        B, NF, s = ctx.dims
        return_rgb, return_alpha, return_depth = ctx.flags
        maps = ctx.maps
        grad_faces = torch.zeros(B, NF, 3, 3)
        grad_textures = None
        if (return_rgb or return_alpha) and ctx.needs_input_grad[0]:
            g_rgb = grad_rgb.contiguous().float() if (return_rgb and grad_rgb is not None) else torch.zeros(B, s, s, 3)
            g_alpha = grad_alpha.contiguous().float() if (return_alpha and grad_alpha is not None) else torch.zeros(B, s, s)
            rgb_map = ctx.rgb_map if return_rgb else torch.zeros(B, s, s, 3)
            alpha_map = (maps["face_index_map"] >= 0).float()
            lib().nr_backward_pixel_map(_fp(faces), _ip(maps["face_index_map"]), _fp(rgb_map), _fp(alpha_map), _fp(g_rgb),
                                        _fp(g_alpha), _fp(grad_faces), B, NF, s, ctx.eps, int(return_rgb), int(return_alpha))
        if return_rgb and ctx.has_tex and ctx.needs_input_grad[1]:
            sidx, swt, ts = ctx.sampling
            grad_textures = torch.zeros(B, NF, ts, ts, ts, 3)
            g = grad_rgb.contiguous().float()
            lib().nr_backward_textures(_ip(maps["face_index_map"]), _fp(swt), _ip(sidx), _fp(g),
                                       _fp(grad_textures), B, NF, s, ts)
        if return_depth and grad_depth is not None and ctx.needs_input_grad[0]:
            g = grad_depth.contiguous().float()
            lib().nr_backward_depth_map(_fp(faces), _fp(maps["depth_map"]), _ip(maps["face_index_map"]),
                                        _fp(maps["face_inv_map"]), _fp(maps["weight_map"]), _fp(g),
                                        _fp(grad_faces), B, NF, s)
        return grad_faces, grad_textures, None, None, None, None, None, None, None, None


def _flip_rows(x, dim):
    idx = torch.arange(x.shape[dim] - 1, -1, -1)
    return x.index_select(dim, idx)


def rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color,
                    return_rgb, return_alpha, return_depth):
    """[nr] rasterize.rasterize_rgbad: 2x supersampling, vertical flip, 2x2 average pooling."""
    s = image_size * 2 if anti_aliasing else image_size
    rgb, alpha, depth = Rasterize.apply(faces, textures, s, near, far, eps, background_color, return_rgb,
                                        return_alpha, return_depth)
    out = {}
    if return_rgb:
        rgb = _flip_rows(rgb.permute(0, 3, 1, 2), 2)
        out["rgb"] = F.avg_pool2d(rgb, kernel_size=(2, 2)) if anti_aliasing else rgb
    if return_alpha:
        alpha = _flip_rows(alpha, 1)
        out["alpha"] = F.avg_pool2d(alpha[:, None], kernel_size=(2, 2))[:, 0] if anti_aliasing else alpha
    if return_depth:
        depth = _flip_rows(depth, 1)
        out["depth"] = F.avg_pool2d(depth[:, None], kernel_size=(2, 2))[:, 0] if anti_aliasing else depth
    return out


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """[nr] projection.projection, K/R/t batched [1|B,3,3] / [1|B,1,3]; the two matmuls use the pinned
    fma chain (Mm3)."""
    mm = mm3
    vertices = mm(vertices, R) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1, k2, p1, p2, k3 = [dist_coeffs[:, None, i] for i in range(5)]
    r = torch.sqrt(x_ ** 2 + y_ ** 2)
    x__ = x_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + 2 * p1 * x_ * y_ + p2 * (r ** 2 + 2 * x_ ** 2)
    y__ = y_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + p1 * (r ** 2 + 2 * y_ ** 2) + 2 * p2 * x_ * y_
    vertices = mm(torch.stack([x__, y__, torch.ones_like(z)], dim=-1), K)
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.) / orig_size
    v = 2 * (v - orig_size / 2.) / orig_size
    return torch.stack([u, v, z], dim=-1)


def vertices_to_faces(vertices, faces):
    """[nr] vertices_to_faces: gather [B,V,3] by int faces [B,NF,3] -> [B,NF,3,3]."""
    bs, nv = vertices.shape[:2]
    faces = faces + (torch.arange(bs, dtype=torch.int32) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces.long()]


class Renderer:
    """[nr] renderer.Renderer restricted to camera_mode='projection' (what renderer.py:47-54 builds)."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=[0, 0, 0], fill_back=True,
                 camera_mode='projection', K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 perspective=True, viewing_angle=30, camera_direction=[0, 0, 1], near=0.1, far=100,
                 light_intensity_ambient=0.5, light_intensity_directional=0.5,
                 light_color_ambient=[1, 1, 1], light_color_directional=[1, 1, 1],
                 light_direction=[0, 1, 0]):
        if camera_mode != 'projection':
            raise ValueError("oracle nr port: only camera_mode='projection' is restated")
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.K, self.R, self.t = K, R, t
        if isinstance(self.t, torch.Tensor) and self.t.dim() == 2:
            self.t = self.t[:, None, :]
        self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros(1, 5)
        self.orig_size = orig_size
        self.near, self.far = near, far
        self.light_intensity_ambient = light_intensity_ambient
        self.light_intensity_directional = light_intensity_directional
        self.light_color_ambient = light_color_ambient
        self.rasterizer_eps = 1e-3

    def _fill_back_faces(self, faces):
        return torch.cat((faces, faces[:, :, [2, 1, 0]]), dim=1).detach()

    def render_depth(self, vertices, faces):
        if self.fill_back:
            faces = self._fill_back_faces(faces)
        vertices = projection(vertices, self.K, self.R, self.t, self.dist_coeffs, self.orig_size)
        faces = vertices_to_faces(vertices, faces)
        # rasterize_depth uses the MODULE defaults near=0.1, far=100, eps=1e-4 (A.2 step 4)
        return rasterize_rgbad(faces, None, self.image_size, self.anti_aliasing, DEFAULT_NEAR, DEFAULT_FAR,
                               DEFAULT_EPS, [0, 0, 0], False, False, True)["depth"]

    def render_rgb(self, vertices, faces, textures):
        if self.fill_back:
            faces = self._fill_back_faces(faces)
            textures = torch.cat((textures, textures.permute((0, 1, 4, 3, 2, 5))), dim=1)
        # [nr] lighting: ambient*colour + directional*relu(n.dir); the reference sets ambient=1,
        # directional=0 (renderer.py:48-49) so the light is exactly (1,1,1)
        if self.light_intensity_directional != 0:
            raise ValueError("oracle nr port: directional lighting is not restated (reference sets 0)")
        light = torch.zeros(3) + self.light_intensity_ambient * torch.tensor(self.light_color_ambient,
                                                                             dtype=torch.float32)
        textures = textures * light
        vertices = projection(vertices, self.K, self.R, self.t, self.dist_coeffs, self.orig_size)
        faces = vertices_to_faces(vertices, faces)
        return rasterize_rgbad(faces, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                               self.rasterizer_eps, self.background_color, True, False, False)["rgb"]
