"""Drop-in for the part of `neural_renderer` that GAN2Shape/renderer/renderer.py uses (the LOWER boundary of the path):

    import g2s_b200.nr_compat as nr          # instead of:  import neural_renderer as nr     (renderer.py:6)
    r = nr.Renderer(camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0., K=K, R=R, t=t,
                    near=near, far=far, image_size=S, orig_size=S, fill_back=True, background_color=[1, 1, 1])   # :47-54
    r.render_depth(vertices, faces)                 # renderer.py:120              -> [B,S,S]
    r.render_rgb(vertices, faces, textures)         # renderer.py:196,230,248,272,275 -> [B,3,S,S]

with the reference's own `get_face_idx` / `get_textures_from_im` tensors as `faces` / `textures`.  Only what the reference
reaches is supported and everything else raises: camera_mode='projection' with R = I, t = 0 (the reference passes identity
extrinsics and moves the vertices itself), ambient light 1 / directional 0, anti-aliasing on, fill_back on, the S x S
grid-mesh topology of utils.py:76-80, texture cubes of size 2 built by utils.py:98-109.

render_depth is differentiable with respect to the vertices (neural_renderer's approximate backward_depth_map gradient);
render_rgb with respect to the vertices (backward_pixel_map) and, through the image the cubes were built from, the textures'
corner colours (backward_textures).  CUDA tensors only; no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .functional import RenderRgbFn, ZBuffer, _f32c, _p, _require_cuda, _stream
from .utils import get_face_idx


class _RenderDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, owner):
        lib = _lib.load()
        v = _f32c(vertices)
        B, S = v.shape[0], owner.image_size
        cam = owner._camera(depth_pass=True)
        zbuf = owner._zbuf.get(B, S, cam.far_z, v.device)
        out = torch.empty(B, S, S, device=v.device, dtype=torch.float32)
        fidx = torch.empty(B, 2 * S, 2 * S, device=v.device, dtype=torch.int32)
        _lib.check(lib.g2s_render_depth_fwd(ctypes.byref(cam), _p(v), B, _p(zbuf), _p(out), _p(fidx), _stream()),
                   "g2s_render_depth_fwd")
        ctx.save_for_backward(v, fidx)
        ctx.owner = owner
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        v, fidx = ctx.saved_tensors
        B, S = v.shape[0], ctx.owner.image_size
        cam = ctx.owner._camera(depth_pass=True)
        ws = torch.empty(_lib.ws_floats(_lib.WS_RASTER_BWD, B, S), device=v.device, dtype=torch.float32)
        gv = torch.empty_like(v)
        _lib.check(lib.g2s_render_depth_bwd(ctypes.byref(cam), _p(v), B, _p(fidx), _p(_f32c(g)), _p(ws), _p(gv), _stream()),
                   "g2s_render_depth_bwd")
        return gv, None


class Renderer:
    """neural_renderer.Renderer as renderer.py:47-54 constructs it (see the module docstring for what is supported)."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=[0, 0, 0], fill_back=True,
                 camera_mode='projection', K=None, R=None, t=None, dist_coeffs=None, orig_size=1024, near=0.1, far=100,
                 light_intensity_ambient=0.5, light_intensity_directional=0.5, **unused):
        if camera_mode != 'projection' or not anti_aliasing or not fill_back:
            raise NotImplementedError("nr_compat: only camera_mode='projection', anti_aliasing=True, fill_back=True "
                                      "(what GAN2Shape/renderer/renderer.py:47-54 uses)")
        if light_intensity_ambient != 1.0 or light_intensity_directional != 0.0:
            raise NotImplementedError("nr_compat: only ambient light 1 / directional light 0 (renderer.py:48-49)")
        if orig_size != image_size:
            raise NotImplementedError("nr_compat: orig_size must equal image_size (renderer.py:52)")
        if dist_coeffs is not None and float(torch.as_tensor(dist_coeffs).abs().max()) != 0.0:
            raise NotImplementedError("nr_compat: lens distortion is not supported (the reference passes none)")
        for name, m, ref in (("R", R, torch.eye(3)), ("t", t, torch.zeros(3))):
            if m is not None and not torch.equal(torch.as_tensor(m).detach().float().cpu().reshape(ref.shape), ref):
                raise NotImplementedError("nr_compat: %s must be the identity extrinsics (renderer.py:33-34)" % name)
        self.image_size, self.background_color = int(image_size), [float(c) for c in background_color]
        self.near, self.far = float(near), float(far)
        self.K = torch.as_tensor(K).detach().float().cpu().reshape(3, 3)
        self.tex_cube_size = 2
        self._zbuf = ZBuffer()
        self._grid_faces = None
        self._checked = set()      # (data_ptr, shape, device) of `faces` tensors already validated
        self._cams = {}

    # the camera struct of the C ABI; render_depth uses neural_renderer's module defaults near=0.1, far=100 (the
    # constructor's near / far only reach render_rgb), exactly like the reference's Renderer._camera
    def _camera(self, depth_pass=False, rgb_pass=False):
        cam = self._cams.get(bool(depth_pass))
        if cam is None:      # built once: the 3x3 inverse and the ctypes fill are host work that every render would repeat
            cam = _lib.Camera()
            K = self.K
            invK = torch.inverse(K)
            for i in range(9):
                cam.K[i] = float(K.reshape(-1)[i])
                cam.K_grid[i] = float(K.reshape(-1)[i])
                cam.inv_K[i] = float(invK.reshape(-1)[i])
            cam.rot_center_depth = 0.0
            cam.near_z, cam.far_z = (0.1, 100.0) if depth_pass else (self.near, self.far)
            cam.clamp_lo, cam.clamp_hi = -3.0e38, 3.0e38
            cam.image_size = self.image_size
            self._cams[bool(depth_pass)] = cam
        return cam

    def _check_mesh(self, vertices, faces):
        _require_cuda(vertices)
        S = self.image_size
        if vertices.dim() != 3 or vertices.shape[1] != S * S or vertices.shape[2] != 3:
            raise RuntimeError("nr_compat: vertices must be [B,%d,3] (the S x S grid mesh)" % (S * S))
        F = 2 * (S - 1) * (S - 1)
        if faces.dim() != 3 or faces.shape[1] != F or faces.shape[2] != 3:
            raise RuntimeError("nr_compat: faces must be get_face_idx(b, S, S): [B,%d,3]" % F)
        key = (faces.data_ptr(), tuple(faces.shape), str(faces.device), faces._version)
        if key not in self._checked:     # the device-to-host copy + compare is a blocking sync: once per faces tensor
            if self._grid_faces is None:
                self._grid_faces = get_face_idx(1, S, S)[0]
            if not torch.equal(faces[0].detach().to("cpu", torch.int32), self._grid_faces):
                raise NotImplementedError("nr_compat: only the grid-mesh topology of utils.py:76-80 is supported")
            if len(self._checked) > 64:
                self._checked.clear()
            self._checked.add(key)

    def render_depth(self, vertices, faces):
        """[B,S*S,3], int32 [B,2(S-1)^2,3] -> [B,S,S] (background = 100, not clamped)."""
        self._check_mesh(vertices, faces)
        return _RenderDepthFn.apply(vertices, self)

    def render_rgb(self, vertices, faces, textures):
        """+ textures [B,2(S-1)^2,2,2,2,C] from get_textures_from_im(im, 2) -> [B,C,S,S] (not clamped)."""
        self._check_mesh(vertices, faces)
        _require_cuda(textures)
        S, B = self.image_size, vertices.shape[0]
        Q = (S - 1) * (S - 1)
        if textures.dim() != 6 or textures.shape[1] != 2 * Q or tuple(textures.shape[2:5]) != (2, 2, 2):
            raise NotImplementedError("nr_compat: textures must be get_textures_from_im(im, tx_size=2) cubes")
        C = textures.shape[5]
        # the cube's unit corners hold the three vertex colours (utils.py:83-95): rebuild the per-pixel image they came from
        t1, t2 = textures[:, :Q], textures[:, Q:]
        im = torch.empty(B, C, S, S, device=textures.device, dtype=torch.float32)
        im[:, :, :S - 1, :S - 1] = t1[:, :, 1, 0, 0].reshape(B, S - 1, S - 1, C).permute(0, 3, 1, 2)    # top-left corners
        im[:, :, :S - 1, S - 1] = t1[:, :, 0, 1, 0].reshape(B, S - 1, S - 1, C)[:, :, -1].permute(0, 2, 1)   # top-right, last column
        im[:, :, S - 1, :S - 1] = t1[:, :, 0, 0, 1].reshape(B, S - 1, S - 1, C)[:, -1].permute(0, 2, 1)      # bottom-left, last row
        im[:, :, S - 1, S - 1] = t2[:, :, 0, 0, 1].reshape(B, S - 1, S - 1, C)[:, -1, -1]                    # bottom-right corner
        return RenderRgbFn.apply(vertices, im, self, False)[0]
