"""Drop-in for the reference's `GAN2Shape.renderer.Renderer` (GAN2Shape/renderer/renderer.py:13-277).

Same constructor, same method names, same argument meaning; every operator is a torch.autograd.Function whose
body is a hand-written sm_100a kernel behind the C ABI in include/g2s_b200.h.  Differences that are deliberate:
  * `get_grid` / `get_face_idx` are never materialised (the reference rebuilds them on the CPU per call);
  * `align_corners` is explicit (the reference calls F.grid_sample without it: torch >= 1.3 means False,
    the torch 1.2 the authors pinned meant True) -- default False = what the reference code does today;
  * `render_chain` exposes the fused projected-view render the callers compose by hand (model.py:243-270).
"""
import ctypes
import math

import torch

from . import _lib, functional as Fn
from .utils import get_transform_matrices, get_lighting_directions

EPS = 1e-7  # renderer.py:10

# neural_renderer module defaults used by render_depth (the ctor's near/far reach only render_rgb)
NR_DEPTH_NEAR, NR_DEPTH_FAR = 0.1, 100.0


class Renderer:
    def __init__(self, cfgs, image_size, min_depth, max_depth, device="cuda", align_corners=False):
        """renderer.py:14-54."""
        self.image_size = image_size
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.rot_center_depth = cfgs.get('rot_center_depth', (self.min_depth + self.max_depth) / 2)
        self.fov = cfgs.get('fov', 10)
        self.tex_cube_size = cfgs.get('tex_cube_size', 2)
        self.renderer_min_depth = cfgs.get('renderer_min_depth', 0.1)
        self.renderer_max_depth = cfgs.get('renderer_max_depth', 10.)
        self.align_corners = bool(align_corners)
        # the fused chain's forward hands its projected vertices to the backward (16 S^2 bytes per view kept until then, 1 GB for
        # 4096 views at 128^2) so that the backward does not project the mesh again; False = recompute instead of keep
        self.share_projection = True
        self.device = torch.device(device)

        fx = (self.image_size - 1) / 2 / (math.tan(self.fov / 2 * math.pi / 180))
        fy = (self.image_size - 1) / 2 / (math.tan(self.fov / 2 * math.pi / 180))
        cx = (self.image_size - 1) / 2
        cy = (self.image_size - 1) / 2
        # K and its inverse are 3x3 one-off host work (the inverse via the same LAPACK path torch-CPU uses)
        K = torch.tensor([[fx, 0., cx], [0., fy, cy], [0., 0., 1.]], dtype=torch.float32)
        self._set_K(K.unsqueeze(0), torch.inverse(K).unsqueeze(0), origin=True)
        self.background_color = [1., 1., 1.]  # renderer.py:54
        self._zbuf = Fn.ZBuffer()
        self._raster_scratch = Fn.RasterScratch()
        self._ctx = {}        # g2s_context per device (created on first use)
        self.rot_mat = None
        self.trans_xyz = None
        _lib.load()  # fail loudly at construction when the CUDA library is missing

    # -- camera state ----------------------------------------------------------------------------------
    def _set_K(self, K_host, inv_K_host, origin=False):
        self._K_host, self._inv_K_host = K_host.clone(), inv_K_host.clone()
        self.K = K_host.to(self.device)
        self.inv_K = inv_K_host.to(self.device)
        if origin:
            self._K_origin_host, self._inv_K_origin_host = K_host.clone(), inv_K_host.clone()
            self.K_origin, self.inv_K_origin = self.K.clone(), self.inv_K.clone()
            # neural_renderer captures K at construction (renderer.py:47-50); downscale_K never reaches it
            self._K_raster_host = K_host.clone()
        self._cams = {}

    def _camera(self, depth_pass=False, rgb_pass=False):
        """struct g2s_camera for a launch.  The grid operators use the current K / inv_K; the rasteriser uses
        the K captured at construction (as nr.Renderer does) with near/far per pass."""
        key = (depth_pass, rgb_pass)
        cam = self._cams.get(key)
        if cam is None:
            cam = _lib.Camera()
            Ksrc = self._K_raster_host if (depth_pass or rgb_pass) else self._K_host
            K = Ksrc.reshape(-1).tolist()
            Kg = self._K_host.reshape(-1).tolist()      # the grid operators always use the current K (renderer.py:82-88)
            iK = self._inv_K_host.reshape(-1).tolist()
            for i in range(9):
                cam.K[i] = K[i]
                cam.K_grid[i] = Kg[i]
                cam.inv_K[i] = iK[i]
            cam.rot_center_depth = self.rot_center_depth
            if rgb_pass:
                cam.near_z, cam.far_z = self.renderer_min_depth, self.renderer_max_depth
            else:
                cam.near_z, cam.far_z = NR_DEPTH_NEAR, NR_DEPTH_FAR
            margin = (self.max_depth - self.min_depth) / 2
            cam.clamp_lo = self.min_depth - margin
            cam.clamp_hi = self.max_depth + margin
            cam.image_size = self.image_size
            self._cams[key] = cam
        return cam

    def _context(self, device):
        """the caller-owned g2s_context of `device` (streams / events of the multi-lane forward)"""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        ctx = self._ctx.get(idx)
        if ctx is None:
            with torch.cuda.device(idx):
                ctx = _lib.Context()
            self._ctx[idx] = ctx
        return ctx

    def downscale_K(self, downscale):
        """renderer.py:56-59 (does not reach the K the rasteriser captured, as in the reference)."""
        if downscale > 1:
            K = torch.cat((self._K_origin_host[:, 0:2] / downscale, self._K_origin_host[:, 2:]), dim=1)
            self._set_K(K, torch.inverse(K[0]).unsqueeze(0))

    def set_transform_matrices(self, view):
        """renderer.py:61-62.  CUDA views go through one kernel (k_view_fwd/bwd); the torch mirror in utils.py is kept
        for host-side (CPU) use of the class."""
        if view.is_cuda:
            self.rot_mat, self.trans_xyz = Fn.ViewToRtFn.apply(view)
        else:
            self.rot_mat, self.trans_xyz = get_transform_matrices(view)

    # -- point helpers kept for API compatibility (plain torch, not on the hot path) -------------------
    def rotate_pts(self, pts, rot_mat):
        """renderer.py:64-69."""
        centroid = torch.tensor([0., 0., self.rot_center_depth], device=pts.device).view(1, 1, 3)
        return (pts - centroid).matmul(rot_mat.transpose(2, 1)) + centroid

    def translate_pts(self, pts, trans_xyz):
        """renderer.py:71-72."""
        return pts + trans_xyz

    # -- hot-path operators -------------------------------------------------------------------------
    def depth_to_3d_grid(self, depth):
        """renderer.py:74-80 -> [B,H,W,3]."""
        return Fn.Grid3dFn.apply(depth, None, None, self, 0)

    def grid_3d_to_2d(self, grid_3d):
        """renderer.py:82-88 -> [B,H,W,2] in [-1,1]."""
        return Fn.Grid3dTo2dFn.apply(grid_3d, self)

    def get_warped_3d_grid(self, depth):
        """renderer.py:90-95 -> [B,H,W,3]."""
        return Fn.Grid3dFn.apply(depth, self.rot_mat, self.trans_xyz, self, 1)

    def get_inv_warped_3d_grid(self, depth):
        """renderer.py:97-102 -> [B,H,W,3]."""
        return Fn.Grid3dFn.apply(depth, self.rot_mat, self.trans_xyz, self, 2)

    def get_warped_2d_grid(self, depth):
        """renderer.py:104-108."""
        return Fn.WarpGridFn.apply(depth, self.rot_mat, self.trans_xyz, self, False)

    def get_inv_warped_2d_grid(self, depth):
        """renderer.py:110-114."""
        return Fn.WarpGridFn.apply(depth, self.rot_mat, self.trans_xyz, self, True)

    def warp_canon_depth(self, canon_depth, return_face_idx=False):
        """renderer.py:116-125.  `return_face_idx=True` also returns the int32 [B,2S,2S] face-index map the
        reference API hides (image orientation, -1 = background)."""
        recon, fidx = Fn.WarpCanonDepthFn.apply(canon_depth, self.rot_mat, self.trans_xyz, self)
        return (recon, fidx) if return_face_idx else recon

    def get_normal_from_depth(self, depth):
        """renderer.py:127-139."""
        return Fn.NormalFromDepthFn.apply(depth, self)

    def grid_sample(self, im, grid, mode='bilinear'):
        """nn.functional.grid_sample as the reference's callers use it (model.py:151, 270)."""
        return Fn.grid_sample(im, grid, mode, self.align_corners)

    def render_chain(self, depth, albedo, view, light, views_per_image=None):
        """The fused projected-view render of model.py:243-270: depth [N,S,S], albedo [N,3,S,S],
        view [N*P,6], raw light [N*P,4] -> (recon_im [N*P,3,S,S], recon_depth [N*P,S,S], face_idx)."""
        N = depth.shape[0]
        B = view.shape[0]
        P = views_per_image if views_per_image is not None else B // N
        if N * P != B:
            raise RuntimeError("render_chain: view must have n_images * views_per_image rows")
        # one autograd node from the raw view / light (sets rot_mat / trans_xyz like set_transform_matrices)
        return Fn.RenderChainViewFn.apply(depth, albedo, view, light, self, P, self.align_corners)

    def render_chain_loss(self, depth, albedo, view, light, target, masks=None, depth_thresh=None, views_per_image=None):
        """render_chain with the step-3 photometric loss of model.py:265-274 taken inside the render:
        loss = PhotometricLoss(recon_im, target, mask=(recon_depth < depth_thresh).unsqueeze(1) * masks); depth_thresh
        defaults to max_depth + (max_depth - min_depth) / 2 (model.py:264-268).  target [N*P,3,S,S], masks [N*P,1,S,S] or
        None.  Returns (loss, recon_im, recon_depth, face_idx): further losses on recon_im / recon_depth (the perceptual
        loss of model.py:275) back-propagate through the same backward call."""
        N = depth.shape[0]
        B = view.shape[0]
        P = views_per_image if views_per_image is not None else B // N
        if N * P != B:
            raise RuntimeError("render_chain_loss: view must have n_images * views_per_image rows")
        if depth_thresh is None:
            depth_thresh = self.max_depth + (self.max_depth - self.min_depth) / 2
        im, rd, fidx, loss = Fn.RenderChainLossViewFn.apply(depth, albedo, view, light, target, masks, self, P,
                                                            self.align_corners, depth_thresh)
        return loss, im, rd, fidx

    def render_pseudo_views(self, depth, albedo, view, light_a, light_b, light_d, mask=None, views_per_image=None):
        """Forward-only twin of render_chain used by sample_pseudo_imgs (model.py:291-328): shade the canonical
        albedo with per-view lighting `shading = a + b * max(0, n.d)` (for the random relighting of model.py:298-309
        pass a = light_a + alpha * rand, b = light_b + rand, d = rand_light_d), warp it to `view` as
        render_given_view(..., grid_sample=True) does (renderer.py:257-264) and warp `mask` [N,1|3,S,S] (None = ones)
        with mode='nearest'.  Returns (pseudo_im [B,3,S,S] clamped to [-1,1], mask [B,1,S,S]); no gradients."""
        N, B = depth.shape[0], view.shape[0]
        P = views_per_image if views_per_image is not None else B // N
        with torch.no_grad():
            self.set_transform_matrices(view)
            light5 = torch.cat([light_a.reshape(B, 1), light_b.reshape(B, 1), light_d.reshape(B, 3)], 1)
            m = mask[:, 0] if mask is not None else None
            im, _, _, mask_out = Fn.RenderChainFn.apply(depth, albedo, self.rot_mat, self.trans_xyz, light5, self, P,
                                                        self.align_corners, m, True)
        return im, mask_out

    # -- sweeps -----------------------------------------------------------------------------------------
    def _view_sample(self, im, depth, view):
        self.set_transform_matrices(view)
        recon_depth = self.warp_canon_depth(depth)
        grid = self.get_inv_warped_2d_grid(recon_depth)
        return Fn.grid_sample(im, grid, 'bilinear', self.align_corners), grid

    def _grid3d(self, depth, crop_mesh=None, v_before=None, rot1=None, v_after=None):
        """depth_to_3d_grid (+crop, inverse warp by v_before, rotation rot1, warp by v_after) -> [B,H*W,3]."""
        lib = _lib.load()
        B, H, W = depth.shape
        dstore, dstride = Fn._batched_image(depth)
        dev = depth.device
        R0 = t0 = R1 = R2 = t2 = None
        if v_before is not None:
            R0, t0 = self._views_to_Rt(v_before)
            R0, t0 = Fn._Rt(R0.detach(), t0.detach(), B)
        if rot1 is not None:
            R1 = Fn._f32c(rot1.detach().expand(B, 3, 3))
        if v_after is not None:
            R2, t2 = self._views_to_Rt(v_after)
            R2, t2 = Fn._Rt(R2.detach(), t2.detach(), B)
        crop = (ctypes.c_int * 4)(*[int(c) for c in crop_mesh]) if crop_mesh is not None else None
        out = torch.empty(B, H * W, 3, device=dev, dtype=torch.float32)
        _lib.check(lib.g2s_grid3d_fwd(ctypes.byref(self._camera()), Fn._p(dstore), dstride, B, H, W, crop,
                                      Fn._p(R0), Fn._p(t0), Fn._p(R1), Fn._p(R2), Fn._p(t2), Fn._p(out),
                                      Fn._stream()), "g2s_grid3d_fwd")
        return out

    def _render_rgb(self, vertices3d, im, clamp=True, return_face_idx=False):
        """nr.Renderer.render_rgb(vertices, get_face_idx, get_textures_from_im(im, tex_cube_size)) +
        clamp(-1,1) (renderer.py:194-196).  Differentiable with respect to `im` (backward_textures) and to the vertices
        (neural_renderer's approximate backward_pixel_map gradient): Fn.RenderRgbFn."""
        out, fidx = Fn.RenderRgbFn.apply(vertices3d, im, self, clamp)
        return (out, fidx) if return_face_idx else out

    @staticmethod
    def _views_to_Rt(view):
        if view.is_cuda:
            return Fn.ViewToRtFn.apply(view)
        return get_transform_matrices(view)

    def _sweep(self, im, depth, angles, v_before=None, v_after=None, grid_sample=False, crop_mesh=None):
        """All `t` rotations of a sweep in ONE batch of b*t views (the reference loops over them, rebuilding faces and
        textures per iteration: renderer.py:169-197).  angles: [t,3] rotation vectors.  -> [b, t, c, h, w]."""
        b, c, h, w = im.shape
        T = angles.shape[0]
        dev = im.device
        angles = angles.to(device=dev, dtype=torch.float32)
        rep = (lambda x: x.expand(T, *x.shape[1:])) if b == 1 else (lambda x: x.repeat_interleave(T, 0))
        depth_bt, im_bt = rep(depth), rep(im)
        if grid_sample:
            view = torch.cat([angles, torch.zeros(T, 3, device=dev)], 1).repeat(b, 1)          # [b*T, 6], (b, t) order
            if v_before is not None:
                view = view - rep(v_before.expand(b, v_before.shape[1]))
            warped = self._view_sample(im_bt, depth_bt, view)[0]
        else:
            R1, _ = self._views_to_Rt(angles)
            R1 = R1.detach().repeat(b, 1, 1)
            vb = rep(v_before.expand(b, v_before.shape[1])) if v_before is not None else None
            va = None
            if v_after is not None:
                if len(v_after.shape) == 3:                                                   # [t, b, 6] (renderer.py:186-187)
                    va = v_after.expand(T, b, v_after.shape[2]).transpose(0, 1).reshape(b * T, -1)
                else:
                    va = rep(v_after.expand(b, v_after.shape[1]))
            verts = self._grid3d(depth_bt, crop_mesh, vb, R1, va)
            warped = self._render_rgb(verts, im if b == 1 else im_bt)      # one image: shared by all T views
        return warped.reshape(b, T, c, h, w)

    def render_yaw(self, im, depth, v_before=None, v_after=None, rotations=None, maxr=90, nsample=9,
                   grid_sample=False, crop_mesh=None):
        """renderer.py:141-198 -> [b, t, c, h, w]."""
        if rotations is None:
            rotations = torch.linspace(-math.pi / 180 * maxr, math.pi / 180 * maxr, nsample)
        rot = torch.as_tensor(rotations, dtype=torch.float32).reshape(-1)
        angles = torch.zeros(rot.shape[0], 3)
        angles[:, 1] = rot.cpu()
        return self._sweep(im, depth, angles, v_before, v_after, grid_sample, crop_mesh)

    def render_view(self, im, depth, v_before=None, rotations=None, maxr=[20, 90], nsample=[5, 9],
                    grid_sample=False):
        """renderer.py:200-250: yaw sweep, then pitch sweep -> [b, t, c, h, w]."""
        rotations_p = torch.linspace(-math.pi / 180 * maxr[0], math.pi / 180 * maxr[0], nsample[0])
        rotations_y = torch.linspace(-math.pi / 180 * maxr[1], math.pi / 180 * maxr[1], nsample[1])
        angles = torch.zeros(nsample[1] + nsample[0], 3)
        angles[:nsample[1], 1] = rotations_y
        angles[nsample[1]:, 0] = rotations_p
        return self._sweep(im, depth, angles, v_before, None, grid_sample, None)

    def render_given_view(self, im, depth, view, mask=None, grid_sample=True):
        """renderer.py:252-277."""
        if grid_sample:
            warped, grid = self._view_sample(im, depth, view)
            if mask is not None:
                return warped, Fn.grid_sample(mask, grid, 'nearest', self.align_corners)
            return warped
        verts = self._grid3d(depth, None, None, None, view)
        warped = self._render_rgb(verts, im)
        if mask is not None:
            return warped, self._render_rgb(verts, mask)
        return warped
