"""oracle/renderer_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch-CPU, fp32) of the reference's renderer hot path:
  /root/reference/GAN2Shape/renderer/renderer.py:13-277   (class Renderer)
  /root/reference/GAN2Shape/renderer/utils.py:22-109      (grid, rotation, faces, textures)
  /root/reference/GAN2Shape/model.py:347-360              (lighting directions, Lambertian shading)
  /root/reference/GAN2Shape/model.py:146-151, 260-270     (mask + bilinear sampling; "chain C")
with the external rasteriser served by oracle/nr_port.py + oracle/nr_raster.c.

Every 3-wide contraction the reference writes as `matmul` goes through nr_port.Mm3: the explicit chain
fma(c,m2, fma(b,m1, a*m0)) that torch-CPU matmul evaluates in the build container (oracle/fma_mm.c), so
values do not depend on the host BLAS;
tests/test_oracle_vs_reference.py pins this against the reference's unmodified code (oracle/ref_shim.py)
in the build container, and tests/golden/*.npz carry the reference's outputs to the GPU box.
Autograd through these ops is the oracle for every gradient.
"""
import math

import torch
import torch.nn.functional as F

from . import nr_port

EPS = 1e-7  # renderer.py:10


mm3 = nr_port.mm3


# ----------------------------------------------------------------------------- utils.py
def get_grid(b, H, W, normalize=True):
    """utils.py:22-30."""
    if normalize:
        h_range = torch.linspace(-1, 1, H)
        w_range = torch.linspace(-1, 1, W)
    else:
        h_range = torch.arange(0, H)
        w_range = torch.arange(0, W)
    yy, xx = torch.meshgrid(h_range, w_range, indexing="ij")
    return torch.stack([xx, yy], -1).repeat(b, 1, 1, 1).float()


def get_rotation_matrix(tx, ty, tz):
    """utils.py:33-49: R = Rz @ (Ry @ Rx)."""
    n = len(tx)
    cx, sx, cy, sy, cz, sz = tx.cos(), tx.sin(), ty.cos(), ty.sin(), tz.cos(), tz.sin()
    zero, one = torch.zeros(n), torch.ones(n)
    m_x = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(n, 3, 3)
    m_y = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(n, 3, 3)
    m_z = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(n, 3, 3)

    def bmm(A, B):
        # torch-CPU bmm on [n,3,3] x [n,3,3] evaluates un-fused, left-to-right (pinned against the
        # reference in tests/test_oracle_vs_reference.py); R is an INPUT of every kernel parity test, so
        # its last bit never decides a face index.
        rows = []
        for i in range(3):
            for j in range(3):
                rows.append((A[:, i, 0] * B[:, 0, j] + A[:, i, 1] * B[:, 1, j]) + A[:, i, 2] * B[:, 2, j])
        return torch.stack(rows, 1).view(n, 3, 3)

    return bmm(m_z, bmm(m_y, m_x))


def get_transform_matrices(view):
    """utils.py:52-73."""
    b = view.size(0)
    if view.size(1) == 6:
        trans_xyz = view[:, 3:].reshape(b, 1, 3)
    elif view.size(1) == 5:
        trans_xyz = torch.cat([view[:, 3:].reshape(b, 1, 2), torch.zeros(b, 1, 1)], 2)
    elif view.size(1) == 3:
        trans_xyz = torch.zeros(b, 1, 3)
    else:
        raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")
    return get_rotation_matrix(view[:, 0], view[:, 1], view[:, 2]), trans_xyz


def get_face_idx(b, h, w):
    """utils.py:76-80: all 'upper-left' triangles, then all 'lower-right' ones."""
    idx_map = torch.arange(h * w).reshape(h, w)
    faces1 = torch.stack([idx_map[:h - 1, :w - 1], idx_map[1:, :w - 1], idx_map[:h - 1, 1:]], -1).reshape(-1, 3)
    faces2 = torch.stack([idx_map[:h - 1, 1:], idx_map[1:, :w - 1], idx_map[1:, 1:]], -1).reshape(-1, 3)
    return torch.cat([faces1, faces2], 0).repeat(b, 1, 1).int()


_CUBE = [[0.5, 0.5, 0.5], [0, 0, 1], [0, 1, 0], [-0.5, 0.5, 0.5],
         [1, 0, 0], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [0, 0, 0]]


def vcolor_to_texture_cube(vcolors):
    """utils.py:83-95: bxcxnx3 vertex colours -> bxnx2x2x2xc cube."""
    b, c, n, f = vcolors.shape
    coeffs = torch.tensor(_CUBE, dtype=torch.float32)
    v = vcolors.permute(0, 2, 3, 1).reshape(b * n, 3, c)  # (b n) 3 c
    return nr_port.mm_k3(coeffs.unsqueeze(0), v).reshape(b, n, 2, 2, 2, c)


def get_textures_from_im(im, tx_size=1):
    """utils.py:98-109."""
    b, c, h, w = im.shape
    if tx_size == 1:
        textures = torch.cat([im[:, :, :h - 1, :w - 1].reshape(b, c, -1), im[:, :, 1:, 1:].reshape(b, c, -1)], 2)
        return textures.transpose(2, 1).reshape(b, -1, 1, 1, 1, c)
    if tx_size == 2:
        t1 = torch.stack([im[:, :, :h - 1, :w - 1], im[:, :, :h - 1, 1:], im[:, :, 1:, :w - 1]], -1).reshape(b, c, -1, 3)
        t2 = torch.stack([im[:, :, 1:, :w - 1], im[:, :, :h - 1, 1:], im[:, :, 1:, 1:]], -1).reshape(b, c, -1, 3)
        return vcolor_to_texture_cube(torch.cat([t1, t2], 2))
    raise NotImplementedError("Currently support texture size of 1 or 2 only.")


# ----------------------------------------------------------------------------- model.py glue
def get_lighting_directions(lighting):
    """model.py:347-353."""
    a = lighting[:, :1] / 2 + 0.5
    b = lighting[:, 1:2] / 2 + 0.5
    d = torch.cat([lighting[:, 2:], torch.ones(lighting.size(0), 1)], 1)
    d = d / ((d ** 2).sum(1, keepdim=True)) ** 0.5
    return a, b, d


def get_shading(normal, lighting_a, lighting_b, lighting_d, albedo):
    """model.py:355-360."""
    diffuse = (normal * lighting_d.view(-1, 1, 1, 3)).sum(3).clamp(min=0).unsqueeze(1)
    shading = lighting_a.view(-1, 1, 1, 1) + lighting_b.view(-1, 1, 1, 1) * diffuse
    texture = (albedo / 2 + 0.5) * shading * 2 - 1
    return diffuse, texture


# ----------------------------------------------------------------------------- renderer.py
class OracleRenderer:
    def __init__(self, cfgs, image_size, min_depth, max_depth, K=None, inv_K=None):
        """renderer.py:14-54.  K / inv_K may be injected so the product and the oracle share them."""
        self.image_size = image_size
        self.min_depth, self.max_depth = min_depth, max_depth
        self.rot_center_depth = cfgs.get('rot_center_depth', (min_depth + max_depth) / 2)
        self.fov = cfgs.get('fov', 10)
        self.tex_cube_size = cfgs.get('tex_cube_size', 2)
        self.renderer_min_depth = cfgs.get('renderer_min_depth', 0.1)
        self.renderer_max_depth = cfgs.get('renderer_max_depth', 10.)
        f = (image_size - 1) / 2 / (math.tan(self.fov / 2 * math.pi / 180))
        c = (image_size - 1) / 2
        Kt = torch.tensor([[f, 0., c], [0., f, c], [0., 0., 1.]], dtype=torch.float32)
        self.K = (Kt if K is None else K.reshape(3, 3).float()).unsqueeze(0)
        self.inv_K = (torch.inverse(Kt) if inv_K is None else inv_K.reshape(3, 3).float()).unsqueeze(0)
        self.K_origin, self.inv_K_origin = self.K.clone(), self.inv_K.clone()
        self.renderer = nr_port.Renderer(
            camera_mode='projection', light_intensity_ambient=1.0, light_intensity_directional=0.,
            K=self.K, R=torch.eye(3).unsqueeze(0), t=torch.zeros(1, 3),
            near=self.renderer_min_depth, far=self.renderer_max_depth,
            image_size=image_size, orig_size=image_size, fill_back=True, background_color=[1, 1, 1])

    def downscale_K(self, downscale):
        """renderer.py:56-59 (does not reach the K captured by the rasteriser)."""
        if downscale > 1:
            self.K = torch.cat((self.K_origin[:, 0:2] / downscale, self.K_origin[:, 2:]), dim=1)
            self.inv_K = torch.inverse(self.K[0]).unsqueeze(0)

    def set_transform_matrices(self, view):
        self.rot_mat, self.trans_xyz = get_transform_matrices(view)

    def rotate_pts(self, pts, rot_mat):
        """renderer.py:64-69."""
        centroid = torch.tensor([0., 0., self.rot_center_depth]).view(1, 1, 3)
        return mm3(pts - centroid, rot_mat) + centroid

    def translate_pts(self, pts, trans_xyz):
        return pts + trans_xyz

    def depth_to_3d_grid(self, depth):
        """renderer.py:74-80."""
        b, h, w = depth.shape
        grid_2d = get_grid(b, h, w, normalize=False)
        depth = depth.unsqueeze(-1)
        grid_3d = torch.cat((grid_2d, torch.ones_like(depth)), dim=3)
        return mm3(grid_3d, self.inv_K) * depth

    def grid_3d_to_2d(self, grid_3d):
        """renderer.py:82-88."""
        b, h, w, _ = grid_3d.shape
        grid_2d = mm3(grid_3d / grid_3d[..., 2:], self.K)[:, :, :, :2]
        WH = torch.tensor([w - 1, h - 1], dtype=torch.float32).view(1, 1, 1, 2)
        return grid_2d / WH * 2. - 1.

    def get_warped_3d_grid(self, depth):
        b, h, w = depth.shape
        g = self.depth_to_3d_grid(depth).reshape(b, -1, 3)
        g = self.translate_pts(self.rotate_pts(g, self.rot_mat), self.trans_xyz)
        return g.reshape(b, h, w, 3)

    def get_inv_warped_3d_grid(self, depth):
        b, h, w = depth.shape
        g = self.depth_to_3d_grid(depth).reshape(b, -1, 3)
        g = self.rotate_pts(self.translate_pts(g, -self.trans_xyz), self.rot_mat.transpose(2, 1))
        return g.reshape(b, h, w, 3)

    def get_warped_2d_grid(self, depth):
        return self.grid_3d_to_2d(self.get_warped_3d_grid(depth))

    def get_inv_warped_2d_grid(self, depth):
        return self.grid_3d_to_2d(self.get_inv_warped_3d_grid(depth))

    def warp_canon_depth(self, canon_depth):
        """renderer.py:116-125."""
        b, h, w = canon_depth.shape
        grid_3d = self.get_warped_3d_grid(canon_depth).reshape(b, -1, 3)
        warped = self.renderer.render_depth(grid_3d, get_face_idx(b, h, w))
        margin = (self.max_depth - self.min_depth) / 2
        return warped.clamp(min=self.min_depth - margin, max=self.max_depth + margin)

    def get_normal_from_depth(self, depth):
        """renderer.py:127-139."""
        b, h, w = depth.shape
        g = self.depth_to_3d_grid(depth)
        tu = g[:, 1:-1, 2:] - g[:, 1:-1, :-2]
        tv = g[:, 2:, 1:-1] - g[:, :-2, 1:-1]
        normal = torch.stack([tu[..., 1] * tv[..., 2] - tu[..., 2] * tv[..., 1],
                              tu[..., 2] * tv[..., 0] - tu[..., 0] * tv[..., 2],
                              tu[..., 0] * tv[..., 1] - tu[..., 1] * tv[..., 0]], 3)
        zero = torch.tensor([0., 0., 1.])
        normal = torch.cat([zero.repeat(b, h - 2, 1, 1), normal, zero.repeat(b, h - 2, 1, 1)], 2)
        normal = torch.cat([zero.repeat(b, 1, w, 1), normal, zero.repeat(b, 1, w, 1)], 1)
        return normal / (((normal ** 2).sum(3, keepdim=True)) ** 0.5 + EPS)

    # -- mesh-texture / sweep operators ------------------------------------------------------------
    def _crop(self, grid_3d, crop_mesh):
        """renderer.py:145-158 (in-place edits of the 3-D grid)."""
        top, bottom, left, right = crop_mesh
        if top > 0:
            grid_3d[:, :top, :, 1] = grid_3d[:, top:top + 1, :, 1].repeat(1, top, 1)
            grid_3d[:, :top, :, 2] = grid_3d[:, top:top + 1, :, 2].repeat(1, top, 1)
        if bottom > 0:
            grid_3d[:, -bottom:, :, 1] = grid_3d[:, -bottom - 1:-bottom, :, 1].repeat(1, bottom, 1)
            grid_3d[:, -bottom:, :, 2] = grid_3d[:, -bottom - 1:-bottom, :, 2].repeat(1, bottom, 1)
        if left > 0:
            grid_3d[:, :, :left, 0] = grid_3d[:, :, left:left + 1, 0].repeat(1, 1, left)
            grid_3d[:, :, :left, 2] = grid_3d[:, :, left:left + 1, 2].repeat(1, 1, left)
        if right > 0:
            grid_3d[:, :, -right:, 0] = grid_3d[:, :, -right - 1:-right, 0].repeat(1, 1, right)
            grid_3d[:, :, -right:, 2] = grid_3d[:, :, -right - 1:-right, 2].repeat(1, 1, right)
        return grid_3d

    def _sample_view(self, im, depth, view, align_corners):
        self.set_transform_matrices(view)
        recon_depth = self.warp_canon_depth(depth)
        grid = self.get_inv_warped_2d_grid(recon_depth)
        return F.grid_sample(im, grid, mode='bilinear', align_corners=align_corners), grid

    def _mesh_view(self, im, grid_3d_i, b, h, w):
        faces = get_face_idx(b, h, w)
        textures = get_textures_from_im(im, tx_size=self.tex_cube_size)
        return self.renderer.render_rgb(grid_3d_i, faces, textures).clamp(min=-1., max=1.)

    def render_yaw(self, im, depth, v_before=None, v_after=None, rotations=None, maxr=90, nsample=9,
                   grid_sample=False, crop_mesh=None, align_corners=False):
        """renderer.py:141-198."""
        b, c, h, w = im.shape
        grid_3d = self.depth_to_3d_grid(depth)
        if crop_mesh is not None:
            grid_3d = self._crop(grid_3d, crop_mesh)
        grid_3d = grid_3d.reshape(b, -1, 3)
        if v_before is not None:
            rot_mat, trans_xyz = get_transform_matrices(v_before)
            grid_3d = self.rotate_pts(self.translate_pts(grid_3d, -trans_xyz), rot_mat.transpose(2, 1))
        if rotations is None:
            rotations = torch.linspace(-math.pi / 180 * maxr, math.pi / 180 * maxr, nsample)
        out = []
        for i, ri in enumerate(rotations):
            if grid_sample:
                view = torch.tensor([0, float(ri), 0, 0, 0, 0]).view(1, 6)
                if v_before is not None:
                    view = view - v_before
                out.append(self._sample_view(im, depth, view, align_corners)[0])
            else:
                rot_mat_i, _ = get_transform_matrices(torch.tensor([0, float(ri), 0]).view(1, 3))
                g = self.rotate_pts(grid_3d, rot_mat_i.repeat(b, 1, 1))
                if v_after is not None:
                    v_after_i = v_after[i] if v_after.dim() == 3 else v_after
                    rot_mat, trans_xyz = get_transform_matrices(v_after_i)
                    g = self.translate_pts(self.rotate_pts(g, rot_mat), trans_xyz)
                out.append(self._mesh_view(im, g, b, h, w))
        return torch.stack(out, 1)

    def render_view(self, im, depth, v_before=None, rotations=None, maxr=[20, 90], nsample=[5, 9],
                    grid_sample=False, align_corners=False):
        """renderer.py:200-250: yaw sweep then pitch sweep."""
        b, c, h, w = im.shape
        grid_3d = self.depth_to_3d_grid(depth).reshape(b, -1, 3)
        if v_before is not None:
            rot_mat, trans_xyz = get_transform_matrices(v_before)
            grid_3d = self.rotate_pts(self.translate_pts(grid_3d, -trans_xyz), rot_mat.transpose(2, 1))
        rot_p = torch.linspace(-math.pi / 180 * maxr[0], math.pi / 180 * maxr[0], nsample[0])
        rot_y = torch.linspace(-math.pi / 180 * maxr[1], math.pi / 180 * maxr[1], nsample[1])
        out = []
        for axis, angles in ((1, rot_y), (0, rot_p)):
            for a in angles:
                r3 = [0., 0., 0.]
                r3[axis] = float(a)
                if grid_sample:
                    view = torch.tensor(r3 + [0., 0., 0.]).view(1, 6)
                    if v_before is not None:
                        view = view - v_before
                    out.append(self._sample_view(im, depth, view, align_corners)[0])
                else:
                    rot_mat_i, _ = get_transform_matrices(torch.tensor(r3).view(1, 3))
                    out.append(self._mesh_view(im, self.rotate_pts(grid_3d, rot_mat_i.repeat(b, 1, 1)), b, h, w))
        return torch.stack(out, 1)

    def render_given_view(self, im, depth, view, mask=None, grid_sample=True, align_corners=False):
        """renderer.py:252-277."""
        b, c, h, w = im.shape
        if grid_sample:
            warped, grid = self._sample_view(im, depth, view, align_corners)
            if mask is not None:
                return warped, F.grid_sample(mask, grid, mode='nearest', align_corners=align_corners)
            return warped
        rot_mat, trans_xyz = get_transform_matrices(view)
        g = self.depth_to_3d_grid(depth).reshape(b, -1, 3)
        g = self.translate_pts(self.rotate_pts(g, rot_mat), trans_xyz)
        warped = self._mesh_view(im, g, b, h, w)
        if mask is not None:
            return warped, self._mesh_view(mask, g, b, h, w)
        return warped

    # -- the metric's unit of work (SURVEY.md 8d "chain C"; model.py:243-270) ----------------------
    def render_chain(self, depth, albedo, view, light, align_corners=False):
        """depth [1,S,S], albedo [1,3,S,S], view [P,6], light [P,4] ->
        recon_im [P,3,S,S], recon_depth [P,S,S] (+ normal, texture, grid for inspection)."""
        P = view.shape[0]
        S = self.image_size
        normal = self.get_normal_from_depth(depth)
        a, b, d = get_lighting_directions(light)
        diffuse, texture = get_shading(normal, a, b, d, albedo)
        self.set_transform_matrices(view)
        recon_depth = self.warp_canon_depth(depth.expand(P, S, S))
        grid = self.get_inv_warped_2d_grid(recon_depth)
        recon_im = F.grid_sample(texture, grid, mode='bilinear', align_corners=align_corners).clamp(min=-1, max=1)
        return dict(recon_im=recon_im, recon_depth=recon_depth, normal=normal, texture=texture, grid=grid,
                    diffuse=diffuse)
