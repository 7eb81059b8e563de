"""Host-side mirror of GAN2Shape/renderer/utils.py (reference file:line in each docstring).

`get_grid` and `get_face_idx` exist for API compatibility only: the CUDA kernels derive pixel coordinates and
the grid-mesh topology from indices, so the hot path never builds (or copies) these tensors, whereas the
reference rebuilds them on the CPU and copies them to the device on every call (renderer.py:76, 119).
"""
import torch


def get_grid(b, H, W, normalize=True, device=None):
    """utils.py:22-30."""
    if normalize:
        h_range = torch.linspace(-1, 1, H, device=device)
        w_range = torch.linspace(-1, 1, W, device=device)
    else:
        h_range = torch.arange(0, H, device=device)
        w_range = torch.arange(0, W, device=device)
    yy, xx = torch.meshgrid(h_range, w_range, indexing="ij")
    return torch.stack([xx, yy], -1).repeat(b, 1, 1, 1).float()


def get_rotation_matrix(tx, ty, tz):
    """utils.py:33-49: R = Rz @ Ry @ Rx (plain torch ops: tiny, differentiable w.r.t. the view)."""
    m_x = torch.zeros((len(tx), 3, 3), device=tx.device, dtype=tx.dtype)
    m_y = torch.zeros((len(tx), 3, 3), device=tx.device, dtype=tx.dtype)
    m_z = torch.zeros((len(tx), 3, 3), device=tx.device, dtype=tx.dtype)
    m_x[:, 1, 1], m_x[:, 1, 2] = tx.cos(), -tx.sin()
    m_x[:, 2, 1], m_x[:, 2, 2] = tx.sin(), tx.cos()
    m_x[:, 0, 0] = 1
    m_y[:, 0, 0], m_y[:, 0, 2] = ty.cos(), ty.sin()
    m_y[:, 2, 0], m_y[:, 2, 2] = -ty.sin(), ty.cos()
    m_y[:, 1, 1] = 1
    m_z[:, 0, 0], m_z[:, 0, 1] = tz.cos(), -tz.sin()
    m_z[:, 1, 0], m_z[:, 1, 1] = tz.sin(), tz.cos()
    m_z[:, 2, 2] = 1
    return torch.matmul(m_z, torch.matmul(m_y, m_x))


def get_transform_matrices(view):
    """utils.py:52-73: view [B,3|5|6] -> (R [B,3,3], t [B,1,3])."""
    b = view.size(0)
    if view.size(1) == 6:
        rx, ry, rz = view[:, 0], view[:, 1], view[:, 2]
        trans_xyz = view[:, 3:].reshape(b, 1, 3)
    elif view.size(1) == 5:
        rx, ry, rz = view[:, 0], view[:, 1], view[:, 2]
        delta_xy = view[:, 3:].reshape(b, 1, 2)
        trans_xyz = torch.cat([delta_xy, torch.zeros(b, 1, 1, device=view.device, dtype=view.dtype)], 2)
    elif view.size(1) == 3:
        rx, ry, rz = view[:, 0], view[:, 1], view[:, 2]
        trans_xyz = torch.zeros(b, 1, 3, device=view.device, dtype=view.dtype)
    else:
        raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")
    return get_rotation_matrix(rx, ry, rz), trans_xyz


def get_face_idx(b, h, w):
    """utils.py:76-80 (compatibility only; the kernels use the closed form in csrc/g2s_math.cuh face_vertices)."""
    idx_map = torch.arange(h * w).reshape(h, w)
    faces1 = torch.stack([idx_map[:h - 1, :w - 1], idx_map[1:, :w - 1], idx_map[:h - 1, 1:]], -1).reshape(-1, 3)
    faces2 = torch.stack([idx_map[:h - 1, 1:], idx_map[1:, :w - 1], idx_map[1:, 1:]], -1).reshape(-1, 3)
    return torch.cat([faces1, faces2], 0).repeat(b, 1, 1).int()


# utils.py:83-95 cube coefficients: with them, trilinear sampling of the 2x2x2 texture on the simplex reproduces the
# barycentric blend of the three vertex colours (k_resolve_rgb evaluates exactly this in registers)
_CUBE = [[0.5, 0.5, 0.5], [0, 0, 1], [0, 1, 0], [-0.5, 0.5, 0.5], [1, 0, 0], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5],
         [0, 0, 0]]


def vcolor_to_texture_cube(vcolors):
    """utils.py:83-95 (compatibility only: the mesh-texture kernels never materialise the [B,F,2,2,2,C] cube)."""
    b, c, n, f = vcolors.shape
    coeffs = torch.tensor(_CUBE, dtype=vcolors.dtype, device=vcolors.device)
    return coeffs.matmul(vcolors.permute(0, 2, 3, 1)).reshape(b, n, 2, 2, 2, c)


def get_textures_from_im(im, tx_size=1):
    """utils.py:98-109 (compatibility only)."""
    b, c, h, w = im.shape
    if tx_size == 1:
        textures = torch.cat([im[:, :, :h - 1, :w - 1].reshape(b, c, -1), im[:, :, 1:, 1:].reshape(b, c, -1)], 2)
        return textures.transpose(2, 1).reshape(b, -1, 1, 1, 1, c)
    if tx_size == 2:
        t1 = torch.stack([im[:, :, :h - 1, :w - 1], im[:, :, :h - 1, 1:], im[:, :, 1:, :w - 1]], -1).reshape(b, c, -1, 3)
        t2 = torch.stack([im[:, :, 1:, :w - 1], im[:, :, :h - 1, 1:], im[:, :, 1:, 1:]], -1).reshape(b, c, -1, 3)
        return vcolor_to_texture_cube(torch.cat([t1, t2], 2))
    raise NotImplementedError("Currently support texture size of 1 or 2 only.")


def mm_normalize(x, min=0, max=1):
    """utils.py:4-10."""
    x_min = x.min()
    return (x - x_min) / (x.max() - x_min) * (max - min) + min


def rand_range(size, min, max):
    """utils.py:13-14."""
    return torch.rand(size) * (max - min) + min


def rand_posneg_range(size, min, max):
    """utils.py:17-19."""
    i = (torch.rand(size) > 0.5).type(torch.float) * 2. - 1.
    return i * rand_range(size, min, max)


def get_lighting_directions(lighting):
    """model.py:347-353: raw light [B,4] -> (ambient a [B,1], diffuse b [B,1], direction d [B,3])."""
    a = lighting[:, :1] / 2 + 0.5
    b = lighting[:, 1:2] / 2 + 0.5
    d = torch.cat([lighting[:, 2:], torch.ones(lighting.size(0), 1, device=lighting.device, dtype=lighting.dtype)], 1)
    d = d / ((d ** 2).sum(1, keepdim=True)) ** 0.5
    return a, b, d
