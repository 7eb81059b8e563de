"""Host-side helpers with the names and results of GAN2Shape/renderer/utils.py (reference file:line in each docstring).

None of this is on the hot path: the CUDA kernels derive pixel coordinates, the grid-mesh topology and the per-face
vertex colours from indices (csrc/g2s_math.cuh `face_vertices`, `k_resolve_rgb`), so the tensors these helpers build are
never materialised there, whereas the reference rebuilds them on the CPU and copies them to the device on every call
(renderer.py:76, 119, 194).  They exist so that code written against the reference's module keeps working, and as the
CPU-side statement of the conventions the kernels implement (tests/test_cabi_and_host.py compares them with the oracle).
"""
import torch


# ---- small numeric helpers (utils.py:4-19) -------------------------------------------------------------------------
def mm_normalize(x, min=0, max=1):
    """utils.py:4-10: affine map of x onto [min, max]."""
    lo, hi = x.min(), x.max()
    return (x - lo) / (hi - lo) * (max - min) + min


def rand_range(size, min, max):
    """utils.py:13-14: U(min, max)."""
    return min + (max - min) * torch.rand(size)


def rand_posneg_range(size, min, max):
    """utils.py:17-19: U(min, max) with a random sign."""
    sign = torch.where(torch.rand(size) > 0.5, 1.0, -1.0)
    return sign * rand_range(size, min, max)


# ---- pixel grid and mesh topology (utils.py:22-30, 76-80) -------------------------------------------------------------
def get_grid(b, H, W, normalize=True, device=None):
    """utils.py:22-30: [b,H,W,2] of (x, y) per pixel -- integer pixel coordinates, or [-1,1] when `normalize`."""
    ys = torch.linspace(-1, 1, H, device=device) if normalize else torch.arange(H, device=device)
    xs = torch.linspace(-1, 1, W, device=device) if normalize else torch.arange(W, device=device)
    grid = torch.empty(H, W, 2, dtype=torch.float32, device=device)
    grid[..., 0] = xs.float().view(1, W)
    grid[..., 1] = ys.float().view(H, 1)
    return grid.unsqueeze(0).repeat(b, 1, 1, 1)


def _quad_corners(h, w):
    """Vertex indices (top-left, top-right, bottom-left, bottom-right) of the (h-1)(w-1) quads, row-major."""
    tl = (torch.arange(h - 1).view(-1, 1) * w + torch.arange(w - 1).view(1, -1)).reshape(-1)
    return tl, tl + 1, tl + w, tl + w + 1


def get_face_idx(b, h, w):
    """utils.py:76-80: int32 [b, 2(h-1)(w-1), 3]; all upper-left triangles (tl, bl, tr) first, then all lower-right ones
    (tr, bl, br).  The kernels use the same closed form (csrc/g2s_math.cuh `face_vertices`)."""
    tl, tr, bl, br = _quad_corners(h, w)
    faces = torch.cat([torch.stack([tl, bl, tr], 1), torch.stack([tr, bl, br], 1)], 0)
    return faces.to(torch.int32).unsqueeze(0).repeat(b, 1, 1)


# ---- view -> rotation / translation (utils.py:33-73) ------------------------------------------------------------------
def _axis_rotation(angle, axis):
    """[n,3,3] rotation about one coordinate axis (0 = x, 1 = y, 2 = z) with the reference's sign convention."""
    c, s = angle.cos(), angle.sin()
    i, j = [(1, 2), (2, 0), (0, 1)][axis]          # the plane that turns: (y,z), (z,x), (x,y)
    m = torch.zeros(angle.shape[0], 3, 3, device=angle.device, dtype=angle.dtype)
    m[:, axis, axis] = 1
    m[:, i, i] = c
    m[:, j, j] = c
    m[:, i, j] = -s
    m[:, j, i] = s
    return m


def get_rotation_matrix(tx, ty, tz):
    """utils.py:33-49: R = Rz Ry Rx (two matmuls in that association, as the reference evaluates it)."""
    return torch.matmul(_axis_rotation(tz, 2), torch.matmul(_axis_rotation(ty, 1), _axis_rotation(tx, 0)))


def get_transform_matrices(view):
    """utils.py:52-73: view [B,3|5|6] = (rx, ry, rz[, dx, dy[, dz]]) -> (R [B,3,3], t [B,1,3]); other widths raise.
    CUDA views go through one kernel instead (functional.ViewToRtFn)."""
    n, width = view.shape
    if width not in (3, 5, 6):
        raise Exception("Unsupported view size. size(1) must be either 3, 5, 6.")
    t = torch.zeros(n, 1, 3, device=view.device, dtype=view.dtype)
    if width > 3:
        t = torch.cat([view[:, 3:].reshape(n, 1, width - 3), t[:, :, :6 - width]], 2)
    return get_rotation_matrix(view[:, 0], view[:, 1], view[:, 2]), t


# ---- vertex colours -> neural_renderer textures (utils.py:83-109) ------------------------------------------------------
def _cube_coefficients(dtype, device):
    """The 8 x 3 table of utils.py:84-93.  Corner (i,j,k) of the 2x2x2 texture cube, in index order 4i+2j+k, weighs the
    three vertex colours by: one half each for (0,0,0); the indicator s = (i,j,k) itself when one bit is set; s - 1/2
    when two are; nothing for (1,1,1).  With these values trilinear sampling on the simplex reproduces the barycentric
    blend of the three colours (what k_resolve_rgb evaluates in registers)."""
    rows = []
    for corner in range(8):
        s = [float((corner >> 2) & 1), float((corner >> 1) & 1), float(corner & 1)]
        ones = int(sum(s))
        rows.append([[0.5, 0.5, 0.5], s, [v - 0.5 for v in s], [0.0, 0.0, 0.0]][ones])
    return torch.tensor(rows, dtype=dtype, device=device)


def vcolor_to_texture_cube(vcolors):
    """utils.py:83-95: vertex colours [b,c,n,3] -> texture cubes [b,n,2,2,2,c] (compatibility only: the mesh-texture
    kernels never materialise them)."""
    b, c, n, _ = vcolors.shape
    cube = _cube_coefficients(vcolors.dtype, vcolors.device)                 # [8,3]
    return torch.matmul(cube, vcolors.permute(0, 2, 3, 1)).reshape(b, n, 2, 2, 2, c)


def get_textures_from_im(im, tx_size=1):
    """utils.py:98-109: per-face textures of the grid mesh from an image [b,c,h,w] (compatibility only).
    tx_size 1: one colour per face (top-left pixel for the upper-left triangles, bottom-right for the others);
    tx_size 2: the texture cube of the three corner colours, in the reference's corner order (tl, tr, bl) / (bl, tr, br)."""
    b, c, h, w = im.shape
    tl, tr, bl, br = _quad_corners(h, w)
    flat = im.reshape(b, c, h * w)
    if tx_size == 1:
        per_face = flat[:, :, torch.cat([tl, br]).to(im.device)]             # [b,c,F]
        return per_face.permute(0, 2, 1).reshape(b, -1, 1, 1, 1, c)
    if tx_size == 2:
        corners = torch.cat([torch.stack([tl, tr, bl], 1), torch.stack([bl, tr, br], 1)], 0).to(im.device)   # [F,3]
        return vcolor_to_texture_cube(flat[:, :, corners])
    raise NotImplementedError("Currently support texture size of 1 or 2 only.")


# ---- light (GAN2Shape/model.py:347-353; the reference keeps it in the model class) -------------------------------------
def get_lighting_directions(lighting):
    """model.py:347-353: raw light [B,4] -> (ambient a [B,1], diffuse b [B,1], unit direction d [B,3])."""
    a, b = lighting[:, 0:1] / 2 + 0.5, lighting[:, 1:2] / 2 + 0.5
    d = torch.nn.functional.pad(lighting[:, 2:], (0, 1), value=1.0)
    return a, b, d / ((d ** 2).sum(1, keepdim=True)) ** 0.5
