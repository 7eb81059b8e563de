"""gan-2d-to-3d_b200 -- B200-native (sm_100a) implementation of GAN2Shape's differentiable depth-map renderer path
(reference: GAN2Shape/renderer).  The directory name is not a Python identifier; import it through the
repo-root shim:  `import g2s_b200`  (g2s_b200.py loads this package under that name).

Public surface (same names as the reference's GAN2Shape.renderer):
    Renderer, get_grid, get_rotation_matrix, get_transform_matrices, get_face_idx
plus the callers either side of the path (`callers`: get_clamped_depth, get_shading, PhotometricLoss with the
validity mask, SmoothLoss -- CUDA kernels behind the same C ABI), get_lighting_directions, and the autograd Functions in
`functional`.
"""
from .utils import (get_grid, get_rotation_matrix, get_transform_matrices, get_face_idx, get_lighting_directions,
                    get_textures_from_im, vcolor_to_texture_cube, mm_normalize, rand_range, rand_posneg_range)
from .callers import get_shading, get_clamped_depth, recon_im_mask, PhotometricLoss, SmoothLoss
from .renderer import Renderer, EPS
from . import functional, callers, graphs, hostio, mesh_export, nr_compat, synthetic, sharding, build as _build  # noqa: F401

__all__ = ["Renderer", "get_grid", "get_rotation_matrix", "get_transform_matrices", "get_face_idx",
           "get_lighting_directions", "get_shading", "get_textures_from_im", "vcolor_to_texture_cube", "mm_normalize",
           "rand_range", "rand_posneg_range", "get_clamped_depth", "recon_im_mask", "PhotometricLoss", "SmoothLoss",
           "callers", "hostio", "mesh_export", "nr_compat", "functional", "graphs", "sharding", "synthetic", "EPS"]
