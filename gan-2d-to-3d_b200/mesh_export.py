"""Mesh / surface export for the consumers of the renderer's outputs.

The reference's only consumers are its plotting helpers (/root/reference/plotting.py:58-131): `plotly_3d_animate(recon_depth,
texture)` draws `go.Surface(z=-1 * recon_depth[0], surfacecolor=texture[0, 0])`, `plt_3d_depth` a matplotlib surface of the
same array.  `surface_arrays` returns exactly those arrays; `depth_to_mesh` turns a depth map (optionally warped by a view, as
warp_canon_depth sees it) into the triangle mesh the rasteriser renders -- vertices from the CUDA `depth_to_3d_grid` /
`get_warped_3d_grid` operators, faces from `get_face_idx` (utils.py:76-80), per-vertex colours from an image -- and
`write_obj` / `write_ply` store it for external viewers (`go.Mesh3d(x, y, z, i, j, k, vertexcolor)` takes the same arrays).
"""
import numpy as np
import torch

from .utils import get_face_idx


def surface_arrays(recon_depth, texture=None):
    """(z, surfacecolor) as plotting.py:58-63 builds them: z = -recon_depth[0] [H,W], surfacecolor = texture[0,0] or None."""
    z = -1.0 * recon_depth[0].detach().float().cpu().numpy()
    color = None if texture is None else texture[0, 0].detach().float().cpu().numpy()
    return z, color


def depth_to_mesh(renderer, depth, image=None, view=None):
    """depth [1|B,S,S] (CUDA) -> list of meshes, one per batch item: dict(vertices float32 [S*S,3], faces int32 [2(S-1)^2,3],
    colors float32 [S*S,C] in [0,1] or None).  With `view` [B,6] the vertices are the warped 3-D grid the rasteriser
    receives (renderer.py:90-95), otherwise the canonical grid (renderer.py:74-80).  `image` [1|B,C,S,S] in [-1,1] gives the
    per-vertex colours (the same image get_textures_from_im would turn into texture cubes)."""
    B, H, W = depth.shape
    with torch.no_grad():
        if view is not None:
            renderer.set_transform_matrices(view)
            grid = renderer.get_warped_3d_grid(depth)
        else:
            grid = renderer.depth_to_3d_grid(depth)
    verts = grid.reshape(B, H * W, 3).float().cpu().numpy()
    faces = get_face_idx(1, H, W)[0].numpy().astype(np.int32)
    out = []
    for b in range(B):
        colors = None
        if image is not None:
            im = image[b if image.shape[0] > 1 else 0].detach().float().cpu()
            colors = (im.reshape(im.shape[0], -1).T.numpy() * 0.5 + 0.5).clip(0.0, 1.0).astype(np.float32)
        out.append(dict(vertices=verts[b], faces=faces, colors=colors))
    return out


def write_obj(path, mesh):
    """Wavefront OBJ with per-vertex colours as the common `v x y z r g b` extension (1-based face indices)."""
    v, f, c = mesh["vertices"], mesh["faces"], mesh.get("colors")
    with open(path, "w") as fh:
        fh.write("# grid mesh of the depth-map renderer: %d vertices, %d faces\n" % (len(v), len(f)))
        for i in range(len(v)):
            if c is not None and c.shape[1] >= 3:
                fh.write("v %.7g %.7g %.7g %.4f %.4f %.4f\n" % (v[i, 0], v[i, 1], v[i, 2], c[i, 0], c[i, 1], c[i, 2]))
            else:
                fh.write("v %.7g %.7g %.7g\n" % (v[i, 0], v[i, 1], v[i, 2]))
        for a, b, d in f:
            fh.write("f %d %d %d\n" % (a + 1, b + 1, d + 1))


def write_ply(path, mesh):
    """Binary little-endian PLY (vertices float32, optional uchar colours, faces as int32 triples)."""
    v, f, c = mesh["vertices"].astype("<f4"), mesh["faces"].astype("<i4"), mesh.get("colors")
    has_c = c is not None and c.shape[1] >= 3
    header = ["ply", "format binary_little_endian 1.0", "element vertex %d" % len(v), "property float x", "property float y",
              "property float z"]
    if has_c:
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    header += ["element face %d" % len(f), "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as fh:
        fh.write(("\n".join(header) + "\n").encode())
        if has_c:
            rec = np.zeros(len(v), dtype=[("p", "<f4", 3), ("c", "u1", 3)])
            rec["p"] = v
            rec["c"] = (c[:, :3] * 255.0 + 0.5).astype(np.uint8)
            fh.write(rec.tobytes())
        else:
            fh.write(v.tobytes())
        frec = np.zeros(len(f), dtype=[("n", "u1"), ("i", "<i4", 3)])
        frec["n"] = 3
        frec["i"] = f
        fh.write(frec.tobytes())


def mesh3d_arrays(mesh):
    """keyword arrays for plotly's go.Mesh3d: x, y, z, i, j, k (+ vertexcolor)."""
    v, f, c = mesh["vertices"], mesh["faces"], mesh.get("colors")
    kw = dict(x=v[:, 0], y=v[:, 1], z=v[:, 2], i=f[:, 0], j=f[:, 1], k=f[:, 2])
    if c is not None:
        kw["vertexcolor"] = c
    return kw
