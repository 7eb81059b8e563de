"""An INDEPENDENT float64 reference for the rasteriser: a 3-D ray caster over the grid mesh (test infrastructure).

neural_renderer (and oracle/nr_raster.c, its restatement) rasterise in 2-D: project the vertices, test sub-pixel centres
against the projected edges, interpolate 1/z with barycentric weights taken from a 3x3 inverse.  This module answers the same
question from the geometric definition instead, sharing no formula with that code path: for every sub-pixel centre cast the
camera ray through it, intersect it with every triangle of the mesh in 3-D camera space (Moeller-Trumbore), and keep the
nearest hit with near < z < far.  For a pinhole camera the perspective-correct depth nr interpolates IS the depth of that
intersection, the covering triangle IS the one the ray hits first, and nr's fill_back rule (the reversed copy of a face has
index f + 2(S-1)^2 and is the one that survives the back-face cull) maps to "was the triangle hit from behind".

What is compared and what is exempt: a sub-pixel is UNAMBIGUOUS when its ray hits its nearest triangle with every
barycentric coordinate > `edge_margin` (it is not within rounding of an edge, where nr's fp32 edge functions decide) and the
runner-up hit is more than `gap_margin` (relative) behind.  On unambiguous sub-pixels the face index must be equal and the
depth must agree to fp32 accuracy; the ambiguous ones are counted and reported.

Conventions: camera looks down +z, pixel (u, v) = K (x/z, y/z, 1); neural_renderer flips v to y-up NDC
and rasterises at is = 2S; in IMAGE orientation (row 0 = top) sub-pixel (r, c) has its centre at pixel coordinates
(u, v) = ((c + 0.5) / 2, (r + 0.5) / 2).
"""
import numpy as np


def grid_faces(S):
    """utils.py:76-80: faces1 = [i(y,x), i(y+1,x), i(y,x+1)], faces2 = [i(y,x+1), i(y+1,x), i(y+1,x+1)] -> [2(S-1)^2, 3]"""
    idx = np.arange(S * S).reshape(S, S)
    f1 = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:]], -1).reshape(-1, 3)
    f2 = np.stack([idx[:-1, 1:], idx[1:, :-1], idx[1:, 1:]], -1).reshape(-1, 3)
    return np.concatenate([f1, f2], 0)


def raycast(verts, S, K, near, far, rows_per_chunk=8):
    """verts [S*S, 3] camera-space vertices of the grid mesh (what nr.render_depth / render_rgb receive), K [3,3].
    Returns a dict of [2S, 2S] maps in image orientation: face (nr numbering, -1 = background), z, bary [2S,2S,3] (weights
    of the face's vertices in nr's vertex order of the ORIGINAL winding), tri (0..2(S-1)^2-1 or -1), edge (smallest
    barycentric coordinate of the winning hit), gap (relative depth gap to the runner-up hit, inf if none), graze (closest
    near-miss of any triangle), orient (relative signed-area magnitude of the winner), nhit."""
    V = np.asarray(verts, np.float64).reshape(-1, 3)
    K = np.asarray(K, np.float64).reshape(3, 3)
    F = grid_faces(S)
    nf, is_ = F.shape[0], 2 * S
    A, B, C = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    E1, E2 = B - A, C - A
    # orientation as neural_renderer sees it: NDC x right, y UP; back-facing <=> (y2-y0)(x1-x0) < (y1-y0)(x2-x0)
    P = V / V[:, 2:3]
    u = K[0, 0] * P[:, 0] + K[0, 1] * P[:, 1] + K[0, 2]
    v = K[1, 0] * P[:, 0] + K[1, 1] * P[:, 1] + K[1, 2]
    xn, yn = 2 * (u - S / 2) / S, 2 * ((S - v) - S / 2) / S
    x0, x1, x2 = xn[F[:, 0]], xn[F[:, 1]], xn[F[:, 2]]
    y0, y1, y2 = yn[F[:, 0]], yn[F[:, 1]], yn[F[:, 2]]
    cross = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)          # >= 0: the original winding faces the camera
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    out = dict(face=np.full((is_, is_), -1, np.int64), tri=np.full((is_, is_), -1, np.int64),
               z=np.full((is_, is_), np.inf), bary=np.zeros((is_, is_, 3)), edge=np.zeros((is_, is_)),
               gap=np.full((is_, is_), np.inf), nhit=np.zeros((is_, is_), np.int64),
               graze=np.full((is_, is_), np.inf), orient=np.full((is_, is_), np.inf))
    cols = (np.arange(is_) + 0.5) / 2
    for r0 in range(0, is_, rows_per_chunk):
        rows = (np.arange(r0, min(r0 + rows_per_chunk, is_)) + 0.5) / 2
        vv, uu = np.meshgrid(rows, cols, indexing="ij")
        D = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1).reshape(-1, 1, 3)      # [n,1,3]
        # Moeller-Trumbore, ray origin 0: t D = A + a E1 + b E2
        Pv = np.cross(D, E2[None])                                 # [n,nf,3]
        det = (Pv * E1[None]).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            T = -A[None]                                           # origin - A
            a = (Pv * T).sum(-1) * inv
            Qv = np.cross(np.broadcast_to(T, Pv.shape), E1[None])
            b = (Qv * D).sum(-1) * inv
            t = (Qv * E2[None]).sum(-1) * inv
        w0 = 1.0 - a - b
        mn = np.minimum(np.minimum(w0, a), b)
        hit = (np.abs(det) > 1e-300) & (mn >= 0) & (t > near) & (t < far) & np.isfinite(t)
        tz = np.where(hit, t, np.inf)
        order = np.argsort(tz, axis=1)[:, :2]
        n = tz.shape[0]
        best, second = order[:, 0], order[:, 1]
        zb, zs = tz[np.arange(n), best], tz[np.arange(n), second]
        ok = np.isfinite(zb)
        rr = slice(r0, min(r0 + rows_per_chunk, is_))
        shp = (rows.shape[0], is_)
        tri = np.where(ok, best, -1)
        back = cross[best] < 0
        out["tri"][rr] = tri.reshape(shp)
        out["face"][rr] = np.where(ok, best + np.where(back, nf, 0), -1).reshape(shp)
        out["z"][rr] = zb.reshape(shp)
        out["bary"][rr] = np.stack([w0[np.arange(n), best], a[np.arange(n), best], b[np.arange(n), best]], -1).reshape(shp + (3,))
        out["edge"][rr] = np.where(ok, mn[np.arange(n), best], 0.0).reshape(shp)
        with np.errstate(invalid="ignore"):
            out["gap"][rr] = np.where(np.isfinite(zs), (zs - zb) / zb, np.inf).reshape(shp)
        out["nhit"][rr] = hit.sum(1).reshape(shp)
        # the closest NEAR-MISS: smallest distance (in barycentric units) by which the ray misses any triangle whose plane it
        # meets in range -- a ray that grazes an edge is decided by nr's fp32 edge functions, not by geometry
        graze = np.where((np.abs(det) > 1e-300) & (t > near) & (t < far) & (mn < 0), -mn, np.inf).min(1)
        out["graze"][rr] = graze.reshape(shp)
        out["orient"][rr] = np.where(ok, np.abs(cross[best]) / (np.abs(cross).max() + 1e-300), np.inf).reshape(shp)
    return out


def unambiguous(rc, edge_margin=1e-4, gap_margin=1e-4):
    """sub-pixels whose winner does not depend on rounding: the nearest hit is strictly inside its triangle and clear of the
    runner-up; background rays must not graze any triangle edge"""
    covered = rc["tri"] >= 0
    clear = rc["graze"] > edge_margin
    return clear & (~covered | ((rc["edge"] > edge_margin) & (rc["gap"] > gap_margin) & (rc["orient"] > 1e-9)))


def depth_gradient_f64(verts_ndc, face_map, grad_depth_map, S):
    """neural_renderer's backward_depth_map evaluated independently in float64 from the projected
    vertices: for every covered sub-pixel of face f with barycentric weights w and depth zp,
        d zp / d z_k     = w_k zp^2 / z_k^2
        d zp / d (x,y)_k = -tmp_l w_k zp^2 is / 2,   tmp_l = -sum_m face_inv[m][l] / z_m
    verts_ndc [S*S,3] (NDC x, y-up, camera z), face_map [is,is] in nr's NATIVE row order (row 0 = bottom) with nr face
    numbering, grad_depth_map [is,is] native.  Returns grad_faces [4(S-1)^2, 3, 3]."""
    Vn = np.asarray(verts_ndc, np.float64)
    F = grid_faces(S)
    nf, is_ = F.shape[0], 2 * S
    F4 = np.concatenate([F, F[:, ::-1]], 0)
    out = np.zeros((2 * nf, 3, 3))
    ys, xs = np.nonzero(face_map >= 0)
    for yi, xi in zip(ys, xs):
        f = face_map[yi, xi]
        tri = Vn[F4[f]]                                            # [3, (x, y, z)]
        p = 0.5 * (tri[:, :2] * is_ + is_ - 1)                     # sub-pixel coordinates
        M = np.stack([p[:, 0], p[:, 1], np.ones(3)], 0)            # columns = vertices
        fi = np.linalg.inv(M)                                      # rows k: w_k = fi[k] . (xi, yi, 1)
        w = fi @ np.array([xi, yi, 1.0])
        w = np.clip(w, 0, 1)
        w = w / w.sum()
        z = tri[:, 2]
        zp = 1.0 / (w / z).sum()
        g = grad_depth_map[yi, xi]
        tmp = -(fi[:, :2] / z[:, None]).sum(0)
        for k in range(3):
            out[f, k, 2] += g * w[k] * zp * zp / (z[k] * z[k])
            out[f, k, 0] += -g * tmp[0] * w[k] * zp * zp * is_ / 2
            out[f, k, 1] += -g * tmp[1] * w[k] * zp * zp * is_ / 2
    return out
