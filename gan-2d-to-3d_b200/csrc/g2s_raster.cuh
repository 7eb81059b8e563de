// g2s_raster.cuh -- device-side building blocks of the grid-mesh rasteriser (sm_100a).
//
// Replaces neural_renderer's forward_face_index_map kernels (brute force: every sub-pixel loops over
// every face) with a grid-aware splat: the mesh is the regular grid of utils.py:76-80, so a CTA owns a
// TILE x TILE block of quads, projects its (TILE+1)^2 vertices once into shared memory, and every
// thread rasterises the (up to) four windings of one quad over their sub-pixel bounding boxes into a
// packed-key z-buffer with 64-bit atomicMin.  Faces whose bounding box is large (the depth-step walls
// at the image border stretch to >100 sub-pixels under yaw) are queued in shared memory and
// rasterised by whole warps afterwards, so one thread never walks a long box alone.
#pragma once
#include "g2s_math.cuh"

namespace g2s {

#ifndef G2S_TILE_H
#define G2S_TILE_H 16
#endif
constexpr int TILE = 16;             // quads per tile row
constexpr int TILE_H = G2S_TILE_H;   // quad rows per tile
constexpr int TV = TILE + 1;         // vertices per tile row
constexpr int TVH = TILE_H + 1;      // vertex rows per tile
constexpr int SPLAT_THREADS = TILE * TILE_H;   // one thread per quad
constexpr int SMALL_BOX = 12;       // boxes up to this many sub-pixels are walked by the owning thread

G2S_HD int imax(int a, int b) { return a > b ? a : b; }
G2S_HD int imin(int a, int b) { return a < b ? a : b; }

struct BBox {
    int x0, x1, y0, y1;
};

// Sub-pixel bounding box (nr-native coordinates: y up) grown by 1/64 px and clipped to the image.
// Returns false for empty boxes and for faces with a non-finite coordinate (those can never win the
// z test: their weights are NaN).
G2S_HD bool tri_bbox(const Tri& f, int is, BBox& bb) {
    const float px0 = ndc_to_pix(f.x0, is), px1 = ndc_to_pix(f.x1, is), px2 = ndc_to_pix(f.x2, is);
    const float py0 = ndc_to_pix(f.y0, is), py1 = ndc_to_pix(f.y1, is), py2 = ndc_to_pix(f.y2, is);
    const float chk = px0 + px1 + px2 + py0 + py1 + py2;
    if (!(fabsf(chk) < 3.0e38f)) return false;
    const float m = 1.0f / 64.0f;
    const float lim = (float)is;
    const float xmin = fmaxf(fminf(px0, fminf(px1, px2)) - m, -1.0f);
    const float xmax = fminf(fmaxf(px0, fmaxf(px1, px2)) + m, lim);
    const float ymin = fmaxf(fminf(py0, fminf(py1, py2)) - m, -1.0f);
    const float ymax = fminf(fmaxf(py0, fmaxf(py1, py2)) + m, lim);
    bb.x0 = imax(0, (int)ceilf(xmin));
    bb.x1 = imin(is - 1, (int)floorf(xmax));
    bb.y0 = imax(0, (int)ceilf(ymin));
    bb.y1 = imin(is - 1, (int)floorf(ymax));
    return bb.x0 <= bb.x1 && bb.y0 <= bb.y1;
}

// The four windings of quad (qy,qx) of a tile whose projected vertices sit in shared memory `sv`
// ([TV*TV][3] = u, v, z).  w: 0 = faces1, 1 = faces2, 2/3 = their fill_back copies (reversed order).
G2S_HD Tri tile_winding(const float* sv, int qy, int qx, int w) {
    const float* a = &sv[(qy * TV + qx) * 3];
    const float* b = &sv[((qy + 1) * TV + qx) * 3];
    const float* c = &sv[(qy * TV + qx + 1) * 3];
    const float* d = &sv[((qy + 1) * TV + qx + 1) * 3];
    switch (w) {
        case 0: return make_tri(a, b, c);
        case 1: return make_tri(c, b, d);
        case 2: return make_tri(c, b, a);
        default: return make_tri(d, b, c);
    }
}

// Evaluate one (face, sub-pixel) pair exactly as [nr] kernel 2 does; lazily builds face_inv.
G2S_HD bool tri_sample(const Tri& f, float* fi, bool& have_fi, int xi, int yi, float xp,
                                           float yp, int is, float near, float far, float w[3], float* zp) {
    if (!tri_contains(f, xp, yp)) return false;
    if (!have_fi) {
        tri_face_inv(f, is, fi);
        have_fi = true;
    }
    return tri_weights_depth(f, fi, xi, yi, near, far, w, zp);
}

}  // namespace g2s
