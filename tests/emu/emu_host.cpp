// tests/emu/emu_host.cpp -- TEST HARNESS, NOT PRODUCT CODE.
//
// Compiles the product's per-element arithmetic headers (gan-2d-to-3d_b200/csrc/g2s_math.cuh,
// g2s_raster.cuh) for the HOST with g++ -ffp-contract=off and runs them in plain serial loops, so the
// arithmetic contract (vertex transform, projection, inside test, weights, z, bounding boxes, key order)
// can be pinned bit-for-bit against the oracle in a container that has no GPU.  It mirrors the
// structure of k_splat / k_resolve / k_resolve_rgb but is never imported, linked or called by the package;
// the kernels themselves are checked against the oracle on the GPU (tests -m gpu).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../gan-2d-to-3d_b200/csrc/g2s_raster.cuh"

using namespace g2s;

extern "C" {

struct EmuCam {
    float K[9], invK[9];
    float rcd, near_z, far_z, clamp_lo, clamp_hi;
    int S;
};

static Cam to_cam(const EmuCam* e) {
    Cam c;
    memcpy(c.K, e->K, sizeof c.K);
    memcpy(c.Kg, e->K, sizeof c.Kg);
    memcpy(c.invK, e->invK, sizeof c.invK);
    c.rcd = e->rcd; c.os = (float)e->S; c.half_os = (float)(e->S / 2.0);
    c.near = e->near_z; c.far = e->far_z; c.clamp_lo = e->clamp_lo; c.clamp_hi = e->clamp_hi; c.S = e->S;
    return c;
}

// projected NDC vertices of one view: [S*S,3]
__attribute__((visibility("default"))) void emu_project(const EmuCam* ec, const float* depth, const float* R,
                                                          const float* t, float* ndc) {
    const Cam cam = to_cam(ec);
    const int S = cam.S;
    for (int y = 0; y < S; y++)
        for (int x = 0; x < S; x++) {
            float ray[3], q[3];
            pixel_ray(cam, x, y, ray);
            warp_point(cam, R, t, ray, depth[y * S + x], q);
            project_ndc(cam, q, &ndc[(y * S + x) * 3]);
        }
}

// splat + resolve of one view from projected vertices; face_idx [2S,2S] image orientation, recon [S,S]
__attribute__((visibility("default"))) void emu_raster(const EmuCam* ec, const float* ndc, int32_t* face_idx,
                                                         float* recon, float* zsub) {
    const Cam cam = to_cam(ec);
    const int S = cam.S, is = 2 * S, Q = (S - 1) * (S - 1);
    unsigned long long* zb = (unsigned long long*)malloc(sizeof(unsigned long long) * is * is);
    for (int i = 0; i < is * is; i++) zb[i] = zkey_empty(cam.far);
    for (int qy = 0; qy < S - 1; qy++)
        for (int qx = 0; qx < S - 1; qx++) {
            const float* a = &ndc[(qy * S + qx) * 3];
            const float* b = &ndc[((qy + 1) * S + qx) * 3];
            const float* c = &ndc[(qy * S + qx + 1) * 3];
            const float* d = &ndc[((qy + 1) * S + qx + 1) * 3];
            for (int w = 0; w < 4; w++) {
                const Tri f = w == 0 ? make_tri(a, b, c) : w == 1 ? make_tri(c, b, d) : w == 2 ? make_tri(c, b, a)
                                                                                                : make_tri(d, b, c);
                if (tri_is_back(f)) continue;
                BBox bb;
                if (!tri_bbox(f, is, bb)) continue;
                float fi[9];
                bool have_fi = false;
                const uint32_t face = (uint32_t)(w * Q + qy * (S - 1) + qx);
                for (int yi = bb.y0; yi <= bb.y1; yi++)
                    for (int xi = bb.x0; xi <= bb.x1; xi++) {
                        float wv[3], zp;
                        if (!tri_sample(f, fi, have_fi, xi, yi, pix_center_ndc(xi, is), pix_center_ndc(yi, is), is,
                                        cam.near, cam.far, wv, &zp))
                            continue;
                        const unsigned long long key = zkey_pack(zp, face);
                        unsigned long long* slot = &zb[(long)(is - 1 - yi) * is + xi];
                        if (key < *slot) *slot = key;
                    }
            }
        }
    for (int i = 0; i < is * is; i++) {
        face_idx[i] = zkey_face(zb[i]);
        if (zsub) zsub[i] = zkey_depth(zb[i]);
    }
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++) {
            const float s = add(add(add(zkey_depth(zb[(2 * i) * is + 2 * j]), zkey_depth(zb[(2 * i) * is + 2 * j + 1])),
                                    zkey_depth(zb[(2 * i + 1) * is + 2 * j])),
                                zkey_depth(zb[(2 * i + 1) * is + 2 * j + 1]));
            recon[i * S + j] = fminf(fmaxf(mul(s, 0.25f), cam.clamp_lo), cam.clamp_hi);
        }
    free(zb);
}

// grids: inverse / forward warped 2-D grid of one view, [H,W,2]
__attribute__((visibility("default"))) void emu_warp_grid(const EmuCam* ec, const float* depth, const float* R,
                                                            const float* t, int H, int W, int inverse, float* grid) {
    const Cam cam = to_cam(ec);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float ray[3], q[3], v[3];
            pixel_ray(cam, x, y, ray);
            if (inverse) inv_warp_point(cam, R, t, ray, depth[y * W + x], q, v);
            else warp_point(cam, R, t, ray, depth[y * W + x], q);
            point_to_grid(cam, q, W, H, &grid[(y * W + x) * 2]);
        }
}

__attribute__((visibility("default"))) void emu_face_vertices(int f, int S, int* vidx) { face_vertices(f, S, vidx); }

}  // extern "C"
