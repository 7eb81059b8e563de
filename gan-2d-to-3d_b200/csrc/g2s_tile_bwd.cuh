// g2s_tile_bwd.cuh -- backward of the rasteriser (neural_renderer's backward_depth_map chained through vertices_to_faces,
// the projection and rotate_pts / translate_pts, GAN2Shape/renderer/renderer.py:90-95, 116-125), FACE-centric and fused.
//
// Round 1 ran three kernels over ~1 GB of scratch: k_project_verts (all projected vertices -> HBM), k_raster_bwd_px (one thread
// per output pixel: rebuild the 3x3 inverse of every distinct face of its 2x2 sub-pixels -- a face was rebuilt by every pixel
// it touched, at 19 active lanes per instruction) and k_vertex_bwd (re-read the per-view vertex gradients).  Here the
// backward mirrors the forward's two stages and keeps everything on chip:
//   k_raster_bwd_tile: one CTA per 16 x 16 block of quads of one view.  Project the tile's vertices into shared memory (the
//     forward's code, same bits), classify the quads exactly as the forward did, build each front face's 3x3 inverse ONCE,
//     find the sub-pixels a face owns by comparing the saved face-index map over the quad's 4 x 4 box jobs (no inside tests:
//     the map already names the owner), evaluate weights / depth of the owned sub-pixels with the forward's exact arithmetic
//     (balanced over the warp through the hit queue) into per-face sums A_m = sum g zp^2 w_m, turn those into the (u, v, z)
//     gradients of the face's vertices ([nr] backward_depth_map) accumulated per tile vertex in shared
//     memory, and push every tile vertex through projection / rotation to grad_depth, grad_R, grad_t before leaving.
//   k_raster_bwd_big: the faces stage 1 defers (long walls, degenerate quads), from a work list, with the row-task machinery of
//     g2s_bigface.cuh and the same per-face evaluation.
// Scratch: the masked quarter gradient g_sub [views, S, S] and the work list; nothing else touches HBM.
#pragma once
#include "g2s_bigface.cuh"

#ifndef G2S_BPHASE
#define G2S_BPHASE 9
#endif

namespace g2s {

struct TileSmemB {
    TileSmem2 t;
    float part[TWARPS][32 * 6];    // per (job lane, triangle) partial sums of g zp^2 w_m of one scan pass
};

// [nr] backward_depth_map for one face from its sums A_m: gradient of vertex m = (-t0 A_m is/2, -t1 A_m is/2, A_m / z_m^2),
// t_l = -sum_m fi[m][l] / z_m.  The six quotients are exact (the x / y gradient is a cancelling sum: the oracle's own
// rounding is part of the answer at the 1e-5 bar, profiles/r01_notes.md).
__device__ __forceinline__ void face_vertex_grads(const float fi[9], const float z[3], bool zbad, const float A[3], float hs,
                                                  float gv[3][3]) {
    float qd[6];
    unsigned bad = 0;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const float yz = rcp_seed(z[m]);
        qd[m] = div_core(fi[3 * m], z[m], yz);
        qd[3 + m] = div_core(fi[3 * m + 1], z[m], yz);
        bad = max(bad, fi[3 * m] == 0.0f ? 0u : range_key(fi[3 * m]));
        bad = max(bad, fi[3 * m + 1] == 0.0f ? 0u : range_key(fi[3 * m + 1]));
    }
    if (bad >= RANGE_SPAN || zbad) {
#pragma unroll
        for (int m = 0; m < 3; m++) { qd[m] = __fdiv_rn(fi[3 * m], z[m]); qd[3 + m] = __fdiv_rn(fi[3 * m + 1], z[m]); }
    }
    const float t0 = -(qd[0] + qd[1] + qd[2]);
    const float t1 = -(qd[3] + qd[4] + qd[5]);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        gv[m][0] = -t0 * A[m] * hs;
        gv[m][1] = -t1 * A[m] * hs;
        gv[m][2] = __fdiv_rn(A[m], z[m] * z[m]);
    }
}

// (u, v, z) NDC gradient of vertex `v` of view `b` -> grad_depth / grad_R | grad_t (acc[12]), or grad_vertices when the
// mesh came as 3-D points ([nr] projection backward + rotate_pts / translate_pts backward)
template <bool FROM_VERTS>
__device__ __forceinline__ void vertex_chain(const Cam& cam, const float* __restrict__ depth_img, const float* __restrict__ Rt,
                                             const float* __restrict__ verts_b, int v, float gu, float gvv, float gz,
                                             float* __restrict__ grad_depth_img, float* __restrict__ grad_verts_b,
                                             float acc[12]) {
    const int S = cam.S;
    if (FROM_VERTS) {
        const float* p = verts_b + (long)v * 3;
        const float zz = __ldg(p + 2) + 1e-9f, iz = 1.0f / zz;
        const float x_ = __ldg(p) * iz, y_ = __ldg(p + 1) * iz;
        const float gup = gu * (2.0f / cam.os), gvp = -gvv * (2.0f / cam.os);
        const float gx_ = gup * cam.K[0] + gvp * cam.K[3], gy_ = gup * cam.K[1] + gvp * cam.K[4];
        float* o = grad_verts_b + (long)v * 3;
        atomicAdd(o, gx_ * iz);
        atomicAdd(o + 1, gy_ * iz);
        atomicAdd(o + 2, gz - (gx_ * x_ + gy_ * y_) * iz);
    } else {
        const int vy = v / S, vx = v - vy * S;
        float ray[3], q[3];
        pixel_ray(cam, vx, vy, ray);
        const float d = __ldg(&depth_img[v]);
        warp_point(cam, Rt, Rt + 9, ray, d, q);
        const float p3[3] = {ray[0] * d, ray[1] * d, ray[2] * d - cam.rcd};
        const float zz = q[2] + 1e-9f, iz = 1.0f / zz;
        const float x_ = q[0] * iz, y_ = q[1] * iz;
        const float gup = gu * (2.0f / cam.os), gvp = -gvv * (2.0f / cam.os);
        const float gx_ = gup * cam.K[0] + gvp * cam.K[3], gy_ = gup * cam.K[1] + gvp * cam.K[4];
        const float gq[3] = {gx_ * iz, gy_ * iz, gz - (gx_ * x_ + gy_ * y_) * iz};
        float gd = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float gvk = gq[0] * Rt[k] + gq[1] * Rt[3 + k] + gq[2] * Rt[6 + k];
            gd += gvk * ray[k];
#pragma unroll
            for (int jj = 0; jj < 3; jj++) acc[3 * jj + k] += gq[jj] * p3[k];
            acc[9 + k] += gq[k];
        }
        atomicAdd(&grad_depth_img[v], gd);
    }
}

// which sub-pixels of the 4 x 4 box at (bx0, by0) does the saved face-index map give to faces fa / fb?  (nw, nh) = columns /
// rows of the box inside the quad's own box.  Bits as scan_job.
__device__ __forceinline__ unsigned owned_job(const int* __restrict__ fmap, int is, int fa, int fb, int bx0, int by0, int nw,
                                              int nh) {
    int f[16];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const bool in = c < nw && r < nh;
            f[r * 4 + c] = in ? __ldg(&fmap[(long)(is - 1 - (by0 + r)) * is + bx0 + c]) : -1;
        }
    }
    unsigned m = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        m |= (f[k] == fa ? 1u : 0u) << k;
        m |= (f[k] == fb ? 1u : 0u) << (16 + k);
    }
    return m;
}

// Backward hit: s w_m = g zp^2 w_m of one owned sub-pixel, summed WITHOUT atomics: the queue lists the hits of one (job,
// triangle) on consecutive lanes, so a segmented warp scan over runs of equal (job lane, triangle) leaves each run's total on
// its last lane, which adds it to that job's partial sums (no two runs of one drain step share a job and a triangle).  The
// owner lanes collect their jobs' partial sums after the pass.  (Shared-memory float atomics are compare-and-swap loops on this
// architecture: three per hit made the first form of this kernel slower than the three kernels it replaced.)
struct BwdHit {
    const TileCtx* cx;
    const float* gsub;     // masked quarter gradient of this view [S,S]
    float* part;           // this warp's partial sums [32 job lanes][2 triangles][3]
    __device__ __forceinline__ void operator()(bool valid, int jl, int tri, int, int xi, int yi, const float fi[9],
                                               const float z[3], uint32_t fw) const {
        const int lane = threadIdx.x & 31;
        const float g = valid ? __ldg(&gsub[((cx->is - 1 - yi) >> 1) * cx->S + (xi >> 1)]) : 0.f;
        float v[3] = {0.f, 0.f, 0.f};
        if (g != 0.f) {
            float w[3], zp = 0.f;
            weights_depth_core(fi, z, (fw >> 31) != 0u, xi, yi, cx->near, cx->far, w, &zp);
            const float s = g * zp * zp;
#pragma unroll
            for (int m = 0; m < 3; m++) v[m] = s * w[m];
        }
        const int key = valid ? jl * 2 + tri : -1 - lane;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ks = __shfl_up_sync(0xffffffffu, key, d);
            const float a0 = __shfl_up_sync(0xffffffffu, v[0], d), a1 = __shfl_up_sync(0xffffffffu, v[1], d),
                        a2 = __shfl_up_sync(0xffffffffu, v[2], d);
            if (lane >= d && ks == key) { v[0] += a0; v[1] += a1; v[2] += a2; }
        }
        const int kn = __shfl_down_sync(0xffffffffu, key, 1);
        if (valid && (lane == 31 || kn != key)) {
            part[key * 3] += v[0]; part[key * 3 + 1] += v[1]; part[key * 3 + 2] += v[2];
        }
    }
};

template <bool FROM_VERTS>
__device__ __forceinline__ void raster_bwd_tile_body(TileSmemB& smb, const Cam& cam, const float* __restrict__ depth_img,
                                                     const float* __restrict__ verts_b, const float* __restrict__ R_b,
                                                     const float* __restrict__ t_b, const int* __restrict__ fmap,
                                                     const float* __restrict__ gsub, const WorkList& wl, int view_in_launch,
                                                     int ty0, int tx0, float* __restrict__ grad_depth_img,
                                                     float* __restrict__ grad_verts_b, float* __restrict__ grad_R_b,
                                                     float* __restrict__ grad_t_b) {
    TileSmem2& sm = smb.t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, S = cam.S, is = 2 * S;
    tile_project2<FROM_VERTS>(cam, depth_img, verts_b, R_b, t_b, ty0, tx0, sm);
    __syncthreads();
#if G2S_BPHASE < 2     // experiment builds: stop the backward tile kernel after a phase (per-phase timing)
    if (sm.vz[tid] == 1.2345e-30f) grad_depth_img[0] = 1.f;
    return;
#endif
    TileCtx cx;
    cx.zb = nullptr; cx.near = cam.near; cx.far = cam.far; cx.is = is; cx.S = S; cx.Q = (S - 1) * (S - 1);
    cx.ty0 = ty0; cx.tx0 = tx0;
    QuadGeom g;
    const int quad = tid, qy = tid >> 4, qx = tid & 15;
    quad_geometry(sm, is, quad, ty0 + qy < S - 1 && tx0 + qx < S - 1, g);
    defer_quads(wl, cx, g, g.cls == QC_DEFER, quad, view_in_launch);
    const bool active = g.cls == QC_SMALL || g.cls == QC_MEDIUM;
    const bool actA = active && (g.fronts & 3u) != 0u, actB = active && (g.fronts & 12u) != 0u;
    float A6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // this lane's two per-face sums A_m = sum g zp^2 w_m
    if (__any_sync(0xffffffffu, active)) {
        float* part = smb.part[warp];
        build_face_records(sm, cx, g, quad, actA, actB);
#if G2S_BPHASE < 3
        if (sm.ftab[warp][lane] == 1.2345e-30f) grad_depth_img[0] = 1.f;
        return;
#endif
        const int njx = active ? (g.uw + 3) >> 2 : 0, njy = active ? (g.uh + 3) >> 2 : 0, nj = njx * njy;
        int total;
        const int jbase = warp_excl_scan(nj, &total);
        uint8_t* jobs = sm.jobs[warp];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nj) jobs[jbase + k] = (uint8_t)(lane | ((k % njx) << 5) | ((k / njx) << 6));
        const uint32_t origin = (uint32_t)g.x0 | ((uint32_t)g.y0 << 12);
        const uint32_t extent = (uint32_t)g.uw | ((uint32_t)g.uh << 8);
        // the two face indices this lane's quad can own (-2: none)
        const int fq = (ty0 + qy) * (S - 1) + tx0 + qx;
        const int faceA = actA ? (((g.fronts & 1u) == 0u) ? 2 * cx.Q : 0) + fq : -2;
        const int faceB = actB ? (((g.fronts & 4u) == 0u) ? 3 * cx.Q : cx.Q) + fq : -2;
        __syncwarp();
        BwdHit hit;
        hit.cx = &cx; hit.gsub = gsub; hit.part = part;
#pragma unroll 1
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < total;
            const unsigned e = valid ? jobs[j] : 0u;
            const int owner = (int)(e & 31u), bx = (int)((e >> 5) & 1u), by = (int)(e >> 6);
            const uint32_t ow = __shfl_sync(0xffffffffu, origin, owner);
            const uint32_t ex = __shfl_sync(0xffffffffu, extent, owner);
            const int fa = __shfl_sync(0xffffffffu, faceA, owner), fb = __shfl_sync(0xffffffffu, faceB, owner);
            unsigned M = 0u;
            if (valid)
                M = owned_job(fmap, is, fa, fb, (int)(ow & 4095u) + 4 * bx, (int)(ow >> 12) + 4 * by, (int)(ex & 255u) - 4 * bx,
                              (int)(ex >> 8) - 4 * by);
#if G2S_BPHASE < 4
            if (M == 0xdeadbeefu) A6[0] = 1.f;
            continue;
#endif
#pragma unroll
            for (int k = 0; k < 6; k++) part[lane * 6 + k] = 0.f;
            __syncwarp();
            drain_rounds(sm, M, e, origin, hit);
            // owners collect the partial sums of their jobs that ran in this pass (a quad's jobs are consecutive in the list)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int jl = jbase + k - j0;
                if (k < nj && jl >= 0 && jl < 32) {
#pragma unroll
                    for (int m = 0; m < 6; m++) A6[m] += part[jl * 6 + m];
                }
            }
            __syncwarp();
        }
    }
#if G2S_BPHASE < 5
    if (A6[0] + A6[3] == 1.2345e-30f) grad_depth_img[0] = 1.f;
    return;
#endif
    // per-face sums -> (u, v, z) gradients of the face's three vertices, left in the face's table entry in the CANONICAL vertex
    // order (first triangle v00, v10, v01; second v01, v10, v11), zeros for faces that own nothing
    {
        const float hs = 0.5f * (float)is;
#pragma unroll
        for (int tri = 0; tri < 2; tri++) {
            const bool act = tri ? actB : actA;
            const int slot = lane * 2 + tri;
            float gv[3][3];
#pragma unroll
            for (int m = 0; m < 3; m++) gv[m][0] = gv[m][1] = gv[m][2] = 0.f;
            float* rec = &sm.ftab[warp][slot * REC_F];
            if (act) {
                const float A[3] = {A6[tri * 3], A6[tri * 3 + 1], A6[tri * 3 + 2]};
                if (A[0] != 0.f || A[1] != 0.f || A[2] != 0.f) {
                    const float4* r4 = reinterpret_cast<const float4*>(rec);
                    const float4 r0 = r4[0], r1 = r4[1], r2 = r4[2];
                    const float fi[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
                    const float z[3] = {r2.y, r2.z, r2.w};
                    face_vertex_grads(fi, z, (sm.fword[warp][slot] >> 31) != 0u, A, hs, gv);
                }
            }
            const bool rev = tri ? (g.fronts & 4u) == 0u : (g.fronts & 1u) == 0u;     // reversed winding: vertices 0 and 2 swap
#pragma unroll
            for (int k = 0; k < 3; k++) {
                rec[k] = rev ? gv[2][k] : gv[0][k];
                rec[3 + k] = gv[1][k];
                rec[6 + k] = rev ? gv[0][k] : gv[2][k];
            }
        }
    }
    __syncwarp();
    // Vertex chain, warp-local (no CTA barrier: the eight warps of a tile finish independently).  This warp owns quad rows 2w
    // and 2w + 1, i.e. vertex rows 2w .. 2w + 2: every one of those 3 x 17 vertices gathers the gradients of the faces of THIS
    // warp around it and goes through the chain.  The chain is linear, so the rows shared with the neighbouring warps (and
    // the tile's border vertices, shared with neighbouring tiles) are simply pushed through once per owner of a share.
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    float Rt[12];
    if (!FROM_VERTS) {
#pragma unroll
        for (int k = 0; k < 9; k++) Rt[k] = __ldg(&R_b[k]);
#pragma unroll
        for (int k = 0; k < 3; k++) Rt[9 + k] = __ldg(&t_b[k]);
    }
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int idx = it * 32 + lane;
        if (idx >= 3 * TV) break;
        const int r = idx / TV, lx = idx - r * TV, ly = 2 * warp + r;
        float gsum[3] = {0.f, 0.f, 0.f};
        // (quad dy, quad dx, triangle, canonical vertex) of the six faces that share a vertex
        const int nb[6][4] = {{0, 0, 0, 0}, {0, -1, 0, 2}, {0, -1, 1, 0}, {-1, 0, 0, 1}, {-1, 0, 1, 1}, {-1, -1, 1, 2}};
#pragma unroll
        for (int n = 0; n < 6; n++) {
            const int qr = r + nb[n][0], qx2 = lx + nb[n][1];          // quad row within the warp: 0 or 1
            if (qr < 0 || qr > 1 || qx2 < 0 || qx2 >= TILE) continue;
            const float* rec = &sm.ftab[warp][((qr * TILE + qx2) * 2 + nb[n][2]) * REC_F + nb[n][3] * 3];
            gsum[0] += rec[0]; gsum[1] += rec[1]; gsum[2] += rec[2];
        }
        const int vy = ty0 + ly, vx = tx0 + lx;
        if ((gsum[0] != 0.f || gsum[1] != 0.f || gsum[2] != 0.f) && vy < S && vx < S)
            vertex_chain<FROM_VERTS>(cam, depth_img, Rt, verts_b, vy * S + vx, gsum[0], gsum[1], gsum[2], grad_depth_img,
                                     grad_verts_b, acc);
    }
    // grad_R | grad_t: warp sums (recursive halving), one global atomic per value and warp
    if (!FROM_VERTS && grad_R_b) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = k < 12 ? acc[k] : 0.f;
        const float sum = warp_reduce_halving<4>(v, lane);
        const int idx = halving_index<4>(lane);
        if ((lane & 1) == 0 && idx < 12 && sum != 0.f) atomicAdd(idx < 9 ? &grad_R_b[idx] : &grad_t_b[idx - 9], sum);
    }
}

// ---- stage 2: the deferred faces --------------------------------------------------------------------------------------
struct BigBwd {
    static constexpr bool kBackward = true;
    struct Row {      // "does the face-index map name this face here" on one sub-pixel row
        const int* rowp;
        int face;
        __device__ __forceinline__ void init(const BigSmem& sm, const BigCtx& c, int slot, const Tri&, int yi) {
            rowp = c.fmap + sm.zoff[slot] + (long)(c.is - 1 - yi) * c.is;
            face = (int)sm.face[slot];
        }
        __device__ __forceinline__ bool test(int xi) const { return __ldg(&rowp[xi]) == face; }
    };
    static __device__ __forceinline__ void hit(BigSmem& sm, const BigCtx& cx, int slot, int xi, int yi) {
        const int S = cx.is >> 1;
        const float g = __ldg(&cx.gsub[(sm.zoff[slot] >> 2) + (unsigned long long)(((cx.is - 1 - yi) >> 1) * S + (xi >> 1))]);
        if (g == 0.f) return;
        float w[3], zp = 0.f;
        record_weights_depth(&sm.ftab[slot * FT_STRIDE], xi, yi, cx.near, cx.far, w, &zp);
        const float s = g * zp * zp;
#pragma unroll
        for (int m = 0; m < 3; m++) atomicAdd(&cx.aacc[slot * 3 + m], s * w[m]);
    }
};

// per-face sums of a batch -> vertex gradients -> grad_depth / grad_R / grad_t (or grad_vertices): one thread per face
template <bool FROM_VERTS>
struct BigBwdFinish {
    Cam cam;
    const float* depth; long dstride; int vpi;
    const float* R; const float* t; const float* verts3d;
    float* grad_depth; long gdstride;
    float* grad_verts; float* grad_R; float* grad_t;
    __device__ __forceinline__ void operator()(BigSmem& sm, const BigCtx& cx, int n, int view0) const {
        const int tid = threadIdx.x, S = cam.S;
        if (tid >= n) return;
        const float A[3] = {cx.aacc[tid * 3], cx.aacc[tid * 3 + 1], cx.aacc[tid * 3 + 2]};
        if (A[0] == 0.f && A[1] == 0.f && A[2] == 0.f) return;
        const float* rec = &sm.ftab[tid * FT_STRIDE];
        float fi[9], z[3], gv[3][3];
#pragma unroll
        for (int k = 0; k < 9; k++) fi[k] = rec[k];
#pragma unroll
        for (int k = 0; k < 3; k++) z[k] = rec[9 + k];
        face_vertex_grads(fi, z, rec[FT_FLAG] != 0.0f, A, 0.5f * (float)cx.is, gv);
        const int b = view0 + (int)(sm.zoff[tid] / ((unsigned long long)cx.is * cx.is));
        int vidx[3];
        face_vertices((int)sm.face[tid], S, vidx);
        float acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.f;
        float Rt[12];
        if (!FROM_VERTS) {
#pragma unroll
            for (int k = 0; k < 9; k++) Rt[k] = __ldg(&R[(long)b * 9 + k]);
#pragma unroll
            for (int k = 0; k < 3; k++) Rt[9 + k] = __ldg(&t[(long)b * 3 + k]);
        }
#pragma unroll 1
        for (int m = 0; m < 3; m++) {
            if (A[m] == 0.f) continue;
            vertex_chain<FROM_VERTS>(cam, FROM_VERTS ? nullptr : depth + (long)(b / vpi) * dstride, Rt,
                                     FROM_VERTS ? verts3d + (long)b * S * S * 3 : nullptr, vidx[m], gv[m][0], gv[m][1], gv[m][2],
                                     FROM_VERTS ? nullptr : grad_depth + (long)(b / vpi) * gdstride,
                                     FROM_VERTS ? grad_verts + (long)b * S * S * 3 : nullptr, acc);
        }
        if (!FROM_VERTS && grad_R) {
#pragma unroll
            for (int k = 0; k < 9; k++) if (acc[k] != 0.f) atomicAdd(&grad_R[(long)b * 9 + k], acc[k]);
#pragma unroll
            for (int k = 0; k < 3; k++) if (acc[9 + k] != 0.f) atomicAdd(&grad_t[(long)b * 3 + k], acc[9 + k]);
        }
    }
};

}  // namespace g2s
