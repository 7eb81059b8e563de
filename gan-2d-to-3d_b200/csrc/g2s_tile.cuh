// g2s_tile.cuh -- forward rasteriser of the grid mesh, stage 1 of 2: k_splat_tile (round 2 rewrite).
//
// Replaces neural_renderer's forward_face_index_map (every sub-pixel loops over every face; reference call site
// GAN2Shape/renderer/renderer.py:120) for the regular-grid mesh of GAN2Shape/renderer/utils.py:76-80.
//
// One CTA owns a TILE x TILE_H block of quads of one view.
//   1. project the tile's (TILE+1) x (TILE_H+1) vertices once into shared memory: NDC x, y, z and the sub-pixel
//      coordinates px, py that both the boxes and the 3x3 face inverses are built from; ONE barrier.  Everything after it is
//      warp-local (a warp owns two rows of 16 quads): the eight warps of a tile drift apart and overlap their phases.
//   2. every lane classifies its quad from the UNION box of its two triangles and builds the 3x3 inverses of its front
//      windings into the warp's face table (12 floats per face, conflict-free 16-byte rows).  Quads whose box exceeds 16 x 16
//      sub-pixels (the long depth-step walls), degenerate and non-finite quads go to a global WORK LIST of faces that the
//      second stage (k_splat_big, g2s_bigface.cuh) rasterises, balanced over the whole GPU.
//   3. the boxes are cut into JOBS of 4 x 4 sub-pixels (one for an interior quad, up to sixteen for a sheared one), listed per
//      warp and scanned one lane per job with straight-line code: the exact inside test `!((yp-yk)*dx < (xp-xk)*dy)` split
//      into per-row and per-column terms -- three predicated compares per (triangle, candidate) -- into two 16-bit hit
//      masks.  No loop bound depends on the largest box of the warp (round 1 / the first form of this file scanned
//      warp-maximum boxes and compacted the medium quads CTA-wide behind a second barrier).
//   4. the hits are evaluated in ROUNDS: every lane contributes at most K_ROUND hits per round to the warp's hit queue
//      (16-bit entries), which the warp drains one lane per hit -- weights, perspective z (seven correctly rounded
//      divisions as residual-corrected products), 64-bit atomicMin into the z-buffer.
// 37 KB of static shared memory (round 1: 106 KB dynamic, 2 CTAs per SM).
#pragma once
#include "g2s_splat.cuh"

namespace g2s {

#ifndef G2S_PHASE
#define G2S_PHASE 9     // < 9: experiment builds that stop k_splat_tile after a phase (per-phase timing, profiles/r02_notes.md)
#endif
constexpr int K_ROUND = 8;                      // hits a lane contributes to one drain round
#ifndef G2S_MAX_BOX
#define G2S_MAX_BOX 16                          // largest quad box (sub-pixels per side) rasterised by stage 1
#endif
constexpr int MAX_BOX = G2S_MAX_BOX;
constexpr int MAX_JOBS = (MAX_BOX / 4) * (MAX_BOX / 4);   // 4 x 4 jobs of one quad
constexpr int TWARPS = SPLAT_THREADS / 32;
constexpr int REC_F = 12;                       // face-table entry: fi[9], z[3]
constexpr int NV_PAD = (TV * TVH + 3) & ~3;

static_assert(TILE == 16 && SPLAT_THREADS == 256, "quad <-> thread mapping below assumes 16 x 16 tiles");

struct TileSmem2 {
    float4 vxy[TV * TVH];                       // NDC x, NDC y, sub-pixel x, sub-pixel y
    float vz[NV_PAD];
    float ftab[TWARPS][64 * REC_F];             // warp-local face table, entry = lane * 2 + triangle
    uint32_t fword[TWARPS][64];                 // face index | z-range verdict << 31
    uint16_t hq[TWARPS][32 * K_ROUND];          // warp-local hit queue: job lane | bit << 5
    uint16_t jobs[TWARPS][32 * MAX_JOBS];       // warp-local job list: owner lane | box column << 5 | box row << 7
};

// Work list of the second stage.  Every word of the z-buffer workspace is the EMPTY key at rest (so that any later call may
// lay the workspace out for another number of views): the three counters are kept BIASED by that value.
struct WorkList {
    unsigned long long* items;                  // (view in launch) << 32 | face index
    unsigned long long* ctr;                    // [0] items queued, [1] next ticket, [2] CTAs done
    unsigned long long bias;                    // zkey_empty(far)
};

// [nr] kernel 1 (tri_face_inv) from the sub-pixel coordinates the tile already holds; same bits as face_record.
__device__ __forceinline__ void face_record_px(float p00, float p01, float p10, float p11, float p20, float p21, float z0,
                                               float z1, float z2, float rec[REC_F], bool* zbad) {
    float a[9];
    a[0] = sub(p11, p21); a[1] = sub(p20, p10); a[2] = sub(mul(p10, p21), mul(p20, p11));
    a[3] = sub(p21, p01); a[4] = sub(p00, p20); a[5] = sub(mul(p20, p01), mul(p00, p21));
    a[6] = sub(p01, p11); a[7] = sub(p10, p00); a[8] = sub(mul(p00, p11), mul(p10, p01));
    const float den = add(add(mul(p20, sub(p01, p11)), mul(p00, sub(p11, p21))), mul(p10, sub(p21, p01)));
    const float y = rcp_seed(den);
#pragma unroll
    for (int k = 0; k < 9; k++) rec[k] = div_core(a[k], den, y);
    // operand ranges from the structure of the values (see face_record)
    const float pmax = fmaxf(fmaxf(fmaxf(fabsf(p00), fabsf(p01)), fmaxf(fabsf(p10), fabsf(p11))), fmaxf(fabsf(p20), fabsf(p21)));
    if (!(pmax < 524288.0f) || range_key(den) >= RANGE_SPAN) {
#pragma unroll
        for (int k = 0; k < 9; k++) rec[k] = __fdiv_rn(a[k], den);
    }
    rec[9] = z0; rec[10] = z1; rec[11] = z2;
    *zbad = max(max(range_key(z0), range_key(z1)), range_key(z2)) >= RANGE_SPAN;
}

template <bool FROM_VERTS>
__device__ __forceinline__ void tile_project2(const Cam& cam, const float* __restrict__ depth_b,
                                              const float* __restrict__ verts_b, const float* __restrict__ R_b,
                                              const float* __restrict__ t_b, int ty0, int tx0, TileSmem2& sm,
                                              float4* __restrict__ proj_view) {
    const int S = cam.S, is = 2 * S;
    constexpr int ROUNDS = (TV * TVH + SPLAT_THREADS - 1) / SPLAT_THREADS;
    float in[ROUNDS][3];
    bool live[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int i = threadIdx.x + r * SPLAT_THREADS;
        const int vy = ty0 + i / TV, vx = tx0 + i % TV;
        live[r] = i < TV * TVH && vy < S && vx < S;
        in[r][0] = in[r][1] = in[r][2] = 0.f;
        if (live[r]) {
            if (FROM_VERTS) {
                const float* p = &verts_b[((long)vy * S + vx) * 3];
                in[r][0] = __ldg(p); in[r][1] = __ldg(p + 1); in[r][2] = __ldg(p + 2);
            } else {
                in[r][0] = __ldg(&depth_b[vy * S + vx]);
            }
        }
    }
    float Rt[12];
    if (!FROM_VERTS) {
#pragma unroll
        for (int k = 0; k < 9; k++) Rt[k] = __ldg(&R_b[k]);
#pragma unroll
        for (int k = 0; k < 3; k++) Rt[9 + k] = __ldg(&t_b[k]);
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int i = threadIdx.x + r * SPLAT_THREADS;
        if (r > 0 && !__any_sync(0xffffffffu, i < TV * TVH)) break;
        if (i >= TV * TVH) continue;
        float ndc[3] = {0.f, 0.f, 0.f};
        if (live[r]) {
            float q[3];
            if (FROM_VERTS) {
                q[0] = in[r][0]; q[1] = in[r][1]; q[2] = in[r][2];
            } else {
                float ray[3];
                pixel_ray(cam, tx0 + i % TV, ty0 + i / TV, ray);
                warp_point(cam, Rt, Rt + 9, ray, in[r][0], q);
            }
            project_ndc(cam, q, ndc);
        }
        const float4 vx4 = make_float4(ndc[0], ndc[1], ndc_to_pix(ndc[0], is), ndc_to_pix(ndc[1], is));
        sm.vxy[i] = vx4;
        sm.vz[i] = ndc[2];
        // handed to the backward (k_raster_bwd_px reads sub-pixel x, y and z): every vertex by the tile that owns it (the 17th
        // row / column of a tile belongs to its neighbour, except on the mesh's last row / column)
        if (proj_view != nullptr && live[r]) {
            const int ly = i / TV, lx = i % TV, vy = ty0 + ly, vx = tx0 + lx;
            if ((ly < TILE_H || vy == S - 1) && (lx < TILE || vx == S - 1))
                proj_view[vy * S + vx] = make_float4(vx4.z, vx4.w, ndc[2], 0.f);
        }
    }
}

__device__ __forceinline__ bool back_facing(float x0, float y0, float x1, float y1, float x2, float y2) {
    return mul(sub(y2, y0), sub(x1, x0)) < mul(sub(y1, y0), sub(x2, x0));   // [nr] kernels 1 and 2
}

enum QuadClass { QC_NONE = 0, QC_SMALL = 1, QC_MEDIUM = 2, QC_DEFER = 3 };

// Geometry of one quad: its four projected vertices, the union box of its two triangles, which winding of each
// triangle faces the camera.  Triangle A = (v00, v10, v01), B = (v01, v10, v11) (utils.py:76-80); the fill_back copy
// of either swaps the first and last vertex.
struct QuadGeom {
    float4 v00, v01, v10, v11;     // x, y (NDC), px, py (sub-pixels)
    int x0, y0, uw, uh;
    unsigned fronts;               // bit 0: A front, 1: A reversed front, 2: B front, 3: B reversed front
    bool finA, finB;
    int cls;
};

__device__ __forceinline__ void quad_geometry(const TileSmem2& sm, int is, int quad, bool quad_ok, QuadGeom& g) {
    const int i00 = (quad >> 4) * TV + (quad & 15);
    g.v00 = sm.vxy[i00]; g.v01 = sm.vxy[i00 + 1]; g.v10 = sm.vxy[i00 + TV]; g.v11 = sm.vxy[i00 + TV + 1];
    const float chkA = g.v00.z + g.v10.z + g.v01.z + g.v00.w + g.v10.w + g.v01.w;
    const float chkB = g.v01.z + g.v10.z + g.v11.z + g.v01.w + g.v10.w + g.v11.w;
    g.finA = fabsf(chkA) < 3.0e38f;     // a face with a non-finite coordinate can never win (NaN weights)
    g.finB = fabsf(chkB) < 3.0e38f;
    // union of the two triangle boxes of tri_bbox (grown by 1/64 px, clipped to the image)
    const float m = 1.0f / 64.0f, lim = (float)is;
    const float xmin = fmaxf(fminf(fminf(g.v00.z, g.v01.z), fminf(g.v10.z, g.v11.z)) - m, -1.0f);
    const float xmax = fminf(fmaxf(fmaxf(g.v00.z, g.v01.z), fmaxf(g.v10.z, g.v11.z)) + m, lim);
    const float ymin = fmaxf(fminf(fminf(g.v00.w, g.v01.w), fminf(g.v10.w, g.v11.w)) - m, -1.0f);
    const float ymax = fminf(fmaxf(fmaxf(g.v00.w, g.v01.w), fmaxf(g.v10.w, g.v11.w)) + m, lim);
    g.x0 = imax(0, (int)ceilf(xmin));
    g.y0 = imax(0, (int)ceilf(ymin));
    g.uw = imin(is - 1, (int)floorf(xmax)) - g.x0 + 1;
    g.uh = imin(is - 1, (int)floorf(ymax)) - g.y0 + 1;
    const bool fA0 = !back_facing(g.v00.x, g.v00.y, g.v10.x, g.v10.y, g.v01.x, g.v01.y);
    const bool fA1 = !back_facing(g.v01.x, g.v01.y, g.v10.x, g.v10.y, g.v00.x, g.v00.y);
    const bool fB0 = !back_facing(g.v01.x, g.v01.y, g.v10.x, g.v10.y, g.v11.x, g.v11.y);
    const bool fB1 = !back_facing(g.v11.x, g.v11.y, g.v10.x, g.v10.y, g.v01.x, g.v01.y);
    g.fronts = g.finA ? ((fA0 ? 1u : 0u) | (fA1 ? 2u : 0u)) : 0u;
    g.fronts |= g.finB ? ((fB0 ? 4u : 0u) | (fB1 ? 8u : 0u)) : 0u;
    if (!quad_ok) g.fronts = 0u;
    const bool dup = (g.fronts & 3u) == 3u || (g.fronts & 12u) == 12u;   // both windings pass: degenerate triangle
    if (g.fronts == 0u) g.cls = QC_NONE;
    else if (!(g.finA && g.finB) || dup) g.cls = QC_DEFER;
    else if (g.uw <= 0 || g.uh <= 0) g.cls = QC_NONE;
    else if (g.uw > MAX_BOX || g.uh > MAX_BOX) g.cls = QC_DEFER;
    else g.cls = (g.uw <= 4 && g.uh <= 4) ? QC_SMALL : QC_MEDIUM;
}

// NDC coordinate of a sub-pixel centre, (2i + 1 - is) / is: an exact product when `is` is a power of two (compile-time: the
// run-time form of PixCenter costs a branch per row and per column of the scan)
template <bool POW2>
struct PixCenterT {
    int is;
    float inv;
    __device__ __forceinline__ void init(int is_) { is = is_; inv = POW2 ? 1.0f / (float)is_ : rcp_seed((float)is_); }
    __device__ __forceinline__ float operator()(int i) const {
        const float n = (float)(2 * i + 1 - is);      // |n| < 2^13, is in [4, 4096]: inside div_core's exact range (or n = 0)
        return POW2 ? __fmul_rn(n, inv) : div_core(n, (float)is, inv);
    }
};

struct TileCtx {
    unsigned long long* zb;        // z-buffer of this view
    float near, far;
    int is, S, Q, ty0, tx0;
};

// per-row / per-column halves of the three edge tests of one triangle (same values as tri_contains)
struct EdgeSet {
    float x0, x1, x2, y0, y1, y2, dx01, dy01, dx12, dy12, dx20, dy20;
    __device__ __forceinline__ void init(float ax, float ay, float bx, float by, float cx, float cy) {
        x0 = ax; y0 = ay; x1 = bx; y1 = by; x2 = cx; y2 = cy;
        dx01 = sub(bx, ax); dy01 = sub(by, ay);
        dx12 = sub(cx, bx); dy12 = sub(cy, by);
        dx20 = sub(ax, cx); dy20 = sub(ay, cy);
    }
    __device__ __forceinline__ void col(float xp, float c[3]) const {
        c[0] = mul(sub(xp, x0), dy01); c[1] = mul(sub(xp, x1), dy12); c[2] = mul(sub(xp, x2), dy20);
    }
    __device__ __forceinline__ void row(float yp, float a[3]) const {
        a[0] = mul(sub(yp, y0), dx01); a[1] = mul(sub(yp, y1), dx12); a[2] = mul(sub(yp, y2), dx20);
    }
};
// mask |= bit  iff  !(a0 < c0) && !(a1 < c1) && !(a2 < c2): three chained predicate compares and one predicated OR (the
// compiler's own form of `cond ? bit : 0` ORed together was three selects and an OR per candidate)
__device__ __forceinline__ void inside_or(unsigned& mask, const float a[3], const float c[3], unsigned bit) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.geu.f32 p, %1, %2;\n\t"
        "setp.geu.and.f32 p, %3, %4, p;\n\t"
        "setp.geu.and.f32 p, %5, %6, p;\n\t"
        "@p or.b32 %0, %0, %7;\n\t}"
        : "+r"(mask) : "f"(a[0]), "f"(c[0]), "f"(a[1]), "f"(c[1]), "f"(a[2]), "f"(c[2]), "r"(bit));
}

// face-table entries (3x3 inverse, depths) of this lane's triangles that scored, into the warp's table
__device__ __forceinline__ void build_face_records(TileSmem2& sm, const TileCtx& cx, const QuadGeom& g, int quad, bool doA,
                                                   bool doB) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool revA = (g.fronts & 1u) == 0u, revB = (g.fronts & 4u) == 0u;
    const int i00 = (quad >> 4) * TV + (quad & 15);
    const int fq = (cx.ty0 + (quad >> 4)) * (cx.S - 1) + cx.tx0 + (quad & 15);
    float rec[REC_F];
    bool zbad;
    if (doA) {
        const int ia = revA ? i00 + 1 : i00, ic = revA ? i00 : i00 + 1, ib = i00 + TV;
        const float4 a = revA ? g.v01 : g.v00, c = revA ? g.v00 : g.v01;
        face_record_px(a.z, a.w, g.v10.z, g.v10.w, c.z, c.w, sm.vz[ia], sm.vz[ib], sm.vz[ic], rec, &zbad);
        float4* dst = reinterpret_cast<float4*>(&sm.ftab[warp][(lane * 2) * REC_F]);
        dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
        dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
        dst[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
        sm.fword[warp][lane * 2] = (uint32_t)((revA ? 2 * cx.Q : 0) + fq) | (zbad ? 0x80000000u : 0u);
    }
    if (doB) {
        const int ia = revB ? i00 + TV + 1 : i00 + 1, ic = revB ? i00 + 1 : i00 + TV + 1, ib = i00 + TV;
        const float4 a = revB ? g.v11 : g.v01, c = revB ? g.v01 : g.v11;
        face_record_px(a.z, a.w, g.v10.z, g.v10.w, c.z, c.w, sm.vz[ia], sm.vz[ib], sm.vz[ic], rec, &zbad);
        float4* dst = reinterpret_cast<float4*>(&sm.ftab[warp][(lane * 2 + 1) * REC_F]);
        dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
        dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
        dst[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
        sm.fword[warp][lane * 2 + 1] = (uint32_t)((revB ? 3 * cx.Q : cx.Q) + fq) | (zbad ? 0x80000000u : 0u);
    }
}

// Rounds: every lane queues at most K_ROUND of the hits of its job (M: bits 0..15 = first triangle, 16..31 = second, bit =
// row * 4 + column of the job's 4 x 4 box) as 16-bit entries `lane | bit << 5`; the warp drains the queue one lane per hit:
// hit(valid, job lane, triangle, slot, xi, yi, fi, z, face word), called by ALL lanes.
// jobword = owner lane | box column << 5 | box row << 7 of this lane's job; origin = x0 | y0 << 12 of this lane's OWN quad box.
template <class Hit>
__device__ __forceinline__ void drain_rounds(TileSmem2& sm, unsigned M, unsigned jobword, uint32_t origin, const Hit& hit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* hq = sm.hq[warp];
    while (true) {
        const int c = min(__popc(M), K_ROUND);
        const int cmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)c);
        if (cmax == 0) break;
        int total;
        int base = warp_excl_scan(c, &total);
#pragma unroll 1
        for (int k = 0; k < cmax; k++) {
            if (k < c) {
                const unsigned bit = (unsigned)(__ffs(M) - 1);
                M &= M - 1u;
                hq[base++] = (uint16_t)((unsigned)lane | (bit << 5));
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < total;
            const unsigned e = valid ? hq[i] : 0u;
            const int jl = (int)(e & 31u);                       // the lane whose job this hit belongs to
            const unsigned jw = __shfl_sync(0xffffffffu, jobword, jl);
            const unsigned owner = jw & 31u;
            const uint32_t ow = __shfl_sync(0xffffffffu, origin, (int)owner);
            const unsigned bit = e >> 5, tri = bit >> 4;
            const int xi = (int)(ow & 4095u) + (int)(((jw >> 5) & 3u) * 4u + (bit & 3u));
            const int yi = (int)((ow >> 12) & 4095u) + (int)(((jw >> 7) & 3u) * 4u + ((bit >> 2) & 3u));
            const int slot = (int)(owner * 2u + tri);            // invalid lanes read slot 0
            const float4* r4 = reinterpret_cast<const float4*>(&sm.ftab[warp][slot * REC_F]);
            const float4 r0 = r4[0], r1 = r4[1], r2 = r4[2];
            const uint32_t fw = sm.fword[warp][slot];
            const float fi[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
            const float z[3] = {r2.y, r2.z, r2.w};
            hit(valid, jl, (int)tri, slot, xi, yi, fi, z, fw);
        }
        __syncwarp();
    }
}

// forward hit: weights, perspective z, 64-bit atomicMin of the packed key
struct FwdHit {
    const TileCtx* cx;
    __device__ __forceinline__ void operator()(bool valid, int, int, int, int xi, int yi, const float fi[9],
                                               const float z[3], uint32_t fw) const {
        float w[3], zp;
        if (valid && weights_depth_core(fi, z, (fw >> 31) != 0u, xi, yi, cx->near, cx->far, w, &zp))
            atomicMin(&cx->zb[(long)(cx->is - 1 - yi) * cx->is + xi], zkey_pack(zp, fw & 0x7fffffffu));
    }
};

// Exact inside tests of both triangles of quad `quad` over the 4 x 4 sub-pixel box at (bx0, by0); `fronts` says which winding
// of each triangle faces the camera, (nw, nh) = columns / rows of the box that belong to the quad's own box.
// Returns the hit mask: bit row * 4 + column (first triangle), 16 + row * 4 + column (second).
template <bool POW2>
__device__ __forceinline__ unsigned scan_job(const TileSmem2& sm, const PixCenterT<POW2>& pc, int quad, unsigned fronts,
                                             int bx0, int by0, int nw, int nh, int is) {
    const int i00 = (quad >> 4) * TV + (quad & 15);
    const float4 v00 = sm.vxy[i00], v01 = sm.vxy[i00 + 1], v10 = sm.vxy[i00 + TV], v11 = sm.vxy[i00 + TV + 1];
    const bool revA = (fronts & 1u) == 0u, revB = (fronts & 4u) == 0u;
    EdgeSet ea, eb;
    ea.init(revA ? v01.x : v00.x, revA ? v01.y : v00.y, v10.x, v10.y, revA ? v00.x : v01.x, revA ? v00.y : v01.y);
    eb.init(revB ? v11.x : v01.x, revB ? v11.y : v01.y, v10.x, v10.y, revB ? v01.x : v11.x, revB ? v01.y : v11.y);
    float ca[4][3], cb[4][3];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float xp = pc(min(bx0 + c, is - 1));
        ea.col(xp, ca[c]);
        eb.col(xp, cb[c]);
    }
    unsigned mA = 0u, mB = 0u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float yp = pc(min(by0 + r, is - 1));
        float ra[3], rb[3];
        ea.row(yp, ra);
        eb.row(yp, rb);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            inside_or(mA, ra, ca[c], 1u << (r * 4 + c));
            inside_or(mB, rb, cb[c], 1u << (r * 4 + c));
        }
    }
    // validity: inside the quad's own box, triangle front-facing in one of its windings
    unsigned vm = ((1u << min(max(nw, 0), 4)) - 1u) * 0x1111u;
    vm &= (1u << (4 * min(max(nh, 0), 4))) - 1u;
    if ((fronts & 3u) == 0u) mA = 0u;
    if ((fronts & 12u) == 0u) mB = 0u;
    return (mA & vm) | ((mB & vm) << 16);
}

// One warp rasterises its 32 quads (lane `active` <=> its quad g has a box of at most MAX_BOX x MAX_BOX sub-pixels and is
// neither degenerate nor non-finite): face records, job list, scan, rounds.  Warp-local: no CTA barrier.
template <bool POW2>
__device__ __forceinline__ void raster_warp(TileSmem2& sm, const TileCtx& cx, const PixCenterT<POW2>& pc, const QuadGeom& g,
                                            bool active, int quad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, is = cx.is;
    if (!__any_sync(0xffffffffu, active)) return;
    build_face_records(sm, cx, g, quad, active && (g.fronts & 3u) != 0u, active && (g.fronts & 12u) != 0u);
    // jobs: the quad's box in 4 x 4 pieces
    const int njx = active ? (g.uw + 3) >> 2 : 0, njy = active ? (g.uh + 3) >> 2 : 0, nj = njx * njy;
    int total;
    const int jbase = warp_excl_scan(nj, &total);
    uint16_t* jobs = sm.jobs[warp];
    // (column, row) of the k-th piece by counting, not by k % njx and k / njx: two integer divisions per job were 2.7 % of the
    // kernel's instructions
#pragma unroll 1
    for (int k = 0, jx = 0, jy = 0; k < nj; k++) {
        jobs[jbase + k] = (uint16_t)(lane | (jx << 5) | (jy << 7));
        if (++jx == njx) { jx = 0; jy++; }
    }
    const uint32_t origin = (uint32_t)g.x0 | ((uint32_t)g.y0 << 12) | (g.fronts << 24);
    const uint32_t extent = (uint32_t)g.uw | ((uint32_t)g.uh << 8);
    __syncwarp();
    FwdHit hit;
    hit.cx = &cx;
#pragma unroll 1
    for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < total;
        const unsigned e = valid ? jobs[j] : 0u;
        const int owner = (int)(e & 31u), bx = (int)((e >> 5) & 3u), by = (int)(e >> 7);
        const uint32_t ow = __shfl_sync(0xffffffffu, origin, owner);
        const uint32_t ex = __shfl_sync(0xffffffffu, extent, owner);
        unsigned M = scan_job<POW2>(sm, pc, (quad & ~31) + owner, ow >> 24, (int)(ow & 4095u) + 4 * bx,
                                    (int)((ow >> 12) & 4095u) + 4 * by, (int)(ex & 255u) - 4 * bx, (int)(ex >> 8) - 4 * by, is);
        if (!valid) M = 0u;
#if G2S_PHASE < 5     // experiment builds (profiles/tools/ab_bench.py): stop after the scan, keep its result alive
        if (M == 0xdeadbeefu) cx.zb[0] = M;
        continue;
#endif
        drain_rounds(sm, M, e, origin, hit);
    }
}

// Append the front windings of a deferred quad to the work list (one global atomic per warp).
__device__ __forceinline__ void defer_quads(const WorkList& wl, const TileCtx& cx, const QuadGeom& g, bool defer, int quad,
                                            int view_in_launch) {
    if (!__any_sync(0xffffffffu, defer)) return;
    const int lane = threadIdx.x & 31;
    const int n = defer ? __popc(g.fronts) : 0;
    int total;
    const int off = warp_excl_scan(n, &total);
    unsigned long long base = 0ull;
    if (lane == 0) base = atomicAdd(wl.ctr, (unsigned long long)total) - wl.bias;
    base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)off;
    if (defer) {
        const int fq = (cx.ty0 + (quad >> 4)) * (cx.S - 1) + cx.tx0 + (quad & 15);
        unsigned f = g.fronts;
        while (f) {
            const int k = __ffs(f) - 1;          // 0: A, 1: A reversed, 2: B, 3: B reversed
            f &= f - 1u;
            const int face = ((k >> 1) + 2 * (k & 1)) * cx.Q + fq;
            wl.items[base++] = ((unsigned long long)view_in_launch << 32) | (unsigned long long)(uint32_t)face;
        }
    }
}

template <bool FROM_VERTS, bool POW2>
__device__ __forceinline__ void splat_tile_body(TileSmem2& sm, const Cam& cam, const float* __restrict__ depth_b,
                                                const float* __restrict__ verts_b, const float* __restrict__ R_b,
                                                const float* __restrict__ t_b, unsigned long long* zb_view,
                                                const WorkList& wl, int view_in_launch, int ty0, int tx0,
                                                float4* __restrict__ proj_view) {
    const int tid = threadIdx.x, S = cam.S, is = 2 * S;
    tile_project2<FROM_VERTS>(cam, depth_b, verts_b, R_b, t_b, ty0, tx0, sm, proj_view);
    __syncthreads();
#if G2S_PHASE < 2
    if (sm.vz[tid] == 1.2345e-30f) zb_view[0] = 0ull;
    return;
#endif
    TileCtx cx;
    cx.zb = zb_view; cx.near = cam.near; cx.far = cam.far; cx.is = is; cx.S = S; cx.Q = (S - 1) * (S - 1);
    cx.ty0 = ty0; cx.tx0 = tx0;
    PixCenterT<POW2> pc;
    pc.init(is);
    QuadGeom g;
    const int qy = tid >> 4, qx = tid & 15;
    quad_geometry(sm, is, tid, ty0 + qy < S - 1 && tx0 + qx < S - 1, g);
    defer_quads(wl, cx, g, g.cls == QC_DEFER, tid, view_in_launch);
#if G2S_PHASE < 3
    if (g.x0 == -12345 && g.uw == 77) zb_view[0] = (unsigned long long)g.fronts;
    return;
#endif
    raster_warp<POW2>(sm, cx, pc, g, g.cls == QC_SMALL || g.cls == QC_MEDIUM, tid);
}

}  // namespace g2s
