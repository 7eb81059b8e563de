// g2s_splat.cuh -- the forward tile rasteriser of the grid mesh (k_splat) and the per-hit arithmetic shared with
// the pixel-centric backward (k_raster_bwd_px in g2s_kernels.cu).
//
// Replaces neural_renderer's forward_face_index_map, where every sub-pixel loops over every face.
//
// One CTA per TILE x TILE block of quads of one view:
//   1. project the tile's (TILE+1)^2 vertices once into shared memory (u, v, z);
//   2. SMALL QUADS (both triangles inside one SB x SB sub-pixel box: every interior quad) never leave their warp: the
//      owning thread scans the box with UNIFORM control flow (loop bounds = warp maxima, per-lane predicates; only the
//      cheap, exact inside test: three edge functions with the loop invariants hoisted) into two 16-bit hit masks; the hits
//      go to the warp's private slice of the hit queue (offsets from a warp scan, no atomics); every lane builds the 3x3
//      inverses of its own two triangles into the shared-memory face table; after a __syncwarp the warp drains its slice,
//      one lane per hit: weights, perspective z, 64-bit atomicMin into the z-buffer.  No CTA barrier in this path;
//   3. everything else (the 1-px depth-step walls stretch to >100 sub-pixels under yaw) is queued; after ONE barrier the
//      table entries of all queued faces are built (one thread per face), whole warps expand the faces into ROW TASKS
//      (face, row, 8-column segment; wide boxes only the segments that overlap a conservative per-row extent), and the
//      tasks are scanned one lane per task in rounds sized to the CTA-wide hit queue, each round drained by consecutive
//      threads.
// The correctly rounded divisions (face table, drain) run fully converged in both paths; in the first per-thread form
// the same code ran at 3-7 active lanes per warp (ncu, profiles/r01_notes.md).
#pragma once
#include "g2s_raster.cuh"

namespace g2s {

#ifdef G2S_COUNT_SLOW     // experiment builds only: how often the queue-overflow slow paths run
__device__ unsigned long long g_slow_counters[2];
#endif

#ifndef G2S_HQ_PER_THREAD
#define G2S_HQ_PER_THREAD 16
#endif
#ifndef G2S_TQ_CAP
#define G2S_TQ_CAP 8192     // 256^2 wall tiles queue up to ~4300 row tasks; overflow falls to a slow inline scan
#endif
#ifndef G2S_FT_SEEDS
#define G2S_FT_SEEDS 0      // 1: the face table also holds the reciprocal seeds of the three z's; 0: recomputed per hit
                            // (same speed, 8 KB less shared memory: spent on the task queue)
#endif
constexpr int HQ_CAP = G2S_HQ_PER_THREAD * SPLAT_THREADS;  // hit-queue entries per drain
constexpr int NSLOT = 2 * TILE * TILE_H;       // (quad, triangle) slots of the face table
constexpr int TQ_CAP = G2S_TQ_CAP;            // queued row tasks per tile
constexpr int FT_FLAG = G2S_FT_SEEDS ? 15 : 12;   // index of the per-face z-range verdict
constexpr int FT_STRIDE = G2S_FT_SEEDS ? 17 : 13;                // fi[9], z[3], rcp_seed(z)[3], pad; ODD so that lanes on consecutive table
                                             // entries hit distinct shared-memory banks (stride 16 was a 16/32-way conflict)
// table entry of a (quad, triangle) code: triangle-major, so that the threads of a warp (consecutive quads) own
// consecutive entries
__device__ __forceinline__ int ft_index(int code) { return (code & 1) * (NSLOT / 2) + ((code & 511) >> 1); }

struct TileSmem {
    float ftab[NSLOT * FT_STRIDE];
    float sv[TV * TVH * 3];
    float recs[NSLOT * 9];       // queued-face records (REC_STRIDE floats each)
    uint32_t tq[TQ_CAP];         // row tasks
    uint32_t hq_pix[HQ_CAP];
    uint16_t hq_code[HQ_CAP];
    uint16_t wq[NSLOT];          // wide faces whose rows a warp expands into tasks
    uint16_t mq[NSLOT];          // other queued faces
    int n_hq, n_tq, n_wq, n_mq;
};

// code = slot | rev << 9, slot = quad * 2 + tri
__device__ __forceinline__ Tri code_tri(const float* sv, int code) {
    const int quad = (code & 511) >> 1;
    return tile_winding(sv, quad / TILE, quad % TILE, (code & 1) + ((code >> 9) << 1));
}
__device__ __forceinline__ int code_face(int code, int Q, int S, int ty0, int tx0) {
    const int quad = (code & 511) >> 1;
    return ((code & 1) + ((code >> 9) << 1)) * Q + (ty0 + quad / TILE) * (S - 1) + tx0 + quad % TILE;
}
__device__ __forceinline__ Tri reversed(const Tri& f) {
    Tri r;
    r.x0 = f.x2; r.y0 = f.y2; r.z0 = f.z2;
    r.x1 = f.x1; r.y1 = f.y1; r.z1 = f.z1;
    r.x2 = f.x0; r.y2 = f.y0; r.z2 = f.z0;
    return r;
}

// Projects the tile's TV x TVH vertices into shared memory.  A thread owns vertex `tid` and, for the first few threads,
// `tid + SPLAT_THREADS`; the per-vertex inputs (depth, or the 3-D point) of BOTH are requested first and the view's R, t
// come as warp-uniform loads, so the tile starts with one memory round trip instead of two and without a barrier.
template <bool FROM_VERTS>
__device__ __forceinline__ void tile_project(const Cam& cam, const float* __restrict__ depth_b,
                                             const float* __restrict__ verts_b, const float* __restrict__ R_b,
                                             const float* __restrict__ t_b, int ty0, int tx0, float* sv) {
    const int S = cam.S;
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
        sv[i * 3 + 0] = ndc[0];
        sv[i * 3 + 1] = ndc[1];
        sv[i * 3 + 2] = ndc[2];
    }
}

// NDC coordinate of a sub-pixel centre, (2i + 1 - is) / is: an exact product when `is` is a power of two
struct PixCenter {
    int is;
    bool pow2;
    float inv;  // 1/is (pow2) or the refined reciprocal seed
    __device__ __forceinline__ void init(int is_) {
        is = is_;
        pow2 = (is_ & (is_ - 1)) == 0;
        inv = pow2 ? 1.0f / (float)is_ : rcp_seed((float)is_);
    }
    __device__ __forceinline__ float operator()(int i) const {
        const float n = (float)(2 * i + 1 - is);
        return pow2 ? __fmul_rn(n, inv) : dvd_y(n, (float)is, inv);
    }
};

__device__ __forceinline__ unsigned range_key(float x) { return (__float_as_uint(x) & 0x7fffffffu) - 0x2B800000u; }
constexpr unsigned RANGE_SPAN = 0x53800000u - 0x2B800000u;   // 2^-40 <= |x| < 2^40

// q = RN(a / b) from a refined reciprocal seed y of b (two residual corrections; see g2s_math.cuh dvd_y); the caller
// checks the operand ranges
__device__ __forceinline__ float div_core(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// [nr] kernel 1 (tri_face_inv) with the nine divisions sharing one reciprocal and ONE merged operand-range check;
// out-of-range operands re-run on the IEEE path.  Same bits as tri_face_inv.
__device__ __forceinline__ void face_record(const Tri& f, int is, float* rec) {
    const float p00 = ndc_to_pix(f.x0, is), p01 = ndc_to_pix(f.y0, is);
    const float p10 = ndc_to_pix(f.x1, is), p11 = ndc_to_pix(f.y1, is);
    const float p20 = ndc_to_pix(f.x2, is), p21 = ndc_to_pix(f.y2, is);
    float a[9];
    a[0] = sub(p11, p21); a[1] = sub(p20, p10); a[2] = sub(mul(p10, p21), mul(p20, p11));
    a[3] = sub(p21, p01); a[4] = sub(p00, p20); a[5] = sub(mul(p20, p01), mul(p00, p21));
    a[6] = sub(p01, p11); a[7] = sub(p10, p00); a[8] = sub(mul(p00, p11), mul(p10, p01));
    const float den = add(add(mul(p20, sub(p01, p11)), mul(p00, sub(p11, p21))), mul(p10, sub(p21, p01)));
    const float y = rcp_seed(den);
#pragma unroll
    for (int k = 0; k < 9; k++) rec[k] = div_core(a[k], den, y);
    // Operand ranges from the structure of the values instead of one test per numerator: every p is a multiple of 2^-25
    // (ndc_to_pix ends in `0.5 * (X - 1)` with X a multiple of 2^-24), so a non-zero numerator -- a difference of p's or of
    // products of p's -- is at least 2^-50 in magnitude; with max|p| < 2^19 it is below 2^39.  div_core is exact for such
    // numerators over a denominator in [2^-40, 2^40) (g2s_selftest_division covers numerators down to 2^-60).
    const float pmax = fmaxf(fmaxf(fmaxf(fabsf(p00), fabsf(p01)), fmaxf(fabsf(p10), fabsf(p11))), fmaxf(fabsf(p20), fabsf(p21)));
    if (!(pmax < 524288.0f) || range_key(den) >= RANGE_SPAN) {
        float fi[9];
        tri_face_inv(f, is, fi);
#pragma unroll
        for (int k = 0; k < 9; k++) rec[k] = fi[k];
    }
    rec[9] = f.z0; rec[10] = f.z1; rec[11] = f.z2;
#if G2S_FT_SEEDS
    rec[12] = rcp_seed(f.z0); rec[13] = rcp_seed(f.z1); rec[14] = rcp_seed(f.z2);
#endif
    // the operand-range verdict on the three z's is taken once per face, not once per hit
    rec[FT_FLAG] = max(max(range_key(f.z0), range_key(f.z1)), range_key(f.z2)) >= RANGE_SPAN ? 1.0f : 0.0f;
}

// Per-hit evaluation from a face record: [nr] kernel 2 after the inside test (clamped, renormalised weights and
// perspective z).  Fast path: the seven divisions run as residual-corrected products with shared / tabulated
// reciprocals and ONE merged operand-range check; anything out of range re-runs on the IEEE path.
__device__ __forceinline__ bool record_weights_depth(const float* rec, int xi, int yi, float near, float far,
                                                     float w[3], float* zp_out) {
    const float fx = (float)xi, fy = (float)yi;
    float wc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float v = add(add(mul(rec[3 * k], fx), mul(rec[3 * k + 1], fy)), rec[3 * k + 2]);
        wc[k] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
    const float w_sum = add(add(add(0.0f, wc[0]), wc[1]), wc[2]);
    const float ys = rcp_seed(w_sum);
    float t[3];
    // Operand ranges of the seven residual-corrected divisions, checked with the structure of the values instead of one
    // generic test per operand (those tests were a third of the per-hit instructions):
    //   wc[k] in [0,1]: only 0 < wc < 2^-38 is out -- as unsigned integers, bits(wc) - 1 wraps 0 to the top;
    //   w_sum in [max wc, 3]: in range as soon as one wc is, 0 when all are (-> IEEE path, 0/0);
    //   w[k] = wc[k] / w_sum in {0} U [2^-40, 1] follows;  z[k]: the per-face verdict rec[FT_FLAG];  s: checked below.
    unsigned lo = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float q = __fmul_rn(wc[k], ys);
        float r = __fmaf_rn(-w_sum, q, wc[k]);
        q = __fmaf_rn(r, ys, q);
        r = __fmaf_rn(-w_sum, q, wc[k]);
        w[k] = __fmaf_rn(r, ys, q);
        const float z = rec[9 + k], yz = G2S_FT_SEEDS ? rec[12 + k] : rcp_seed(z);
        q = __fmul_rn(w[k], yz);
        r = __fmaf_rn(-z, q, w[k]);
        q = __fmaf_rn(r, yz, q);
        r = __fmaf_rn(-z, q, w[k]);
        t[k] = __fmaf_rn(r, yz, q);
        lo = min(lo, __float_as_uint(wc[k]) - 1u);
    }
    const float s = add(add(t[0], t[1]), t[2]);
    float zp;
    {
        const float y = rcp_seed(s);
        float q = y;  // 1 * y
        float r = __fmaf_rn(-s, q, 1.0f);
        q = __fmaf_rn(r, y, q);
        r = __fmaf_rn(-s, q, 1.0f);
        zp = __fmaf_rn(r, y, q);
    }
    const bool fast = lo >= 0x2C800000u - 1u && w_sum > 0.0f && rec[FT_FLAG] == 0.0f && range_key(s) < RANGE_SPAN;
    if (!fast) {   // some operand outside its range: IEEE path
        Tri f;
        f.z0 = rec[9]; f.z1 = rec[10]; f.z2 = rec[11];
        float fi[9];
#pragma unroll
        for (int k = 0; k < 9; k++) fi[k] = rec[k];
        return tri_weights_depth(f, fi, xi, yi, near, far, w, zp_out);
    }
    if (zp <= near || far <= zp) return false;
    *zp_out = zp;
    return true;
}

// warp-aggregated slot allocation in a shared-memory queue (callable from divergent code)
__device__ __forceinline__ int queue_alloc(int* counter) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// What the forward kernel does with candidates and hits: candidate test = inside test; hit = atomicMin of the
// packed key.  (The backward needs no scan: the face-index map already names the owner of every sub-pixel.)
struct FwdOps {
    unsigned long long* zb;
    float near, far;
    int is;
    PixCenter pc;
    // candidate scan state of one face: [nr] kernel 2 inside test with the loop invariants hoisted
    struct Scan {
        float x0, x1, x2, dx01, dy01, dx12, dy12, dx20, dy20, y0, y1, y2, a0, a1, a2;
        __device__ __forceinline__ void init(const FwdOps&, const Tri& f, int) {
            x0 = f.x0; x1 = f.x1; x2 = f.x2; y0 = f.y0; y1 = f.y1; y2 = f.y2;
            dx01 = sub(f.x1, f.x0); dy01 = sub(f.y1, f.y0);
            dx12 = sub(f.x2, f.x1); dy12 = sub(f.y2, f.y1);
            dx20 = sub(f.x0, f.x2); dy20 = sub(f.y0, f.y2);
        }
        __device__ __forceinline__ void row_y(float yp) {
            a0 = mul(sub(yp, y0), dx01); a1 = mul(sub(yp, y1), dx12); a2 = mul(sub(yp, y2), dx20);
        }
        __device__ __forceinline__ bool test_x(float xp) const {
            return !(a0 < mul(sub(xp, x0), dy01)) && !(a1 < mul(sub(xp, x1), dy12)) && !(a2 < mul(sub(xp, x2), dy20));
        }
        // per-column halves of the three edge tests, and the test from them (same values as test_x)
        __device__ __forceinline__ void col_terms(float xp, float c[3]) const {
            c[0] = mul(sub(xp, x0), dy01); c[1] = mul(sub(xp, x1), dy12); c[2] = mul(sub(xp, x2), dy20);
        }
        __device__ __forceinline__ bool test_terms(const float c[3]) const {
            return !(a0 < c[0]) && !(a1 < c[1]) && !(a2 < c[2]);
        }
        __device__ __forceinline__ void row(const FwdOps& o, int yi) { row_y(o.pc(yi)); }
        __device__ __forceinline__ bool test(const FwdOps& o, int xi) const { return test_x(o.pc(xi)); }
    };
    // both triangles of a quad scanned over one box: bit 0 = first triangle, bit 1 = second
    struct QuadScan {
        Scan A, B;
        __device__ __forceinline__ void init(const FwdOps& o, const Tri& fA, int faceA, const Tri& fB, int faceB) {
            A.init(o, fA, faceA); B.init(o, fB, faceB);
        }
        __device__ __forceinline__ void row(const FwdOps& o, int yi) {
            const float yp = o.pc(yi);
            A.row_y(yp); B.row_y(yp);
        }
        __device__ __forceinline__ unsigned test(const FwdOps& o, int xi) const {
            const float xp = o.pc(xi);
            return (A.test_x(xp) ? 1u : 0u) | (B.test_x(xp) ? 2u : 0u);
        }
    };
    __device__ __forceinline__ void hit(const float* rec, int code, int face, int xi, int yi) const {
        float w[3], zp;
        if (record_weights_depth(rec, xi, yi, near, far, w, &zp))
            atomicMin(&zb[(long)(is - 1 - yi) * is + xi], zkey_pack(zp, (uint32_t)face));
    }
    __device__ __forceinline__ void hit_direct(const float* rec, int code, int face, int xi, int yi) const {
        hit(rec, code, face, xi, yi);
    }
};

// Rare path: evaluate one hit without the queues / face table (queue overflow, or both windings of one
// triangle front-facing, which only happens for degenerate triangles).
template <class Ops>
__device__ __noinline__ void hit_inline(const Ops& ops, const Tri& f, int code, int face, int xi, int yi, int is) {
#ifdef G2S_COUNT_SLOW
    atomicAdd(&g_slow_counters[0], 1ull);
#endif
    float rec[16];
    face_record(f, is, rec);
    ops.hit_direct(rec, code, face, xi, yi);
}

template <class Ops>
__device__ __forceinline__ bool push_hit(TileSmem& sm, const Ops& ops, const Tri& f, int code, int face, int xi,
                                         int yi, int is) {
    const int slot = queue_alloc(&sm.n_hq);
    if (slot < HQ_CAP) {
        sm.hq_pix[slot] = ((uint32_t)yi << 16) | (uint32_t)xi;
        sm.hq_code[slot] = (uint16_t)code;
        return true;
    }
    hit_inline(ops, f, code, face, xi, yi, is);
    return false;
}

// warp-wide exclusive prefix sum of a small per-lane count + total
__device__ __forceinline__ int warp_excl_scan(int v, int* total) {
    const int lane = threadIdx.x & 31;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

constexpr int SB = 4;   // small boxes: at most SB x SB sub-pixels, scanned with uniform control flow

// conservative x-extent [xa, xb] (sub-pixel columns) of the triangle on sub-pixel row yi, from the pixel-space
// vertices; falls back to the whole box when the row misses every edge numerically
__device__ __forceinline__ void row_extent(const float px[3], const float py[3], int yi, const BBox& bb, int* xa,
                                           int* xb) {
    const float y = (float)yi;
    float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const float ax = px[e], ay = py[e], bx = px[(e + 1) % 3], by = py[(e + 1) % 3];
        const float ymin = fminf(ay, by), ymax = fmaxf(ay, by);
        if (y < ymin - 0.75f || y > ymax + 0.75f) continue;
        // the edge's x-range over the band [y - 0.75, y + 0.75]
        const float dy = by - ay;
        float x0 = fminf(ax, bx), x1 = fmaxf(ax, bx);
        if (fabsf(dy) > 1e-3f) {
            const float inv = 1.0f / dy;
            const float ta = fminf(fmaxf((y - 0.75f - ay) * inv, 0.f), 1.f), tb = fminf(fmaxf((y + 0.75f - ay) * inv, 0.f), 1.f);
            const float xa_ = ax + ta * (bx - ax), xb_ = ax + tb * (bx - ax);
            x0 = fminf(xa_, xb_); x1 = fmaxf(xa_, xb_);
        }
        lo = fminf(lo, x0); hi = fmaxf(hi, x1);
    }
    if (lo > hi) { *xa = bb.x0; *xb = bb.x1; return; }
    *xa = max(bb.x0, (int)floorf(lo - 1.0f));
    *xb = min(bb.x1, (int)ceilf(hi + 1.0f));
}

template <class Ops>
__device__ __noinline__ void scan_degenerate(const Ops& ops, const Tri fr, const BBox bb, int code_r, int Q, int S,
                                             int ty0, int tx0) {
    const int face_r = code_face(code_r, Q, S, ty0, tx0), is = 2 * S;
    typename Ops::Scan sr;
    sr.init(ops, fr, face_r);
    for (int yi = bb.y0; yi <= bb.y1; yi++) {
        sr.row(ops, yi);
        for (int xi = bb.x0; xi <= bb.x1; xi++)
            if (sr.test(ops, xi)) hit_inline(ops, fr, code_r, face_r, xi, yi, is);
    }
}

// One triangle of a quad: which winding is front-facing, its box, whether both windings pass (degenerate).
struct TriClass {
    Tri f;        // in the order of the front winding
    BBox bb;
    int rev;
    bool act, dup;
};
__device__ __forceinline__ TriClass classify(const float* sv, int qy, int qx, int tri, int is, bool quad_ok) {
    TriClass c;
    c.f = tile_winding(sv, qy, qx, tri);
    c.bb.x0 = c.bb.y0 = 0; c.bb.x1 = c.bb.y1 = -1;
    const bool boxed = quad_ok && tri_bbox(c.f, is, c.bb);   // the fill_back copy has the same box
    const bool front0 = boxed && !tri_is_back(c.f);
    const bool front1 = boxed && !tri_is_back(reversed(c.f));
    c.rev = front0 ? 0 : 1;
    if (c.rev) c.f = reversed(c.f);
    c.act = front0 || front1;
    c.dup = front0 && front1;
    return c;
}

// queued (medium / large box) face record, aliased on the face table (which is only written after these phases)
constexpr int REC_STRIDE = 9;   // x0,y0,x1,y1,x2,y2, box x (x0 | x1<<16), box y, code
__device__ __forceinline__ void rec_store(float* recs, int slot, const TriClass& c, int code) {
    float* r = &recs[slot * REC_STRIDE];
    r[0] = c.f.x0; r[1] = c.f.y0; r[2] = c.f.x1; r[3] = c.f.y1; r[4] = c.f.x2; r[5] = c.f.y2;
    r[6] = __uint_as_float((uint32_t)c.bb.x0 | ((uint32_t)c.bb.x1 << 16));
    r[7] = __uint_as_float((uint32_t)c.bb.y0 | ((uint32_t)c.bb.y1 << 16));
    r[8] = __uint_as_float((uint32_t)code);
}
__device__ __forceinline__ void rec_load(const float* recs, int slot, Tri& f, BBox& bb) {
    const float* r = &recs[slot * REC_STRIDE];
    f.x0 = r[0]; f.y0 = r[1]; f.x1 = r[2]; f.y1 = r[3]; f.x2 = r[4]; f.y2 = r[5];
    f.z0 = f.z1 = f.z2 = 0.f;
    const uint32_t bx = __float_as_uint(r[6]), by = __float_as_uint(r[7]);
    bb.x0 = (int)(bx & 0xffffu); bb.x1 = (int)(bx >> 16);
    bb.y0 = (int)(by & 0xffffu); bb.y1 = (int)(by >> 16);
}


// hits of one warp step -> hit queue (one atomic per warp); every lane calls with its mask (bit k = hit at column
// x0 + k of row yi) and code
template <class Ops>
__device__ __forceinline__ void push_row_masks(TileSmem& sm, const Ops& ops, unsigned mask, int x0, int yi, int code,
                                               int face, int is) {
    int total;
    const int off = warp_excl_scan(__popc(mask), &total);
    if (total == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&sm.n_hq, total);
    base = __shfl_sync(0xffffffffu, base, 0) + off;
    while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        if (base < HQ_CAP) {
            sm.hq_pix[base] = ((uint32_t)yi << 16) | (uint32_t)(x0 + bit);
            sm.hq_code[base] = (uint16_t)code;
        } else {
            hit_inline(ops, code_tri(sm.sv, code), code, face, x0 + bit, yi, is);
        }
        base++;
    }
}

// Rare path (task queue full): scan one row segment on the spot.
template <class Ops>
__device__ __noinline__ void scan_row_inline(const Ops& ops, const Tri f, int code, int face, int yi, int xa, int xb,
                                             int is) {
#ifdef G2S_COUNT_SLOW
    atomicAdd(&g_slow_counters[1], 1ull);
#endif
    typename Ops::Scan sc;
    sc.init(ops, f, face);
    sc.row(ops, yi);
    for (int xi = xa; xi <= xb; xi++)
        if (sc.test(ops, xi)) hit_inline(ops, f, code, face, xi, yi, is);
}

// Queue the rows of one face as tasks of at most 8 columns; wide boxes only queue the segments that overlap the
// conservative per-row extent of the triangle.
template <class Ops>
__device__ __noinline__ void push_row_tasks(TileSmem& sm, const Ops& ops, float* recs, const TriClass c, int code,
                                            int face, int is) {
    rec_store(recs, code & 511, c, code);
    // the rows are expanded into tasks later by whole warps (expand_queued_faces): wide boxes one lane per row with
    // per-row extents, the others one lane per face
    const int bw = c.bb.x1 - c.bb.x0 + 1;
    if (bw > 16) sm.wq[atomicAdd(&sm.n_wq, 1)] = (uint16_t)code;
    else sm.mq[atomicAdd(&sm.n_mq, 1)] = (uint16_t)code;
}

// Rows of the queued wide faces -> row tasks: one warp per face, one lane per row; a row only queues the 8-column
// segments that overlap the conservative extent of the triangle on that row.
template <class Ops>
__device__ __forceinline__ void expand_queued_faces(TileSmem& sm, const Ops& ops, const float* recs, int Q, int S, int ty0,
                                                    int tx0) {
    const int is = 2 * S, lane = threadIdx.x & 31, nw = sm.n_wq, nm = sm.n_mq;
    // medium boxes: one lane per face, all rows x ceil(width / 8) segments
    for (int e0 = (threadIdx.x >> 5) * 32; e0 < nm; e0 += SPLAT_THREADS) {
        const int e = e0 + lane;
        const int code = e < nm ? sm.mq[e] : 0;
        Tri f;
        BBox bb;
        rec_load(recs, code & 511, f, bb);
        const int nseg = (bb.x1 - bb.x0 + 8) >> 3, n = e < nm ? (bb.y1 - bb.y0 + 1) * nseg : 0;
        int total;
        const int off = warp_excl_scan(n, &total);
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.n_tq, total);
        base = __shfl_sync(0xffffffffu, base, 0) + off;
        for (int k = 0, ry = 0, sg = 0; k < n; k++) {
            if (base + k < TQ_CAP) {
                sm.tq[base + k] = (uint32_t)code | ((uint32_t)ry << 10) | ((uint32_t)sg << 22);
            } else {
                const int xa = bb.x0 + sg * 8;
                scan_row_inline(ops, code_tri(sm.sv, code), code, code_face(code, Q, S, ty0, tx0), bb.y0 + ry, xa,
                                min(xa + 7, bb.x1), is);
            }
            if (++sg == nseg) { sg = 0; ry++; }
        }
    }
    for (int e = threadIdx.x >> 5; e < nw; e += SPLAT_THREADS / 32) {
        const int code = sm.wq[e];
        Tri f;
        BBox bb;
        rec_load(recs, code & 511, f, bb);
        const float px[3] = {ndc_to_pix(f.x0, is), ndc_to_pix(f.x1, is), ndc_to_pix(f.x2, is)};
        const float py[3] = {ndc_to_pix(f.y0, is), ndc_to_pix(f.y1, is), ndc_to_pix(f.y2, is)};
        const int bh = bb.y1 - bb.y0 + 1;
        for (int r0 = 0; r0 < bh; r0 += 32) {
            const int ry = r0 + lane;
            int s0 = 0, n = 0;
            if (ry < bh) {
                int xa, xb;
                row_extent(px, py, bb.y0 + ry, bb, &xa, &xb);
                if (xa <= xb) { s0 = (xa - bb.x0) >> 3; n = ((xb - bb.x0) >> 3) - s0 + 1; }
            }
            int total;
            const int off = warp_excl_scan(n, &total);
            if (total == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.n_tq, total);
            base = __shfl_sync(0xffffffffu, base, 0) + off;
            for (int k = 0; k < n; k++) {
                if (base + k < TQ_CAP) {
                    sm.tq[base + k] = (uint32_t)code | ((uint32_t)ry << 10) | ((uint32_t)(s0 + k) << 22);
                } else {
                    const int xa = bb.x0 + (s0 + k) * 8;
                    scan_row_inline(ops, code_tri(sm.sv, code), code, code_face(code, Q, S, ty0, tx0), bb.y0 + ry, xa,
                                    min(xa + 7, bb.x1), is);
                }
            }
        }
    }
}

// Steps 2-4 of the header comment.
template <class Ops>
__device__ __forceinline__ void tile_rasterise(TileSmem& sm, const Ops& ops, const Cam& cam, int ty0, int tx0) {
    const int tid = threadIdx.x, S = cam.S, is = 2 * S, Q = (S - 1) * (S - 1), lane = tid & 31;
    const int qy = tid / TILE, qx = tid % TILE;
    const bool quad_ok = ty0 + qy < S - 1 && tx0 + qx < S - 1;
    float* recs = sm.recs;
    int tile_queues = 1;
    // ---- quads whose two triangles fit one SB x SB box: scanned by the owning thread with uniform control flow
    {
        TriClass A = classify(sm.sv, qy, qx, 0, is, quad_ok), B = classify(sm.sv, qy, qx, 1, is, quad_ok);
        const int codeA = (tid * 2) | (A.rev << 9), codeB = (tid * 2 + 1) | (B.rev << 9);
        const int faceA = A.act ? code_face(codeA, Q, S, ty0, tx0) : -2, faceB = B.act ? code_face(codeB, Q, S, ty0, tx0) : -2;
        BBox u;
        u.x0 = min(A.act ? A.bb.x0 : 1 << 20, B.act ? B.bb.x0 : 1 << 20);
        u.y0 = min(A.act ? A.bb.y0 : 1 << 20, B.act ? B.bb.y0 : 1 << 20);
        u.x1 = max(A.act ? A.bb.x1 : -1, B.act ? B.bb.x1 : -1);
        u.y1 = max(A.act ? A.bb.y1 : -1, B.act ? B.bb.y1 : -1);
        const int uw = u.x1 - u.x0 + 1, uh = u.y1 - u.y0 + 1;
        const bool any_act = A.act || B.act;
        const bool small = any_act && uw <= SB && uh <= SB;
        // Does the tile queue anything at all?  Asked HERE, where the eight warps are still in lockstep (the barrier is
        // nearly free), not after the warp-local phases, where they have drifted apart: tiles of small quads only (the
        // interior) then finish warp by warp without another barrier.
        tile_queues = __syncthreads_or(any_act && !small);
        if (any_act && !small) {
            // everything else becomes ROW TASKS: (face, row, 8-column segment), one lane each in the next phase
            if (A.act) push_row_tasks(sm, ops, recs, A, codeA, faceA, is);
            if (B.act) push_row_tasks(sm, ops, recs, B, codeB, faceB, is);
        }
        // Hoisted scan: the inside test `!((yp-yk)*dx < (xp-xk)*dy)` splits into a per-row and a per-column term per
        // edge, so a candidate costs six compares.  Loops are fully unrolled over the SB x SB box; rows / columns
        // beyond the warp-wide maximum are skipped with uniform branches; per-lane validity is one mask at the end.
        const int mh = (int)__reduce_max_sync(0xffffffffu, small ? (unsigned)uh : 0u);
        const int mw = (int)__reduce_max_sync(0xffffffffu, small ? (unsigned)uw : 0u);
        unsigned maskA = 0, maskB = 0;
        if (mh > 0) {
            typename Ops::Scan sa, sb;
            sa.init(ops, A.f, faceA);
            sb.init(ops, B.f, faceB);
            float ca[SB][3], cb[SB][3];
#pragma unroll
            for (int rx = 0; rx < SB; rx++) {
                if (rx < mw) {
                    const float xp = ops.pc(small ? min(u.x0 + rx, is - 1) : 0);
                    sa.col_terms(xp, ca[rx]);
                    sb.col_terms(xp, cb[rx]);
                }
            }
#pragma unroll
            for (int ry = 0; ry < SB; ry++) {
                if (ry < mh) {
                    const float yp = ops.pc(small ? min(u.y0 + ry, is - 1) : 0);
                    sa.row_y(yp);
                    sb.row_y(yp);
#pragma unroll
                    for (int rx = 0; rx < SB; rx++) {
                        if (rx < mw) {
                            maskA |= (sa.test_terms(ca[rx]) ? 1u : 0u) << (ry * SB + rx);
                            maskB |= (sb.test_terms(cb[rx]) ? 1u : 0u) << (ry * SB + rx);
                        }
                    }
                }
            }
            // per-lane validity: inside this quad's box, triangle active
            unsigned vm = 0;
            if (small) {
                const unsigned rowbits = (1u << uw) - 1u;
                vm = rowbits | (rowbits << SB) | (rowbits << (2 * SB)) | (rowbits << (3 * SB));
                vm &= (uh >= SB) ? 0xffffffffu : ((1u << (uh * SB)) - 1u);
            }
            maskA &= A.act ? vm : 0u;
            maskB &= B.act ? vm : 0u;
        }
        // Small quads never leave their warp: the hits go to the warp's PRIVATE slice of the hit queue (offsets from a
        // warp scan, no shared-memory atomics), every lane builds the table entries of its own two triangles from the
        // registers it already holds, and the warp drains its slice after a __syncwarp -- no CTA barrier between scan,
        // table and drain, so the eight warps of a tile drift apart and overlap their phases (the CTA-wide form spent
        // 23 % of its warp-cycles at barriers, profiles/r01_notes.md).
        {
            constexpr int WQ_CAP = HQ_CAP / (SPLAT_THREADS / 32);
            uint32_t* wq_pix = sm.hq_pix + (tid >> 5) * WQ_CAP;
            uint16_t* wq_code = sm.hq_code + (tid >> 5) * WQ_CAP;
            int total;
            warp_excl_scan(__popc(maskA) + __popc(maskB), &total);
            if (total) {
                if (maskA) face_record(A.f, is, &sm.ftab[ft_index(tid * 2) * FT_STRIDE]);
                if (maskB) face_record(B.f, is, &sm.ftab[ft_index(tid * 2 + 1) * FT_STRIDE]);
                // one pass when the warp's hits fit its slice; otherwise the first triangles, then the second ones (a
                // triangle has at most SB*SB hits), each in two halves of its box if the slice is smaller still: some
                // split always fits, so there is no overflow path here
                static_assert(WQ_CAP >= 32 * SB * SB / 2, "a warp's slice must hold half a triangle box per lane");
                constexpr bool HALVES = WQ_CAP < 32 * SB * SB;
                constexpr unsigned LOW = (1u << (SB * SB / 2)) - 1u;
                const int npass = total <= WQ_CAP ? 1 : (HALVES ? 4 : 2);
#pragma unroll 1
                for (int pass = 0; pass < npass; pass++) {
                    unsigned mA = maskA, mB = maskB;
                    if (npass > 1) {
                        const int tri = HALVES ? pass >> 1 : pass;
                        if (tri == 0) mB = 0u; else mA = 0u;
                        if (HALVES) { const unsigned keep = (pass & 1) ? ~LOW : LOW; mA &= keep; mB &= keep; }
                    }
                    int nh;
                    int base = warp_excl_scan(__popc(mA) + __popc(mB), &nh);
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        unsigned m = k ? mB : mA;
                        const int code = k ? codeB : codeA;
                        while (m) {
                            const int bit = __ffs(m) - 1;
                            m &= m - 1;
                            wq_pix[base] = ((uint32_t)(u.y0 + bit / SB) << 16) | (uint32_t)(u.x0 + (bit & (SB - 1)));
                            wq_code[base] = (uint16_t)code;
                            base++;
                        }
                    }
                    __syncwarp();
                    for (int i = lane; i < nh; i += 32) {
                        const int code = wq_code[i];
                        const uint32_t pix = wq_pix[i];
                        ops.hit(&sm.ftab[ft_index(code) * FT_STRIDE], code, code_face(code, Q, S, ty0, tx0),
                                (int)(pix & 0xffffu), (int)(pix >> 16));
                    }
                    __syncwarp();
                }
            }
        }
        // degenerate triangles whose two windings both pass the back-face test (rounding): the reversed copy
        // bypasses the queues and the face table, whose slot the first winding owns
        if (A.dup) scan_degenerate(ops, reversed(A.f), A.bb, (tid * 2) | (1 << 9), Q, S, ty0, tx0);
        if (B.dup) scan_degenerate(ops, reversed(B.f), B.bb, (tid * 2 + 1) | (1 << 9), Q, S, ty0, tx0);
    }
    // ---- rounds: scan as many queued row tasks as are guaranteed to fit the hit queue (8 hits per task at most),
    // build the table entries of the faces that scored for the first time, drain.  One round for ordinary tiles;
    // tiles full of long wall faces take several.
    if (!tile_queues) return;   // interior tiles: nothing but small quads, no barrier
    __syncthreads();
    // Table entries of ALL queued faces now, one thread per face (from the top thread down: the expansion below keeps the
    // low warps busy), instead of lazily for the faces that scored in a round: that phase kept one or two warps busy while
    // six waited at its barrier, every round.
    {
        const int nw_ = sm.n_wq, nq_ = nw_ + sm.n_mq;
        for (int e = SPLAT_THREADS - 1 - tid; e < nq_; e += SPLAT_THREADS) {
            const int code = e < nw_ ? sm.wq[e] : sm.mq[e - nw_];
            face_record(code_tri(sm.sv, code), is, &sm.ftab[ft_index(code) * FT_STRIDE]);
        }
    }
    expand_queued_faces(sm, ops, recs, Q, S, ty0, tx0);
    const uint32_t* tq = sm.tq;
    int t0 = 0;
    while (true) {
        __syncthreads();
        const int nt = min(sm.n_tq, TQ_CAP);
        const int t1 = min(nt, t0 + (HQ_CAP - min(sm.n_hq, HQ_CAP)) / 8);
        // row tasks: one lane per (face, row, 8-column segment); uniform 8-column scan
        for (int i0 = t0 + (tid >> 5) * 32; i0 < t1; i0 += SPLAT_THREADS) {
            const int i = i0 + lane;
            const bool valid = i < t1;
            const uint32_t task = valid ? tq[i] : 0u;
            const int code = (int)(task & 1023u), ry = (int)((task >> 10) & 4095u), seg = (int)(task >> 22);
            Tri f;
            BBox bb;
            rec_load(recs, code & 511, f, bb);
            const int face = code_face(code, Q, S, ty0, tx0);
            typename Ops::Scan sc;
            sc.init(ops, f, face);
            const int yi = valid ? bb.y0 + ry : 0, x0 = valid ? bb.x0 + seg * 8 : 0;
            const int ncol = valid ? min(8, bb.x1 - x0 + 1) : 0;
            sc.row(ops, yi);
            unsigned mask = 0;
#pragma unroll
            for (int rx = 0; rx < 8; rx++) {
                const bool in = rx < ncol && sc.test(ops, min(x0 + rx, is - 1));
                mask |= (in ? 1u : 0u) << rx;
            }
            push_row_masks(sm, ops, mask, x0, yi, code, face, is);
        }
        __syncthreads();
        // drain the hit queue: one thread per hit
        const int nh = min(sm.n_hq, HQ_CAP);
        for (int i = tid; i < nh; i += SPLAT_THREADS) {
            const int code = sm.hq_code[i];
            const uint32_t pix = sm.hq_pix[i];
            ops.hit(&sm.ftab[ft_index(code) * FT_STRIDE], code, code_face(code, Q, S, ty0, tx0), (int)(pix & 0xffffu),
                    (int)(pix >> 16));
        }
        t0 = t1;
        if (t0 >= nt) break;
        __syncthreads();
        if (tid == 0) sm.n_hq = 0;
    }
}

}  // namespace g2s
