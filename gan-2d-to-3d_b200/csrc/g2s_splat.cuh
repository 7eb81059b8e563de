// g2s_splat.cuh -- the two tile rasterisation kernels (forward splat, backward gather) of the grid mesh.
//
// Replaces neural_renderer's forward_face_index_map (every sub-pixel loops over every face) and
// backward_depth_map (one float atomicAdd x 9 per covered sub-pixel into grad_faces[B,F,3,3]).
//
// Structure (both kernels): one CTA per TILE x TILE block of quads of one view.
//   1. project the tile's (TILE+1)^2 vertices once into shared memory (u, v, z);
//   2. one thread per quad walks the sub-pixel boxes of its front-facing windings and runs only the
//      cheap candidate test (forward: the three edge functions; backward: face_idx == this face).
//      Candidates that pass are appended to a HIT QUEUE in shared memory; faces with a large box go to a
//      second queue and are scanned by whole warps;
//   3. the faces that own at least one hit get their 3x3 inverse computed ONCE, by consecutive threads,
//      into a shared-memory face table;
//   4. the hit queue is drained by consecutive threads: weights, perspective z, then the 64-bit
//      atomicMin into the z-buffer (forward) or the per-face gradient accumulators (backward).
// Steps 3 and 4 hold the IEEE divisions -- ~90% of the instructions -- and run fully converged; in the
// naive per-thread form the same code ran at 3-7 active lanes per warp (ncu, profiles/).
#pragma once
#include "g2s_raster.cuh"

namespace g2s {

constexpr int HQ_CAP = 2048;                 // hit-queue entries per drain
constexpr int NSLOT = 2 * TILE * TILE;       // (quad, triangle) slots of the face table
constexpr int FT_STRIDE = 16;                // fi[9], z[3], rcp_seed(z)[3], pad

struct TileSmem {
    float ftab[NSLOT * FT_STRIDE];
    float sv[TV * TV * 3];
    uint32_t hq_pix[HQ_CAP];
    uint16_t hq_code[HQ_CAP];
    uint16_t fq[NSLOT];
    uint16_t lq[2 * NSLOT];
    float sRt[12];
    int n_hq, n_fq, n_lq;
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

template <bool FROM_VERTS>
__device__ __forceinline__ void tile_project(const Cam& cam, const float* __restrict__ depth_b,
                                             const float* __restrict__ verts_b, const float* sRt, int ty0,
                                             int tx0, float* sv) {
    const int S = cam.S;
    for (int i = threadIdx.x; i < TV * TV; i += SPLAT_THREADS) {
        const int vy = ty0 + i / TV, vx = tx0 + i % TV;
        float ndc[3] = {0.f, 0.f, 0.f};
        if (vy < S && vx < S) {
            float q[3];
            if (FROM_VERTS) {
                const float* p = &verts_b[((long)vy * S + vx) * 3];
                q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
            } else {
                float ray[3];
                pixel_ray(cam, vx, vy, ray);
                warp_point(cam, sRt, sRt + 9, ray, depth_b[vy * S + vx], q);
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

__device__ __forceinline__ void face_record(const Tri& f, int is, float* rec) {
    float fi[9];
    tri_face_inv(f, is, fi);
#pragma unroll
    for (int k = 0; k < 9; k++) rec[k] = fi[k];
    rec[9] = f.z0; rec[10] = f.z1; rec[11] = f.z2;
    rec[12] = rcp_seed(f.z0); rec[13] = rcp_seed(f.z1); rec[14] = rcp_seed(f.z2);
}

__device__ __forceinline__ unsigned range_key(float x) { return (__float_as_uint(x) & 0x7fffffffu) - 0x2B800000u; }

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
    unsigned bad = range_key(w_sum);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float q = __fmul_rn(wc[k], ys);
        float r = __fmaf_rn(-w_sum, q, wc[k]);
        q = __fmaf_rn(r, ys, q);
        r = __fmaf_rn(-w_sum, q, wc[k]);
        w[k] = __fmaf_rn(r, ys, q);
        const float z = rec[9 + k], yz = rec[12 + k];
        q = __fmul_rn(w[k], yz);
        r = __fmaf_rn(-z, q, w[k]);
        q = __fmaf_rn(r, yz, q);
        r = __fmaf_rn(-z, q, w[k]);
        t[k] = __fmaf_rn(r, yz, q);
        bad = max(bad, range_key(z));
        bad = max(bad, wc[k] == 0.0f ? 0u : range_key(wc[k]));
        bad = max(bad, w[k] == 0.0f ? 0u : range_key(w[k]));
    }
    const float s = add(add(t[0], t[1]), t[2]);
    bad = max(bad, range_key(s));
    float zp;
    {
        const float y = rcp_seed(s);
        float q = y;  // 1 * y
        float r = __fmaf_rn(-s, q, 1.0f);
        q = __fmaf_rn(r, y, q);
        r = __fmaf_rn(-s, q, 1.0f);
        zp = __fmaf_rn(r, y, q);
    }
    if (bad >= 0x53800000u - 0x2B800000u) {   // some operand outside [2^-40, 2^40): IEEE path
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

// What a kernel does with candidates and hits.
//   FWD: candidate test = inside test; hit = atomicMin of the packed key.
//   BWD: candidate test = face-index map lookup; hit = accumulate g * zp^2 * w_k per face.
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
        __device__ __forceinline__ void row(const FwdOps& o, int yi) {
            const float yp = o.pc(yi);
            a0 = mul(sub(yp, y0), dx01); a1 = mul(sub(yp, y1), dx12); a2 = mul(sub(yp, y2), dx20);
        }
        __device__ __forceinline__ bool test(const FwdOps& o, int xi) const {
            const float xp = o.pc(xi);
            return !(a0 < mul(sub(xp, x0), dy01)) && !(a1 < mul(sub(xp, x1), dy12)) && !(a2 < mul(sub(xp, x2), dy20));
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

// scatter one face's accumulated A_k = sum g * zp^2 * w_k to its three vertices' (u,v,z) gradient
// accumulators in shared memory ([nr] backward_depth_map, factored per face)
__device__ __forceinline__ void face_scatter(const float* rec, const float A[3], int is, float* sg, int code) {
    const float z[3] = {rec[9], rec[10], rec[11]};
    // tmp[l] = -sum_m face_inv[m][l] / z_m
    const float t0 = -(rec[0] / z[0] + rec[3] / z[1] + rec[6] / z[2]);
    const float t1 = -(rec[1] / z[0] + rec[4] / z[1] + rec[7] / z[2]);
    const float hs = 0.5f * (float)is;
    const int quad = (code & 511) >> 1, qy = quad / TILE, qx = quad % TILE, w = (code & 1) + ((code >> 9) << 1);
    const int a = (qy * TV + qx) * 3, b = ((qy + 1) * TV + qx) * 3, c = (qy * TV + qx + 1) * 3,
              d = ((qy + 1) * TV + qx + 1) * 3;
    int v[3];
    switch (w) {
        case 0: v[0] = a; v[1] = b; v[2] = c; break;
        case 1: v[0] = c; v[1] = b; v[2] = d; break;
        case 2: v[0] = c; v[1] = b; v[2] = a; break;
        default: v[0] = d; v[1] = b; v[2] = c; break;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (A[k] == 0.f) continue;
        atomicAdd(&sg[v[k] + 0], -t0 * A[k] * hs);
        atomicAdd(&sg[v[k] + 1], -t1 * A[k] * hs);
        atomicAdd(&sg[v[k] + 2], A[k] / (z[k] * z[k]));
    }
}

struct BwdOps {
    const int* fmap;
    const float* gsub;
    float* sA;   // [NSLOT][3] per-face accumulators (shared memory)
    float* sg;   // [TV*TV][3] vertex (u,v,z) gradient accumulators (shared memory)
    float near, far;
    int is, S;
    struct Scan {
        const int* row_ptr;
        int face;
        __device__ __forceinline__ void init(const BwdOps&, const Tri&, int face_) { face = face_; }
        __device__ __forceinline__ void row(const BwdOps& o, int yi) { row_ptr = o.fmap + (long)(o.is - 1 - yi) * o.is; }
        __device__ __forceinline__ bool test(const BwdOps&, int xi) const { return __ldg(&row_ptr[xi]) == face; }
    };
    __device__ __forceinline__ bool contrib(const float* rec, int xi, int yi, float A[3]) const {
        const int r = is - 1 - yi;
        const float g = gsub[(r >> 1) * S + (xi >> 1)];
        if (g == 0.f) return false;
        float w[3], zp = 0.f;
        record_weights_depth(rec, xi, yi, near, far, w, &zp);
        const float s = g * zp * zp;
        A[0] = s * w[0]; A[1] = s * w[1]; A[2] = s * w[2];
        return true;
    }
    __device__ __forceinline__ void hit(const float* rec, int code, int face, int xi, int yi) const {
        float A[3];
        if (!contrib(rec, xi, yi, A)) return;
        float* acc = &sA[(code & 511) * 3];
        atomicAdd(&acc[0], A[0]);
        atomicAdd(&acc[1], A[1]);
        atomicAdd(&acc[2], A[2]);
    }
    __device__ __forceinline__ void hit_direct(const float* rec, int code, int face, int xi, int yi) const {
        float A[3];
        if (contrib(rec, xi, yi, A)) face_scatter(rec, A, is, sg, code);
    }
};

// Rare path: evaluate one hit without the queues / face table (queue overflow, or both windings of one
// triangle front-facing, which only happens for degenerate triangles).
template <class Ops>
__device__ __noinline__ void hit_inline(const Ops& ops, const Tri& f, int code, int face, int xi, int yi, int is) {
    float rec[FT_STRIDE];
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

// Steps 2-4 of the header comment.
template <class Ops>
__device__ __forceinline__ void tile_rasterise(TileSmem& sm, const Ops& ops, const Cam& cam, int ty0, int tx0) {
    const int tid = threadIdx.x, S = cam.S, is = 2 * S, Q = (S - 1) * (S - 1), lane = tid & 31;
    const int qy = tid / TILE, qx = tid % TILE;
    const bool quad_ok = ty0 + qy < S - 1 && tx0 + qx < S - 1;
    // ---- small boxes: every thread scans the (at most) SB x SB box of its triangle's front winding with uniform
    // control flow (loop bounds = warp maxima, per-lane predicates), collecting a hit mask in a register
#pragma unroll 1
    for (int tri = 0; tri < 2; tri++) {
        Tri f = tile_winding(sm.sv, qy, qx, tri);
        BBox bb;
        const bool boxed = quad_ok && tri_bbox(f, is, bb);   // the fill_back copy has the same box
        const bool front0 = boxed && !tri_is_back(f);
        const bool front1 = boxed && !tri_is_back(reversed(f));
        const int rev = front0 ? 0 : 1;
        if (rev) f = reversed(f);
        const bool front = front0 || front1;
        const int bw = boxed ? bb.x1 - bb.x0 + 1 : 0, bh = boxed ? bb.y1 - bb.y0 + 1 : 0;
        const bool large = front && (bw > SB || bh > SB);
        const bool small = front && !large;
        const int code = (tid * 2 + tri) | (rev << 9);
        const int face = code_face(code, Q, S, ty0, tx0);
        if (large) sm.lq[atomicAdd(&sm.n_lq, 1)] = (uint16_t)code;
        typename Ops::Scan sc;
        sc.init(ops, f, face);
        const int mh = (int)__reduce_max_sync(0xffffffffu, small ? (unsigned)bh : 0u);
        const int mw = (int)__reduce_max_sync(0xffffffffu, small ? (unsigned)bw : 0u);
        unsigned mask = 0;
        for (int ry = 0; ry < mh; ry++) {
            const int yi = min(bb.y0 + ry, is - 1);
            sc.row(ops, small ? yi : 0);
            for (int rx = 0; rx < mw; rx++) {
                const int xi = min(bb.x0 + rx, is - 1);
                const bool in = small && ry < bh && rx < bw && sc.test(ops, small ? xi : 0);
                mask |= (in ? 1u : 0u) << (ry * SB + rx);
            }
        }
        // queue the hits: one atomic per warp
        int total;
        const int off = warp_excl_scan(__popc(mask), &total);
        if (total) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.n_hq, total);
            base = __shfl_sync(0xffffffffu, base, 0) + off;
            unsigned m = mask;
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int xi = bb.x0 + (bit & (SB - 1)), yi = bb.y0 + bit / SB;
                if (base < HQ_CAP) {
                    sm.hq_pix[base] = ((uint32_t)yi << 16) | (uint32_t)xi;
                    sm.hq_code[base] = (uint16_t)code;
                } else {
                    hit_inline(ops, f, code, face, xi, yi, is);
                }
                base++;
            }
            const unsigned owners = __ballot_sync(0xffffffffu, mask != 0);
            int fbase = 0;
            if (lane == 0) fbase = atomicAdd(&sm.n_fq, __popc(owners));
            fbase = __shfl_sync(0xffffffffu, fbase, 0);
            if (mask) sm.fq[fbase + __popc(owners & ((1u << lane) - 1u))] = (uint16_t)code;
        }
        if (front0 && front1) {
            // degenerate triangle whose two windings both pass the back-face test (rounding): the reversed copy
            // bypasses the queues and the face table, whose slot the first winding owns
            const Tri fr = reversed(f);
            const int code_r = (tid * 2 + tri) | (1 << 9);
            const int face_r = code_face(code_r, Q, S, ty0, tx0);
            typename Ops::Scan sr;
            sr.init(ops, fr, face_r);
            for (int yi = bb.y0; yi <= bb.y1; yi++) {
                sr.row(ops, yi);
                for (int xi = bb.x0; xi <= bb.x1; xi++)
                    if (sr.test(ops, xi)) hit_inline(ops, fr, code_r, face_r, xi, yi, is);
            }
        }
    }
    __syncthreads();
    // ---- large boxes: one warp per face; wide boxes row by row over a conservative per-row extent, narrow ones
    // flattened over the box
    const int nl = sm.n_lq;
    for (int e = tid >> 5; e < nl; e += SPLAT_THREADS / 32) {
        const int code = sm.lq[e];
        const Tri f = code_tri(sm.sv, code);
        const int face = code_face(code, Q, S, ty0, tx0);
        BBox bb;
        tri_bbox(f, is, bb);
        typename Ops::Scan sc;
        sc.init(ops, f, face);
        const int bw = bb.x1 - bb.x0 + 1, n = bw * (bb.y1 - bb.y0 + 1);
        bool any = false;
        if (bw >= 24) {
            const float px[3] = {ndc_to_pix(f.x0, is), ndc_to_pix(f.x1, is), ndc_to_pix(f.x2, is)};
            const float py[3] = {ndc_to_pix(f.y0, is), ndc_to_pix(f.y1, is), ndc_to_pix(f.y2, is)};
            for (int yi = bb.y0; yi <= bb.y1; yi++) {
                int xa, xb;
                row_extent(px, py, yi, bb, &xa, &xb);
                sc.row(ops, yi);
                for (int xi = xa + lane; xi <= xb; xi += 32)
                    if (sc.test(ops, xi)) any |= push_hit(sm, ops, f, code, face, xi, yi, is);
            }
        } else {
            const float inv_bw = 1.0f / (float)bw;
            for (int idx = lane; idx < n; idx += 32) {
                const int ry = (int)(((float)idx + 0.5f) * inv_bw);
                sc.row(ops, bb.y0 + ry);
                if (sc.test(ops, bb.x0 + idx - ry * bw))
                    any |= push_hit(sm, ops, f, code, face, bb.x0 + idx - ry * bw, bb.y0 + ry, is);
            }
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0 && any) sm.fq[atomicAdd(&sm.n_fq, 1)] = (uint16_t)code;
    }
    __syncthreads();
    // ---- face table: one thread per face that owns a hit
    const int nf = sm.n_fq;
    for (int i = tid; i < nf; i += SPLAT_THREADS) {
        const int code = sm.fq[i];
        face_record(code_tri(sm.sv, code), is, &sm.ftab[(code & 511) * FT_STRIDE]);
    }
    __syncthreads();
    // ---- drain the hit queue: one thread per hit
    const int nh = min(sm.n_hq, HQ_CAP);
    for (int i = tid; i < nh; i += SPLAT_THREADS) {
        const int code = sm.hq_code[i];
        const uint32_t pix = sm.hq_pix[i];
        ops.hit(&sm.ftab[(code & 511) * FT_STRIDE], code, code_face(code, Q, S, ty0, tx0), (int)(pix & 0xffffu),
                (int)(pix >> 16));
    }
}

}  // namespace g2s
