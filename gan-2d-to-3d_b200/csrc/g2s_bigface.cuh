// g2s_bigface.cuh -- forward rasteriser of the grid mesh, stage 2 of 2: k_splat_big.
//
// Stage 1 (g2s_tile.cuh) rasterises every quad whose two triangles fit an 8 x 8 sub-pixel box and appends the rest -- the
// 1-px depth-step walls of GAN2Shape/model.py:341-344 stretch to >100 sub-pixels under yaw; degenerate triangles whose two
// windings both pass the back-face test; quads with a non-finite vertex -- to a global work list of FACES.  This kernel is
// persistent: CTAs pull batches of up to 256 faces from the list until it is empty, so the long faces are balanced over the whole
// GPU instead of sitting in the border tiles' CTAs (round 1: border tiles took 3.8 of k_splat's 10 ms and scaled worse than
// 4x from 128^2 to 256^2).  Per batch:
//   1. one thread per face: re-project its three vertices (the same exact arithmetic as stage 1), sub-pixel box, 3x3 inverse
//      -> shared-memory face table;
//   2. faces are expanded into ROW TASKS (face, row, 8-column segment) by whole warps; wide boxes only queue the segments
//      that overlap a conservative per-row extent of the triangle;
//   3. the tasks are scanned one lane per task (exact inside test, uniform 8-column loop) in rounds sized so that the hit
//      queue cannot overflow, each round drained one thread per hit: weights, perspective z, 64-bit atomicMin.
// Every consumed list entry is set back to the EMPTY key and the last CTA to finish resets the (biased) counters, so the
// workspace is uniformly EMPTY at rest.
#pragma once
#include "g2s_tile.cuh"

namespace g2s {

#ifndef G2S_BIG_THREADS
#define G2S_BIG_THREADS 256                // measured (128^2 / 256^2, ms per 4096 / 1024 views): 256 x 2 CTAs per SM 2.14 / 2.78,
#endif                                     // 128 x 4: 2.92 / 4.06, 64 x 8: 4.67 / 10.7 -- the per-batch phases amortise over large batches
#ifndef G2S_BIG_CTAS
#define G2S_BIG_CTAS 2                     // resident CTAs per SM the kernel is sized for (registers, shared memory, grid)
#endif
constexpr int BIG_THREADS = G2S_BIG_THREADS;
constexpr int BIG_HQ = 16 * BIG_THREADS;   // hit-queue entries per drain
constexpr int BIG_TQ = 32 * BIG_THREADS;   // row tasks per batch; overflow falls to an inline scan
constexpr int BREC = 11;                   // x0,y0,x1,y1,x2,y2 (NDC), box x, box y, + pad to an odd stride

struct BigSmem {
    float ftab[BIG_THREADS * FT_STRIDE];   // fi[9], z[3], z-range verdict
    float recs[BIG_THREADS * BREC];
    uint32_t tq[BIG_TQ];                   // slot | row << 8 | segment << 20
    uint32_t hq_pix[BIG_HQ];
    uint16_t hq_slot[BIG_HQ];
    uint16_t wq[BIG_THREADS], mq[BIG_THREADS];
    unsigned long long zoff[BIG_THREADS];  // offset of the face's view in the z-buffer
    uint32_t face[BIG_THREADS];
    int n_hq, n_tq, n_wq, n_mq;
    long long batch;
};

struct BigCtx {
    unsigned long long* zbuf;      // z-buffer of the launch
    float near, far;
    int is;
    PixCenter pc;
};

// [nr] kernel 2 inside test of one triangle with the loop invariants hoisted
struct RowScanT {
    float x0, x1, x2, dx01, dy01, dx12, dy12, dx20, dy20, y0, y1, y2, a0, a1, a2;
    __device__ __forceinline__ void init(const Tri& f) {
        x0 = f.x0; x1 = f.x1; x2 = f.x2; y0 = f.y0; y1 = f.y1; y2 = f.y2;
        dx01 = sub(f.x1, f.x0); dy01 = sub(f.y1, f.y0);
        dx12 = sub(f.x2, f.x1); dy12 = sub(f.y2, f.y1);
        dx20 = sub(f.x0, f.x2); dy20 = sub(f.y0, f.y2);
    }
    __device__ __forceinline__ void row(float yp) {
        a0 = mul(sub(yp, y0), dx01); a1 = mul(sub(yp, y1), dx12); a2 = mul(sub(yp, y2), dx20);
    }
    __device__ __forceinline__ bool test(float xp) const {
        return !(a0 < mul(sub(xp, x0), dy01)) && !(a1 < mul(sub(xp, x1), dy12)) && !(a2 < mul(sub(xp, x2), dy20));
    }
};

// What stage 2 does with a (face, sub-pixel) pair: candidate test = [nr] kernel 2's exact inside test, hit = weights,
// perspective z, 64-bit atomicMin.
struct BigFwd {
    struct Row {      // inside test of one face on one sub-pixel row
        RowScanT sc;
        const BigCtx* cx;
        __device__ __forceinline__ void init(const BigSmem& sm, const BigCtx& c, int slot, const Tri& f, int yi) {
            cx = &c;
            sc.init(f);
            sc.row(c.pc(yi));
        }
        __device__ __forceinline__ bool test(int xi) const { return sc.test(cx->pc(xi)); }
    };
    static __device__ __forceinline__ void hit(BigSmem& sm, const BigCtx& cx, int slot, int xi, int yi) {
        float w[3], zp;
        if (record_weights_depth(&sm.ftab[slot * FT_STRIDE], xi, yi, cx.near, cx.far, w, &zp))
            atomicMin(&cx.zbuf[sm.zoff[slot] + (unsigned long long)((long)(cx.is - 1 - yi) * cx.is + xi)],
                      zkey_pack(zp, sm.face[slot]));
    }
};

__device__ __forceinline__ void big_rec_load(const float* recs, int slot, Tri& f, BBox& bb) {
    const float* r = &recs[slot * BREC];
    f.x0 = r[0]; f.y0 = r[1]; f.x1 = r[2]; f.y1 = r[3]; f.x2 = r[4]; f.y2 = r[5];
    f.z0 = f.z1 = f.z2 = 0.f;
    const uint32_t bx = __float_as_uint(r[6]), by = __float_as_uint(r[7]);
    bb.x0 = (int)(bx & 0xffffu); bb.x1 = (int)(bx >> 16);
    bb.y0 = (int)(by & 0xffffu); bb.y1 = (int)(by >> 16);
}

// Rare path (task queue full): scan one row segment and evaluate its hits on the spot.
template <class P>
__device__ __noinline__ void big_scan_row_inline(BigSmem& sm, const BigCtx& cx, int slot, int yi, int xa, int xb) {
    Tri f;
    BBox bb;
    big_rec_load(sm.recs, slot, f, bb);
    typename P::Row row;
    row.init(sm, cx, slot, f, yi);
    for (int xi = xa; xi <= xb; xi++)
        if (row.test(xi)) P::hit(sm, cx, slot, xi, yi);
}

template <class P>
__device__ __forceinline__ void big_push_task(BigSmem& sm, const BigCtx& cx, int pos, int slot, int ry, int seg,
                                              const BBox& bb) {
    if (pos < BIG_TQ) {
        sm.tq[pos] = (uint32_t)slot | ((uint32_t)ry << 8) | ((uint32_t)seg << 20);
    } else {
        const int xa = bb.x0 + seg * 8;
        big_scan_row_inline<P>(sm, cx, slot, bb.y0 + ry, xa, min(xa + 7, bb.x1));
    }
}

// Rows of the batch's faces -> row tasks.  Medium boxes (<= 16 columns): one lane per face, all rows x segments; wide
// boxes: one warp per face, one lane per row, only the segments that overlap the conservative extent of the triangle.
template <class P>
__device__ __forceinline__ void big_expand(BigSmem& sm, const BigCtx& cx) {
    const int is = cx.is, lane = threadIdx.x & 31, nw = sm.n_wq, nm = sm.n_mq;
    for (int e0 = (threadIdx.x >> 5) * 32; e0 < nm; e0 += BIG_THREADS) {
        const int e = e0 + lane;
        const int slot = e < nm ? sm.mq[e] : 0;
        Tri f;
        BBox bb;
        big_rec_load(sm.recs, slot, f, bb);
        const int nseg = (bb.x1 - bb.x0 + 8) >> 3, n = e < nm ? (bb.y1 - bb.y0 + 1) * nseg : 0;
        int total;
        const int off = warp_excl_scan(n, &total);
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.n_tq, total);
        base = __shfl_sync(0xffffffffu, base, 0) + off;
        for (int k = 0, ry = 0, sg = 0; k < n; k++) {
            big_push_task<P>(sm, cx, base + k, slot, ry, sg, bb);
            if (++sg == nseg) { sg = 0; ry++; }
        }
    }
    for (int e = threadIdx.x >> 5; e < nw; e += BIG_THREADS / 32) {
        const int slot = sm.wq[e];
        Tri f;
        BBox bb;
        big_rec_load(sm.recs, slot, f, bb);
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
            for (int k = 0; k < n; k++) big_push_task<P>(sm, cx, base + k, slot, ry, s0 + k, bb);
        }
    }
}

// hits of one warp step -> hit queue (one shared-memory atomic per warp); bit k of `mask` = hit at column x0 + k of row yi
template <class P>
__device__ __forceinline__ void big_push_hits(BigSmem& sm, const BigCtx& cx, unsigned mask, int x0, int yi, int slot) {
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
        if (base < BIG_HQ) {
            sm.hq_pix[base] = ((uint32_t)yi << 16) | (uint32_t)(x0 + bit);
            sm.hq_slot[base] = (uint16_t)slot;
        } else {
            P::hit(sm, cx, slot, x0 + bit, yi);
        }
        base++;
    }
}

// expand the batch's faces into row tasks, scan them one lane per task, drain the hits one thread per hit
template <class P>
__device__ __forceinline__ void big_rounds(BigSmem& sm, const BigCtx& cx) {
    const int tid = threadIdx.x, lane = tid & 31, is = cx.is;
    big_expand<P>(sm, cx);
    int t0 = 0;
    while (true) {
        __syncthreads();
        const int nt = min(sm.n_tq, BIG_TQ);
        const int t1 = min(nt, t0 + (BIG_HQ - min(sm.n_hq, BIG_HQ)) / 8);
        for (int i0 = t0 + (tid >> 5) * 32; i0 < t1; i0 += BIG_THREADS) {
            const int i = i0 + lane;
            const bool valid = i < t1;
            const uint32_t task = valid ? sm.tq[i] : 0u;
            const int slot = (int)(task & 255u), ry = (int)((task >> 8) & 4095u), seg = (int)(task >> 20);
            Tri f;
            BBox bb;
            big_rec_load(sm.recs, slot, f, bb);
            const int yi = valid ? bb.y0 + ry : 0, x0 = valid ? bb.x0 + seg * 8 : 0;
            const int ncol = valid ? min(8, bb.x1 - x0 + 1) : 0;
            typename P::Row row;
            row.init(sm, cx, slot, f, yi);
            unsigned mask = 0;
#pragma unroll
            for (int rx = 0; rx < 8; rx++) {
                const bool in = rx < ncol && row.test(min(x0 + rx, is - 1));
                mask |= (in ? 1u : 0u) << rx;
            }
            big_push_hits<P>(sm, cx, mask, x0, yi, slot);
        }
        __syncthreads();
        const int nh = min(sm.n_hq, BIG_HQ);
        for (int i = tid; i < nh; i += BIG_THREADS) {
            const uint32_t pix = sm.hq_pix[i];
            P::hit(sm, cx, sm.hq_slot[i], (int)(pix & 0xffffu), (int)(pix >> 16));
        }
        t0 = t1;
        if (t0 >= nt) break;
        __syncthreads();
        if (tid == 0) sm.n_hq = 0;
    }
}

// projected (u, v, z) of vertex `v` of view `b`
template <bool FROM_VERTS>
__device__ __forceinline__ void project_vertex(const Cam& cam, const float* __restrict__ depth, long dstride, int vpi,
                                               const float* __restrict__ R, const float* __restrict__ t,
                                               const float* __restrict__ verts3d, int b, int v, float ndc[3]) {
    const int S = cam.S;
    float q[3];
    if (FROM_VERTS) {
        const float* p = verts3d + ((long)b * S * S + v) * 3;
        q[0] = __ldg(p); q[1] = __ldg(p + 1); q[2] = __ldg(p + 2);
    } else {
        float Rt[12], ray[3];
#pragma unroll
        for (int k = 0; k < 9; k++) Rt[k] = __ldg(&R[(long)b * 9 + k]);
#pragma unroll
        for (int k = 0; k < 3; k++) Rt[9 + k] = __ldg(&t[(long)b * 3 + k]);
        const int vy = v / S, vx = v - vy * S;
        pixel_ray(cam, vx, vy, ray);
        warp_point(cam, Rt, Rt + 9, ray, __ldg(&depth[(long)(b / vpi) * dstride + v]), q);
    }
    project_ndc(cam, q, ndc);
}

template <bool FROM_VERTS>
__device__ __forceinline__ void splat_big_body(BigSmem& sm, const Cam& cam, const float* __restrict__ depth, long dstride,
                                               int vpi, const float* __restrict__ R, const float* __restrict__ t,
                                               const float* __restrict__ verts3d, unsigned long long* zbuf,
                                               const WorkList& wl, int view0) {
    typedef BigFwd P;
    const int tid = threadIdx.x, S = cam.S, is = 2 * S;
    BigCtx cx;
    cx.zbuf = zbuf; cx.near = cam.near; cx.far = cam.far; cx.is = is;
    cx.pc.init(is);
    const long long count = (long long)(*(volatile unsigned long long*)&wl.ctr[0] - wl.bias);   // written by stage 1
    // faces per ticket: the list spread over all CTAs of the grid (a full 256-face batch of wall faces keeps one CTA busy
    // for ~60 us while most of the grid idles), at least 32 (one face per lane of the expansion warps)
    const int per = (int)min((long long)BIG_THREADS, max(32ll, (count + gridDim.x - 1) / (long long)gridDim.x));
    while (true) {
        __syncthreads();     // the previous batch is done with the shared-memory state
        if (tid == 0) {
            sm.batch = (long long)(atomicAdd(&wl.ctr[1], (unsigned long long)per) - wl.bias);
            sm.n_hq = sm.n_tq = sm.n_wq = sm.n_mq = 0;
        }
        __syncthreads();
        const long long start = sm.batch;
        if (start >= count) break;
        const int n = (int)(count - start < per ? count - start : per);
        // one thread per (face, vertex): re-project the three vertices of every face of the batch
        for (int k = tid; k < 3 * n; k += BIG_THREADS) {
            const int e = k / 3, m = k - 3 * e;
            const unsigned long long item = wl.items[start + e];
            int vidx[3];
            face_vertices((int)(uint32_t)item, S, vidx);
            float nd[3];
            project_vertex<FROM_VERTS>(cam, depth, dstride, vpi, R, t, verts3d, view0 + (int)(item >> 32), vidx[m], nd);
            float* dst = &sm.recs[e * BREC];       // x0,y0,x1,y1,x2,y2 | z0,z1,z2 parked in 8..10 until the record is built
            dst[2 * m] = nd[0]; dst[2 * m + 1] = nd[1]; dst[8 + m] = nd[2];
        }
        __syncthreads();
        if (tid < n) {
            const unsigned long long item = wl.items[start + tid];
            wl.items[start + tid] = wl.bias;
            const int bl = (int)(item >> 32);
            const uint32_t face = (uint32_t)item;
            float* r = &sm.recs[tid * BREC];
            Tri f;
            f.x0 = r[0]; f.y0 = r[1]; f.x1 = r[2]; f.y1 = r[3]; f.x2 = r[4]; f.y2 = r[5];
            f.z0 = r[8]; f.z1 = r[9]; f.z2 = r[10];
            BBox bb;
            sm.zoff[tid] = (unsigned long long)bl * (unsigned long long)is * (unsigned long long)is;
            sm.face[tid] = face;
            if (tri_bbox(f, is, bb)) {
                r[6] = __uint_as_float((uint32_t)bb.x0 | ((uint32_t)bb.x1 << 16));
                r[7] = __uint_as_float((uint32_t)bb.y0 | ((uint32_t)bb.y1 << 16));
                face_record(f, is, &sm.ftab[tid * FT_STRIDE]);
                if (bb.x1 - bb.x0 + 1 > 16) sm.wq[atomicAdd(&sm.n_wq, 1)] = (uint16_t)tid;
                else sm.mq[atomicAdd(&sm.n_mq, 1)] = (uint16_t)tid;
            }
        }
        __syncthreads();
        big_rounds<P>(sm, cx);
    }
    // the last CTA to leave puts the counters back to their rest value
    if (tid == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(&wl.ctr[2], 1ull) - wl.bias;
        if (done == (unsigned long long)gridDim.x - 1ull) {
            wl.ctr[0] = wl.bias; wl.ctr[1] = wl.bias; wl.ctr[2] = wl.bias;
        }
    }
}

}  // namespace g2s
