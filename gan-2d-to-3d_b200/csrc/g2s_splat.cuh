// g2s_splat.cuh -- per-face / per-hit arithmetic shared by the two stages of the forward rasteriser (g2s_tile.cuh,
// g2s_bigface.cuh) and by the pixel-centric backward (k_raster_bwd_px in g2s_kernels.cu): the 3x3 face inverse and the
// clamped, renormalised weights + perspective z of neural_renderer's forward_face_index_map, with every division a
// correctly rounded quotient built from a shared reciprocal (bit-identical to the IEEE formulation: g2s_selftest_raster).
#pragma once
#include "g2s_raster.cuh"

namespace g2s {

constexpr int FT_FLAG = 12;     // index of the per-face z-range verdict in a face record
constexpr int FT_STRIDE = 13;   // fi[9], z[3], verdict; ODD so that lanes on consecutive records hit distinct banks

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
    // the operand-range verdict on the three z's is taken once per face, not once per hit
    rec[FT_FLAG] = max(max(range_key(f.z0), range_key(f.z1)), range_key(f.z2)) >= RANGE_SPAN ? 1.0f : 0.0f;
}

// Per-hit evaluation from a face record: [nr] kernel 2 after the inside test (clamped, renormalised weights and
// perspective z).  Fast path: the seven divisions run as residual-corrected products with shared / tabulated
// reciprocals and ONE merged operand-range check; anything out of range re-runs on the IEEE path.
// fi = the 3x3 face inverse, z = the three vertex depths, zbad = the per-face operand-range verdict on z.
__device__ __forceinline__ bool weights_depth_core(const float fi[9], const float z[3], bool zbad, int xi, int yi,
                                                   float near, float far, float w[3], float* zp_out) {
    const float fx = (float)xi, fy = (float)yi;
    float wc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float v = add(add(mul(fi[3 * k], fx), mul(fi[3 * k + 1], fy)), fi[3 * k + 2]);
        wc[k] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
    const float w_sum = add(add(add(0.0f, wc[0]), wc[1]), wc[2]);
    const float ys = rcp_seed(w_sum);
    float t[3];
    // Operand ranges of the seven residual-corrected divisions, checked with the structure of the values instead of one
    // generic test per operand (those tests were a third of the per-hit instructions):
    //   wc[k] in [0,1]: only 0 < wc < 2^-38 is out -- as unsigned integers, bits(wc) - 1 wraps 0 to the top;
    //   w_sum in [max wc, 3]: in range as soon as one wc is, 0 when all are (-> IEEE path, 0/0);
    //   w[k] = wc[k] / w_sum in {0} U [2^-40, 1] follows;  z[k]: the per-face verdict zbad;  s: checked below.
    unsigned lo = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float q = __fmul_rn(wc[k], ys);
        float r = __fmaf_rn(-w_sum, q, wc[k]);
        q = __fmaf_rn(r, ys, q);
        r = __fmaf_rn(-w_sum, q, wc[k]);
        w[k] = __fmaf_rn(r, ys, q);
        const float yz = rcp_seed(z[k]);
        q = __fmul_rn(w[k], yz);
        r = __fmaf_rn(-z[k], q, w[k]);
        q = __fmaf_rn(r, yz, q);
        r = __fmaf_rn(-z[k], q, w[k]);
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
    const bool fast = lo >= 0x2C800000u - 1u && w_sum > 0.0f && !zbad && range_key(s) < RANGE_SPAN;
    if (!fast) {   // some operand outside its range: IEEE path
        Tri f;
        f.z0 = z[0]; f.z1 = z[1]; f.z2 = z[2];
        return tri_weights_depth(f, fi, xi, yi, near, far, w, zp_out);
    }
    if (zp <= near || far <= zp) return false;
    *zp_out = zp;
    return true;
}

// the same from a record in memory: rec[0..8] = fi, rec[9..11] = z, rec[FT_FLAG] = the z-range verdict
__device__ __forceinline__ bool record_weights_depth(const float* rec, int xi, int yi, float near, float far,
                                                     float w[3], float* zp_out) {
    float fi[9], z[3];
#pragma unroll
    for (int k = 0; k < 9; k++) fi[k] = rec[k];
#pragma unroll
    for (int k = 0; k < 3; k++) z[k] = rec[9 + k];
    return weights_depth_core(fi, z, rec[FT_FLAG] != 0.0f, xi, yi, near, far, w, zp_out);
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

// Warp sums of 2^LOG per-lane values by RECURSIVE HALVING: at every step a lane hands one half of its values to its
// partner and keeps the other, so 2^LOG values cost 2^LOG - 1 + (5 - LOG) shuffles instead of 5 * 2^LOG butterflies
// (16 instead of 80 for the 12 entries of grad_R | grad_t; the butterflies were 25 % of k_render_bwd_pixel's
// instructions, profiles/r01_notes.md).  Afterwards every lane holds the warp total of value number halving_index().
template <int LOG>
__device__ __forceinline__ int halving_index(int lane) { return (lane >> (5 - LOG)) & ((1 << LOG) - 1); }
template <int LOG>
__device__ __forceinline__ float warp_reduce_halving(float (&v)[1 << LOG], int lane) {
#pragma unroll
    for (int s = 0; s < LOG; s++) {
        const int h = (1 << LOG) >> (s + 1), bit = 16 >> s;
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < h; k++) {
            const float send = up ? v[k] : v[k + h], keep = up ? v[k + h] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    float r = v[0];
#pragma unroll
    for (int bit = 16 >> LOG; bit > 0; bit >>= 1) r += __shfl_xor_sync(0xffffffffu, r, bit);
    return r;
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

}  // namespace g2s
