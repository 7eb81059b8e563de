// g2s_callers.cuh -- kernels + C ABI for the callers either side of the renderer path:
//   * get_clamped_depth / rescale_depth            GAN2Shape/model.py:85-86, 337-345   (the depth prologue)
//   * get_shading (diffuse + texture)              GAN2Shape/model.py:355-360          (stand-alone form)
//   * validity mask + PhotometricLoss              GAN2Shape/model.py:146-150, 265-269; GAN2Shape/losses.py:39-51
//   * SmoothLoss                                   GAN2Shape/losses.py:54-79
// Included at the end of g2s_kernels.cu (one translation unit: shares Launch / the kernel ids).
//
// All of these are HBM-bound streaming / reduction kernels.  Reductions are deterministic: per-block partial sums
// (fp32 inside a thread, fp64 across threads) written to a caller-owned workspace, finished in a fixed order.
#pragma once

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1184;   // 148 SMs x 8 resident CTAs of 256 threads

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;   // valid in thread 0
}

// ---- get_clamped_depth ------------------------------------------------------------------------------------------
// model.py:337-345: d = raw - mean(raw); depth = rescale(tanh(d)); border columns (2 left, 2 right) are blended with the
// literal weight 1.02 of the reference's F.pad(value=1.02): depth * (1 - 1.02) + 1.02 * border_depth.
constexpr int CD_PARTS = 32;   // partial sums per group

__device__ __forceinline__ float cd_rescale(float th, float lo, float hi) {
    return (1.f + th) / 2.f * hi + (1.f - th) / 2.f * lo;   // model.py:85-86
}
__device__ __forceinline__ float cd_border_w(int x, int W, int clamp_border) {
    return (clamp_border && (x < 2 || x >= W - 2)) ? 1.02f : 0.f;
}
__device__ __forceinline__ double cd_group_total(const double* parts) {
    double s = 0.0;
    for (int p = 0; p < CD_PARTS; p++) s += parts[p];
    return s;
}

// MODE 0: sum of raw; MODE 1: sum of d(loss)/d(centred depth)
template <int MODE>
__global__ void __launch_bounds__(RED_THREADS) k_cd_partial(const float* __restrict__ raw, const float* __restrict__ mean,
                                                            const float* __restrict__ g_depth, long group_elems, int W,
                                                            float lo, float hi, int clamp_border, double* __restrict__ parts) {
    __shared__ double sh[RED_THREADS / 32];
    const int g = blockIdx.y, p = blockIdx.x;
    const long per = (group_elems + CD_PARTS - 1) / CD_PARTS, i0 = p * per, i1 = min(group_elems, i0 + per);
    const float* r = raw + (long)g * group_elems;
    float acc = 0.f;
    if (MODE == 0) {
        for (long i = i0 + threadIdx.x; i < i1; i += RED_THREADS) acc += r[i];
    } else {
        const float m = mean[g], half = (hi - lo) * 0.5f;
        const float* gd = g_depth + (long)g * group_elems;
        for (long i = i0 + threadIdx.x; i < i1; i += RED_THREADS) {
            const float th = tanhf(r[i] - m);
            acc += gd[i] * (1.f - cd_border_w((int)(i % W), W, clamp_border)) * half * (1.f - th * th);
        }
    }
    const double s = block_sum((double)acc, sh);
    if (threadIdx.x == 0) parts[g * CD_PARTS + p] = s;
}

__global__ void __launch_bounds__(RED_THREADS) k_cd_apply_fwd(const float* __restrict__ raw, const double* __restrict__ parts,
                                                              long group_elems, int W, float lo, float hi, float border_depth,
                                                              int clamp_border, float* __restrict__ mean_out,
                                                              float* __restrict__ depth) {
    const int g = blockIdx.y;
    const float m = (float)(cd_group_total(parts + g * CD_PARTS) / (double)group_elems);
    if (blockIdx.x == 0 && threadIdx.x == 0) mean_out[g] = m;
    const long i = (long)blockIdx.x * RED_THREADS + threadIdx.x;
    if (i >= group_elems) return;
    const float d = cd_rescale(tanhf(raw[(long)g * group_elems + i] - m), lo, hi);
    const float bw = cd_border_w((int)(i % W), W, clamp_border);
    depth[(long)g * group_elems + i] = clamp_border ? d * (1.f - bw) + bw * border_depth : d;
}

__global__ void __launch_bounds__(RED_THREADS) k_cd_apply_bwd(const float* __restrict__ raw, const float* __restrict__ mean,
                                                              const float* __restrict__ g_depth, const double* __restrict__ parts,
                                                              long group_elems, int W, float lo, float hi, int clamp_border,
                                                              float* __restrict__ g_raw) {
    const int g = blockIdx.y;
    const float gm = (float)(cd_group_total(parts + g * CD_PARTS) / (double)group_elems);
    const long i = (long)blockIdx.x * RED_THREADS + threadIdx.x;
    if (i >= group_elems) return;
    const float th = tanhf(raw[(long)g * group_elems + i] - mean[g]);
    const float gd = g_depth[(long)g * group_elems + i] * (1.f - cd_border_w((int)(i % W), W, clamp_border)) *
                     ((hi - lo) * 0.5f) * (1.f - th * th);
    g_raw[(long)g * group_elems + i] = gd - gm;
}

// ---- get_shading --------------------------------------------------------------------------------------------------
// model.py:355-360.  normal [*,HW,3] and albedo [*,3,HW] may be shared by all views (stride 0).
__global__ void __launch_bounds__(PIX_THREADS) k_shading_fwd(const float* __restrict__ normal, long n_stride,
                                                             const float* __restrict__ light5, const float* __restrict__ albedo,
                                                             long a_stride, int HW, float* __restrict__ diffuse,
                                                             float* __restrict__ texture) {
    const int b = blockIdx.y, i = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (i >= HW) return;
    const float* n = normal + b * n_stride + (long)i * 3;
    const float* L = light5 + b * 5;
    const float dif = fmaxf(n[0] * L[2] + n[1] * L[3] + n[2] * L[4], 0.f);
    const float sh = L[0] + L[1] * dif;
    if (diffuse) diffuse[(long)b * HW + i] = dif;
    const float* al = albedo + b * a_stride + i;
#pragma unroll
    for (int c = 0; c < 3; c++) texture[((long)b * 3 + c) * HW + i] = (al[(long)c * HW] / 2.f + 0.5f) * sh * 2.f - 1.f;
}

// grad_normal / grad_albedo are ACCUMULATED with atomics (stride 0 = summed over the views), grad_light5 [B,5] too
__global__ void __launch_bounds__(PIX_THREADS) k_shading_bwd(const float* __restrict__ normal, long n_stride,
                                                             const float* __restrict__ light5, const float* __restrict__ albedo,
                                                             long a_stride, int HW, const float* __restrict__ g_diffuse,
                                                             const float* __restrict__ g_texture, float* __restrict__ g_normal,
                                                             long gn_stride, float* __restrict__ g_light5,
                                                             float* __restrict__ g_albedo, long ga_stride) {
    __shared__ double sh5[5][PIX_THREADS / 32];
    const int b = blockIdx.y, i = blockIdx.x * PIX_THREADS + threadIdx.x;
    float gl[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (i < HW) {
        const float* n = normal + b * n_stride + (long)i * 3;
        const float* L = light5 + b * 5;
        const float dot = n[0] * L[2] + n[1] * L[3] + n[2] * L[4];
        const float dif = fmaxf(dot, 0.f), sh = L[0] + L[1] * dif;
        const float* al = albedo + b * a_stride + i;
        float g_sh = 0.f;
        if (g_texture) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float gt = g_texture[((long)b * 3 + c) * HW + i];
                const float a01 = al[(long)c * HW] / 2.f + 0.5f;
                g_sh += gt * a01 * 2.f;
                if (g_albedo) atomicAdd(&g_albedo[b * ga_stride + (long)c * HW + i], gt * sh);   // (1/2) * sh * 2
            }
        }
        float g_dif = g_sh * L[1] + (g_diffuse ? g_diffuse[(long)b * HW + i] : 0.f);
        if (!(dot > 0.f)) g_dif = 0.f;    // clamp(min=0): zero gradient at and below 0
        gl[0] = g_sh; gl[1] = g_sh * dif;
        gl[2] = g_dif * n[0]; gl[3] = g_dif * n[1]; gl[4] = g_dif * n[2];
        if (g_normal) {
            float* gn = g_normal + b * gn_stride + (long)i * 3;
            atomicAdd(&gn[0], g_dif * L[2]); atomicAdd(&gn[1], g_dif * L[3]); atomicAdd(&gn[2], g_dif * L[4]);
        }
    }
    if (g_light5) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            double v = (double)gl[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sh5[k][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            double s = 0.0;
            for (int w = 0; w < PIX_THREADS / 32; w++) s += sh5[threadIdx.x][w];
            atomicAdd(&g_light5[b * 5 + threadIdx.x], (float)s);
        }
    }
}

// ---- validity mask + PhotometricLoss ---------------------------------------------------------------------------
// mask[b,i] = (recon_depth ? recon_depth[b,i] < thresh : 1) * (mask_in ? mask_in[b,i] : 1)      model.py:146-150, 265-269
// loss = sum(|im1 - im2| * mask) / sum(mask.expand_as(loss))                                      losses.py:42-50
// One thread = 4 consecutive pixels (16-byte loads) x C channels; grid-stride over B * HW / 4 quads.
struct PhotoArgs {
    const float *im1, *im2, *recon_depth, *mask_in;
    long im2_stride;     // floats between batch items of im2 (0 = one target for all views)
    float thresh;
    int B, C, HW;
    const float* sigma;  // conf_sigma [B, sigma_C, HW] (losses.py:44-45) or NULL; sigma_C = 1 (shared by the channels) or C
    int sigma_C;
};
constexpr float PHOTO_EPS = 1e-7f;      // PhotometricLoss.EPS, losses.py:40
constexpr float PHOTO_SQRT2 = 1.41421356237309515f;   // float(2 ** 0.5)
// losses.py:45: |d| * 2**0.5 / (sigma + EPS) + log(sigma + EPS)
__device__ __forceinline__ float photo_sigma_term(float ad, float sg) {
    const float se = sg + PHOTO_EPS;
    return ad * PHOTO_SQRT2 / se + logf(se);
}

__device__ __forceinline__ float4 photo_mask4(const PhotoArgs& a, int b, int i) {
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (a.recon_depth) {
        const float4 d = __ldcs(reinterpret_cast<const float4*>(a.recon_depth + (long)b * a.HW + i));
        m.x = d.x < a.thresh ? 1.f : 0.f; m.y = d.y < a.thresh ? 1.f : 0.f;
        m.z = d.z < a.thresh ? 1.f : 0.f; m.w = d.w < a.thresh ? 1.f : 0.f;
    }
    if (a.mask_in) {
        const float4 k = __ldcs(reinterpret_cast<const float4*>(a.mask_in + (long)b * a.HW + i));
        m.x *= k.x; m.y *= k.y; m.z *= k.z; m.w *= k.w;
    }
    return m;
}
__device__ __forceinline__ float photo_mask1(const PhotoArgs& a, int b, int i) {
    float m = 1.f;
    if (a.recon_depth) m = a.recon_depth[(long)b * a.HW + i] < a.thresh ? 1.f : 0.f;
    if (a.mask_in) m *= a.mask_in[(long)b * a.HW + i];
    return m;
}

template <bool VEC>
__global__ void __launch_bounds__(RED_THREADS) k_photo_fwd(PhotoArgs a, double* __restrict__ parts) {
    __shared__ double sh[RED_THREADS / 32];
    double num = 0.0, den = 0.0;   // fp32 inside one quad, fp64 across quads
    if (VEC) {
        const int qpi = a.HW / 4;
        const long nq = (long)a.B * qpi;
        for (long q = (long)blockIdx.x * RED_THREADS + threadIdx.x; q < nq; q += (long)gridDim.x * RED_THREADS) {
            const int b = (int)(q / qpi), i = (int)(q % qpi) * 4;
            const float4 m = photo_mask4(a, b, i);
            den += (double)((m.x + m.y) + (m.z + m.w));
            float nq4 = 0.f;
            for (int c = 0; c < a.C; c++) {
                const float4 x = __ldcs(reinterpret_cast<const float4*>(a.im1 + ((long)b * a.C + c) * a.HW + i));
                const float4 y = __ldcs(reinterpret_cast<const float4*>(a.im2 + b * a.im2_stride + (long)c * a.HW + i));
                nq4 += (fabsf(x.x - y.x) * m.x + fabsf(x.y - y.y) * m.y) + (fabsf(x.z - y.z) * m.z + fabsf(x.w - y.w) * m.w);
            }
            num += (double)nq4;
        }
    } else {
        const long n = (long)a.B * a.HW;
        for (long q = (long)blockIdx.x * RED_THREADS + threadIdx.x; q < n; q += (long)gridDim.x * RED_THREADS) {
            const int b = (int)(q / a.HW), i = (int)(q % a.HW);
            const float m = photo_mask1(a, b, i);
            den += (double)m;
            float n1 = 0.f;
            for (int c = 0; c < a.C; c++) {
                float l = fabsf(a.im1[((long)b * a.C + c) * a.HW + i] - a.im2[b * a.im2_stride + (long)c * a.HW + i]);
                if (a.sigma) l = photo_sigma_term(l, a.sigma[((long)b * a.sigma_C + (a.sigma_C == 1 ? 0 : c)) * a.HW + i]);
                n1 += l * m;
            }
            num += (double)n1;
        }
    }
    const double sn = block_sum(num, sh);
    const double sd = block_sum(den, sh);
    if (threadIdx.x == 0) { parts[2 * blockIdx.x] = sn; parts[2 * blockIdx.x + 1] = sd; }
}

// out[0] = loss, out[1] = numerator, out[2] = denominator (mask summed over the C channels)
__global__ void __launch_bounds__(RED_THREADS) k_photo_finish(const double* __restrict__ parts, int nparts, int C,
                                                              float* __restrict__ out) {
    __shared__ double sh[RED_THREADS / 32];
    double n = 0.0, d = 0.0;
    for (int p = threadIdx.x; p < nparts; p += RED_THREADS) { n += parts[2 * p]; d += parts[2 * p + 1]; }
    n = block_sum(n, sh);
    d = block_sum(d, sh);
    if (threadIdx.x == 0) {
        d *= (double)C;
        out[0] = (float)(n / d); out[1] = (float)n; out[2] = (float)d;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(RED_THREADS) k_photo_bwd(PhotoArgs a, const float* __restrict__ sums,
                                                           const float* __restrict__ g_loss, float* __restrict__ g_im1,
                                                           float* __restrict__ g_im2, float* __restrict__ g_sigma) {
    const float s = g_loss[0] / sums[2];
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    if (VEC) {
        const int qpi = a.HW / 4;
        const long nq = (long)a.B * qpi;
        for (long q = (long)blockIdx.x * RED_THREADS + threadIdx.x; q < nq; q += (long)gridDim.x * RED_THREADS) {
            const int b = (int)(q / qpi), i = (int)(q % qpi) * 4;
            const float4 m = photo_mask4(a, b, i);
            for (int c = 0; c < a.C; c++) {
                const long o1 = ((long)b * a.C + c) * a.HW + i;
                const float4 x = __ldcs(reinterpret_cast<const float4*>(a.im1 + o1));
                const float4 y = __ldcs(reinterpret_cast<const float4*>(a.im2 + b * a.im2_stride + (long)c * a.HW + i));
                const float4 g = make_float4(sgn(x.x - y.x) * m.x * s, sgn(x.y - y.y) * m.y * s, sgn(x.z - y.z) * m.z * s,
                                             sgn(x.w - y.w) * m.w * s);
                if (g_im1) __stcs(reinterpret_cast<float4*>(g_im1 + o1), g);
                if (g_im2) __stcs(reinterpret_cast<float4*>(g_im2 + o1), make_float4(-g.x, -g.y, -g.z, -g.w));
            }
        }
    } else {
        const long n = (long)a.B * a.HW;
        for (long q = (long)blockIdx.x * RED_THREADS + threadIdx.x; q < n; q += (long)gridDim.x * RED_THREADS) {
            const int b = (int)(q / a.HW), i = (int)(q % a.HW);
            const float m = photo_mask1(a, b, i);
            float gs_shared = 0.f;
            for (int c = 0; c < a.C; c++) {
                const long o1 = ((long)b * a.C + c) * a.HW + i;
                const float d = a.im1[o1] - a.im2[b * a.im2_stride + (long)c * a.HW + i];
                float g = sgn(d) * m * s;
                if (a.sigma) {      // d/d|d| = 2**0.5 / (sigma + EPS);  d/dsigma = -|d| 2**0.5 / (sigma + EPS)^2 + 1 / (sigma + EPS)
                    const long os = ((long)b * a.sigma_C + (a.sigma_C == 1 ? 0 : c)) * a.HW + i;
                    const float se = a.sigma[os] + PHOTO_EPS;
                    g = g * PHOTO_SQRT2 / se;
                    const float gsg = m * s * (1.f / se - fabsf(d) * PHOTO_SQRT2 / (se * se));
                    if (g_sigma) { if (a.sigma_C == 1) gs_shared += gsg; else g_sigma[os] = gsg; }
                }
                if (g_im1) g_im1[o1] = g;
                if (g_im2) g_im2[o1] = -g;
            }
            if (a.sigma && g_sigma && a.sigma_C == 1) g_sigma[(long)b * a.HW + i] = gs_shared;
        }
    }
}

// ---- SmoothLoss -------------------------------------------------------------------------------------------------
// losses.py:54-79 for one map [M,H,W]: mean|dx2| + mean|dxdy| + mean|dydx| + mean|dy2|, every difference evaluated in the
// reference's order (differences of first differences).  Terms are anchored at their top-left pixel.
struct SmoothMap {
    const float* p;
    int H, W;
    __device__ __forceinline__ float at(int y, int x) const { return __ldg(p + (long)y * W + x); }
    __device__ __forceinline__ float dx(int y, int x) const { return at(y, x + 1) - at(y, x); }
    __device__ __forceinline__ float dy(int y, int x) const { return at(y + 1, x) - at(y, x); }
    // second differences; `ok` = the anchor is inside the term's domain
    __device__ __forceinline__ float dx2(int y, int x, bool* ok) const {
        *ok = y >= 0 && x >= 0 && y < H && x < W - 2;
        return *ok ? dx(y, x + 1) - dx(y, x) : 0.f;
    }
    __device__ __forceinline__ float dxdy(int y, int x, bool* ok) const {
        *ok = y >= 0 && x >= 0 && y < H - 1 && x < W - 1;
        return *ok ? dx(y + 1, x) - dx(y, x) : 0.f;
    }
    __device__ __forceinline__ float dydx(int y, int x, bool* ok) const {
        *ok = y >= 0 && x >= 0 && y < H - 1 && x < W - 1;
        return *ok ? dy(y, x + 1) - dy(y, x) : 0.f;
    }
    __device__ __forceinline__ float dy2(int y, int x, bool* ok) const {
        *ok = y >= 0 && x >= 0 && y < H - 2 && x < W;
        return *ok ? dy(y + 1, x) - dy(y, x) : 0.f;
    }
};

__global__ void __launch_bounds__(RED_THREADS) k_smooth_fwd(const float* __restrict__ map, int M, int H, int W,
                                                            double* __restrict__ parts) {
    __shared__ double sh[RED_THREADS / 32];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const long n = (long)M * H * W;
    for (long q = (long)blockIdx.x * RED_THREADS + threadIdx.x; q < n; q += (long)gridDim.x * RED_THREADS) {
        const int m = (int)(q / ((long)H * W)), r = (int)(q % ((long)H * W)), y = r / W, x = r % W;
        const SmoothMap s{map + (long)m * H * W, H, W};
        bool ok;
        acc[0] += (double)fabsf(s.dx2(y, x, &ok));
        acc[1] += (double)fabsf(s.dxdy(y, x, &ok));
        acc[2] += (double)fabsf(s.dydx(y, x, &ok));
        acc[3] += (double)fabsf(s.dy2(y, x, &ok));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double v = block_sum(acc[k], sh);
        if (threadIdx.x == 0) parts[4 * blockIdx.x + k] = v;
    }
}

// out[0] = loss, out[1..4] = the four means
__global__ void __launch_bounds__(RED_THREADS) k_smooth_finish(const double* __restrict__ parts, int nparts, int M, int H, int W,
                                                               float* __restrict__ out) {
    __shared__ double sh[RED_THREADS / 32];
    const double cnt[4] = {(double)M * H * (W - 2), (double)M * (H - 1) * (W - 1), (double)M * (H - 1) * (W - 1),
                           (double)M * (H - 2) * W};
    float loss = 0.f;
    for (int k = 0; k < 4; k++) {
        double v = 0.0;
        for (int p = threadIdx.x; p < nparts; p += RED_THREADS) v += parts[4 * p + k];
        v = block_sum(v, sh);
        if (threadIdx.x == 0) {
            const float mean = (float)(v / cnt[k]);
            out[1 + k] = mean;
            loss += mean;     // losses.py:67-70: fp32 sum of the four means, in this order
        }
    }
    if (threadIdx.x == 0) out[0] = loss;
}

__global__ void __launch_bounds__(PIX_THREADS) k_smooth_bwd(const float* __restrict__ map, int M, int H, int W,
                                                            const float* __restrict__ g_loss, float* __restrict__ g_map) {
    const long n = (long)M * H * W, q = (long)blockIdx.x * PIX_THREADS + threadIdx.x;
    if (q >= n) return;
    const int m = (int)(q / ((long)H * W)), r = (int)(q % ((long)H * W)), y = r / W, x = r % W;
    const SmoothMap s{map + (long)m * H * W, H, W};
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    const float g = g_loss[0];
    const float w0 = g / ((float)M * H * (W - 2)), w1 = g / ((float)M * (H - 1) * (W - 1)), w3 = g / ((float)M * (H - 2) * W);
    bool ok;
    float acc = 0.f;
    // dx2(y,x') = p(x'+2) - 2 p(x'+1) + p(x')
    acc += w0 * (sgn(s.dx2(y, x, &ok)) - 2.f * sgn(s.dx2(y, x - 1, &ok)) + sgn(s.dx2(y, x - 2, &ok)));
    // dxdy / dydx (y',x') = p(y'+1,x'+1) - p(y'+1,x') - p(y',x'+1) + p(y',x')
    acc += w1 * (sgn(s.dxdy(y, x, &ok)) - sgn(s.dxdy(y, x - 1, &ok)) - sgn(s.dxdy(y - 1, x, &ok)) + sgn(s.dxdy(y - 1, x - 1, &ok)));
    acc += w1 * (sgn(s.dydx(y, x, &ok)) - sgn(s.dydx(y, x - 1, &ok)) - sgn(s.dydx(y - 1, x, &ok)) + sgn(s.dydx(y - 1, x - 1, &ok)));
    acc += w3 * (sgn(s.dy2(y, x, &ok)) - 2.f * sgn(s.dy2(y - 1, x, &ok)) + sgn(s.dy2(y - 2, x, &ok)));
    g_map[q] = acc;
}

inline int red_blocks(long items) {
    long b = (items + RED_THREADS - 1) / RED_THREADS;
    return (int)(b < 1 ? 1 : (b > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : b));
}
inline bool aligned16(const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" {

size_t g2s_reduce_ws_bytes(void) { return sizeof(double) * 4 * RED_MAX_BLOCKS; }

int g2s_clamped_depth_fwd(const float* depth_raw, int n_groups, int maps_per_group, int H, int W, float min_depth,
                          float max_depth, float border_depth, int clamp_border, void* reduce_ws, float* mean_out,
                          float* depth, void* stream) {
    if (!depth_raw || !reduce_ws || !mean_out || !depth) return G2S_ERR_NULL;
    if (n_groups <= 0 || n_groups > 65535 || maps_per_group <= 0 || H <= 0 || W < 4) return G2S_ERR_SHAPE;
    if ((size_t)n_groups * CD_PARTS * sizeof(double) > g2s_reduce_ws_bytes()) return G2S_ERR_SHAPE;
    const long ge = (long)maps_per_group * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    { Launch l_(K_CLAMPED_DEPTH, st);
      k_cd_partial<0><<<dim3(CD_PARTS, n_groups), RED_THREADS, 0, st>>>(depth_raw, nullptr, nullptr, ge, W, min_depth, max_depth,
                                                                         clamp_border, (double*)reduce_ws); }
    { Launch l_(K_CLAMPED_DEPTH, st);
      k_cd_apply_fwd<<<dim3((unsigned)((ge + RED_THREADS - 1) / RED_THREADS), n_groups), RED_THREADS, 0, st>>>(
          depth_raw, (const double*)reduce_ws, ge, W, min_depth, max_depth, border_depth, clamp_border, mean_out, depth); }
    return launch_status();
}

int g2s_clamped_depth_bwd(const float* depth_raw, const float* mean, const float* grad_depth, int n_groups,
                          int maps_per_group, int H, int W, float min_depth, float max_depth, int clamp_border,
                          void* reduce_ws, float* grad_raw, void* stream) {
    if (!depth_raw || !mean || !grad_depth || !reduce_ws || !grad_raw) return G2S_ERR_NULL;
    if (n_groups <= 0 || n_groups > 65535 || maps_per_group <= 0 || H <= 0 || W < 4) return G2S_ERR_SHAPE;
    if ((size_t)n_groups * CD_PARTS * sizeof(double) > g2s_reduce_ws_bytes()) return G2S_ERR_SHAPE;
    const long ge = (long)maps_per_group * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    { Launch l_(K_CLAMPED_DEPTH, st);
      k_cd_partial<1><<<dim3(CD_PARTS, n_groups), RED_THREADS, 0, st>>>(depth_raw, mean, grad_depth, ge, W, min_depth, max_depth,
                                                                         clamp_border, (double*)reduce_ws); }
    { Launch l_(K_CLAMPED_DEPTH, st);
      k_cd_apply_bwd<<<dim3((unsigned)((ge + RED_THREADS - 1) / RED_THREADS), n_groups), RED_THREADS, 0, st>>>(
          depth_raw, mean, grad_depth, (const double*)reduce_ws, ge, W, min_depth, max_depth, clamp_border, grad_raw); }
    return launch_status();
}

int g2s_shading_fwd(const float* normal, long normal_view_stride, const float* light5, const float* albedo,
                    long albedo_view_stride, int B, int HW, float* diffuse, float* texture, void* stream) {
    if (!normal || !light5 || !albedo || !texture) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || HW <= 0) return G2S_ERR_SHAPE;
    { Launch l_(K_SHADING, (cudaStream_t)stream);
      k_shading_fwd<<<pix_grid(HW, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(normal, normal_view_stride, light5, albedo,
                                                                              albedo_view_stride, HW, diffuse, texture); }
    return launch_status();
}

int g2s_shading_bwd(const float* normal, long normal_view_stride, const float* light5, const float* albedo,
                    long albedo_view_stride, int B, int HW, const float* grad_diffuse, const float* grad_texture,
                    float* grad_normal, long grad_normal_view_stride, float* grad_light5, float* grad_albedo,
                    long grad_albedo_view_stride, void* stream) {
    if (!normal || !light5 || !albedo) return G2S_ERR_NULL;
    if (!grad_diffuse && !grad_texture) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || HW <= 0) return G2S_ERR_SHAPE;
    { Launch l_(K_SHADING, (cudaStream_t)stream);
      k_shading_bwd<<<pix_grid(HW, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
          normal, normal_view_stride, light5, albedo, albedo_view_stride, HW, grad_diffuse, grad_texture, grad_normal,
          grad_normal_view_stride, grad_light5, grad_albedo, grad_albedo_view_stride); }
    return launch_status();
}

int g2s_photometric_fwd(const float* im1, const float* im2, long im2_batch_stride, const float* recon_depth,
                        float depth_thresh, const float* mask_in, const float* conf_sigma, int sigma_channels, int B, int C,
                        int HW, void* reduce_ws, float* out3, void* stream) {
    if (!im1 || !im2 || !reduce_ws || !out3) return G2S_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return G2S_ERR_SHAPE;
    if (conf_sigma && sigma_channels != 1 && sigma_channels != C) return G2S_ERR_SHAPE;
    const PhotoArgs a{im1, im2, recon_depth, mask_in, im2_batch_stride, depth_thresh, B, C, HW, conf_sigma, sigma_channels};
    // conf_sigma (never used by the reference's callers) takes the scalar path
    const bool vec = !conf_sigma && HW % 4 == 0 && im2_batch_stride % 4 == 0 && aligned16(im1) && aligned16(im2) &&
                     aligned16(recon_depth) && aligned16(mask_in);
    const int nb = red_blocks(vec ? (long)B * HW / 4 : (long)B * HW);
    cudaStream_t st = (cudaStream_t)stream;
    { Launch l_(K_PHOTOMETRIC, st);
      if (vec) k_photo_fwd<true><<<nb, RED_THREADS, 0, st>>>(a, (double*)reduce_ws);
      else k_photo_fwd<false><<<nb, RED_THREADS, 0, st>>>(a, (double*)reduce_ws); }
    { Launch l_(K_PHOTOMETRIC, st); k_photo_finish<<<1, RED_THREADS, 0, st>>>((const double*)reduce_ws, nb, C, out3); }
    return launch_status();
}

int g2s_photometric_bwd(const float* im1, const float* im2, long im2_batch_stride, const float* recon_depth,
                        float depth_thresh, const float* mask_in, const float* conf_sigma, int sigma_channels, int B, int C,
                        int HW, const float* sums3, const float* grad_loss, float* grad_im1, float* grad_im2,
                        float* grad_sigma, void* stream) {
    if (!im1 || !im2 || !sums3 || !grad_loss) return G2S_ERR_NULL;
    if (!grad_im1 && !grad_im2 && !grad_sigma) return G2S_ERR_NULL;
    if (grad_sigma && !conf_sigma) return G2S_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return G2S_ERR_SHAPE;
    if (conf_sigma && sigma_channels != 1 && sigma_channels != C) return G2S_ERR_SHAPE;
    if (grad_im2 && im2_batch_stride != (long)C * HW) return G2S_ERR_UNSUPPORTED;   // a broadcast target gets no gradient here
    const PhotoArgs a{im1, im2, recon_depth, mask_in, im2_batch_stride, depth_thresh, B, C, HW, conf_sigma, sigma_channels};
    const bool vec = !conf_sigma && HW % 4 == 0 && im2_batch_stride % 4 == 0 && aligned16(im1) && aligned16(im2) &&
                     aligned16(recon_depth) && aligned16(mask_in) && aligned16(grad_im1) && aligned16(grad_im2);
    const int nb = red_blocks(vec ? (long)B * HW / 4 : (long)B * HW);
    cudaStream_t st = (cudaStream_t)stream;
    { Launch l_(K_PHOTOMETRIC, st);
      if (vec) k_photo_bwd<true><<<nb, RED_THREADS, 0, st>>>(a, sums3, grad_loss, grad_im1, grad_im2, grad_sigma);
      else k_photo_bwd<false><<<nb, RED_THREADS, 0, st>>>(a, sums3, grad_loss, grad_im1, grad_im2, grad_sigma); }
    return launch_status();
}

int g2s_smooth_fwd(const float* map, int M, int H, int W, void* reduce_ws, float* out5, void* stream) {
    if (!map || !reduce_ws || !out5) return G2S_ERR_NULL;
    if (M <= 0 || H < 3 || W < 3) return G2S_ERR_SHAPE;
    const int nb = red_blocks((long)M * H * W);
    cudaStream_t st = (cudaStream_t)stream;
    { Launch l_(K_SMOOTH, st); k_smooth_fwd<<<nb, RED_THREADS, 0, st>>>(map, M, H, W, (double*)reduce_ws); }
    { Launch l_(K_SMOOTH, st); k_smooth_finish<<<1, RED_THREADS, 0, st>>>((const double*)reduce_ws, nb, M, H, W, out5); }
    return launch_status();
}

int g2s_smooth_bwd(const float* map, int M, int H, int W, const float* grad_loss, float* grad_map, void* stream) {
    if (!map || !grad_loss || !grad_map) return G2S_ERR_NULL;
    if (M <= 0 || H < 3 || W < 3) return G2S_ERR_SHAPE;
    const long n = (long)M * H * W;
    { Launch l_(K_SMOOTH, (cudaStream_t)stream);
      k_smooth_bwd<<<(unsigned)((n + PIX_THREADS - 1) / PIX_THREADS), PIX_THREADS, 0, (cudaStream_t)stream>>>(map, M, H, W,
                                                                                                             grad_loss, grad_map); }
    return launch_status();
}

}  // extern "C"
