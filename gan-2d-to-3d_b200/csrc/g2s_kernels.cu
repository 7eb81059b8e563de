// g2s_kernels.cu -- sm_100a kernels + C ABI (include/g2s_b200.h) of the depth-map renderer path.
//
// Reference being replaced: GAN2Shape/renderer/renderer.py:61-139, 252-277, GAN2Shape/renderer/utils.py,
// GAN2Shape/model.py:146-151, 257-270, 347-360 and the external neural_renderer rasteriser
// (renderer.py:47-54, 120).  See DESIGN.md for the kernel list, data layout and rooflines.
//
// Nothing here is a dense contraction: no tensor cores.  The kernels are HBM/L2- and fp32-ALU-bound;
// the design rules are coalesced 8/16-byte accesses, shared-memory staging of the projected vertex
// tiles, 64-bit RED.MIN for the z-buffer and warp-level reductions for the per-view gradients.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/g2s_b200.h"
#include "g2s_bigface.cuh"

using namespace g2s;

namespace {

constexpr int PIX_THREADS = 256;

Cam make_cam(const g2s_camera* c) {
    Cam k;
    for (int i = 0; i < 9; i++) {
        k.K[i] = c->K[i];
        k.Kg[i] = c->K_grid[i];
        k.invK[i] = c->inv_K[i];
    }
    k.rcd = c->rot_center_depth;
    k.os = (float)c->image_size;
    k.half_os = (float)(c->image_size / 2.0);
    k.near = c->near_z;
    k.far = c->far_z;
    k.clamp_lo = c->clamp_lo;
    k.clamp_hi = c->clamp_hi;
    k.S = c->image_size;
    return k;
}


// ---- instrumentation: launch counter + optional CUDA-event timing of every kernel (bench.py / tests) ----
enum KernelId { K_ZINIT, K_SPLAT, K_SPLAT_BIG, K_RESOLVE, K_RESOLVE_FUSED, K_GRID_FWD, K_GRID_BWD, K_NORMAL_FWD, K_NORMAL_BWD,
                K_SAMPLE_FWD, K_SAMPLE_BWD, K_CLAMP_GRAD, K_RASTER_BWD, K_BWD_PIXEL, K_BWD_TEX, K_GRID3D, K_RESOLVE_RGB,
                K_PROJECT, K_VERTEX_BWD, K_VIEW, K_LIGHT, K_CLAMPED_DEPTH, K_SHADING, K_PHOTOMETRIC, K_SMOOTH, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"k_zbuf_init", "k_splat_tile", "k_splat_big", "k_resolve", "k_resolve_fused", "k_warp_grid_fwd",
                                           "k_warp_grid_bwd", "k_normal_fwd", "k_normal_bwd", "k_sample_fwd",
                                           "k_sample_bwd", "k_clamp_grad", "k_raster_bwd_px", "k_render_bwd_pixel",
                                           "k_render_bwd_tex", "k_grid3d", "k_resolve_rgb", "k_project_verts",
                                           "k_vertex_bwd", "k_view_fwd/bwd", "k_light_fwd/bwd", "k_clamped_depth",
                                           "k_shading_fwd/bwd", "k_photometric", "k_smooth"};
std::atomic<long> g_launches{0};
struct ProfRec { int id; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;

struct Launch {
    int id; cudaStream_t st; bool rec; cudaEvent_t a, b;
    Launch(int id_, cudaStream_t st_) : id(id_), st(st_), rec(false) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (g_prof_on) {
            std::lock_guard<std::mutex> lk(g_prof_mu);
            auto get = [&]() { cudaEvent_t e; if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); }
                               else cudaEventCreate(&e); return e; };
            a = get(); b = get(); rec = true;
            cudaEventRecord(a, st);
        }
    }
    ~Launch() {
        if (rec) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_prof_recs.push_back({id, a, b});
        }
    }
};

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? G2S_OK : G2S_ERR_LAUNCH; }

constexpr int MAX_LANES = 4;
constexpr int LOSS_STAGE_BLOCKS = 512;   // first-stage blocks of the fused loss's fixed-order sum

// Sum the per-thread grad_R (9) | grad_t (3) contributions over the block and atomically add the totals to the view's
// grad_R / grad_t.  All threads must call.
template <int THREADS>
__device__ __forceinline__ void block_accumulate_Rt(const float (&acc)[12], float* grad_R, float* grad_t) {
    __shared__ float red[(THREADS / 32) * 12];
    const int lin = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = lin & 31, warp = lin >> 5;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = k < 12 ? acc[k] : 0.f;
    const float s = warp_reduce_halving<4>(v, lane);
    const int idx = halving_index<4>(lane);
    if ((lane & 1) == 0 && idx < 12) red[warp * 12 + idx] = s;
    __syncthreads();
    if (lin < 12) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) tot += red[w * 12 + lin];
        if (tot != 0.f) atomicAdd(lin < 9 ? &grad_R[lin] : &grad_t[lin - 9], tot);
    }
}

// ------------------------------------------------------------------------------------------------
// z-buffer
__global__ void k_zbuf_init(unsigned long long* zb, long n, unsigned long long key) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) zb[i] = key;
}

// image of view b = b / views_per_image as a multiply-high by ceil(2^32 / vpi) (host: vpi_magic(); 0 = divide)
__device__ __forceinline__ int view_image(int b, int vpi, unsigned magic) {
    return magic ? (int)__umulhi((unsigned)b, magic) : (vpi == 1 ? b : b / vpi);
}

// host: the multiply-high constant of view_image() for a call of n_views views (0 = the kernel divides)
inline unsigned vpi_magic(int vpi, long n_views) {
    if (vpi < 2) return 0u;
    const unsigned long long m = ((1ull << 32) + (unsigned long long)vpi - 1) / (unsigned long long)vpi;
    const unsigned long long e = m * (unsigned long long)vpi - (1ull << 32);      // < vpi
    // umulhi(n, m) == n / vpi for every n with n * e < 2^32
    return (unsigned long long)n_views * (e + 1) < (1ull << 32) && m < (1ull << 32) ? (unsigned)m : 0u;
}

// Forward rasterisation of the grid mesh into the packed-key z-buffer, stage 1 (g2s_tile.cuh): one CTA per 16 x 16 block
// of quads of one view, grid = (views, tile columns, tile rows).  4 CTAs per SM (64 registers), 37 KB of static shared memory.
#ifndef G2S_TILE_MINBLOCKS
#define G2S_TILE_MINBLOCKS 4
#endif
template <bool FROM_VERTS, bool POW2>
__global__ void __launch_bounds__(SPLAT_THREADS, G2S_TILE_MINBLOCKS)
k_splat_tile(const Cam cam, const float* __restrict__ depth, long dstride, int vpi, const float* __restrict__ R,
             const float* __restrict__ t, const float* __restrict__ verts3d, unsigned long long* __restrict__ zbuf,
             const WorkList wl, unsigned magic, int view0, float4* __restrict__ proj_out) {
    __shared__ TileSmem2 sm;
    // grid = (views, tile columns, tile rows): no per-thread integer division to find the tile or the image
    const int bl = blockIdx.x, b = view0 + bl, S = cam.S, is = 2 * S;
    const int tile_y = blockIdx.z, tile_x = blockIdx.y;
    splat_tile_body<FROM_VERTS, POW2>(sm, cam, FROM_VERTS ? nullptr : depth + (long)view_image(b, vpi, magic) * dstride,
                                FROM_VERTS ? verts3d + (long)b * S * S * 3 : nullptr, FROM_VERTS ? nullptr : R + (long)b * 9,
                                FROM_VERTS ? nullptr : t + (long)b * 3, zbuf + (long)bl * is * is, wl, bl, tile_y * TILE_H,
                                tile_x * TILE, proj_out ? proj_out + (long)b * S * S : nullptr);
}

// stage 2 (g2s_bigface.cuh): persistent CTAs pull the faces stage 1 deferred (long walls, degenerate quads) from the work list
template <bool FROM_VERTS>
__global__ void __launch_bounds__(BIG_THREADS, G2S_BIG_CTAS)
k_splat_big(const Cam cam, const float* __restrict__ depth, long dstride, int vpi, const float* __restrict__ R,
            const float* __restrict__ t, const float* __restrict__ verts3d, unsigned long long* __restrict__ zbuf,
            const WorkList wl, int view0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem& sm = *reinterpret_cast<BigSmem*>(smem_raw);
    splat_big_body<FROM_VERTS>(sm, cam, depth, dstride, vpi, R, t, verts3d, zbuf, wl, view0);
}

// ------------------------------------------------------------------------------------------------
// shading + bilinear sampling (model.py:355-360, 270)

__device__ __forceinline__ bool bilinear_setup(float gx, float gy, int W, int H, int align, int& x0, int& y0,
                                               float& tx, float& ty) {
    const float ix = grid_unnormalize(gx, W, align), iy = grid_unnormalize(gy, H, align);
    if (!(ix > -1.0f && ix < (float)W && iy > -1.0f && iy < (float)H)) return false;
    const float fx = floorf(ix), fy = floorf(iy);
    x0 = (int)fx; y0 = (int)fy;
    tx = ix - fx; ty = iy - fy;
    return true;
}

struct FusedArgs {
    const float* R;
    const float* t;
    const float* light;   // [n_views,5]
    const float* normal;  // packed texel map [n_images,S,S,8]: n0 n1 n2 a0 a1 a2 - -
    const float* albedo;  // [n_images,3,S,S]
    float* recon_im;      // [n_views,3,S,S]
    int vpi;
    int align;
    int view0;            // first view of this launch (chunked launches)
    const float* mask_in; // optional [n_images,S,S] canonical mask (NULL = all ones)
    float* mask_out;      // optional [n_views,S,S]: grid_sample(mask, grid, mode='nearest') (renderer.py:263)
    // fused masked photometric loss (model.py:265-274 + losses.py:39-51), all optional
    const float* target;  // [n_views,3,S,S] the images the renders are compared with
    const float* vmask;   // [n_views,S,S] per-view masks or NULL
    float thresh;         // validity: recon_depth < thresh
    double* loss_parts;   // forward: [n_views * CTAs per view][2] partial (numerator, mask count)
    const float* sums3;   // backward: {loss, numerator, denominator} of the forward
    const float* gloss;   // backward: d(total) / d(loss), a device scalar
    unsigned vpi_magic;   // ceil(2^32 / vpi) when view / vpi == umulhi(view, magic) for every view of the call, else 0
};


// pixel-kernel blocks (threads = output pixels): the z-buffer resolve streams rows (64 x 4); the two backward pixel
// kernels gather vertices / texels around a 2-D patch and run faster on square blocks (measured, profiles/r01_notes.md)
#ifndef G2S_PBX
#define G2S_PBX 64
#define G2S_PBY 4
#endif
#ifndef G2S_BPX
#define G2S_BPX 16
#define G2S_BPY 8
#endif
constexpr int PBX = G2S_PBX, PBY = G2S_PBY;   // k_resolve
constexpr int BPX = G2S_BPX, BPY = G2S_BPY;   // k_render_bwd_pixel

// The 4 bilinear taps of one sample, clamped so that every load is unconditional (one round trip);
// out-of-range taps get weight 0 (zeros padding).
struct Taps {
    int p[4];
    float w[4];
    float wx[4], wy[4];   // per-tap x / y weights (for the grid gradient)
    bool any;
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int W, int H, int align) {
    Taps tp;
    int x0 = 0, y0 = 0;
    float tx = 0.f, ty = 0.f;
    tp.any = bilinear_setup(gx, gy, W, H, align, x0, y0, tx, ty);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        const bool ok = tp.any && xx >= 0 && xx < W && yy >= 0 && yy < H;
        tp.p[k] = min(max(yy, 0), H - 1) * W + min(max(xx, 0), W - 1);
        tp.wx[k] = ok ? ((k & 1) ? tx : 1.f - tx) : 0.f;
        tp.wy[k] = ok ? ((k >> 1) ? ty : 1.f - ty) : 0.f;
        tp.w[k] = tp.wx[k] * tp.wy[k];
    }
    return tp;
}

// shaded texture of the 4 taps: tex[k][c] = (albedo/2+.5) * (a + b*max(0, n.l)) * 2 - 1.  The per-image normal map
// and albedo are packed as 8 floats per texel (n0 n1 n2 a0 a1 a2 - -) by k_normal_fwd, so a tap is two
// 16-byte loads instead of six scalar ones.
constexpr int TEXEL = 8;
__device__ __forceinline__ void shade_taps(const float* __restrict__ pack, const Taps& tp, const float* L,
                                           float tex[4][3]) {
    float4 lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4* q = reinterpret_cast<const float4*>(pack + (long)tp.p[k] * TEXEL);
        lo[k] = __ldg(q);
        hi[k] = __ldg(q + 1);
    }
    // All eight loads are in flight before anything waits on one of them: ONE round trip.  The empty asm "uses" every loaded
    // value, so the scheduler cannot sink a tap's loads below the shading of the previous tap (it did, after an unrelated
    // change to the camera struct: long-scoreboard stalls 34 % -> 67 %, k_resolve 1.87 -> 2.51 ms; profiles/r02_notes.md).
    asm volatile("" : "+f"(lo[0].x), "+f"(lo[0].y), "+f"(lo[0].z), "+f"(lo[0].w), "+f"(hi[0].x), "+f"(hi[0].y),
                      "+f"(lo[1].x), "+f"(lo[1].y), "+f"(lo[1].z), "+f"(lo[1].w), "+f"(hi[1].x), "+f"(hi[1].y),
                      "+f"(lo[2].x), "+f"(lo[2].y), "+f"(lo[2].z), "+f"(lo[2].w), "+f"(hi[2].x), "+f"(hi[2].y),
                      "+f"(lo[3].x), "+f"(lo[3].y), "+f"(lo[3].z), "+f"(lo[3].w), "+f"(hi[3].x), "+f"(hi[3].y));
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float ndl = lo[k].x * L[2] + lo[k].y * L[3] + lo[k].z * L[4];
        const float sh = L[0] + L[1] * fmaxf(ndl, 0.f);
        tex[k][0] = (lo[k].w * 0.5f + 0.5f) * sh * 2.0f - 1.0f;
        tex[k][1] = (hi[k].x * 0.5f + 0.5f) * sh * 2.0f - 1.0f;
        tex[k][2] = (hi[k].y * 0.5f + 0.5f) * sh * 2.0f - 1.0f;
    }
}

// z-buffer resolve: face-index map, flip + 2x2 mean + clamp -> recon_depth, z-buffer reset; when FUSED
// also inverse warp grid + shaded bilinear sampling -> recon_im.  One thread per output pixel,
// block = 64 columns x 4 rows, grid = (cols, rows, views).
// (LOSS: 5 CTAs/SM = 48 registers like the plain kernel; left alone the compiler takes 61-72 and the kernel 1.98 ms instead of 1.62)
template <bool FUSED, bool LOSS = false>
__global__ void __launch_bounds__(PBX * PBY, LOSS ? 5 : 0)
k_resolve(const Cam cam, unsigned long long* __restrict__ zbuf, float* __restrict__ recon_depth,
          int* __restrict__ face_idx, const FusedArgs fa) {
    const int S = cam.S, is = 2 * S, bl = blockIdx.z, b = fa.view0 + bl;
    const int j = blockIdx.x * PBX + threadIdx.x, i = blockIdx.y * PBY + threadIdx.y;
    const bool inside = j < S && i < S;
    unsigned long long* zb = zbuf + (long)bl * is * is;
    ulonglong2* r0 = reinterpret_cast<ulonglong2*>(zb + (long)(2 * i) * is + 2 * j);
    ulonglong2* r1 = reinterpret_cast<ulonglong2*>(zb + (long)(2 * i + 1) * is + 2 * j);
    // z-buffer traffic stays in L2 (.cg); the outputs are written once and not re-read by this pass: streaming
    // stores (.cs, evict-first) keep them from pushing the z-buffer chunk out of L2.  The key loads are issued BEFORE
    // anything waits on them, together with the view's R, t, light: one round trip where there were three (profiles/r01_notes.md).
    ulonglong2 k0 = make_ulonglong2(0ull, 0ull), k1 = k0;
    if (inside) { k0 = __ldcg(r0); k1 = __ldcg(r1); }
    // LOSS: the target pixel and its mask are only PREFETCHED here (to L2) and loaded where they are used -- four more live
    // registers across the whole body spill at the 48 this variant is held to; nobody returns before the block sum at the end
    float l_num = 0.f, l_den = 0.f;
    if (LOSS && inside) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&fa.target[((long)b * 3 + c) * S * S + i * S + j]));
        if (fa.vmask) asm volatile("prefetch.global.L2 [%0];" ::"l"(&fa.vmask[(long)b * S * S + i * S + j]));
    }
    // R, t, light of the view: warp-uniform loads (one L1 transaction per warp, broadcast), no shared memory and no barrier
    float sview[17];
    if (FUSED) {
#pragma unroll
        for (int k = 0; k < 9; k++) sview[k] = __ldg(&fa.R[b * 9 + k]);
#pragma unroll
        for (int k = 0; k < 3; k++) sview[9 + k] = __ldg(&fa.t[b * 3 + k]);
#pragma unroll
        for (int k = 0; k < 5; k++) sview[12 + k] = __ldg(&fa.light[b * 5 + k]);
    }
    if (!LOSS && !inside) return;
    if (inside) {
    const int pix = i * S + j;
    if (face_idx) {
        int* fo = face_idx + (long)b * is * is;
        __stcs(reinterpret_cast<int2*>(fo + (long)(2 * i) * is + 2 * j), make_int2(zkey_face(k0.x), zkey_face(k0.y)));
        __stcs(reinterpret_cast<int2*>(fo + (long)(2 * i + 1) * is + 2 * j), make_int2(zkey_face(k1.x), zkey_face(k1.y)));
    }
    const float sum = add(add(add(zkey_depth(k0.x), zkey_depth(k0.y)), zkey_depth(k1.x)), zkey_depth(k1.y));
    const float rd = fminf(fmaxf(mul(sum, 0.25f), cam.clamp_lo), cam.clamp_hi);
    __stcs(&recon_depth[(long)b * S * S + pix], rd);
    // The z-buffer reset.  Its data is made to depend on `rd` (through a mask that is zero for every valid launch but that the
    // compiler cannot see through), so that it cannot be scheduled before the key loads have RETURNED: when ptxas placed the
    // first of the two stores in front of the 17 view loads, that store -- to the very line a key load was still waiting
    // on -- held the load / store unit until the load came back, the view loads went out a round trip late, long-scoreboard
    // stalls rose from 34 % to 67 % of the samples and the kernel from 1.87 to 2.51 ms (profiles/r02_notes.md).
    {
        const unsigned zmask = (unsigned)(fa.view0 >> 31);
        const unsigned long long empty = zkey_empty(cam.far) + (unsigned long long)(__float_as_uint(rd) & zmask);
        __stcg(r0, make_ulonglong2(empty, empty));
        __stcg(r1, make_ulonglong2(empty, empty));
    }
    if (FUSED) {
        // (the LOSS variant sits at its register cap: the multiply-high form spills there, the division does not)
        const int img = LOSS ? b / fa.vpi : view_image(b, fa.vpi, fa.vpi_magic);
        float ray[3], q[3], v[3], g[2];
        pixel_ray(cam, j, i, ray);
        inv_warp_point(cam, sview, sview + 9, ray, rd, q, v);
        point_to_grid(cam, q, S, S, g);
        const Taps tp = make_taps(g[0], g[1], S, S, fa.align);
        float tex[4][3];
        shade_taps(fa.normal + (long)img * S * S * TEXEL, tp, sview + 12, tex);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float o = tp.w[0] * tex[0][c] + tp.w[1] * tex[1][c] + tp.w[2] * tex[2][c] + tp.w[3] * tex[3][c];
            const float oc = fminf(fmaxf(o, -1.f), 1.f);
            __stcs(&fa.recon_im[((long)b * 3 + c) * S * S + pix], oc);
            if (LOSS) l_num += fabsf(oc - __ldcs(&fa.target[((long)b * 3 + c) * S * S + pix]));
        }
        if (LOSS) {      // model.py:268-269: (recon_depth < max_depth + margin) * masks
            const float vm = fa.vmask ? __ldcs(&fa.vmask[(long)b * S * S + pix]) : 1.f;
            l_den = (rd < fa.thresh ? 1.f : 0.f) * vm;
            l_num *= l_den;
        }
        if (fa.mask_out) {   // nearest-neighbour warp of the canonical mask with the same grid (sample_pseudo_imgs)
            const float ix = nearbyintf(grid_unnormalize(g[0], S, fa.align)), iy = nearbyintf(grid_unnormalize(g[1], S, fa.align));
            const bool ok = ix >= 0.f && ix < (float)S && iy >= 0.f && iy < (float)S;
            float m = 0.f;
            if (ok) m = fa.mask_in ? __ldg(&fa.mask_in[(long)img * S * S + (int)iy * S + (int)ix]) : 1.0f;
            __stcs(&fa.mask_out[(long)b * S * S + pix], m);
        }
    }
    }
    if (LOSS) {      // one (numerator, mask count) pair per CTA, summed in a fixed order by k_photo_finish: deterministic
        __shared__ double sh[2][PBX * PBY / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l_num += __shfl_xor_sync(0xffffffffu, l_num, o);
            l_den += __shfl_xor_sync(0xffffffffu, l_den, o);
        }
        const int lin = threadIdx.y * PBX + threadIdx.x;
        if ((lin & 31) == 0) { sh[0][lin >> 5] = (double)l_num; sh[1][lin >> 5] = (double)l_den; }
        __syncthreads();
        if (lin == 0) {
            double n = 0.0, d = 0.0;
#pragma unroll
            for (int w = 0; w < PBX * PBY / 32; w++) { n += sh[0][w]; d += sh[1][w]; }
            const long cta = ((long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            fa.loss_parts[2 * cta] = n;
            fa.loss_parts[2 * cta + 1] = d;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// standalone per-pixel operators

__global__ void __launch_bounds__(PIX_THREADS)
k_warp_grid_fwd(const Cam cam, const float* __restrict__ depth, long dstride, const float* __restrict__ R,
                const float* __restrict__ t, int H, int W, int inverse, float* __restrict__ grid) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    float Rm[9], tv[3], ray[3], q[3], v[3], g[2];
#pragma unroll
    for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R[b * 9 + k]);
#pragma unroll
    for (int k = 0; k < 3; k++) tv[k] = __ldg(&t[b * 3 + k]);
    pixel_ray(cam, x, y, ray);
    const float d = depth[(long)b * dstride + pix];
    if (inverse) inv_warp_point(cam, Rm, tv, ray, d, q, v);
    else warp_point(cam, Rm, tv, ray, d, q);
    point_to_grid(cam, q, W, H, g);
    reinterpret_cast<float2*>(grid)[(long)b * H * W + pix] = make_float2(g[0], g[1]);
}

// Backward of (depth -> sampling grid) for one pixel.  Returns d(loss)/d(depth) and accumulates the
// pixel's contribution to grad_R (acc[0..9)) and grad_t (acc[9..12)).
__device__ __forceinline__ float warp_grid_bwd_pixel(const Cam& cam, const float* Rm, const float* tv, int x, int y,
                                                     int W, int H, float d, int inverse, float Gx, float Gy,
                                                     float acc[12]) {
    float ray[3], q[3], v[3];
    pixel_ray(cam, x, y, ray);
    if (inverse) {
        inv_warp_point(cam, Rm, tv, ray, d, q, v);
    } else {
        warp_point(cam, Rm, tv, ray, d, q);
        v[0] = ray[0] * d; v[1] = ray[1] * d; v[2] = ray[2] * d - cam.rcd;
    }
    const float iz = 1.0f / q[2];
    const float nx = q[0] * iz, ny = q[1] * iz;
    const float dpx = Gx * 2.0f / (float)(W - 1), dpy = Gy * 2.0f / (float)(H - 1);
    const float dnx = dpx * cam.Kg[0] + dpy * cam.Kg[3], dny = dpx * cam.Kg[1] + dpy * cam.Kg[4];
    float dq[3] = {dnx * iz, dny * iz, -(dnx * nx + dny * ny) * iz};
    float dv[3];
    if (inverse) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dv[k] = dq[0] * Rm[3 * k] + dq[1] * Rm[3 * k + 1] + dq[2] * Rm[3 * k + 2];
#pragma unroll
            for (int j = 0; j < 3; j++) acc[3 * k + j] += v[k] * dq[j];
            acc[9 + k] -= dv[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dv[k] = dq[0] * Rm[k] + dq[1] * Rm[3 + k] + dq[2] * Rm[6 + k];
#pragma unroll
            for (int j = 0; j < 3; j++) acc[3 * j + k] += dq[j] * v[k];
            acc[9 + k] += dq[k];
        }
    }
    return dv[0] * ray[0] + dv[1] * ray[1] + dv[2] * ray[2];
}

__global__ void __launch_bounds__(PIX_THREADS)
k_warp_grid_bwd(const Cam cam, const float* __restrict__ depth, long dstride, const float* __restrict__ R,
                const float* __restrict__ t, int H, int W, int inverse, const float* __restrict__ grad_grid,
                float* __restrict__ grad_depth, float* __restrict__ grad_R, float* __restrict__ grad_t) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    if (pix < H * W) {
        const int y = pix / W, x = pix - y * W;
        float Rm[9], tv[3];
#pragma unroll
        for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R[b * 9 + k]);
#pragma unroll
        for (int k = 0; k < 3; k++) tv[k] = __ldg(&t[b * 3 + k]);
        const float2 G = reinterpret_cast<const float2*>(grad_grid)[(long)b * H * W + pix];
        grad_depth[(long)b * H * W + pix] =
            warp_grid_bwd_pixel(cam, Rm, tv, x, y, W, H, depth[(long)b * dstride + pix], inverse, G.x, G.y, acc);
    }
    if (grad_R) {
        block_accumulate_Rt<PIX_THREADS>(acc, grad_R + b * 9, grad_t + b * 3);
    }
}

// 3-D point of pixel (x,y): ray * depth
__device__ __forceinline__ void depth_point(const Cam& cam, const float* __restrict__ dimg, int W, int x, int y,
                                            float p[3]) {
    float ray[3];
    pixel_ray(cam, x, y, ray);
    const float d = dimg[y * W + x];
    p[0] = mul(ray[0], d); p[1] = mul(ray[1], d); p[2] = mul(ray[2], d);
}

__global__ void __launch_bounds__(PIX_THREADS)
k_normal_fwd(const Cam cam, const float* __restrict__ depth, int H, int W, float* __restrict__ normal, int nstride,
             const float* __restrict__ albedo = nullptr) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const float* dimg = depth + (long)b * H * W;
    float n[3] = {0.f, 0.f, dvd(1.0f, add(1.0f, 1e-7f))};
    if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
        float pl[3], pr[3], pu[3], pd[3], len;
        depth_point(cam, dimg, W, x - 1, y, pl);
        depth_point(cam, dimg, W, x + 1, y, pr);
        depth_point(cam, dimg, W, x, y - 1, pu);
        depth_point(cam, dimg, W, x, y + 1, pd);
        normal_from_points(pl, pr, pu, pd, n, &len);
    }
    float* o = normal + ((long)b * H * W + pix) * nstride;
    if (albedo) {      // the fused chain's packed texel (nstride == TEXEL): normal xyz, albedo rgb, 2 pad -- two 16-byte stores
        const float* a = albedo + (long)b * 3 * H * W;
        reinterpret_cast<float4*>(o)[0] = make_float4(n[0], n[1], n[2], a[pix]);
        reinterpret_cast<float4*>(o)[1] = make_float4(a[H * W + pix], a[2 * H * W + pix], 0.f, 0.f);
        return;
    }
    o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
}

// d(loss)/d(tu), d(loss)/d(tv) of the normal at interior pixel (x,y) given d(loss)/d(normal)
__device__ __forceinline__ void normal_tangent_grads(const Cam& cam, const float* __restrict__ dimg,
                                                     const float* __restrict__ gnimg, int H, int W, int x, int y,
                                                     float gtu[3], float gtv[3]) {
    gtu[0] = gtu[1] = gtu[2] = 0.f;
    gtv[0] = gtv[1] = gtv[2] = 0.f;
    if (x < 1 || x > W - 2 || y < 1 || y > H - 2) return;
    float pl[3], pr[3], pu[3], pd[3];
    depth_point(cam, dimg, W, x - 1, y, pl);
    depth_point(cam, dimg, W, x + 1, y, pr);
    depth_point(cam, dimg, W, x, y - 1, pu);
    depth_point(cam, dimg, W, x, y + 1, pd);
    const float tu[3] = {pr[0] - pl[0], pr[1] - pl[1], pr[2] - pl[2]};
    const float tv[3] = {pd[0] - pu[0], pd[1] - pu[1], pd[2] - pu[2]};
    const float c[3] = {tu[1] * tv[2] - tu[2] * tv[1], tu[2] * tv[0] - tu[0] * tv[2], tu[0] * tv[1] - tu[1] * tv[0]};
    const float len = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const float den = len + 1e-7f;
    const float* gn = gnimg + (y * W + x) * 3;
    const float g0 = gn[0], g1 = gn[1], g2 = gn[2];
    // n = c / den, den = |c| + eps
    const float gdotc = g0 * c[0] + g1 * c[1] + g2 * c[2];
    const float s = len > 0.f ? gdotc / (den * den * len) : 0.f;
    const float gc[3] = {g0 / den - s * c[0], g1 / den - s * c[1], g2 / den - s * c[2]};
    gtu[0] = tv[1] * gc[2] - tv[2] * gc[1];
    gtu[1] = tv[2] * gc[0] - tv[0] * gc[2];
    gtu[2] = tv[0] * gc[1] - tv[1] * gc[0];
    gtv[0] = gc[1] * tu[2] - gc[2] * tu[1];
    gtv[1] = gc[2] * tu[0] - gc[0] * tu[2];
    gtv[2] = gc[0] * tu[1] - gc[1] * tu[0];
}

// gather form of the normal backward: pixel (x,y) is the +tu endpoint of (x-1,y), the -tu endpoint of
// (x+1,y), the +tv endpoint of (x,y-1) and the -tv endpoint of (x,y+1)
__global__ void __launch_bounds__(PIX_THREADS)
k_normal_bwd(const Cam cam, const float* __restrict__ depth, int H, int W, const float* __restrict__ grad_normal,
             float* __restrict__ grad_depth, int accumulate) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const float* dimg = depth + (long)b * H * W;
    const float* gn = grad_normal + (long)b * H * W * 3;
    float gp[3] = {0.f, 0.f, 0.f}, a[3], c[3];
    if (x >= 1) {
        normal_tangent_grads(cam, dimg, gn, H, W, x - 1, y, a, c);
        gp[0] += a[0]; gp[1] += a[1]; gp[2] += a[2];
    }
    if (x <= W - 2) {
        normal_tangent_grads(cam, dimg, gn, H, W, x + 1, y, a, c);
        gp[0] -= a[0]; gp[1] -= a[1]; gp[2] -= a[2];
    }
    if (y >= 1) {
        normal_tangent_grads(cam, dimg, gn, H, W, x, y - 1, a, c);
        gp[0] += c[0]; gp[1] += c[1]; gp[2] += c[2];
    }
    if (y <= H - 2) {
        normal_tangent_grads(cam, dimg, gn, H, W, x, y + 1, a, c);
        gp[0] -= c[0]; gp[1] -= c[1]; gp[2] -= c[2];
    }
    float ray[3];
    pixel_ray(cam, x, y, ray);
    const float gd = gp[0] * ray[0] + gp[1] * ray[1] + gp[2] * ray[2];
    float* o = grad_depth + (long)b * H * W + pix;
    *o = accumulate ? *o + gd : gd;
}

__global__ void __launch_bounds__(PIX_THREADS)
k_sample_fwd(const float* __restrict__ input, long istride, const float* __restrict__ grid, int C, int H, int W,
             int Ho, int Wo, int mode, int align, float* __restrict__ out) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= Ho * Wo) return;
    const float2 g = reinterpret_cast<const float2*>(grid)[(long)b * Ho * Wo + pix];
    const float* in = input + (long)b * istride;
    float* o = out + (long)b * C * Ho * Wo + pix;
    if (mode == 1) {
        const float ix = nearbyintf(grid_unnormalize(g.x, W, align)), iy = nearbyintf(grid_unnormalize(g.y, H, align));
        const bool ok = ix >= 0.f && ix < (float)W && iy >= 0.f && iy < (float)H;
        const int p = ok ? (int)iy * W + (int)ix : 0;
        for (int c = 0; c < C; c++) o[(long)c * Ho * Wo] = ok ? in[(long)c * H * W + p] : 0.f;
        return;
    }
    int x0, y0;
    float tx, ty;
    const bool any = bilinear_setup(g.x, g.y, W, H, align, x0, y0, tx, ty);
    for (int c = 0; c < C; c++) {
        float acc = 0.f;
        if (any) {
#pragma unroll
            for (int tap = 0; tap < 4; tap++) {
                const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                acc += ((tap & 1) ? tx : 1.f - tx) * ((tap >> 1) ? ty : 1.f - ty) * in[(long)c * H * W + yy * W + xx];
            }
        }
        o[(long)c * Ho * Wo] = acc;
    }
}

__global__ void __launch_bounds__(PIX_THREADS)
k_sample_bwd(const float* __restrict__ input, long istride, const float* __restrict__ grid,
             const float* __restrict__ grad_out, int C, int H, int W, int Ho, int Wo, int mode, int align,
             float* __restrict__ grad_input, long gistride, float* __restrict__ grad_grid) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= Ho * Wo) return;
    const float2 g = reinterpret_cast<const float2*>(grid)[(long)b * Ho * Wo + pix];
    const float* in = input + (long)b * istride;
    const float* go = grad_out + (long)b * C * Ho * Wo + pix;
    float* gi = grad_input ? grad_input + (long)b * gistride : nullptr;
    float gix = 0.f, giy = 0.f;
    if (mode == 1) {
        const float ix = nearbyintf(grid_unnormalize(g.x, W, align)), iy = nearbyintf(grid_unnormalize(g.y, H, align));
        if (gi && ix >= 0.f && ix < (float)W && iy >= 0.f && iy < (float)H) {
            const int p = (int)iy * W + (int)ix;
            for (int c = 0; c < C; c++) atomicAdd(&gi[(long)c * H * W + p], go[(long)c * Ho * Wo]);
        }
    } else {
        int x0, y0;
        float tx, ty;
        if (bilinear_setup(g.x, g.y, W, H, align, x0, y0, tx, ty)) {
            for (int c = 0; c < C; c++) {
                const float G = go[(long)c * Ho * Wo];
#pragma unroll
                for (int tap = 0; tap < 4; tap++) {
                    const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
                    if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                    const float wx = (tap & 1) ? tx : 1.f - tx, wy = (tap >> 1) ? ty : 1.f - ty;
                    const long p = (long)c * H * W + yy * W + xx;
                    if (gi) atomicAdd(&gi[p], wx * wy * G);
                    const float val = in[p];
                    gix += ((tap & 1) ? 1.f : -1.f) * wy * val * G;
                    giy += ((tap >> 1) ? 1.f : -1.f) * wx * val * G;
                }
            }
        }
    }
    if (grad_grid) {
        const float mx = align ? (float)(W - 1) * 0.5f : (float)W * 0.5f;
        const float my = align ? (float)(H - 1) * 0.5f : (float)H * 0.5f;
        reinterpret_cast<float2*>(grad_grid)[(long)b * Ho * Wo + pix] = make_float2(gix * mx, giy * my);
    }
}

// ------------------------------------------------------------------------------------------------
// backward of warp_canon_depth

// g_sub = d(loss)/d(each of the 4 sub-pixel depths) of an output pixel: clamp mask * g / 4
__global__ void k_clamp_grad(const float* __restrict__ recon_depth, const float* __restrict__ grad, float lo, float hi,
                             long n, float* __restrict__ g_sub) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rd = recon_depth[i];
    g_sub[i] = (rd > lo && rd < hi) ? 0.25f * grad[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// view [B, 3|5|6] -> R = Rz Ry Rx [B,3,3], t [B,3]  (utils.py:33-73) and raw light [B,4] -> (a, b, dx, dy, dz)
// (model.py:347-353): one thread per view.  The reference spends ~40 tiny ATen launches on each.
__device__ __forceinline__ void view_fwd_one(const float* __restrict__ view, int vw, int B, float* __restrict__ R, float* __restrict__ t, int b) {
    const float* v = view + (long)b * vw;
    float sx, cx, sy, cy, sz, cz;
    sincosf(v[0], &sx, &cx);
    sincosf(v[1], &sy, &cy);
    sincosf(v[2], &sz, &cz);
    float* r = R + (long)b * 9;
    r[0] = mul(cz, cy); r[1] = add(mul(cz, mul(sy, sx)), mul(-sz, cx)); r[2] = add(mul(cz, mul(sy, cx)), mul(-sz, -sx));
    r[3] = mul(sz, cy); r[4] = add(mul(sz, mul(sy, sx)), mul(cz, cx));  r[5] = add(mul(sz, mul(sy, cx)), mul(cz, -sx));
    r[6] = -sy;         r[7] = mul(cy, sx);                             r[8] = mul(cy, cx);
    float* tt = t + (long)b * 3;
    tt[0] = vw >= 5 ? v[3] : 0.f;
    tt[1] = vw >= 5 ? v[4] : 0.f;
    tt[2] = vw >= 6 ? v[5] : 0.f;
}
__global__ void k_view_fwd(const float* __restrict__ view, int vw, int B, float* __restrict__ R, float* __restrict__ t) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    view_fwd_one(view, vw, B, R, t, b);
}

__device__ __forceinline__ void view_bwd_one(const float* __restrict__ view, int vw, int B, const float* __restrict__ gR,
                           const float* __restrict__ gt, float* __restrict__ gview, int b) {
    const float* v = view + (long)b * vw;
    float sx, cx, sy, cy, sz, cz;
    sincosf(v[0], &sx, &cx);
    sincosf(v[1], &sy, &cy);
    sincosf(v[2], &sz, &cz);
    float g[9];
#pragma unroll
    for (int k = 0; k < 9; k++) g[k] = gR ? gR[(long)b * 9 + k] : 0.f;
    const float g_sx = g[1] * cz * sy + g[2] * sz + g[4] * sz * sy - g[5] * cz + g[7] * cy;
    const float g_cx = -g[1] * sz + g[2] * cz * sy + g[4] * cz + g[5] * sz * sy + g[8] * cy;
    const float g_sy = g[1] * cz * sx + g[2] * cz * cx + g[4] * sz * sx + g[5] * sz * cx - g[6];
    const float g_cy = g[0] * cz + g[3] * sz + g[7] * sx + g[8] * cx;
    const float g_sz = -g[1] * cx + g[2] * sx + g[3] * cy + g[4] * sy * sx + g[5] * sy * cx;
    const float g_cz = g[0] * cy + g[1] * sy * sx + g[2] * sy * cx + g[4] * cx - g[5] * sx;
    float* o = gview + (long)b * vw;
    o[0] = g_sx * cx - g_cx * sx;
    o[1] = g_sy * cy - g_cy * sy;
    o[2] = g_sz * cz - g_cz * sz;
    if (vw >= 5) { o[3] = gt ? gt[(long)b * 3] : 0.f; o[4] = gt ? gt[(long)b * 3 + 1] : 0.f; }
    if (vw >= 6) o[5] = gt ? gt[(long)b * 3 + 2] : 0.f;
}
__global__ void k_view_bwd(const float* __restrict__ view, int vw, int B, const float* __restrict__ gR,
                           const float* __restrict__ gt, float* __restrict__ gview) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    view_bwd_one(view, vw, B, gR, gt, gview, b);
}

__device__ __forceinline__ void light_fwd_one(const float* __restrict__ light, int B, float* __restrict__ light5, int b) {
    const float* l = light + (long)b * 4;
    float* o = light5 + (long)b * 5;
    o[0] = add(dvd(l[0], 2.0f), 0.5f);
    o[1] = add(dvd(l[1], 2.0f), 0.5f);
    const float n = sqrt_(add(add(mul(l[2], l[2]), mul(l[3], l[3])), 1.0f));
    o[2] = dvd(l[2], n); o[3] = dvd(l[3], n); o[4] = dvd(1.0f, n);
}
__global__ void k_light_fwd(const float* __restrict__ light, int B, float* __restrict__ light5) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    light_fwd_one(light, B, light5, b);
}

__device__ __forceinline__ void light_bwd_one(const float* __restrict__ light, int B, const float* __restrict__ g5,
                            float* __restrict__ glight, int b) {
    const float* l = light + (long)b * 4;
    const float* g = g5 + (long)b * 5;
    const float n = sqrtf(l[2] * l[2] + l[3] * l[3] + 1.0f), inv = 1.0f / n;
    const float d0 = l[2] * inv, d1 = l[3] * inv, d2 = inv;
    const float dot = d0 * g[2] + d1 * g[3] + d2 * g[4];
    float* o = glight + (long)b * 4;
    o[0] = 0.5f * g[0];
    o[1] = 0.5f * g[1];
    o[2] = (g[2] - d0 * dot) * inv;
    o[3] = (g[3] - d1 * dot) * inv;
}
__global__ void k_light_bwd(const float* __restrict__ light, int B, const float* __restrict__ g5,
                            float* __restrict__ glight) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    light_bwd_one(light, B, g5, glight, b);
}

// view -> (R, t) and light -> (a, b, direction) of the same views in one launch (the one-image steps are launch-bound)
__global__ void k_view_light_fwd(const float* __restrict__ view, int vw, const float* __restrict__ light, int B,
                                 float* __restrict__ R, float* __restrict__ t, float* __restrict__ light5) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    view_fwd_one(view, vw, B, R, t, b);
    light_fwd_one(light, B, light5, b);
}

__global__ void k_view_light_bwd(const float* __restrict__ view, int vw, const float* __restrict__ light, int B,
                                 const float* __restrict__ gR, const float* __restrict__ gt, const float* __restrict__ g5,
                                 float* __restrict__ gview, float* __restrict__ glight) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    view_bwd_one(view, vw, B, gR, gt, gview, b);
    light_bwd_one(light, B, g5, glight, b);
}

// ------------------------------------------------------------------------------------------------
// Pixel-centric backward of the rasteriser.  The face-index map already says which face owns every sub-pixel, so
// the backward needs no candidate scan at all: one thread per OUTPUT pixel reads its 2x2 face indices, rebuilds each
// distinct face's 3x3 inverse from the projected vertices (k_project_verts, an L2-resident scratch), evaluates the
// weights / z of the sub-pixels that face owns, and adds the face's (u,v,z) vertex gradients ([nr]
// backward_depth_map, factored per face) to a per-view vertex-gradient scratch; k_vertex_bwd then pushes every
// vertex through projection / rotation to grad_depth, grad_R, grad_t.  neural_renderer does 9 float atomics per
// covered sub-pixel into grad_faces[B,F,3,3] and leaves the gather to autograd (index_put over 6 faces per vertex).

// v / S and v % S for 0 <= v < S*S, S <= 2048, without the generic integer-division sequence (~22 instructions per thread in
// the one-thread-per-vertex kernels): float estimate (within one of the quotient) + one fix-up
__device__ __forceinline__ void div_side(int v, int S, int* q_out, int* r_out) {
    int q = __float2int_rz(__fmul_rz((float)v, __frcp_rz((float)S)));
    int r = v - q * S;
    if (r >= S) { q++; r -= S; }
    else if (r < 0) { q--; r += S; }
    *q_out = q; *r_out = r;
}


// projected (u, v, z) of every vertex of the chunk's views: proj [chunk, S*S] float4 = (u, v, z, -) in NDC or, with PIX, the
// sub-pixel coordinates ndc_to_pix(u), ndc_to_pix(v) the face inverse is built from (what k_raster_bwd_px wants)
template <bool PIX>
__global__ void __launch_bounds__(PIX_THREADS)
k_project_verts(const Cam cam, const float* __restrict__ depth, long dstride, int vpi, const float* __restrict__ R,
                const float* __restrict__ t, int view0, float* __restrict__ proj, float* __restrict__ vgrad_zero) {
    __shared__ float sRt[12];
    __shared__ int s_img;
    const int S = cam.S, bl = blockIdx.y, b = view0 + bl;
    if (threadIdx.x < 9) sRt[threadIdx.x] = R[b * 9 + threadIdx.x];
    else if (threadIdx.x < 12) sRt[threadIdx.x] = t[b * 3 + threadIdx.x - 9];
    else if (threadIdx.x == 32) s_img = b / vpi;          // one division per CTA, not one per thread
    __syncthreads();
    const int v = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (v >= S * S) return;
    int vy, vx;
    div_side(v, S, &vy, &vx);
    float ray[3], q[3], ndc[3];
    pixel_ray(cam, vx, vy, ray);
    warp_point(cam, sRt, sRt + 9, ray, depth[(long)s_img * dstride + v], q);
    project_ndc(cam, q, ndc);
    if (PIX) { ndc[0] = ndc_to_pix(ndc[0], 2 * S); ndc[1] = ndc_to_pix(ndc[1], 2 * S); }
    reinterpret_cast<float4*>(proj)[(long)bl * S * S + v] = make_float4(ndc[0], ndc[1], ndc[2], 0.f);
    // the vertex-gradient scratch of the same vertex starts at zero (a 1 GB cudaMemset per step otherwise)
    reinterpret_cast<float4*>(vgrad_zero)[(long)bl * S * S + v] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// the same for vertices given as 3-D points (the neural_renderer-level entry g2s_render_depth_*): projection only
template <bool PIX>
__global__ void __launch_bounds__(PIX_THREADS)
k_project_points(const Cam cam, const float* __restrict__ verts3d, float* __restrict__ proj, float* __restrict__ vgrad_zero) {
    const int S = cam.S, bl = blockIdx.y;
    const int v = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (v >= S * S) return;
    const float* p = verts3d + ((long)bl * S * S + v) * 3;
    const float q[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    float ndc[3];
    project_ndc(cam, q, ndc);
    if (PIX) { ndc[0] = ndc_to_pix(ndc[0], 2 * S); ndc[1] = ndc_to_pix(ndc[1], 2 * S); }
    reinterpret_cast<float4*>(proj)[(long)bl * S * S + v] = make_float4(ndc[0], ndc[1], ndc[2], 0.f);
    reinterpret_cast<float4*>(vgrad_zero)[(long)bl * S * S + v] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// (u,v,z) NDC gradient -> gradient of the 3-D vertex ([nr] projection backward); WRITES grad_verts [chunk,S*S,3]
__global__ void __launch_bounds__(PIX_THREADS)
k_points_bwd(const Cam cam, const float* __restrict__ verts3d, const float* __restrict__ vgrad,
             float* __restrict__ grad_verts) {
    const int S = cam.S, bl = blockIdx.y;
    const int v = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (v >= S * S) return;
    const float4 gp = __ldcs(reinterpret_cast<const float4*>(vgrad) + (long)bl * S * S + v);
    const float* p = verts3d + ((long)bl * S * S + v) * 3;
    const float zz = __ldg(p + 2) + 1e-9f, iz = 1.0f / zz;
    const float x_ = __ldg(p) * iz, y_ = __ldg(p + 1) * iz;
    const float gup = gp.x * (2.0f / cam.os), gvp = -gp.y * (2.0f / cam.os);
    const float gx_ = gup * cam.K[0] + gvp * cam.K[3], gy_ = gup * cam.K[1] + gvp * cam.K[4];
    float* o = grad_verts + ((long)bl * S * S + v) * 3;
    o[0] = gx_ * iz;
    o[1] = gy_ * iz;
    o[2] = gp.z - (gx_ * x_ + gy_ * y_) * iz;
}

#ifndef G2S_RB_EXACT
#define G2S_RB_EXACT 0
#endif
// One (pixel, face) item: the face's 3x3 inverse from the projected vertices (sub-pixel coordinates), the owned
// sub-pixels' weights / z, the three 16-byte vertex reductions.  `mine` = which of the pixel's 2x2 sub-pixels (bit k:
// column k & 1, row k >> 1) the face owns.
// Arithmetic: the face inverse and the clamped weights are the forward's own operation sequence (the inverse is
// ill-conditioned -- fi * x + fi * y + fi cancels to a weight in [0,1] from terms of size ~is/2 -- so a 1-ulp change in fi is a
// 1e-5 change in a weight), and so are the six quotients of the x / y gradient (see below); the normalisation of the weights,
// the perspective z and the z gradient have no cancellation and run on refined reciprocals (MUFU.RCP + one Newton step,
// ~1 ulp) instead of correctly rounded quotients: gradients hold 1e-5, not bits.
__device__ __forceinline__ void raster_bwd_item(const Cam& cam, const float4* __restrict__ pv, float4* __restrict__ vg, int S,
                                                int is, int face, unsigned mine, int j, int i, float g) {
    const float hs = 0.5f * (float)is;
    int vidx[3];
    face_vertices(face, S, vidx);
    const float4 q0 = __ldg(&pv[vidx[0]]), q1 = __ldg(&pv[vidx[1]]), q2 = __ldg(&pv[vidx[2]]);
#if G2S_RB_EXACT
    float rec[16];
    bool zb;
    face_record_px(q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q0.z, q1.z, q2.z, rec, &zb);
    rec[FT_FLAG] = zb ? 1.0f : 0.0f;
    float A[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    while (mine) {
        const int k2 = __ffs(mine) - 1;
        mine &= mine - 1;
        const int xi = 2 * j + (k2 & 1), yi = is - 1 - (2 * i + (k2 >> 1));
        float w[3], zp = 0.f;
        record_weights_depth(rec, xi, yi, cam.near, cam.far, w, &zp);
        const float s = g * zp * zp;
        A[0] += s * w[0]; A[1] += s * w[1]; A[2] += s * w[2];
    }
    // [nr] backward_depth_map: tmp[l] = -sum_m face_inv[m][l] / z_m
    const float z[3] = {rec[9], rec[10], rec[11]};
    float qd[6];
    unsigned bad = 0;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const float yz = rcp_seed(z[m]);
        qd[m] = div_core(rec[3 * m], z[m], yz);
        qd[3 + m] = div_core(rec[3 * m + 1], z[m], yz);
        bad = max(bad, rec[3 * m] == 0.0f ? 0u : range_key(rec[3 * m]));
        bad = max(bad, rec[3 * m + 1] == 0.0f ? 0u : range_key(rec[3 * m + 1]));
    }
    if (bad >= RANGE_SPAN || rec[FT_FLAG] != 0.0f) {
#pragma unroll
        for (int m = 0; m < 3; m++) { qd[m] = __fdiv_rn(rec[3 * m], z[m]); qd[3 + m] = __fdiv_rn(rec[3 * m + 1], z[m]); }
    }
    const float t0 = -(qd[0] + qd[1] + qd[2]);
    const float t1 = -(qd[3] + qd[4] + qd[5]);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (A[m] == 0.f) continue;
        atomicAdd(&vg[vidx[m]], make_float4(-t0 * A[m] * hs, -t1 * A[m] * hs, __fdiv_rn(A[m], z[m] * z[m]), 0.f));
    }
#else
    float fi[REC_F];
    bool zb;
    face_record_px(q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q0.z, q1.z, q2.z, fi, &zb);
    const float rz[3] = {rcp_seed(q0.z), rcp_seed(q1.z), rcp_seed(q2.z)};
    float A[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    while (mine) {
        const int k2 = __ffs(mine) - 1;
        mine &= mine - 1;
        const float fx = (float)(2 * j + (k2 & 1)), fy = (float)(is - 1 - (2 * i + (k2 >> 1)));
        float wc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {      // [nr] kernel 2: the forward's clamped weights, same operations
            const float v = add(add(mul(fi[3 * k], fx), mul(fi[3 * k + 1], fy)), fi[3 * k + 2]);
            wc[k] = fminf(fmaxf(v, 0.0f), 1.0f);
        }
        const float ys = rcp_seed(add(add(add(0.0f, wc[0]), wc[1]), wc[2]));
        const float w0 = wc[0] * ys, w1 = wc[1] * ys, w2 = wc[2] * ys;
        const float zp = rcp_seed(w0 * rz[0] + w1 * rz[1] + w2 * rz[2]);
        const float s = g * zp * zp;
        A[0] += s * w0; A[1] += s * w1; A[2] += s * w2;
    }
    // [nr] backward_depth_map: tmp[l] = -sum_m face_inv[m][l] / z_m.  The columns of the inverse sum to zero (the weights sum to
    // one), so with three nearly equal z's this sum cancels to its last bits: correctly rounded quotients, summed in the
    // reference's order, or the result is off by 1e-4 (measured with reciprocals: profiles/r02_notes.md)
    const float t0 = -add(add(dvd(fi[0], q0.z), dvd(fi[3], q1.z)), dvd(fi[6], q2.z)) * hs;
    const float t1 = -add(add(dvd(fi[1], q0.z), dvd(fi[4], q1.z)), dvd(fi[7], q2.z)) * hs;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (A[m] == 0.f) continue;
        // one 16-byte vector reduction per vertex instead of three scalar ones
        atomicAdd(&vg[vidx[m]], make_float4(-t0 * A[m], -t1 * A[m], A[m] * rz[m] * rz[m], 0.f));
    }
#endif
}

// A CTA of four warps covers 16 x 32 output pixels, a warp 16 x 8 of them in four groups of 16 x 2.  A pixel holds 0-4
// distinct faces (2.9 on average where covered, none where the view left the frame), so a loop "one thread per pixel, one
// trip per face" runs at 19 of 32 lanes (profiles/r01_notes.md).  Instead every lane lists its pixel's (face, owned
// sub-pixels) ITEMS in a warp-private queue and the warp takes them off 32 at a time: all lanes busy in every round but
// the last of a warp's 128 pixels.  Warp-local throughout (no CTA barrier).
#ifndef G2S_RB_GROUPS
#define G2S_RB_GROUPS 4
#endif
constexpr int RB_GROUPS = G2S_RB_GROUPS, RBX = 16, RBY = 8 * RB_GROUPS, RB_THREADS = 128;
#ifndef G2S_RB_DUAL
#define G2S_RB_DUAL 1
#endif
// Items that own ONE sub-pixel (most of them) and items that own several are queued from the two ends of one buffer and
// taken off in separate rounds: the per-item loop over owned sub-pixels then runs one trip in the rounds of singles instead
// of the warp-wide maximum in every round.
constexpr int RB_QCAP = 32 * 4 + 2 * 32;     // one group's worst case on top of two remainders below 32
struct RasterBwdSmem {
    int face[RB_THREADS / 32][RB_QCAP];
    unsigned short meta[RB_THREADS / 32][RB_QCAP];    // pixel within the warp's 32 * RB_GROUPS (8 bits) | owned sub-pixels << 8
    float g[RB_THREADS / 32][32 * RB_GROUPS];
};

#ifndef G2S_RB_MINB
#define G2S_RB_MINB 8
#endif
__global__ void __launch_bounds__(RB_THREADS, G2S_RB_MINB)
k_raster_bwd_px(const Cam cam, const int* __restrict__ face_idx, const float* __restrict__ g_sub,
                const float* __restrict__ proj, float* __restrict__ vgrad, int view0) {
    __shared__ RasterBwdSmem sm;
    const int S = cam.S, is = 2 * S, bl = blockIdx.z, b = view0 + bl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * RBX + (lane & 15);
    const int i_w = blockIdx.y * RBY + warp * (2 * RB_GROUPS);
    if (i_w >= S) return;                                                            // whole warp below the image
    const int* fm = face_idx + (long)b * is * is;
    const float4* pv = reinterpret_cast<const float4*>(proj) + (long)bl * S * S;   // projected vertices, packed uvz-
    float4* vg = reinterpret_cast<float4*>(vgrad) + (long)bl * S * S;              // vertex gradients, packed uvz-
    int* qf = sm.face[warp];
    unsigned short* qm = sm.meta[warp];
    float* gs = sm.g[warp];
    // a group's inputs are requested one group ahead
    float g_next;
    int2 r0_next, r1_next;
    auto request = [&](int grp) {
        const int i = i_w + 2 * grp + (lane >> 4);
        const bool in = j < S && i < S;
        g_next = in ? __ldg(&g_sub[(long)bl * S * S + i * S + j]) : 0.f;
        r0_next = r1_next = make_int2(-1, -1);
        if (in) {
            r0_next = __ldg(reinterpret_cast<const int2*>(fm + (long)(2 * i) * is + 2 * j));
            r1_next = __ldg(reinterpret_cast<const int2*>(fm + (long)(2 * i + 1) * is + 2 * j));
        }
    };
    request(0);
    int n_one = 0, n_many = 0;      // queued singles (slots [0, n_one)) and multi-sub-pixel items (slots (CAP - 1 - n_many, CAP - 1])
    // ONE copy of the item body (four unrolled copies stall on instruction fetch, profiles/r01_notes.md)
#pragma unroll 1
    for (int grp = 0; grp < RB_GROUPS; grp++) {
        const float gq = g_next;
        const bool on = gq != 0.f;
        const int f0 = on ? r0_next.x : -1, f1 = on ? r0_next.y : -1, f2 = on ? r1_next.x : -1, f3 = on ? r1_next.y : -1;
        if (grp + 1 < RB_GROUPS) request(grp + 1);
        // first occurrence of every distinct face of the 2x2 block, and the sub-pixels it owns
        const bool a0 = f0 >= 0, a1 = f1 >= 0 && f1 != f0, a2 = f2 >= 0 && f2 != f0 && f2 != f1,
                   a3 = f3 >= 0 && f3 != f0 && f3 != f1 && f3 != f2;
        const unsigned m0 = 1u | (f1 == f0 ? 2u : 0u) | (f2 == f0 ? 4u : 0u) | (f3 == f0 ? 8u : 0u);
        const unsigned m1 = 2u | (f2 == f1 ? 4u : 0u) | (f3 == f1 ? 8u : 0u);
        const unsigned m2 = 4u | (f3 == f2 ? 8u : 0u);
        const bool s0 = G2S_RB_DUAL && (m0 & (m0 - 1u)) != 0u, s1 = G2S_RB_DUAL && (m1 & (m1 - 1u)) != 0u,
                   s2 = G2S_RB_DUAL && (m2 & (m2 - 1u)) != 0u;      // owns several sub-pixels
        const int c_many = (int)(a0 && s0) + (int)(a1 && s1) + (int)(a2 && s2);
        const int c_one = (int)a0 + (int)a1 + (int)a2 + (int)a3 - c_many;
        int tot;
        const int ex = warp_excl_scan(c_one | c_many << 8, &tot);
        int p_one = n_one + (ex & 255), p_many = RB_QCAP - 1 - (n_many + (ex >> 8));
        const unsigned pid = (unsigned)(grp * 32 + lane);
        gs[pid] = gq;
        if (a0) { const int q = s0 ? p_many-- : p_one++; qf[q] = f0; qm[q] = (unsigned short)(pid | m0 << 8); }
        if (a1) { const int q = s1 ? p_many-- : p_one++; qf[q] = f1; qm[q] = (unsigned short)(pid | m1 << 8); }
        if (a2) { const int q = s2 ? p_many-- : p_one++; qf[q] = f2; qm[q] = (unsigned short)(pid | m2 << 8); }
        if (a3) { qf[p_one] = f3; qm[p_one] = (unsigned short)(pid | 8u << 8); }
        n_one += tot & 255;
        n_many += tot >> 8;
        __syncwarp();
        // full rounds off the ends of the queue; what stays (< 32 of either kind) waits for the next group, the last group
        // drains it in mixed rounds
        const bool last = grp == RB_GROUPS - 1;
#pragma unroll 1
        while (true) {
            int src = -1;
            if (n_one >= 32) { n_one -= 32; src = n_one + lane; }
            else if (n_many >= 32) { n_many -= 32; src = RB_QCAP - 1 - (n_many + lane); }
            else if (last && n_one + n_many > 0) {
                const int t_one = n_one, t_many = min(n_many, 32 - t_one);
                n_one = 0; n_many -= t_many;
                if (lane < t_one) src = lane;
                else if (lane - t_one < t_many) src = RB_QCAP - 1 - (n_many + lane - t_one);
            } else break;
            int face = -1;
            unsigned meta = 0;
            if (src >= 0) { face = qf[src]; meta = qm[src]; }
            const unsigned p = meta & 255u;
            const float g = gs[p];
            __syncwarp();                 // the slots are free for the next group's items
            if (face >= 0)
                raster_bwd_item(cam, pv, vg, S, is, face, meta >> 8, blockIdx.x * RBX + (int)(p & 15u),
                                i_w + 2 * (int)(p >> 5) + (int)((p >> 4) & 1u), g);
        }
    }
}

// vertex chain: (u,v,z) NDC gradient -> 3-D point -> depth, R, t.  One thread per vertex.
// REZERO: what was consumed is set back to zero (the scratch is then zero at rest and nobody has to clear 1 GB of it per
// step); the store follows the test on the loaded value, so the load it could collide with has already returned.
#ifndef G2S_VB_VPT
#define G2S_VB_VPT 8
#endif
constexpr int VB_VPT = G2S_VB_VPT;
template <bool REZERO>
__global__ void __launch_bounds__(PIX_THREADS)
k_vertex_bwd(const Cam cam, const float* __restrict__ depth, long dstride, int vpi, const float* __restrict__ R,
             const float* __restrict__ t, int view0, float* __restrict__ vgrad, float* __restrict__ grad_depth,
             long gdstride, float* __restrict__ grad_R, float* __restrict__ grad_t) {
    __shared__ float sRt[12];
    __shared__ int s_img;
    const int S = cam.S, bl = blockIdx.y, b = view0 + bl;
    if (threadIdx.x < 9) sRt[threadIdx.x] = R[b * 9 + threadIdx.x];
    else if (threadIdx.x < 12) sRt[threadIdx.x] = t[b * 3 + threadIdx.x - 9];
    else if (threadIdx.x == 32) s_img = b / vpi;
    __syncthreads();
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    // VB_VPT vertices per thread: the 12-value block sum behind grad_R / grad_t (a sixth of this kernel's instructions) is paid
    // once per VB_VPT vertices (1 / 2 / 4 / 8 / 16 per thread: 0.77 / 0.80 / 0.67 / 0.61 / 0.58 ms per 4096 views).  The same loop in
    // k_render_bwd_pixel spills at its 40-register cap (1.78 -> 2.26 ms) and was not kept.
#pragma unroll 1
    for (int it = 0; it < VB_VPT; it++) {
    const int v = (blockIdx.x * VB_VPT + it) * PIX_THREADS + threadIdx.x;
    if (v < S * S) {
        float4* gptr = reinterpret_cast<float4*>(vgrad) + (long)bl * S * S + v;
        const float4 gp = __ldcs(gptr);
        const float gu = gp.x, gv = gp.y, gz = gp.z;
        if (gu != 0.f || gv != 0.f || gz != 0.f) {
            if (REZERO) __stcs(gptr, make_float4(0.f, 0.f, 0.f, 0.f));
            int vy, vx;
            div_side(v, S, &vy, &vx);
            float ray[3], q[3];
            pixel_ray(cam, vx, vy, ray);
            const float d = depth[(long)s_img * dstride + v];
            warp_point(cam, sRt, sRt + 9, ray, d, q);
            const float p3[3] = {ray[0] * d, ray[1] * d, ray[2] * d - cam.rcd};
            const float zz = q[2] + 1e-9f, iz = 1.0f / zz;
            const float x_ = q[0] * iz, y_ = q[1] * iz;
            const float gup = gu * (2.0f / cam.os), gvp = -gv * (2.0f / cam.os);
            const float gx_ = gup * cam.K[0] + gvp * cam.K[3], gy_ = gup * cam.K[1] + gvp * cam.K[4];
            const float gq[3] = {gx_ * iz, gy_ * iz, gz - (gx_ * x_ + gy_ * y_) * iz};
            float gd = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float gvk = gq[0] * sRt[k] + gq[1] * sRt[3 + k] + gq[2] * sRt[6 + k];
                gd += gvk * ray[k];
#pragma unroll
                for (int jj = 0; jj < 3; jj++) acc[3 * jj + k] += gq[jj] * p3[k];
                acc[9 + k] += gq[k];
            }
            float* o = &grad_depth[(long)s_img * gdstride + v];
            if (vpi == 1 && gdstride != 0) *o += gd;   // one view per depth map: this thread is the only writer
            else atomicAdd(o, gd);
        }
    }
    }
    if (grad_R) {
        block_accumulate_Rt<PIX_THREADS>(acc, grad_R + b * 9, grad_t + b * 3);
    }
}

// ------------------------------------------------------------------------------------------------
// fused render backward, pixel stage: clamp(-1,1) -> shaded bilinear sampling -> inverse warp grid.
// Writes per-view texture gradients (atomics into grad_tex_ws), the masked quarter gradient of
// recon_depth (g_sub) and accumulates grad_R / grad_t.  grad_tex / g_sub are indexed by the LOCAL view
// (chunked launches), everything else by the global view fa.view0 + blockIdx.z.
// 12 CTAs of 128 threads per SM (40 registers): measured best of 1 / 12 / 16 (54-64 / 40 / 32 registers + spills)
// LOSS: the fused photometric loss's cotangent is formed here (a variant of its own: the extra live values do not fit the 40
// registers of the plain kernel without spills)
template <bool LOSS>
__global__ void __launch_bounds__(BPX * BPY, LOSS ? 10 : 12)
k_render_bwd_pixel(const Cam cam, const FusedArgs fa, const float* __restrict__ recon_depth,
                   const float* __restrict__ grad_recon_im, const float* __restrict__ grad_recon_depth,
                   float* __restrict__ g_sub, float* __restrict__ grad_tex, float* __restrict__ grad_R,
                   float* __restrict__ grad_t) {
    __shared__ float sview[17];
    __shared__ int s_img;
    const int S = cam.S, bl = blockIdx.z, b = fa.view0 + bl;
    const int j = blockIdx.x * BPX + threadIdx.x, i = blockIdx.y * BPY + threadIdx.y;
    const bool inside = j < S && i < S;
    const int pix = i * S + j;
    // per-pixel inputs are requested before the barrier that publishes R, t, light: one round trip instead of two
    float rd = 0.f, G[3] = {0.f, 0.f, 0.f};
    if (inside) {
        rd = recon_depth[(long)b * S * S + pix];
        if (!LOSS || grad_recon_im) {
#pragma unroll
            for (int c = 0; c < 3; c++) G[c] = __ldcs(&grad_recon_im[((long)b * 3 + c) * S * S + pix]);
        }
    }
    // fused photometric loss: its cotangent sign(recon_im - target) * mask * g_loss / den is formed here from the re-computed
    // render instead of travelling through memory (k_photo_bwd's write, autograd's add, this kernel's read)
    float tg[3] = {0.f, 0.f, 0.f}, lscale = 0.f;
    if (LOSS && inside) {
#pragma unroll
        for (int c = 0; c < 3; c++) tg[c] = __ldcs(&fa.target[((long)b * 3 + c) * S * S + pix]);
        const float vm = fa.vmask ? __ldcs(&fa.vmask[(long)b * S * S + pix]) : 1.f;
        lscale = (rd < fa.thresh ? 1.f : 0.f) * vm * (__ldg(fa.gloss) / __ldg(&fa.sums3[2]));
    }
    {
        const int k = threadIdx.y * BPX + threadIdx.x;
        if (k < 9) sview[k] = fa.R[b * 9 + k];
        else if (k < 12) sview[k] = fa.t[b * 3 + k - 9];
        else if (k < 17) sview[k] = fa.light[b * 5 + k - 12];
        else if (k == 32) s_img = b / fa.vpi;          // one division per CTA, not one per thread
        __syncthreads();
    }
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    if (inside) {
        const int img = s_img;
        float ray[3], q[3], v[3], g[2];
        pixel_ray(cam, j, i, ray);
        inv_warp_point(cam, sview, sview + 9, ray, rd, q, v);
        point_to_grid(cam, q, S, S, g);
        const Taps tp = make_taps(g[0], g[1], S, S, fa.align);
        float tex[4][3];
        shade_taps(fa.normal + (long)img * S * S * TEXEL, tp, sview + 12, tex);
        float gix = 0.f, giy = 0.f;
        // per-view texture gradient, packed rgb- per texel: one 16-byte vector reduction per tap instead of three scalar ones
        float4* gt_b = reinterpret_cast<float4*>(grad_tex) + (long)bl * S * S;
        float Gc[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float o = tp.w[0] * tex[0][c] + tp.w[1] * tex[1][c] + tp.w[2] * tex[2][c] + tp.w[3] * tex[3][c];
            // clamp(-1,1) passes the gradient where -1 <= x <= 1 (so the clamped render equals o wherever it matters)
            float Gl = G[c];
            if (LOSS) {
                const float dlt = o - tg[c];
                Gl += dlt > 0.f ? lscale : (dlt < 0.f ? -lscale : 0.f);
            }
            Gc[c] = (o >= -1.f && o <= 1.f) ? Gl : 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                gix += ((k & 1) ? 1.f : -1.f) * tp.wy[k] * tex[k][c] * Gc[c];
                giy += ((k >> 1) ? 1.f : -1.f) * tp.wx[k] * tex[k][c] * Gc[c];
            }
        }
        if (Gc[0] != 0.f || Gc[1] != 0.f || Gc[2] != 0.f) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (tp.w[k] != 0.f) atomicAdd(&gt_b[tp.p[k]], make_float4(tp.w[k] * Gc[0], tp.w[k] * Gc[1], tp.w[k] * Gc[2], 0.f));
        }
        const float mult = fa.align ? (float)(S - 1) * 0.5f : (float)S * 0.5f;
        float gd = warp_grid_bwd_pixel(cam, sview, sview + 9, j, i, S, S, rd, 1, gix * mult, giy * mult, acc);
        if (grad_recon_depth) gd += grad_recon_depth[(long)b * S * S + pix];
        g_sub[(long)bl * S * S + pix] = (rd > cam.clamp_lo && rd < cam.clamp_hi) ? 0.25f * gd : 0.f;
    }
    block_accumulate_Rt<BPX * BPY>(acc, grad_R + b * 9, grad_t + b * 3);
}

// fused render backward, texture stage: per image pixel, loop over the image's views that fall in this
// chunk [view0, view0+nviews) and turn the per-view texture gradients into grad_albedo, grad_normal
// (registers, accumulated with a plain += : one thread owns one (image, pixel)) and grad_light (warp
// reduction + one atomic per warp and view).  grid = (pixel blocks, images spanned by the chunk).
__global__ void __launch_bounds__(PIX_THREADS)
k_render_bwd_tex(int S, const FusedArgs fa, int nviews, float* __restrict__ grad_tex,
                 float* __restrict__ grad_albedo, float* __restrict__ grad_normal, float* __restrict__ grad_light) {
    const int img = fa.view0 / fa.vpi + blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    const bool live = pix < S * S;
    const int p = live ? pix : 0;
    const float* nimg = fa.normal + (long)img * S * S * TEXEL;
    const float* aimg = fa.albedo + (long)img * S * S * 3;
    const float n0 = nimg[p * TEXEL], n1 = nimg[p * TEXEL + 1], n2 = nimg[p * TEXEL + 2];
    const float al[3] = {aimg[p], aimg[S * S + p], aimg[2 * S * S + p]};
    float ga[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f};
    // the image's views inside this chunk, split over gridDim.z groups of threads (an image with many views -- the face
    // config has 1024 -- would otherwise be a serial loop on too few threads)
    int b0 = max(img * fa.vpi, fa.view0), b1 = min((img + 1) * fa.vpi, fa.view0 + nviews);
    if (gridDim.z > 1) {
        const int per = (b1 - b0 + (int)gridDim.z - 1) / (int)gridDim.z;
        b0 += (int)blockIdx.z * per;
        b1 = min(b1, b0 + per);
    }
    for (int b = b0; b < b1; b++) {
        const float la = __ldg(&fa.light[b * 5]), lb = __ldg(&fa.light[b * 5 + 1]), dx = __ldg(&fa.light[b * 5 + 2]),
                    dy = __ldg(&fa.light[b * 5 + 3]), dz = __ldg(&fa.light[b * 5 + 4]);
        float T[3] = {0.f, 0.f, 0.f};
        if (live) {
            float4* tp = reinterpret_cast<float4*>(grad_tex) + (long)(b - fa.view0) * S * S + p;
            const float4 t4 = __ldcs(tp);
            T[0] = t4.x; T[1] = t4.y; T[2] = t4.z;
        }
        const float ndl = n0 * dx + n1 * dy + n2 * dz;
        const float diff = fmaxf(ndl, 0.f);
        const float sh = la + lb * diff;
        // tex_c = (al_c/2 + .5) * sh * 2 - 1
        float dsh = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            ga[c] += T[c] * sh;
            dsh += T[c] * (al[c] + 1.0f);
        }
        const float ddiff = ndl >= 0.f ? dsh * lb : 0.f;
        gn[0] += ddiff * dx; gn[1] += ddiff * dy; gn[2] += ddiff * dz;
        float lg[8] = {dsh, dsh * diff, ddiff * n0, ddiff * n1, ddiff * n2, 0.f, 0.f, 0.f};
        const int lane = threadIdx.x & 31, li = halving_index<3>(lane);
        const float ls = warp_reduce_halving<3>(lg, lane);
        if ((lane & 3) == 0 && li < 5 && ls != 0.f) atomicAdd(&grad_light[b * 5 + li], ls);
    }
    // What was consumed is left at zero for the next chunk's reductions (instead of a 1 GB cudaMemset per step) -- in a loop of
    // its own, after the sums above have waited for every load: a store to a line that a load is still waiting on holds the
    // load / store unit until the load returns (inside the loop above it serialised the loads: 0.46 -> 4.8 ms).
    if (live) {
        for (int b = b0; b < b1; b++)
            __stcs(reinterpret_cast<float4*>(grad_tex) + (long)(b - fa.view0) * S * S + p, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (live && b1 > b0) {
        float* o = grad_normal + ((long)img * S * S + p) * 3;
        if (gridDim.z > 1) {       // several groups share the (image, pixel)
#pragma unroll
            for (int c = 0; c < 3; c++) atomicAdd(&grad_albedo[((long)img * 3 + c) * S * S + p], ga[c]);
            atomicAdd(&o[0], gn[0]); atomicAdd(&o[1], gn[1]); atomicAdd(&o[2], gn[2]);
        } else {                   // one thread owns the (image, pixel): plain accumulation over the chunks
#pragma unroll
            for (int c = 0; c < 3; c++) grad_albedo[((long)img * 3 + c) * S * S + p] += ga[c];
            o[0] += gn[0]; o[1] += gn[1]; o[2] += gn[2];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mesh-texture branch (render_yaw / render_view / render_given_view with grid_sample=False)

// rotate about the centroid (0,0,rcd): mm3(p - c0, M) + c0 with M = R (transpose = 0) or R^T (1)
__device__ __forceinline__ void rotate_point(const Cam& cam, const float* R, int transpose, float p[3]) {
    const float v0 = p[0], v1 = p[1], v2 = sub(p[2], cam.rcd);
    float q[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float m0 = transpose ? R[j] : R[3 * j], m1 = transpose ? R[3 + j] : R[3 * j + 1],
                    m2 = transpose ? R[6 + j] : R[3 * j + 2];
        float acc = mul(v0, m0);
        acc = fma_(v1, m1, acc);
        acc = fma_(v2, m2, acc);
        q[j] = acc;
    }
    p[0] = add(q[0], 0.0f); p[1] = add(q[1], 0.0f); p[2] = add(q[2], cam.rcd);
}

struct Crop {
    int on, top, bottom, left, right;
};

// depth_to_3d_grid + crop + inverse warp (R0,t0) + rotation R1 + warp (R2,t2): renderer.py:143-192
__global__ void __launch_bounds__(PIX_THREADS)
k_grid3d(const Cam cam, const float* __restrict__ depth, long dstride, int H, int W, const Crop crop,
         const float* __restrict__ R0, const float* __restrict__ t0, const float* __restrict__ R1,
         const float* __restrict__ R2, const float* __restrict__ t2, float* __restrict__ out) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const float* dimg = depth + (long)b * dstride;
    float p[3];
    if (crop.on) {
        const int ys = min(max(y, crop.top), H - 1 - crop.bottom), xs = min(max(x, crop.left), W - 1 - crop.right);
        float a[3];
        depth_point(cam, dimg, W, xs, y, a);
        p[0] = a[0];
        depth_point(cam, dimg, W, x, ys, a);
        p[1] = a[1];
        depth_point(cam, dimg, W, xs, ys, a);
        p[2] = a[2];
    } else {
        depth_point(cam, dimg, W, x, y, p);
    }
    if (R0) {
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = add(p[k], -__ldg(&t0[b * 3 + k]));
        float Rm[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R0[b * 9 + k]);
        rotate_point(cam, Rm, 1, p);
    }
    if (R1) {
        float Rm[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R1[b * 9 + k]);
        rotate_point(cam, Rm, 0, p);
    }
    if (R2) {
        float Rm[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R2[b * 9 + k]);
        rotate_point(cam, Rm, 0, p);
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = add(p[k], __ldg(&t2[b * 3 + k]));
    }
    float* o = out + ((long)b * H * W + pix) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

// Backward of depth_to_3d_grid (mode 0) / get_warped_3d_grid (1) / get_inv_warped_3d_grid (2): renderer.py:74-102
__global__ void __launch_bounds__(PIX_THREADS)
k_grid3d_bwd(const Cam cam, const float* __restrict__ depth, long dstride, int H, int W, int mode, const float* __restrict__ R,
             const float* __restrict__ t, const float* __restrict__ grad_out, float* __restrict__ grad_depth,
             float* __restrict__ grad_R, float* __restrict__ grad_t) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    if (pix < H * W) {
        const int y = pix / W, x = pix - y * W;
        float ray[3];
        pixel_ray(cam, x, y, ray);
        const float* go = grad_out + ((long)b * H * W + pix) * 3;
        const float g[3] = {go[0], go[1], go[2]};
        float gd;
        if (mode == 0) {
            gd = g[0] * ray[0] + g[1] * ray[1] + g[2] * ray[2];
        } else {
            float Rm[9], tv[3];
#pragma unroll
            for (int k = 0; k < 9; k++) Rm[k] = __ldg(&R[b * 9 + k]);
#pragma unroll
            for (int k = 0; k < 3; k++) tv[k] = __ldg(&t[b * 3 + k]);
            const float d = depth[(long)b * dstride + pix];
            float dv[3];
            if (mode == 1) {          // q = R v + c0 + t,  v = ray d - c0
                const float v[3] = {ray[0] * d, ray[1] * d, ray[2] * d - cam.rcd};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    dv[k] = g[0] * Rm[k] + g[1] * Rm[3 + k] + g[2] * Rm[6 + k];
#pragma unroll
                    for (int j = 0; j < 3; j++) acc[3 * j + k] += g[j] * v[k];
                    acc[9 + k] += g[k];
                }
            } else {                  // q = R^T v + c0,  v = ray d - t - c0
                const float v[3] = {ray[0] * d - tv[0], ray[1] * d - tv[1], ray[2] * d - tv[2] - cam.rcd};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    dv[k] = g[0] * Rm[3 * k] + g[1] * Rm[3 * k + 1] + g[2] * Rm[3 * k + 2];
#pragma unroll
                    for (int j = 0; j < 3; j++) acc[3 * k + j] += v[k] * g[j];
                    acc[9 + k] -= dv[k];
                }
            }
            gd = dv[0] * ray[0] + dv[1] * ray[1] + dv[2] * ray[2];
        }
        grad_depth[(long)b * H * W + pix] = gd;
    }
    if (grad_R) {
        block_accumulate_Rt<PIX_THREADS>(acc, grad_R + b * 9, grad_t + b * 3);
    }
}

// grid_3d_to_2d: renderer.py:82-88
__global__ void __launch_bounds__(PIX_THREADS)
k_grid_3d_to_2d_fwd(const Cam cam, const float* __restrict__ grid3d, int H, int W, float* __restrict__ grid) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const float* p = grid3d + ((long)b * H * W + pix) * 3;
    const float q[3] = {p[0], p[1], p[2]};
    float g[2];
    point_to_grid(cam, q, W, H, g);
    reinterpret_cast<float2*>(grid)[(long)b * H * W + pix] = make_float2(g[0], g[1]);
}

__global__ void __launch_bounds__(PIX_THREADS)
k_grid_3d_to_2d_bwd(const Cam cam, const float* __restrict__ grid3d, int H, int W, const float* __restrict__ grad_grid,
                    float* __restrict__ grad_grid3d) {
    const int b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const float* p = grid3d + ((long)b * H * W + pix) * 3;
    const float2 G = reinterpret_cast<const float2*>(grad_grid)[(long)b * H * W + pix];
    const float iz = 1.0f / p[2], nx = p[0] * iz, ny = p[1] * iz;
    const float dpx = G.x * 2.0f / (float)(W - 1), dpy = G.y * 2.0f / (float)(H - 1);
    // (px, py) = K_grid (nx, ny, nz) with nz = z / z = 1: no gradient through nz
    const float dnx = dpx * cam.Kg[0] + dpy * cam.Kg[3], dny = dpx * cam.Kg[1] + dpy * cam.Kg[4];
    float* o = grad_grid3d + ((long)b * H * W + pix) * 3;
    o[0] = dnx * iz;
    o[1] = dny * iz;
    o[2] = -(dnx * nx + dny * ny) * iz;
}

struct Bg {
    float c[4];
};

// utils.py:83-95 cube coefficients (rows i = i0*4 + i1*2 + i2)
__device__ __constant__ float kCube[8][3] = {{0.5f, 0.5f, 0.5f}, {0.f, 0.f, 1.f}, {0.f, 1.f, 0.f}, {-0.5f, 0.5f, 0.5f},
                                             {1.f, 0.f, 0.f}, {0.5f, -0.5f, 0.5f}, {0.5f, 0.5f, -0.5f}, {0.f, 0.f, 0.f}};

// colour of one covered sub-pixel: [nr] forward_texture_sampling with the 2^3 cube of utils.py:98-109
// `coef` (optional): the colour is LINEAR in the three vertex colours, out[c] = sum_j coef[j] * im[c][cidx[j]]; [nr]
// backward_textures chained through get_textures_from_im (utils.py:98-109) is the transpose of that map.
template <int C>
__device__ __forceinline__ void rgb_subpixel(const Cam& cam, const float* __restrict__ verts_b,
                                             const float* __restrict__ im_b, int face, int xi, int yi, float zp_key,
                                             float eps, float out[C], float* coef = nullptr, int* cidx = nullptr) {
    const int S = cam.S, is = 2 * S, Q = (S - 1) * (S - 1);
    int vidx[3];
    face_vertices(face, S, vidx);
    float ndc[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float* p = &verts_b[(long)vidx[k] * 3];
        const float q[3] = {p[0], p[1], p[2]};
        project_ndc(cam, q, ndc[k]);
    }
    const Tri f = make_tri(ndc[0], ndc[1], ndc[2]);
    float fi[9], w[3], zp = zp_key;
    tri_face_inv(f, is, fi);
    tri_weights_depth(f, fi, xi, yi, cam.near, cam.far, w, &zp);
    const float z[3] = {f.z0, f.z1, f.z2};
    float tif[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float v = mul(mul(w[k], 1.0f), dvd(zp, z[k]));
        v = fmaxf(v, 0.f);
        v = fminf(v, sub(1.0f, eps));
        tif[k] = v;
    }
    // vertex colours in the order get_textures_from_im stacks them
    const bool rev = face >= 2 * Q;
    int fb = rev ? face - 2 * Q : face;
    const bool second = fb >= Q;
    if (second) fb -= Q;
    const int qy = fb / (S - 1), qx = fb - qy * (S - 1);
    const int p00 = qy * S + qx;
    int cp[3];
    if (!second) { cp[0] = p00; cp[1] = p00 + 1; cp[2] = p00 + S; }
    else { cp[0] = p00 + S; cp[1] = p00 + 1; cp[2] = p00 + S + 1; }
#pragma unroll
    for (int c = 0; c < C; c++) {
        const float v0 = im_b[(long)c * S * S + cp[0]], v1 = im_b[(long)c * S * S + cp[1]],
                    v2 = im_b[(long)c * S * S + cp[2]];
        float acc = 0.f;
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            float wt = 1.f;
            int ti[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float fr = sub(tif[k], (float)(int)tif[k]);
                if (((pn >> k) & 1) == 0) { wt = mul(wt, sub(1.f, fr)); ti[k] = (int)tif[k]; }
                else { wt = mul(wt, fr); ti[k] = (int)tif[k] + 1; }
            }
            // fill_back copies see the cube with axes 0 and 2 swapped
            const int ci = rev ? ti[2] * 4 + ti[1] * 2 + ti[0] : ti[0] * 4 + ti[1] * 2 + ti[2];
            float tex = mul(kCube[ci][0], v0);
            tex = fma_(kCube[ci][1], v1, tex);
            tex = fma_(kCube[ci][2], v2, tex);
            acc = add(acc, mul(wt, tex));        // un-fused, in corner order: the same bits as [nr] forward_texture_sampling
        }
        out[c] = acc;
    }
    if (coef) {
        coef[0] = coef[1] = coef[2] = 0.f;
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            float wt = 1.f;
            int ti[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float fr = tif[k] - (float)(int)tif[k];
                if (((pn >> k) & 1) == 0) { wt *= 1.f - fr; ti[k] = (int)tif[k]; }
                else { wt *= fr; ti[k] = (int)tif[k] + 1; }
            }
            const int ci = rev ? ti[2] * 4 + ti[1] * 2 + ti[0] : ti[0] * 4 + ti[1] * 2 + ti[2];
#pragma unroll
            for (int j = 0; j < 3; j++) coef[j] += wt * kCube[ci][j];
        }
        cidx[0] = cp[0]; cidx[1] = cp[1]; cidx[2] = cp[2];
    }
}

// Backward of k_resolve_rgb with respect to the per-vertex colours `im`: [nr] backward_textures + the autograd of
// get_textures_from_im / fill_back / the 2x2 mean / clamp(-1,1).  One thread per output pixel; grad_im is ACCUMULATED with
// atomics (im may be shared by all views).  The vertex (geometry) gradient of an rgb render -- [nr] backward_pixel_map,
// an approximate edge gradient -- is not built (the reference never differentiates render_rgb, SURVEY.md 8a').
template <int C>
__global__ void __launch_bounds__(PIX_THREADS)
k_resolve_rgb_bwd(const Cam cam, const int* __restrict__ face_idx, const float* __restrict__ verts3d,
                  const float* __restrict__ im, long imstride, const Bg bg, float eps, int clampv,
                  const float* __restrict__ grad_rgb, float* __restrict__ grad_im, long gstride) {
    const int S = cam.S, is = 2 * S, b = blockIdx.y;
    const int pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= S * S) return;
    const int i = pix / S, j = pix - i * S;
    const int* fm = face_idx + (long)b * is * is;
    const int2 r0 = *reinterpret_cast<const int2*>(fm + (long)(2 * i) * is + 2 * j);
    const int2 r1 = *reinterpret_cast<const int2*>(fm + (long)(2 * i + 1) * is + 2 * j);
    const int faces[4] = {r0.x, r0.y, r1.x, r1.y};
    const float* verts_b = verts3d + (long)b * S * S * 3;
    const float* im_b = im + (long)b * imstride;
    float sum[C], coef[4][3];
    int cidx[4][3];
#pragma unroll
    for (int c = 0; c < C; c++) sum[c] = 0.f;
#pragma unroll 1
    for (int sp = 0; sp < 4; sp++) {
        float col[C];
        coef[sp][0] = coef[sp][1] = coef[sp][2] = 0.f;
        cidx[sp][0] = cidx[sp][1] = cidx[sp][2] = 0;
        if (faces[sp] < 0) {
#pragma unroll
            for (int c = 0; c < C; c++) col[c] = bg.c[c];
        } else {
            const int r = 2 * i + (sp >> 1), xi = 2 * j + (sp & 1);
            rgb_subpixel<C>(cam, verts_b, im_b, faces[sp], xi, is - 1 - r, 0.f, eps, col, coef[sp], cidx[sp]);
        }
#pragma unroll
        for (int c = 0; c < C; c++) sum[c] += col[c];
    }
    float* g_b = grad_im + (long)b * gstride;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const float v = sum[c] * 0.25f;
        float g = grad_rgb[((long)b * C + c) * S * S + pix] * 0.25f;
        if (clampv && !(v >= -1.f && v <= 1.f)) g = 0.f;
        if (g == 0.f) continue;
#pragma unroll
        for (int sp = 0; sp < 4; sp++) {
            if (faces[sp] < 0) continue;
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (coef[sp][k] != 0.f) atomicAdd(&g_b[(long)c * S * S + cidx[sp][k]], g * coef[sp][k]);
        }
    }
}

// ---- geometry gradient of an rgb render: [nr] backward_pixel_map ---------------------------------
// supersampled colour map [n,2S,2S,4] (image orientation; what nr keeps as rgb_map) ...
__global__ void __launch_bounds__(PIX_THREADS)
k_rgb_map(const Cam cam, const int* __restrict__ face_idx, const float* __restrict__ verts3d, const float* __restrict__ im,
          long imstride, const Bg bg, float eps, float* __restrict__ rgb_map) {
    const int S = cam.S, is = 2 * S, b = blockIdx.y;
    const long sp = (long)blockIdx.x * PIX_THREADS + threadIdx.x;
    if (sp >= (long)is * is) return;
    const int r = (int)(sp / is), c = (int)(sp - (long)r * is);
    const int face = face_idx[(long)b * is * is + sp];
    float col[3] = {bg.c[0], bg.c[1], bg.c[2]};
    if (face >= 0) rgb_subpixel<3>(cam, verts3d + (long)b * S * S * 3, im + (long)b * imstride, face, c, is - 1 - r, 0.f, eps, col);
    reinterpret_cast<float4*>(rgb_map)[(long)b * is * is + sp] = make_float4(col[0], col[1], col[2], 0.f);
}

// ... and the gradient every sub-pixel of an output pixel receives: a quarter of the pixel's cotangent, zero where the
// clamp(-1,1) of the 2x2 mean is active
__global__ void __launch_bounds__(PIX_THREADS)
k_rgb_gquarter(int S, const float* __restrict__ rgb_map, const float* __restrict__ grad_rgb, int clampv, float* __restrict__ g4) {
    const int is = 2 * S, b = blockIdx.y, pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= S * S) return;
    const int i = pix / S, j = pix - i * S;
    const float4* m = reinterpret_cast<const float4*>(rgb_map) + (long)b * is * is;
    const float4 a = m[(long)(2 * i) * is + 2 * j], bq = m[(long)(2 * i) * is + 2 * j + 1], c = m[(long)(2 * i + 1) * is + 2 * j],
                 d = m[(long)(2 * i + 1) * is + 2 * j + 1];
    const float v[3] = {(a.x + bq.x + c.x + d.x) * 0.25f, (a.y + bq.y + c.y + d.y) * 0.25f, (a.z + bq.z + c.z + d.z) * 0.25f};
    float g[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        g[k] = grad_rgb[((long)b * 3 + k) * S * S + pix] * 0.25f;
        if (clampv && !(v[k] >= -1.f && v[k] <= 1.f)) g[k] = 0.f;
    }
    reinterpret_cast<float4*>(g4)[(long)b * S * S + pix] = make_float4(g[0], g[1], g[2], 0.f);
}

// One thread per face (fill_back copies included): for each edge and each axis walk the integer positions between the edge's
// end points; "out" pass from the pixel just outside the edge to the image border (only when the pixel just inside belongs
// to this face), "in" pass from the pixel just inside to the opposite edge over this face's own pixels; a visited pixel whose
// colour difference to the pixel across the edge correlates positively with its gradient pulls the edge's two vertices by
// diff_grad / dist.  The face's x / y gradients go to the per-view vertex-gradient scratch (uvz-).
__global__ void __launch_bounds__(128)
k_backward_pixel_map(const Cam cam, const int* __restrict__ face_idx, const float* __restrict__ proj,
                     const float* __restrict__ rgb_map, const float* __restrict__ g4, float eps, float* __restrict__ vgrad) {
    const int S = cam.S, is = 2 * S, b = blockIdx.y, Q = (S - 1) * (S - 1);
    const int fn = blockIdx.x * 128 + threadIdx.x;
    if (fn >= 4 * Q) return;
    int vidx[3];
    face_vertices(fn, S, vidx);
    const float4* pv = reinterpret_cast<const float4*>(proj) + (long)b * S * S;
    float face[9];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const float4 q = __ldg(&pv[vidx[m]]);
        face[3 * m] = q.x; face[3 * m + 1] = q.y; face[3 * m + 2] = q.z;
    }
    if (mul(sub(face[7], face[1]), sub(face[3], face[0])) < mul(sub(face[4], face[1]), sub(face[6], face[0]))) return;
    const int* fmap = face_idx + (long)b * is * is;
    const float4* cmap = reinterpret_cast<const float4*>(rgb_map) + (long)b * is * is;
    const float4* gmap = reinterpret_cast<const float4*>(g4) + (long)b * S * S;
    // maps are stored in image orientation: nr's row yi (counting upwards) is row is - 1 - yi
    auto at = [&](int xi, int yi) { return (long)(is - 1 - yi) * is + xi; };
    auto gat = [&](int xi, int yi) { return (long)((is - 1 - yi) >> 1) * S + (xi >> 1); };
    float grad_face[9];
#pragma unroll
    for (int k = 0; k < 9; k++) grad_face[k] = 0.f;
    // arithmetic as the original evaluates it: un-fused fp32, `* 2. / is` in double
    const float fis = (float)is;
    const double dis = (double)is;
#pragma unroll 1
    for (int edge_num = 0; edge_num < 3; edge_num++) {
        const int pi0 = edge_num, pi1 = (edge_num + 1) % 3, pi2 = (edge_num + 2) % 3;
        const float pp[3][2] = {{ndc_to_pix(face[3 * pi0], is), ndc_to_pix(face[3 * pi0 + 1], is)},
                                {ndc_to_pix(face[3 * pi1], is), ndc_to_pix(face[3 * pi1 + 1], is)},
                                {ndc_to_pix(face[3 * pi2], is), ndc_to_pix(face[3 * pi2 + 1], is)}};
#pragma unroll 1
        for (int axis = 0; axis < 2; axis++) {
            float p[3][2];
#pragma unroll
            for (int num = 0; num < 3; num++) { p[num][0] = pp[num][axis]; p[num][1] = pp[num][1 - axis]; }
            const int direction = axis == 0 ? (p[0][0] < p[1][0] ? -1 : 1) : (p[0][0] < p[1][0] ? 1 : -1);
            const int d0_from = (int)fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f);
            const int d0_to = (int)fminf(fmaxf(p[0][0], p[1][0]), fis - 1.f);
#pragma unroll 1
            for (int d0 = d0_from; d0 <= d0_to; d0++) {
                const float fd0 = (float)d0;
                const float d1_cross = add(mul(dvd(sub(p[1][1], p[0][1]), sub(p[1][0], p[0][0])), sub(fd0, p[0][0])), p[0][1]);
                if (!(fabsf(d1_cross) < 1.0e9f)) continue;          // zero-extent edge on an integer position: 0/0
                const int d1_in = 0 < direction ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
                const int d1_out = d1_in + direction;
                if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
                const int xout = axis == 0 ? d0 : d1_out, yout = axis == 0 ? d1_out : d0;
                const float4 rgb_in = cmap[at(xin, yin)], rgb_out = cmap[at(xout, yout)];
                auto pull = [&](float diff_grad, int d1) {
                    const float along = sub((float)d1, d1_cross), span = sub(p[1][0], p[0][0]);
                    if (p[1][0] != fd0) {
                        float dist = (float)((double)mul(dvd(span, sub(p[1][0], fd0)), along) * 2. / dis);
                        dist = 0.f < dist ? add(dist, eps) : sub(dist, eps);
                        grad_face[pi0 * 3 + (1 - axis)] = sub(grad_face[pi0 * 3 + (1 - axis)], dvd(diff_grad, dist));
                    }
                    if (p[0][0] != fd0) {
                        float dist = (float)((double)mul(dvd(span, sub(fd0, p[0][0])), along) * 2. / dis);
                        dist = 0.f < dist ? add(dist, eps) : sub(dist, eps);
                        grad_face[pi1 * 3 + (1 - axis)] = sub(grad_face[pi1 * 3 + (1 - axis)], dvd(diff_grad, dist));
                    }
                };
                // out
                if (fmap[at(xin, yin)] == fn) {
                    const int d1_limit = 0 < direction ? is - 1 : 0;
                    const int d1_from = max(min(d1_out, d1_limit), 0), d1_to = min(max(d1_out, d1_limit), is - 1);
#pragma unroll 1
                    for (int d1 = d1_from; d1 <= d1_to; d1++) {
                        const int x = axis == 0 ? d0 : d1, y = axis == 0 ? d1 : d0;
                        const float4 cc = cmap[at(x, y)], gg = gmap[gat(x, y)];
                        const float diff_grad = add(add(add(0.f, mul(sub(cc.x, rgb_in.x), gg.x)), mul(sub(cc.y, rgb_in.y), gg.y)),
                                                    mul(sub(cc.z, rgb_in.z), gg.z));
                        if (diff_grad <= 0.f) continue;
                        pull(diff_grad, d1);
                    }
                }
                // in
                {
                    float d0_cross2;
                    if (mul(sub(fd0, p[0][0]), sub(fd0, p[2][0])) < 0.f)
                        d0_cross2 = add(mul(dvd(sub(p[2][1], p[0][1]), sub(p[2][0], p[0][0])), sub(fd0, p[0][0])), p[0][1]);
                    else
                        d0_cross2 = add(mul(dvd(sub(p[1][1], p[2][1]), sub(p[1][0], p[2][0])), sub(fd0, p[2][0])), p[2][1]);
                    if (!(fabsf(d0_cross2) < 1.0e9f)) continue;
                    const int d1_limit = 0 < direction ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
                    const int d1_from = max(min(d1_in, d1_limit), 0), d1_to = min(max(d1_in, d1_limit), is - 1);
#pragma unroll 1
                    for (int d1 = d1_from; d1 <= d1_to; d1++) {
                        const int x = axis == 0 ? d0 : d1, y = axis == 0 ? d1 : d0;
                        if (fmap[at(x, y)] != fn) continue;
                        const float4 cc = cmap[at(x, y)], gg = gmap[gat(x, y)];
                        const float diff_grad = add(add(add(0.f, mul(sub(cc.x, rgb_out.x), gg.x)), mul(sub(cc.y, rgb_out.y), gg.y)),
                                                    mul(sub(cc.z, rgb_out.z), gg.z));
                        if (diff_grad <= 0.f) continue;
                        pull(diff_grad, d1);
                    }
                }
            }
        }
    }
    float4* vg = reinterpret_cast<float4*>(vgrad) + (long)b * S * S;
#pragma unroll
    for (int m = 0; m < 3; m++)
        if (grad_face[3 * m] != 0.f || grad_face[3 * m + 1] != 0.f)
            atomicAdd(&vg[vidx[m]], make_float4(grad_face[3 * m], grad_face[3 * m + 1], 0.f, 0.f));
}

template <int C>
__global__ void __launch_bounds__(PIX_THREADS)
k_resolve_rgb(const Cam cam, unsigned long long* __restrict__ zbuf, const float* __restrict__ verts3d,
              const float* __restrict__ im, long imstride, const Bg bg, float eps, int clampv,
              float* __restrict__ rgb, int* __restrict__ face_idx) {
    const int S = cam.S, is = 2 * S, b = blockIdx.y;
    const int pix = blockIdx.x * PIX_THREADS + threadIdx.x;
    if (pix >= S * S) return;
    const int i = pix / S, j = pix - i * S;
    unsigned long long* zb = zbuf + (long)b * is * is;
    ulonglong2* r0 = reinterpret_cast<ulonglong2*>(zb + (long)(2 * i) * is + 2 * j);
    ulonglong2* r1 = reinterpret_cast<ulonglong2*>(zb + (long)(2 * i + 1) * is + 2 * j);
    const ulonglong2 k0 = *r0, k1 = *r1;
    const unsigned long long empty = zkey_empty(cam.far);
    *r0 = make_ulonglong2(empty, empty);
    *r1 = make_ulonglong2(empty, empty);
    const unsigned long long keys[4] = {k0.x, k0.y, k1.x, k1.y};
    if (face_idx) {
        int* fo = face_idx + (long)b * is * is;
        *reinterpret_cast<int2*>(fo + (long)(2 * i) * is + 2 * j) = make_int2(zkey_face(k0.x), zkey_face(k0.y));
        *reinterpret_cast<int2*>(fo + (long)(2 * i + 1) * is + 2 * j) = make_int2(zkey_face(k1.x), zkey_face(k1.y));
    }
    const float* verts_b = verts3d + (long)b * S * S * 3;
    const float* im_b = im + (long)b * imstride;
    float sum[C];
#pragma unroll
    for (int c = 0; c < C; c++) sum[c] = 0.f;
#pragma unroll 1
    for (int sp = 0; sp < 4; sp++) {
        const int face = zkey_face(keys[sp]);
        float col[C];
        if (face < 0) {
#pragma unroll
            for (int c = 0; c < C; c++) col[c] = bg.c[c];
        } else {
            const int r = 2 * i + (sp >> 1), xi = 2 * j + (sp & 1);
            rgb_subpixel<C>(cam, verts_b, im_b, face, xi, is - 1 - r, zkey_depth(keys[sp]), eps, col);
        }
#pragma unroll
        for (int c = 0; c < C; c++) sum[c] += col[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        float v = sum[c] * 0.25f;
        if (clampv) v = fminf(fmaxf(v, -1.f), 1.f);
        rgb[((long)b * C + c) * S * S + pix] = v;
    }
}

// per-device facts the launchers need: SM count, and the > 48 KB dynamic shared-memory opt-in of k_splat_big (the
// attribute is per device: a process that uses a second GPU must opt in there as well)
struct DeviceInfo { bool ready; int sms; };
inline const DeviceInfo* device_info() {
    static std::mutex mu;
    static DeviceInfo info[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    DeviceInfo& d = info[dev];
    if (!d.ready) {
        if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.sms <= 0) return nullptr;
        if (cudaFuncSetAttribute(k_splat_big<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(k_splat_big<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)) != cudaSuccess)
            return nullptr;
        d.ready = true;
    }
    return &d;
}

// Layout of a z-buffer workspace laid out for `cap` views: keys [cap, 2S, 2S] | work list [cap * 4 (S-1)^2] | 8 counter
// words.  Every word is the EMPTY key at rest (g2s_zbuffer_init; every forward leaves it so), whatever `cap` it is later
// laid out for.
inline size_t ws_words(long cap, int S) { return (size_t)cap * (4ul * S * S + 4ul * (S - 1) * (S - 1)) + 8ul; }
inline WorkList ws_worklist(unsigned long long* ws, long cap, int S, float far_z) {
    WorkList wl;
    wl.items = ws + (size_t)cap * 4ul * S * S;
    wl.ctr = wl.items + (size_t)cap * 4ul * (S - 1) * (S - 1);
    wl.bias = zkey_empty(far_z);
    return wl;
}

// both stages of the forward rasteriser for `nv` views (view0 .. view0 + nv - 1) into workspace `ws` (capacity >= nv views)
template <bool FROM_VERTS>
inline int launch_splat(const Cam& c, const float* depth, long dstride, int vpi, const float* R, const float* t,
                        const float* verts3d, unsigned long long* ws, long cap, int nv, int view0, cudaStream_t st,
                        float4* proj_out = nullptr) {
    const DeviceInfo* di = device_info();
    if (!di) return G2S_ERR_LAUNCH;
    const int S = c.S, tiles = (S - 1 + TILE - 1) / TILE, tiles_y = (S - 1 + TILE_H - 1) / TILE_H;
    const WorkList wl = ws_worklist(ws, cap, S, c.far);
    { Launch l_(K_SPLAT, st);
      const dim3 grid(nv, tiles, tiles_y);
      const unsigned magic = vpi_magic(vpi, (long)view0 + nv);
      if (((2 * S) & (2 * S - 1)) == 0)     // power-of-two side: sub-pixel centres are exact products
          k_splat_tile<FROM_VERTS, true><<<grid, SPLAT_THREADS, 0, st>>>(c, depth, dstride, vpi, R, t, verts3d, ws, wl, magic, view0, proj_out);
      else
          k_splat_tile<FROM_VERTS, false><<<grid, SPLAT_THREADS, 0, st>>>(c, depth, dstride, vpi, R, t, verts3d, ws, wl, magic, view0, proj_out); }
    { Launch l_(K_SPLAT_BIG, st);
      k_splat_big<FROM_VERTS><<<di->sms * G2S_BIG_CTAS, BIG_THREADS, sizeof(BigSmem), st>>>(c, depth, dstride, vpi, R, t, verts3d, ws, wl,
                                                                                view0); }
    return G2S_OK;
}

// compares the shared-reciprocal division (g2s_math.cuh dvd_y) with __fdiv_rn bit for bit on pseudo-random operands
__global__ void k_selftest_division(unsigned long long per_thread, unsigned seed, unsigned long long* mismatches) {
    unsigned long long x = (unsigned long long)seed * 0x9E3779B97F4A7C15ull +
                           ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1) * 0xBF58476D1CE4E5B9ull;
    unsigned long long bad = 0;
    for (unsigned long long i = 0; i < per_thread; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
        // sign | exponent in [127-44, 127+44] | random mantissa, with a share of special mantissas
        unsigned ma = lo & 0x7fffffu, mb = hi & 0x7fffffu;
        const unsigned sel = (lo >> 23) & 7u;
        if (sel == 0) ma = 0; else if (sel == 1) ma = 0x7fffffu; else if (sel == 2) mb = 0x7fffffu; else if (sel == 3) mb = 0;
        const unsigned ea = 127u - 44u + ((hi >> 23) & 0xffu) % 89u, eb = 127u - 44u + ((lo >> 26) ^ (hi >> 25)) % 89u;
        const float a = __uint_as_float(((lo >> 31) << 31) | (ea << 23) | ma);
        const float b = __uint_as_float(((hi >> 31) << 31) | (eb << 23) | mb);
        const float q0 = __fdiv_rn(a, b), q1 = dvd_y(a, b, rcp_seed(b));
        bad += __float_as_uint(q0) != __float_as_uint(q1);
        // the unguarded core with a small numerator (face_record relies on it down to 2^-50): a * 2^-16 in [2^-60, 2^28]
        if (range_key(b) < RANGE_SPAN) {
            const float as = a * 1.52587890625e-05f;
            bad += __float_as_uint(__fdiv_rn(as, b)) != __float_as_uint(div_core(as, b, rcp_seed(b)));
        }
        const float z0 = __fdiv_rn(0.0f, b), z1 = dvd_y(0.0f, b, rcp_seed(b));
        bad += (z0 != z1);
    }
    if (bad) atomicAdd(mismatches, bad);
}

// compares the rasteriser's fast per-face / per-hit arithmetic (face_record + record_weights_depth: shared reciprocals,
// structural operand-range guards) with the plain IEEE formulation (tri_face_inv + tri_weights_depth) bit for bit on
// pseudo-random triangles whose coordinates and depths span ordinary AND extreme magnitudes (so that the guards and
// their IEEE fall-backs are exercised): mismatching (accept, w[3], zp) results are counted
__global__ void k_selftest_raster(unsigned long long per_thread, unsigned seed, int is, unsigned long long* mismatches) {
    unsigned long long x = (unsigned long long)seed * 0x9E3779B97F4A7C15ull +
                           ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1) * 0xD6E8FEB86659FD93ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (unsigned)(x >> 20); };
    auto unit = [&]() { return (float)(rnd() & 0xffffffu) * (1.0f / 16777216.0f); };   // [0,1)
    unsigned long long bad = 0;
    for (unsigned long long it = 0; it < per_thread; it++) {
        // a small triangle somewhere on (or a little off) the screen, in NDC; every 8th one degenerate / extreme
        const unsigned kind = rnd() & 7u;
        const float cx = unit() * 2.2f - 1.1f, cy = unit() * 2.2f - 1.1f;
        const float ext = kind == 1 ? 1e-6f : (kind == 2 ? 0.5f : 6.0f / (float)is);
        float vx[3], vy[3], vz[3];
        for (int k = 0; k < 3; k++) {
            vx[k] = cx + (unit() - 0.5f) * ext;
            vy[k] = cy + (unit() - 0.5f) * ext;
            vz[k] = 0.8f + 0.4f * unit();
        }
        if (kind == 3) { vx[2] = vx[1]; vy[2] = vy[1]; }                       // zero area
        if (kind == 4) vz[rnd() % 3] = __uint_as_float((rnd() % 254u + 1u) << 23);   // any power of two as a depth
        if (kind == 5) { const float sc = __uint_as_float((127u - 60u + rnd() % 120u) << 23); for (int k = 0; k < 3; k++) vz[k] *= sc; }
        if (kind == 6) vz[rnd() % 3] *= -1.0f;
        const float a0[3] = {vx[0], vy[0], vz[0]}, a1[3] = {vx[1], vy[1], vz[1]}, a2[3] = {vx[2], vy[2], vz[2]};
        const Tri f = make_tri(a0, a1, a2);
        float rec[16], fi[9];
        face_record(f, is, rec);
        tri_face_inv(f, is, fi);
        for (int k = 0; k < 9; k++) bad += __float_as_uint(rec[k]) != __float_as_uint(fi[k]);
        // sub-pixels around the triangle (inside and outside: the clamps and the all-zero case)
        const int px = (int)floorf(ndc_to_pix(cx, is)), py = (int)floorf(ndc_to_pix(cy, is));
        for (int s2 = 0; s2 < 4; s2++) {
            const int xi = px + (int)(rnd() % 7u) - 3, yi = py + (int)(rnd() % 7u) - 3;
            float w0[3] = {0.f, 0.f, 0.f}, w1[3] = {0.f, 0.f, 0.f}, z0 = 0.f, z1 = 0.f;
            const float nr = kind == 5 ? 0.0f : 0.1f, fr = kind == 5 ? 3.0e38f : 100.0f;
            const bool ok0 = record_weights_depth(rec, xi, yi, nr, fr, w0, &z0);
            const bool ok1 = tri_weights_depth(f, fi, xi, yi, nr, fr, w1, &z1);
            bool same = ok0 == ok1;
            if (ok0 && ok1) {
                same = __float_as_uint(z0) == __float_as_uint(z1);
                for (int k = 0; k < 3; k++) same = same && __float_as_uint(w0[k]) == __float_as_uint(w1[k]);
            }
            bad += same ? 0 : 1;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// div_side (v / S, v % S through a float estimate + fix-up) for EVERY vertex index of an S x S mesh, and view_image (view / vpi as a
// multiply-high by the host's constant) for EVERY view of a call, against plain integer division
__global__ void k_selftest_index_math(int S, int vpi, long n_views, unsigned magic, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long v = t0; v < (long)S * S; v += stride) {
        int q, r;
        div_side((int)v, S, &q, &r);
        bad += (q != (int)(v / S)) + (r != (int)(v % S));
    }
    for (long b = t0; b < n_views; b += stride) bad += view_image((int)b, vpi, magic) != (int)(b / vpi);
    if (bad) atomicAdd(mismatches, bad);
}

// face_vertices' device path (reciprocal estimate + one fix-up) against plain integer arithmetic, for EVERY face index of
// an S x S grid mesh including the fill_back copies
__global__ void k_selftest_face_vertices(int S, unsigned long long* mismatches) {
    const long Q = (long)(S - 1) * (S - 1), n = 4 * Q;
    unsigned long long bad = 0;
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (long)gridDim.x * blockDim.x) {
        int v[3];
        face_vertices((int)f, S, v);
        long g = f;
        const bool rev = g >= 2 * Q;
        if (rev) g -= 2 * Q;
        const bool second = g >= Q;
        if (second) g -= Q;
        const long qy = g / (S - 1), qx = g % (S - 1), v00 = qy * S + qx;
        long a = second ? v00 + 1 : v00, b = v00 + S, c = second ? v00 + S + 1 : v00 + 1;
        if (rev) { const long t = a; a = c; c = t; }
        bad += (v[0] != a) + (v[1] != b) + (v[2] != c);
    }
    if (bad) atomicAdd(mismatches, bad);
}

inline dim3 pix_grid2(int S, int views, int bx = PBX, int by = PBY) { return dim3((S + bx - 1) / bx, (S + by - 1) / by, views); }

// Views per forward chunk: 256 MB of z-buffer keys (512 views at 128^2, 128 at 256^2), at least 8.  Round 1 kept a chunk's keys
// inside the 126 MB L2 (24 MB per chunk) for k_resolve; with the round-2 kernels the sweeps say otherwise (profiles/r02_notes.md:
// 48 -> 256 -> 512 views per chunk at 128^2: 14.5 -> 14.1 ms mid-round, 11.81 -> 11.75 ms with the final kernels): fewer, longer
// launches win, the keys stream through L2 either way.
inline int chunk_views_for(int S, int cap) {
    long v = (256L << 20) / (32L * S * S);
    if (v < 8) v = 8;
    if (v > cap) v = cap;
    return (int)v;
}

inline dim3 pix_grid(long npix, int batch) { return dim3((unsigned)((npix + PIX_THREADS - 1) / PIX_THREADS), batch); }

inline bool bad_size(int S) { return S < 2 || S > 2048; }

// the masked-quarter-gradient plane of a raster workspace laid out for nv views (see launch_raster_bwd)
inline float* raster_ws_gsub(float* raster_ws, int nv, int S) { return raster_ws + (size_t)nv * 8 * S * S; }

// Backward of the rasteriser for `nv` views.  raster_ws [nv, 9, S, S] floats: projected vertices (uvz-, 16-byte texels) |
// vertex gradients (uvz-) | masked quarter gradient g_sub [nv, S, S].  Mesh from (depth, R, t) or, when verts3d is given, from
// 3-D points (gradient -> grad_verts [nv, S*S, 3], written).
inline void launch_raster_project(const Cam& c, const float* depth, long dstride, int vpi, const float* R, const float* t,
                                  const float* verts3d, float* raster_ws, int nv, int view0, cudaStream_t st) {
    const int S = c.S;
    float* proj = raster_ws;
    float* vgrad = proj + (size_t)nv * 4 * S * S;
    Launch l_(K_PROJECT, st);     // also zeroes vgrad
    if (verts3d) k_project_points<true><<<pix_grid((long)S * S, nv), PIX_THREADS, 0, st>>>(c, verts3d, proj, vgrad);
    else k_project_verts<true><<<pix_grid((long)S * S, nv), PIX_THREADS, 0, st>>>(c, depth, dstride, vpi, R, t, view0, proj, vgrad);
}

inline void launch_raster_gather(const Cam& c, const float* depth, long dstride, int vpi, const float* R, const float* t,
                                 const float* verts3d, const int* face_idx, float* raster_ws, int nv, int view0,
                                 float* grad_depth, long gdstride, float* grad_verts, float* grad_R, float* grad_t,
                                 cudaStream_t st, const float* proj_ext = nullptr, int layout_views = 0) {
    const int S = c.S;
    // with the forward's projected vertices the vertex gradients take the front of the workspace (zero at rest, k_vertex_bwd
    // leaves it so), otherwise they follow our own projection of these nv views (zeroed by the projection kernel)
    const float* proj = proj_ext ? proj_ext : raster_ws;
    float* vgrad = proj_ext ? raster_ws : raster_ws + (size_t)nv * 4 * S * S;
    // the masked quarter gradient sits behind 8 planes of `layout_views` views: the views of this launch, or -- with the
    // zero-at-rest front -- the workspace's capacity, so that a short last chunk cannot write into that front
    float* g_sub = raster_ws_gsub(raster_ws, layout_views > 0 ? layout_views : nv, S);
    { Launch l_(K_RASTER_BWD, st);
      k_raster_bwd_px<<<pix_grid2(S, nv, RBX, RBY), RB_THREADS, 0, st>>>(c, face_idx, g_sub, proj, vgrad, view0); }
    { Launch l_(K_VERTEX_BWD, st);
      if (verts3d) k_points_bwd<<<pix_grid((long)S * S, nv), PIX_THREADS, 0, st>>>(c, verts3d, vgrad, grad_verts);
      else if (proj_ext) k_vertex_bwd<true><<<pix_grid(((long)S * S + VB_VPT - 1) / VB_VPT, nv), PIX_THREADS, 0, st>>>(c, depth, dstride, vpi, R, t, view0, vgrad,
                                                                                            grad_depth, gdstride, grad_R, grad_t);
      else k_vertex_bwd<false><<<pix_grid(((long)S * S + VB_VPT - 1) / VB_VPT, nv), PIX_THREADS, 0, st>>>(c, depth, dstride, vpi, R, t, view0, vgrad, grad_depth,
                                                                        gdstride, grad_R, grad_t); }
}

inline void launch_raster_bwd(const Cam& c, const float* depth, long dstride, int vpi, const float* R, const float* t,
                              const float* verts3d, const int* face_idx, float* raster_ws, int nv, int view0, float* grad_depth,
                              long gdstride, float* grad_verts, float* grad_R, float* grad_t, cudaStream_t st) {
    launch_raster_project(c, depth, dstride, vpi, R, t, verts3d, raster_ws, nv, view0, st);
    launch_raster_gather(c, depth, dstride, vpi, R, t, verts3d, face_idx, raster_ws, nv, view0, grad_depth, gdstride, grad_verts,
                         grad_R, grad_t, st);
}

}  // namespace

// caller-owned context (include/g2s_b200.h): the streams / events of the multi-lane forward and the tuning read from the
// environment at creation
struct g2s_context {
    int device;
    int lanes;                       // forward lanes the host asked for through the environment (0 = library default)
    bool no_pipeline;
    cudaStream_t aux[MAX_LANES];     // aux[0] unused: lane 0 is the caller's stream
    cudaEvent_t ev_fork, ev_join[MAX_LANES];
};

// =================================================================================================
extern "C" {

int g2s_context_create(g2s_context** out) {
    if (!out) return G2S_ERR_NULL;
    *out = nullptr;
    g2s_context* c = new (std::nothrow) g2s_context();
    if (!c) return G2S_ERR_LAUNCH;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return G2S_ERR_LAUNCH; }
    const char* e = getenv("G2S_FWD_LANES");
    c->lanes = e ? atoi(e) : 0;
    c->no_pipeline = getenv("G2S_NO_PIPELINE") != nullptr;
    bool ok = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 1; k < MAX_LANES && ok; k++)
        ok = cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { g2s_context_destroy(c); return G2S_ERR_LAUNCH; }
    *out = c;
    return G2S_OK;
}

void g2s_context_destroy(g2s_context* c) {
    if (!c) return;
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int k = 1; k < MAX_LANES; k++) {
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
        if (c->aux[k]) cudaStreamDestroy(c->aux[k]);
    }
    delete c;
}

size_t g2s_workspace_bytes(int kind, int n, int image_size) {
    if (n <= 0 || bad_size(image_size)) return 0;
    const size_t S2 = (size_t)image_size * image_size, f = sizeof(float);
    switch (kind) {
        case G2S_WS_ZBUFFER: return g2s_zbuffer_bytes(n, image_size);
        case G2S_WS_RASTER_BWD: return (size_t)n * 9 * S2 * f;
        case G2S_WS_TEX_BWD: return (size_t)n * 4 * S2 * f;
        case G2S_WS_TEXELS: return (size_t)n * 8 * S2 * f;
        case G2S_WS_GRAD_NORMAL: return (size_t)n * 3 * S2 * f;
        case G2S_WS_RGB_MAP: return (size_t)n * 20 * S2 * f;      // colour map [n,2S,2S,4] + quarter gradient [n,S,S,4]
        case G2S_WS_LOSS: { const dim3 g = pix_grid2(image_size, 1); return ((size_t)n * g.x * g.y + LOSS_STAGE_BLOCKS) * 2 * sizeof(double); }
        default: return 0;
    }
}

int g2s_version(void) { return 202; }

const char* g2s_error_string(int code) {
    switch (code) {
        case G2S_OK: return "ok";
        case G2S_ERR_NULL: return "required pointer is NULL";
        case G2S_ERR_SHAPE: return "size out of the supported range";
        case G2S_ERR_LAUNCH: return "CUDA launch failure";
        case G2S_ERR_UNSUPPORTED: return "unsupported mode";
        default: return "unknown error";
    }
}

size_t g2s_zbuffer_bytes(int n_views, int image_size) {
    if (n_views <= 0 || bad_size(image_size)) return 0;
    // keys + work list + counters, with room for the counter words of MAX_LANES separately laid out parts
    return (ws_words(n_views, image_size) + 8ul * MAX_LANES) * sizeof(unsigned long long);
}

int g2s_zbuffer_init(void* zbuf, int n_views, int image_size, float far_z, void* stream) {
    if (!zbuf) return G2S_ERR_NULL;
    if (n_views <= 0 || bad_size(image_size)) return G2S_ERR_SHAPE;
    const long n = (long)(g2s_zbuffer_bytes(n_views, image_size) / sizeof(unsigned long long));
    const int blocks = (int)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
    { Launch l_(K_ZINIT, (cudaStream_t)stream); k_zbuf_init<<<blocks, 256, 0, (cudaStream_t)stream>>>((unsigned long long*)zbuf, n, zkey_empty(far_z)); }
    return launch_status();
}

int g2s_warp_depth_fwd(const g2s_camera* cam, const float* depth, long depth_view_stride, const float* R,
                       const float* t, int n_views, void* zbuf, float* recon_depth, int32_t* face_idx, void* stream) {
    if (!cam || !depth || !R || !t || !zbuf || !recon_depth) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size)) return G2S_ERR_SHAPE;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_splat<false>(c, depth, depth_view_stride, 1, R, t, nullptr, (unsigned long long*)zbuf, n_views, n_views, 0, st))
        return rc;
    FusedArgs fa = {};
    { Launch l_(K_RESOLVE, st); k_resolve<false><<<pix_grid2(S, n_views), dim3(PBX, PBY), 0, st>>>(c, (unsigned long long*)zbuf, recon_depth,
                                                                             face_idx, fa); }
    return launch_status();
}

int g2s_warp_depth_bwd(const g2s_camera* cam, const float* depth, long depth_view_stride, const float* R,
                       const float* t, int n_views, const int32_t* face_idx, const float* recon_depth,
                       const float* grad_recon_depth, float* grad_sub_ws, float* grad_depth,
                       long grad_depth_view_stride, float* grad_R, float* grad_t, void* stream) {
    if (!cam || !depth || !R || !t || !face_idx || !recon_depth || !grad_recon_depth || !grad_sub_ws || !grad_depth)
        return G2S_ERR_NULL;
    if ((grad_R == nullptr) != (grad_t == nullptr)) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size)) return G2S_ERR_SHAPE;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    const long n = (long)n_views * S * S;
    { Launch l_(K_CLAMP_GRAD, st); k_clamp_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(recon_depth, grad_recon_depth, c.clamp_lo, c.clamp_hi, n,
                                                              raster_ws_gsub(grad_sub_ws, n_views, S)); }
    launch_raster_bwd(c, depth, depth_view_stride, 1, R, t, nullptr, face_idx, grad_sub_ws, n_views, 0, grad_depth,
                      grad_depth_view_stride, nullptr, grad_R, grad_t, st);
    return launch_status();
}

int g2s_warp_grid_fwd(const g2s_camera* cam, const float* depth, long depth_view_stride, const float* R,
                      const float* t, int B, int H, int W, int inverse, float* grid, void* stream) {
    if (!cam || !depth || !R || !t || !grid) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    { Launch l_(K_GRID_FWD, (cudaStream_t)stream); k_warp_grid_fwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
        make_cam(cam), depth, depth_view_stride, R, t, H, W, inverse, grid); }
    return launch_status();
}

int g2s_warp_grid_bwd(const g2s_camera* cam, const float* depth, long depth_view_stride, const float* R,
                      const float* t, int B, int H, int W, int inverse, const float* grad_grid, float* grad_depth,
                      float* grad_R, float* grad_t, void* stream) {
    if (!cam || !depth || !R || !t || !grad_grid || !grad_depth) return G2S_ERR_NULL;
    if ((grad_R == nullptr) != (grad_t == nullptr)) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    { Launch l_(K_GRID_BWD, (cudaStream_t)stream); k_warp_grid_bwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
        make_cam(cam), depth, depth_view_stride, R, t, H, W, inverse, grad_grid, grad_depth, grad_R, grad_t); }
    return launch_status();
}

int g2s_normal_fwd(const g2s_camera* cam, const float* depth, int B, int H, int W, float* normal, void* stream) {
    if (!cam || !depth || !normal) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 3 || W < 3) return G2S_ERR_SHAPE;
    { Launch l_(K_NORMAL_FWD, (cudaStream_t)stream); k_normal_fwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(make_cam(cam), depth, H, W, normal, 3); }
    return launch_status();
}

int g2s_normal_bwd(const g2s_camera* cam, const float* depth, int B, int H, int W, const float* grad_normal,
                   float* grad_depth, int accumulate, void* stream) {
    if (!cam || !depth || !grad_normal || !grad_depth) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 3 || W < 3) return G2S_ERR_SHAPE;
    { Launch l_(K_NORMAL_BWD, (cudaStream_t)stream); k_normal_bwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(make_cam(cam), depth, H, W,
                                                                                     grad_normal, grad_depth, accumulate); }
    return launch_status();
}

int g2s_sample_fwd(const float* input, long input_batch_stride, const float* grid, int B, int C, int H, int W, int Ho,
                   int Wo, int mode, int align_corners, float* out, void* stream) {
    if (!input || !grid || !out) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || C <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return G2S_ERR_SHAPE;
    if (mode != 0 && mode != 1) return G2S_ERR_UNSUPPORTED;
    { Launch l_(K_SAMPLE_FWD, (cudaStream_t)stream); k_sample_fwd<<<pix_grid((long)Ho * Wo, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
        input, input_batch_stride, grid, C, H, W, Ho, Wo, mode, align_corners, out); }
    return launch_status();
}

int g2s_sample_bwd(const float* input, long input_batch_stride, const float* grid, const float* grad_out, int B, int C,
                   int H, int W, int Ho, int Wo, int mode, int align_corners, float* grad_input,
                   long grad_input_batch_stride, float* grad_grid, void* stream) {
    if (!input || !grid || !grad_out) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || C <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return G2S_ERR_SHAPE;
    if (mode != 0 && mode != 1) return G2S_ERR_UNSUPPORTED;
    { Launch l_(K_SAMPLE_BWD, (cudaStream_t)stream); k_sample_bwd<<<pix_grid((long)Ho * Wo, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
        input, input_batch_stride, grid, grad_out, C, H, W, Ho, Wo, mode, align_corners, grad_input,
        grad_input_batch_stride, grad_grid); }
    return launch_status();
}

// backward scratch does not need L2 residency as much as it needs long launches (measured: fewer, larger launches
// win -- 1024 -> 2048 views per chunk at 128^2: 11.91 -> 11.82 ms per step, face config 12.75 -> 12.55): cap the chunk by
// scratch memory only (~2 GB)
int g2s_chunk_views_bwd(int image_size) {
    if (bad_size(image_size)) return 0;
    static const long env128 = [] { const char* e = getenv("G2S_CHUNK_VIEWS_BWD_128"); return e ? atol(e) : 0L; }();
    if (env128 > 0) {
        const long v = env128 * 128L * 128L / ((long)image_size * image_size);
        if (v >= 1) return (int)v;
    }
    const long per_view = 13L * image_size * image_size * 4;   // 9 S^2 (raster scratch) + 4 S^2 (packed texture gradient)
    long v = (2L << 30) / per_view;
    return (int)(v < 1 ? 1 : v);
}

int g2s_chunk_views(int image_size) {
    if (bad_size(image_size)) return 0;
    // tuning override (views per chunk at 128^2; scaled by (128/S)^2): G2S_CHUNK_VIEWS_128, read once
    static const long env128 = [] { const char* e = getenv("G2S_CHUNK_VIEWS_128"); return e ? atol(e) : 0L; }();
    if (env128 > 0) {
        const long v = env128 * 128L * 128L / ((long)image_size * image_size);
        if (v >= 1) return (int)v;
    }
    return chunk_views_for(image_size, 1 << 20);
}

namespace {
// defined in g2s_callers.cuh (same translation unit, included at the end)
__global__ void __launch_bounds__(256) k_photo_finish(const double* __restrict__ parts, int nparts, int C, float* __restrict__ out);

// first stage of the fixed-order sum of the per-CTA (numerator, mask count) pairs: block k sums pairs [k * per, (k + 1) * per)
__global__ void __launch_bounds__(256) k_loss_stage(const double2* __restrict__ parts, long nparts, long per, double2* __restrict__ out) {
    __shared__ double sh[2][8];
    const long lo = (long)blockIdx.x * per, hi = lo + per < nparts ? lo + per : nparts;
    double n = 0.0, d = 0.0;
    for (long p = lo + threadIdx.x; p < hi; p += 256) { const double2 v = parts[p]; n += v.x; d += v.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = n; sh[1][threadIdx.x >> 5] = d; }
    __syncthreads();
    if (threadIdx.x == 0) {
        n = d = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) { n += sh[0][w]; d += sh[1][w]; }
        out[blockIdx.x] = make_double2(n, d);
    }
}
}

static int fused_fwd_impl(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R, const float* t,
                          const float* light, int n_images, int views_per_image, int align_corners, void* zbuf,
                          int ws_views, float* normal_ws, float* recon_im, float* recon_depth, int32_t* face_idx,
                          const float* mask_in, float* mask_out, const g2s_photo_loss* loss, void* loss_ws, float* out3,
                          float* proj_ws, void* stream) {
    if (!cam || !depth || !albedo || !R || !t || !light || !zbuf || !normal_ws || !recon_im || !recon_depth)
        return G2S_ERR_NULL;
    if (loss && (!loss->target || !loss_ws || !out3)) return G2S_ERR_NULL;
    const long n_views = (long)n_images * views_per_image;
    if (n_images <= 0 || views_per_image <= 0 || n_views > (1L << 30) || ws_views <= 0 || bad_size(cam->image_size))
        return G2S_ERR_SHAPE;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n_images; i0 += 32768) {
        const int ni = n_images - i0 < 32768 ? n_images - i0 : 32768;
        Launch l_(K_NORMAL_FWD, st);
        k_normal_fwd<<<pix_grid((long)S * S, ni), PIX_THREADS, 0, st>>>(c, depth + (long)i0 * S * S, S, S,
                                                                        normal_ws + (long)i0 * S * S * TEXEL, TEXEL,
                                                                        albedo + (long)i0 * 3 * S * S);
    }
    // Chunked to bound the z-buffer workspace.  When the workspace holds two recommended chunks the chunks alternate between
    // its halves on two streams (the caller's and one of the context's, forked and joined with events): one chunk's
    // rasteriser (issue-bound) runs beside the other's resolve (bandwidth-bound) and fills the SMs its tail leaves idle.
    // Not while per-kernel timing is on.
    const int rec = g2s_chunk_views(S);
    int nl = (int)(ws_views / rec);                    // lanes the workspace has room for
    if (nl > MAX_LANES) nl = MAX_LANES;
    if (ctx && ctx->lanes >= 1 && ctx->lanes < nl) nl = ctx->lanes;
    if (!ctx || ctx->no_pipeline || nl < 1 || n_views <= rec || g_prof_on) nl = 1;
    if (nl > 1) {      // the context's streams belong to one device
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess || dev != ctx->device) nl = 1;
    }
    const int per = nl > 1 ? rec : (ws_views >= rec ? rec : ws_views);   // views per chunk
    const int chunk = per < 32768 ? per : 32768;
    cudaStream_t lanes[MAX_LANES] = {st, st, st, st};
    if (nl > 1) {
        cudaEventRecord(ctx->ev_fork, st);
        for (int k = 1; k < nl; k++) {
            lanes[k] = ctx->aux[k];
            cudaStreamWaitEvent(lanes[k], ctx->ev_fork, 0);
        }
    }
    // equal launches: 128 views in chunks of 12 would end in a launch of 8 (the tail of the face config on 8 GPUs)
    const long nch = (n_views + chunk - 1) / chunk;
    const int bal = (int)((n_views + nch - 1) / nch);
    int lane = 0, rc_lane = 0;
    for (long v0 = 0; v0 < n_views; v0 += bal, lane = (lane + 1) % nl) {
        const int nv = (int)(n_views - v0 < bal ? n_views - v0 : bal);
        cudaStream_t ls = lanes[lane];
        unsigned long long* zb = (unsigned long long*)zbuf + (size_t)lane * ws_words(chunk, S);   // this lane's part
        if (int rc = launch_splat<false>(c, depth, (long)S * S, views_per_image, R, t, nullptr, zb, chunk, nv, (int)v0, ls,
                                          reinterpret_cast<float4*>(proj_ws))) {
            rc_lane = rc;      // the lanes are still joined back below
            break;
        }
        FusedArgs fa = {R, t, light, normal_ws, albedo, recon_im, views_per_image, align_corners, (int)v0, mask_in, mask_out};
        fa.vpi_magic = vpi_magic(views_per_image, n_views);
        if (loss) {
            fa.target = loss->target; fa.vmask = loss->view_mask; fa.thresh = loss->depth_thresh;
            fa.loss_parts = (double*)loss_ws;
        }
        { Launch l_(K_RESOLVE_FUSED, ls);
          if (loss) k_resolve<true, true><<<pix_grid2(S, nv), dim3(PBX, PBY), 0, ls>>>(c, zb, recon_depth, face_idx, fa);
          else k_resolve<true><<<pix_grid2(S, nv), dim3(PBX, PBY), 0, ls>>>(c, zb, recon_depth, face_idx, fa); }
    }
    if (nl > 1) {
        for (int k = 1; k < nl; k++) {
            cudaEventRecord(ctx->ev_join[k], lanes[k]);
            cudaStreamWaitEvent(st, ctx->ev_join[k], 0);
        }
    }
    if (loss && !rc_lane) {      // every CTA of every view wrote its pair: sum them in a fixed order
        const dim3 g = pix_grid2(S, 1);
        const long nparts = n_views * g.x * g.y;
        double2* stage = (double2*)loss_ws + nparts;       // LOSS_STAGE_BLOCKS pairs behind the per-CTA pairs
        const long per = (nparts + LOSS_STAGE_BLOCKS - 1) / LOSS_STAGE_BLOCKS;
        const int nb = (int)((nparts + per - 1) / per);
        Launch l_(K_PHOTOMETRIC, st);
        k_loss_stage<<<nb, 256, 0, st>>>((const double2*)loss_ws, nparts, per, stage);
        k_photo_finish<<<1, 256, 0, st>>>((const double*)stage, nb, 3, out3);
    }
    return rc_lane ? rc_lane : launch_status();
}

int g2s_render_fused_fwd(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R, const float* t,
                         const float* light, int n_images, int views_per_image, int align_corners, void* zbuf,
                         int ws_views, float* normal_ws, float* recon_im, float* recon_depth, int32_t* face_idx,
                         const float* mask_in, float* mask_out, float* proj_ws, void* stream) {
    return fused_fwd_impl(ctx, cam, depth, albedo, R, t, light, n_images, views_per_image, align_corners, zbuf, ws_views, normal_ws,
                          recon_im, recon_depth, face_idx, mask_in, mask_out, nullptr, nullptr, nullptr, proj_ws, stream);
}

int g2s_render_fused_loss_fwd(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R,
                              const float* t, const float* light, int n_images, int views_per_image, int align_corners,
                              void* zbuf, int ws_views, float* normal_ws, float* recon_im, float* recon_depth,
                              int32_t* face_idx, const g2s_photo_loss* loss, void* loss_ws, float* out3, float* proj_ws,
                              void* stream) {
    if (!loss) return G2S_ERR_NULL;
    return fused_fwd_impl(ctx, cam, depth, albedo, R, t, light, n_images, views_per_image, align_corners, zbuf, ws_views, normal_ws,
                          recon_im, recon_depth, face_idx, nullptr, nullptr, loss, loss_ws, out3, proj_ws, stream);
}

namespace {
struct ZeroList { float* p[6]; size_t n[6]; };
// zero up to six float arrays in one launch
__global__ void __launch_bounds__(256) k_zero_list(const ZeroList zl) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float* __restrict__ p = zl.p[k];
        const size_t n = zl.n[k];
        for (size_t i = t0; i < n; i += stride) p[i] = 0.f;
    }
}
}

static int fused_bwd_impl(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R, const float* t,
                          const float* light, int n_images, int views_per_image, int align_corners,
                          const float* normal_ws, const float* recon_depth, const int32_t* face_idx,
                          const float* grad_recon_im, const float* grad_recon_depth, const g2s_photo_loss* loss,
                          const float* sums3, const float* grad_loss, const float* proj_ws, int ws_views, float* grad_sub_ws,
                          float* grad_tex_ws, float* grad_normal_ws, float* grad_depth, float* grad_albedo, float* grad_R,
                          float* grad_t, float* grad_light, void* stream) {
    if (!cam || !depth || !albedo || !R || !t || !light || !normal_ws || !recon_depth || !face_idx ||
        !grad_sub_ws || !grad_tex_ws || !grad_normal_ws || !grad_depth || !grad_albedo || !grad_R || !grad_t ||
        !grad_light)
        return G2S_ERR_NULL;
    if (!grad_recon_im && !loss) return G2S_ERR_NULL;
    if (loss && (!loss->target || !sums3 || !grad_loss)) return G2S_ERR_NULL;
    const long n_views = (long)n_images * views_per_image;
    if (n_images <= 0 || views_per_image <= 0 || n_views > (1L << 30) || ws_views <= 0 || bad_size(cam->image_size))
        return G2S_ERR_SHAPE;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t img_f = (size_t)S * S;
    {   // the six accumulators start at zero: one launch instead of six memset nodes (the single-image step is launch-bound)
        ZeroList zl;
        float* ptrs[6] = {grad_R, grad_t, grad_light, grad_depth, grad_albedo, grad_normal_ws};
        const size_t cnt[6] = {(size_t)n_views * 9, (size_t)n_views * 3, (size_t)n_views * 5, (size_t)n_images * img_f,
                               (size_t)n_images * 3 * img_f, (size_t)n_images * 3 * img_f};
        size_t total = 0;
        for (int k = 0; k < 6; k++) { zl.p[k] = ptrs[k]; zl.n[k] = cnt[k]; total += cnt[k]; }
        const long blocks = (long)((total + 4 * 256 - 1) / (4 * 256));
        k_zero_list<<<(unsigned)(blocks < 148 * 16 ? (blocks < 1 ? 1 : blocks) : 148 * 16), 256, 0, st>>>(zl);
    }
    const int chunk = ws_views < 32768 ? ws_views : 32768;
    // the per-view texture-gradient scratch is zeroed once; k_render_bwd_tex leaves what it consumed at zero
    cudaMemsetAsync(grad_tex_ws, 0, sizeof(float) * (size_t)(n_views < chunk ? n_views : chunk) * 4 * img_f, st);
    // Two lanes inside a chunk: the bandwidth-bound kernels (vertex projection, texture-gradient reduction) run on the
    // context's side stream under the issue-bound ones (pixel stage, raster gather) of the caller's stream:
    //   caller:  k_render_bwd_pixel ............ k_raster_bwd_px  k_vertex_bwd
    //   side:    k_project_verts     (wait pixel) k_render_bwd_tex
    // Fork / join with the context's events (capturable in a CUDA graph); without a context everything is serial.
    // (not for small launches: the four cross-stream hand-offs cost more than the overlap gains below ~2 M pixels per chunk --
    // the single-image step went from 0.26 to 0.49 ms with them)
    bool two = ctx != nullptr && !ctx->no_pipeline && !g_prof_on &&      // per-kernel timing wants the kernels alone
               (n_views < ws_views ? n_views : ws_views) * (long)cam->image_size * cam->image_size >= (1L << 21);
    if (two) {      // the context's streams belong to one device
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess || dev != ctx->device) two = false;
    }
    cudaStream_t sd = two ? ctx->aux[1] : st;
    // equal launches (4096 views in chunks of 2520 would be 2520 + 1576)
    const long nch = (n_views + chunk - 1) / chunk;
    const int bal = (int)((n_views + nch - 1) / nch);
    for (long v0 = 0; v0 < n_views; v0 += bal) {
        const int nv = (int)(n_views - v0 < bal ? n_views - v0 : bal);
        // workspace layout: per launch [nv x 4 | nv x 4 | nv x 1]; with the forward's projection the FRONT (vertex gradients, zero at
        // rest) is laid out for the workspace's capacity, whatever this chunk's size
        const int lay = proj_ws ? chunk : nv;
        FusedArgs fa = {R, t, light, normal_ws, albedo, nullptr, views_per_image, align_corners, (int)v0, nullptr, nullptr};
        if (loss) {
            fa.target = loss->target; fa.vmask = loss->view_mask; fa.thresh = loss->depth_thresh;
            fa.sums3 = sums3; fa.gloss = grad_loss;
        }
        if (two) {
            cudaEventRecord(ctx->ev_fork, st);             // after the memsets / the previous chunk's gather
            cudaStreamWaitEvent(sd, ctx->ev_fork, 0);
            if (!proj_ws) launch_raster_project(c, depth, (long)S * S, views_per_image, R, t, nullptr, grad_sub_ws, nv, (int)v0, sd);
            cudaEventRecord(ctx->ev_join[1], sd);
        }
        { Launch l_(K_BWD_PIXEL, st);
          float* gsub = raster_ws_gsub(grad_sub_ws, lay, S);
          if (loss) k_render_bwd_pixel<true><<<pix_grid2(S, nv, BPX, BPY), dim3(BPX, BPY), 0, st>>>(c, fa, recon_depth, grad_recon_im, grad_recon_depth,
                                                                         gsub, grad_tex_ws, grad_R, grad_t);
          else k_render_bwd_pixel<false><<<pix_grid2(S, nv, BPX, BPY), dim3(BPX, BPY), 0, st>>>(c, fa, recon_depth, grad_recon_im, grad_recon_depth,
                                                                         gsub, grad_tex_ws, grad_R, grad_t); }
        if (two) {
            cudaEventRecord(ctx->ev_join[2], st);
            cudaStreamWaitEvent(sd, ctx->ev_join[2], 0);
        }
        const int img_lo = (int)(v0 / views_per_image), img_hi = (int)((v0 + nv - 1) / views_per_image);
        dim3 tex_grid = pix_grid((long)S * S, img_hi - img_lo + 1);
        {   // enough CTAs to fill the GPU: split an image's views over up to 16 groups when there are few (image, pixel) blocks
            const long ctas = (long)tex_grid.x * tex_grid.y;
            long groups = ctas >= 2048 ? 1 : (2048 + ctas - 1) / ctas;
            const long vmax = (views_per_image < nv ? views_per_image : nv) / 8;
            if (groups > vmax) groups = vmax;
            if (groups > 16) groups = 16;
            tex_grid.z = (unsigned)(groups < 1 ? 1 : groups);
        }
        { Launch l_(K_BWD_TEX, sd);
          k_render_bwd_tex<<<tex_grid, PIX_THREADS, 0, sd>>>(S, fa, nv, grad_tex_ws, grad_albedo,
                                                                                             grad_normal_ws, grad_light); }
        if (two) {
            cudaEventRecord(ctx->ev_join[3], sd);
            cudaStreamWaitEvent(st, ctx->ev_join[1], 0);   // the projected vertices
        } else if (!proj_ws) {
            launch_raster_project(c, depth, (long)S * S, views_per_image, R, t, nullptr, grad_sub_ws, nv, (int)v0, st);
        }
        launch_raster_gather(c, depth, (long)S * S, views_per_image, R, t, nullptr, face_idx, grad_sub_ws, nv, (int)v0, grad_depth,
                             (long)S * S, nullptr, grad_R, grad_t, st, proj_ws ? proj_ws + (size_t)v0 * 4 * img_f : nullptr, lay);
        if (two) cudaStreamWaitEvent(st, ctx->ev_join[3], 0);   // the texture scratch is free for the next chunk; join
    }
    for (int i0 = 0; i0 < n_images; i0 += 32768) {
        const int ni = n_images - i0 < 32768 ? n_images - i0 : 32768;
        Launch l_(K_NORMAL_BWD, st);
        k_normal_bwd<<<pix_grid((long)S * S, ni), PIX_THREADS, 0, st>>>(c, depth + (long)i0 * S * S, S, S,
                                                                        grad_normal_ws + (long)i0 * S * S * 3,
                                                                        grad_depth + (long)i0 * S * S, 1);
    }
    return launch_status();
}

int g2s_render_fused_bwd(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R, const float* t,
                         const float* light, int n_images, int views_per_image, int align_corners,
                         const float* normal_ws, const float* recon_depth, const int32_t* face_idx,
                         const float* grad_recon_im, const float* grad_recon_depth, const float* proj_ws, int ws_views,
                         float* grad_sub_ws, float* grad_tex_ws, float* grad_normal_ws, float* grad_depth, float* grad_albedo,
                         float* grad_R, float* grad_t, float* grad_light, void* stream) {
    if (!grad_recon_im) return G2S_ERR_NULL;
    return fused_bwd_impl(ctx, cam, depth, albedo, R, t, light, n_images, views_per_image, align_corners, normal_ws, recon_depth,
                          face_idx, grad_recon_im, grad_recon_depth, nullptr, nullptr, nullptr, proj_ws, ws_views, grad_sub_ws, grad_tex_ws,
                          grad_normal_ws, grad_depth, grad_albedo, grad_R, grad_t, grad_light, stream);
}

int g2s_render_fused_loss_bwd(g2s_context* ctx, const g2s_camera* cam, const float* depth, const float* albedo, const float* R,
                              const float* t, const float* light, int n_images, int views_per_image, int align_corners,
                              const float* normal_ws, const float* recon_depth, const int32_t* face_idx,
                              const float* grad_recon_im, const float* grad_recon_depth, const g2s_photo_loss* loss,
                              const float* sums3, const float* grad_loss, const float* proj_ws, int ws_views,
                              float* grad_sub_ws, float* grad_tex_ws, float* grad_normal_ws, float* grad_depth,
                              float* grad_albedo, float* grad_R, float* grad_t, float* grad_light, void* stream) {
    if (!loss) return G2S_ERR_NULL;
    return fused_bwd_impl(ctx, cam, depth, albedo, R, t, light, n_images, views_per_image, align_corners, normal_ws, recon_depth,
                          face_idx, grad_recon_im, grad_recon_depth, loss, sums3, grad_loss, proj_ws, ws_views, grad_sub_ws, grad_tex_ws,
                          grad_normal_ws, grad_depth, grad_albedo, grad_R, grad_t, grad_light, stream);
}

int g2s_view_fwd(const float* view, int view_width, int B, float* R, float* t, void* stream) {
    if (!view || !R || !t) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    if (view_width != 3 && view_width != 5 && view_width != 6) return G2S_ERR_UNSUPPORTED;   // utils.py:70-71
    { Launch l_(K_VIEW, (cudaStream_t)stream);
      k_view_fwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view, view_width, B, R, t); }
    return launch_status();
}

int g2s_view_bwd(const float* view, int view_width, int B, const float* grad_R, const float* grad_t, float* grad_view,
                 void* stream) {
    if (!view || !grad_view) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    if (view_width != 3 && view_width != 5 && view_width != 6) return G2S_ERR_UNSUPPORTED;
    { Launch l_(K_VIEW, (cudaStream_t)stream);
      k_view_bwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view, view_width, B, grad_R, grad_t, grad_view); }
    return launch_status();
}

int g2s_light_fwd(const float* light, int B, float* light5, void* stream) {
    if (!light || !light5) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    { Launch l_(K_LIGHT, (cudaStream_t)stream);
      k_light_fwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(light, B, light5); }
    return launch_status();
}

int g2s_light_bwd(const float* light, int B, const float* grad_light5, float* grad_light, void* stream) {
    if (!light || !grad_light5 || !grad_light) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    { Launch l_(K_LIGHT, (cudaStream_t)stream);
      k_light_bwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(light, B, grad_light5, grad_light); }
    return launch_status();
}

int g2s_view_light_fwd(const float* view, int view_width, const float* light, int B, float* R, float* t, float* light5,
                       void* stream) {
    if (!view || !light || !R || !t || !light5) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    if (view_width != 3 && view_width != 5 && view_width != 6) return G2S_ERR_UNSUPPORTED;   // utils.py:70-71
    { Launch l_(K_VIEW, (cudaStream_t)stream);
      k_view_light_fwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view, view_width, light, B, R, t, light5); }
    return launch_status();
}

int g2s_view_light_bwd(const float* view, int view_width, const float* light, int B, const float* grad_R, const float* grad_t,
                       const float* grad_light5, float* grad_view, float* grad_light, void* stream) {
    if (!view || !light || !grad_light5 || !grad_view || !grad_light) return G2S_ERR_NULL;
    if (B <= 0) return G2S_ERR_SHAPE;
    if (view_width != 3 && view_width != 5 && view_width != 6) return G2S_ERR_UNSUPPORTED;
    { Launch l_(K_VIEW, (cudaStream_t)stream);
      k_view_light_bwd<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view, view_width, light, B, grad_R, grad_t, grad_light5,
                                                                          grad_view, grad_light); }
    return launch_status();
}

int g2s_grid3d_fwd(const g2s_camera* cam, const float* depth, long depth_view_stride, int B, int H, int W,
                   const int* crop, const float* R0, const float* t0, const float* R1, const float* R2, const float* t2,
                   float* out, void* stream) {
    if (!cam || !depth || !out) return G2S_ERR_NULL;
    if ((R0 == nullptr) != (t0 == nullptr) || (R2 == nullptr) != (t2 == nullptr)) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    Crop c = {0, 0, 0, 0, 0};
    if (crop) {
        c.on = 1; c.top = crop[0]; c.bottom = crop[1]; c.left = crop[2]; c.right = crop[3];
        if (c.top < 0 || c.bottom < 0 || c.left < 0 || c.right < 0 || c.top + c.bottom >= H || c.left + c.right >= W)
            return G2S_ERR_SHAPE;
    }
    { Launch l_(K_GRID3D, (cudaStream_t)stream); k_grid3d<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(make_cam(cam), depth, depth_view_stride,
                                                                                 H, W, c, R0, t0, R1, R2, t2, out); }
    return launch_status();
}

int g2s_grid3d_bwd(const g2s_camera* cam, const float* depth, long depth_view_stride, int B, int H, int W, int mode,
                   const float* R, const float* t, const float* grad_out, float* grad_depth, float* grad_R, float* grad_t,
                   void* stream) {
    if (!cam || !depth || !grad_out || !grad_depth) return G2S_ERR_NULL;
    if (mode < 0 || mode > 2) return G2S_ERR_UNSUPPORTED;
    if (mode != 0 && (!R || !t)) return G2S_ERR_NULL;
    if ((grad_R == nullptr) != (grad_t == nullptr)) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    { Launch l_(K_GRID3D, (cudaStream_t)stream);
      k_grid3d_bwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(
          make_cam(cam), depth, depth_view_stride, H, W, mode, R, t, grad_out, grad_depth, mode == 0 ? nullptr : grad_R,
          mode == 0 ? nullptr : grad_t); }
    return launch_status();
}

int g2s_grid_3d_to_2d_fwd(const g2s_camera* cam, const float* grid3d, int B, int H, int W, float* grid, void* stream) {
    if (!cam || !grid3d || !grid) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    { Launch l_(K_GRID_FWD, (cudaStream_t)stream);
      k_grid_3d_to_2d_fwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(make_cam(cam), grid3d, H, W, grid); }
    return launch_status();
}

int g2s_grid_3d_to_2d_bwd(const g2s_camera* cam, const float* grid3d, int B, int H, int W, const float* grad_grid,
                          float* grad_grid3d, void* stream) {
    if (!cam || !grid3d || !grad_grid || !grad_grid3d) return G2S_ERR_NULL;
    if (B <= 0 || B > 65535 || H < 2 || W < 2) return G2S_ERR_SHAPE;
    { Launch l_(K_GRID_BWD, (cudaStream_t)stream);
      k_grid_3d_to_2d_bwd<<<pix_grid((long)H * W, B), PIX_THREADS, 0, (cudaStream_t)stream>>>(make_cam(cam), grid3d, H, W, grad_grid,
                                                                                            grad_grid3d); }
    return launch_status();
}

int g2s_render_rgb_fwd(const g2s_camera* cam, const float* vertices3d, const float* im, long im_view_stride,
                       int n_views, int C, int tex_cube_size, const float* bg, int clamp, void* zbuf, float* rgb,
                       int32_t* face_idx, void* stream) {
    if (!cam || !vertices3d || !im || !bg || !zbuf || !rgb) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size) || C < 1 || C > 4) return G2S_ERR_SHAPE;
    if (tex_cube_size != 2) return G2S_ERR_UNSUPPORTED;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_splat<true>(c, nullptr, 0, 1, nullptr, nullptr, vertices3d, (unsigned long long*)zbuf, n_views, n_views, 0, st))
        return rc;
    Bg b4 = {{0.f, 0.f, 0.f, 0.f}};
    for (int i = 0; i < C; i++) b4.c[i] = bg[i];
    const float eps = 1e-3f;  // nr.Renderer.rasterizer_eps
    unsigned long long* zb = (unsigned long long*)zbuf;
    const dim3 g = pix_grid((long)S * S, n_views);
    switch (C) {
        case 1: { Launch l_(K_RESOLVE_RGB, st); k_resolve_rgb<1><<<g, PIX_THREADS, 0, st>>>(c, zb, vertices3d, im, im_view_stride, b4, eps, clamp, rgb, face_idx); } break;
        case 2: { Launch l_(K_RESOLVE_RGB, st); k_resolve_rgb<2><<<g, PIX_THREADS, 0, st>>>(c, zb, vertices3d, im, im_view_stride, b4, eps, clamp, rgb, face_idx); } break;
        case 3: { Launch l_(K_RESOLVE_RGB, st); k_resolve_rgb<3><<<g, PIX_THREADS, 0, st>>>(c, zb, vertices3d, im, im_view_stride, b4, eps, clamp, rgb, face_idx); } break;
        default: { Launch l_(K_RESOLVE_RGB, st); k_resolve_rgb<4><<<g, PIX_THREADS, 0, st>>>(c, zb, vertices3d, im, im_view_stride, b4, eps, clamp, rgb, face_idx); } break;
    }
    return launch_status();
}

int g2s_render_depth_fwd(const g2s_camera* cam, const float* vertices3d, int n_views, void* zbuf, float* depth_out,
                         int32_t* face_idx, void* stream) {
    if (!cam || !vertices3d || !zbuf || !depth_out) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size)) return G2S_ERR_SHAPE;
    Cam c = make_cam(cam);
    c.clamp_lo = -3.402823466e38f;     // nr.render_depth does not clamp (renderer.py:122-124 does, afterwards)
    c.clamp_hi = 3.402823466e38f;
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_splat<true>(c, nullptr, 0, 1, nullptr, nullptr, vertices3d, (unsigned long long*)zbuf, n_views, n_views, 0, st))
        return rc;
    FusedArgs fa = {};
    { Launch l_(K_RESOLVE, st); k_resolve<false><<<pix_grid2(S, n_views), dim3(PBX, PBY), 0, st>>>(c, (unsigned long long*)zbuf, depth_out,
                                                                             face_idx, fa); }
    return launch_status();
}

int g2s_render_depth_bwd(const g2s_camera* cam, const float* vertices3d, int n_views, const int32_t* face_idx,
                         const float* grad_depth_out, float* raster_ws, float* grad_vertices, void* stream) {
    if (!cam || !vertices3d || !face_idx || !grad_depth_out || !raster_ws || !grad_vertices) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size)) return G2S_ERR_SHAPE;
    const Cam c = make_cam(cam);
    const int S = c.S;
    cudaStream_t st = (cudaStream_t)stream;
    float* g_sub = raster_ws_gsub(raster_ws, n_views, S);
    const long n = (long)n_views * S * S;
    // flip + 2x2 mean backward: every sub-pixel of an output pixel gets a quarter of its cotangent (no clamp here)
    { Launch l_(K_CLAMP_GRAD, st); k_clamp_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(grad_depth_out, grad_depth_out, -3.402823466e38f,
                                                              3.402823466e38f, n, g_sub); }
    launch_raster_bwd(c, nullptr, 0, 1, nullptr, nullptr, vertices3d, face_idx, raster_ws, n_views, 0, nullptr, 0, grad_vertices,
                      nullptr, nullptr, st);
    return launch_status();
}

int g2s_render_rgb_bwd(const g2s_camera* cam, const float* vertices3d, const float* im, long im_view_stride, int n_views,
                       int C, int tex_cube_size, const float* bg, int clamp, const int32_t* face_idx, const float* grad_rgb,
                       float* grad_im, long grad_im_view_stride, float* rgb_ws, float* raster_ws, float* grad_vertices,
                       void* stream) {
    if (!cam || !vertices3d || !im || !bg || !face_idx || !grad_rgb) return G2S_ERR_NULL;
    if (!grad_im && !grad_vertices) return G2S_ERR_NULL;
    if (grad_vertices && (!rgb_ws || !raster_ws)) return G2S_ERR_NULL;
    if (n_views <= 0 || n_views > 65535 || bad_size(cam->image_size) || C < 1 || C > 4) return G2S_ERR_SHAPE;
    if (tex_cube_size != 2 || (grad_vertices && C != 3)) return G2S_ERR_UNSUPPORTED;
    const Cam c = make_cam(cam);
    const int S = c.S, is = 2 * S;
    cudaStream_t st = (cudaStream_t)stream;
    Bg b4 = {{0.f, 0.f, 0.f, 0.f}};
    for (int i = 0; i < C; i++) b4.c[i] = bg[i];
    const float eps = 1e-3f;  // nr.Renderer.rasterizer_eps
    const dim3 g = pix_grid((long)S * S, n_views);
    if (grad_im) {
        Launch l_(K_RESOLVE_RGB, st);
        switch (C) {
            case 1: k_resolve_rgb_bwd<1><<<g, PIX_THREADS, 0, st>>>(c, face_idx, vertices3d, im, im_view_stride, b4, eps, clamp, grad_rgb, grad_im, grad_im_view_stride); break;
            case 2: k_resolve_rgb_bwd<2><<<g, PIX_THREADS, 0, st>>>(c, face_idx, vertices3d, im, im_view_stride, b4, eps, clamp, grad_rgb, grad_im, grad_im_view_stride); break;
            case 3: k_resolve_rgb_bwd<3><<<g, PIX_THREADS, 0, st>>>(c, face_idx, vertices3d, im, im_view_stride, b4, eps, clamp, grad_rgb, grad_im, grad_im_view_stride); break;
            default: k_resolve_rgb_bwd<4><<<g, PIX_THREADS, 0, st>>>(c, face_idx, vertices3d, im, im_view_stride, b4, eps, clamp, grad_rgb, grad_im, grad_im_view_stride); break;
        }
    }
    if (grad_vertices) {
        const size_t img = (size_t)S * S;
        float* rgb_map = rgb_ws;                                  // [n, 2S, 2S, 4]
        float* g4 = rgb_ws + (size_t)n_views * 16 * img;          // [n, S, S, 4]
        float* proj = raster_ws;
        float* vgrad = proj + (size_t)n_views * 4 * img;
        Launch l_(K_RESOLVE_RGB, st);
        k_rgb_map<<<pix_grid((long)is * is, n_views), PIX_THREADS, 0, st>>>(c, face_idx, vertices3d, im, im_view_stride, b4, eps, rgb_map);
        k_rgb_gquarter<<<g, PIX_THREADS, 0, st>>>(S, rgb_map, grad_rgb, clamp, g4);
        k_project_points<false><<<g, PIX_THREADS, 0, st>>>(c, vertices3d, proj, vgrad);      // also zeroes vgrad
        const int Q4 = 4 * (S - 1) * (S - 1);
        k_backward_pixel_map<<<dim3((Q4 + 127) / 128, n_views), 128, 0, st>>>(c, face_idx, proj, rgb_map, g4, eps, vgrad);
        k_points_bwd<<<g, PIX_THREADS, 0, st>>>(c, vertices3d, vgrad, grad_vertices);
    }
    return launch_status();
}

int g2s_selftest_division(unsigned long long n_pairs, unsigned seed, unsigned long long* mismatches_dev, void* stream) {
    if (!mismatches_dev) return G2S_ERR_NULL;
    const int blocks = 148 * 8, threads = 256;
    const unsigned long long per = (n_pairs + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    k_selftest_division<<<blocks, threads, 0, (cudaStream_t)stream>>>(per, seed, mismatches_dev);
    return launch_status();
}

int g2s_selftest_index_math(int image_size, int views_per_image, long n_views, unsigned long long* mismatches_dev,
                            void* stream) {
    if (!mismatches_dev) return G2S_ERR_NULL;
    if (bad_size(image_size) || views_per_image <= 0 || n_views <= 0 || n_views > (1L << 30)) return G2S_ERR_SHAPE;
    k_selftest_index_math<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(image_size, views_per_image, n_views,
                                                                   vpi_magic(views_per_image, n_views), mismatches_dev);
    return launch_status();
}

int g2s_selftest_face_vertices(int image_size, unsigned long long* mismatches_dev, void* stream) {
    if (!mismatches_dev) return G2S_ERR_NULL;
    if (bad_size(image_size)) return G2S_ERR_SHAPE;
    k_selftest_face_vertices<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(image_size, mismatches_dev);
    return launch_status();
}

int g2s_selftest_raster(unsigned long long n_triangles, unsigned seed, int image_size, unsigned long long* mismatches_dev,
                        void* stream) {
    if (!mismatches_dev) return G2S_ERR_NULL;
    if (bad_size(image_size)) return G2S_ERR_SHAPE;
    const int blocks = 148 * 4, threads = 128;
    const unsigned long long per = (n_triangles + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    k_selftest_raster<<<blocks, threads, 0, (cudaStream_t)stream>>>(per, seed, 2 * image_size, mismatches_dev);
    return launch_status();
}

#ifdef G2S_COUNT_SLOW
int g2s_debug_slow_counters(unsigned long long* out2) {
    return cudaMemcpyFromSymbol(out2, g2s::g_slow_counters, 16) == cudaSuccess ? G2S_OK : G2S_ERR_LAUNCH;
}
#endif

long g2s_launch_count(void) { return g_launches.load(); }

int g2s_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return G2S_OK;
}

int g2s_profile_read(int max_kernels, const char** names, float* total_ms, int* launches) {
    if (!names || !total_ms || !launches) return G2S_ERR_NULL;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    float ms[K_COUNT] = {0};
    int cnt[K_COUNT] = {0};
    for (auto& r : g_prof_recs) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.id] += t;
        cnt[r.id]++;
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    int n = 0;
    for (int k = 0; k < K_COUNT && n < max_kernels; k++)
        if (cnt[k]) { names[n] = kKernelNames[k]; total_ms[n] = ms[k]; launches[n] = cnt[k]; n++; }
    return n;
}

}  // extern "C"

#include "g2s_callers.cuh"
