/*
 * g2s_b200.h -- C ABI of libg2s_b200.so: the sm_100a CUDA implementation of GAN2Shape's
 * depth-map renderer path (reference: GAN2Shape/renderer/renderer.py, GAN2Shape/renderer/utils.py and
 * the external `neural_renderer` rasteriser it calls).
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it is the camera struct
 *     (host memory, copied by value into the launch);
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated or freed here;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no host sync;
 *   - re-entrant; the only process-wide state is the instrumentation at the end of this header and a per-device cache of
 *     device attributes (SM count, the shared-memory opt-in of one kernel).  Streams and events the library needs for
 *     its own pipelining live in a caller-owned g2s_context;
 *   - returns G2S_OK (0) or a negative G2S_ERR_* code; never throws.  The reference's neural_renderer
 *     extension reports the same class of failures through AT_ASSERTM -> RuntimeError
 *     (CHECK_CUDA / CHECK_CONTIGUOUS); the Python host layer turns non-zero codes into RuntimeError.
 *   - all tensors are contiguous fp32 (or int32 where stated), batch-major, row-major.
 *
 * Geometry conventions: S = image_size; the rasteriser works at is = 2S sub-pixels per side
 * (neural_renderer anti_aliasing=True, the reference keeps the default); the mesh is the regular grid
 * of utils.py:76-80 over an S x S depth map with fill_back=True: 4(S-1)^2 faces.  Face-index maps
 * are int32 [n_views, 2S, 2S] in IMAGE orientation (row 0 = top), -1 where no face covers the
 * sub-pixel.
 */
#ifndef G2S_B200_H
#define G2S_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2S_OK 0
#define G2S_ERR_NULL (-1)        /* a required pointer is NULL */
#define G2S_ERR_SHAPE (-2)       /* a size is out of the supported range */
#define G2S_ERR_LAUNCH (-3)      /* cudaGetLastError() reported a launch failure */
#define G2S_ERR_UNSUPPORTED (-4) /* mode / flag value not implemented */

/* State of one Renderer object: renderer.py:14-54 (K, inv_K, rot_center_depth, image_size) plus the
 * rasteriser constants neural_renderer would hold (near/far) and warp_canon_depth's clamp range
 * (renderer.py:122-124). */
typedef struct g2s_camera {
    float K[9];
    float inv_K[9];
    float rot_center_depth;
    float near_z;    /* render_depth: 0.1 (nr module default); render_rgb: renderer_min_depth */
    float far_z;     /* render_depth: 100 (nr module default); render_rgb: renderer_max_depth */
    float clamp_lo;  /* min_depth - margin */
    float clamp_hi;  /* max_depth + margin */
    int32_t image_size;
    float K_grid[9]; /* the CURRENT K of the grid operators (renderer.py:82-88 uses self.K, which downscale_K rescales), while
                      * K above is what the rasteriser projects with: neural_renderer captures K at construction
                      * (renderer.py:47-50) and downscale_K never reaches it.  Equal to K unless downscale_K was called. */
} g2s_camera;

int g2s_version(void);
const char *g2s_error_string(int code);

/* ---- caller-owned context -------------------------------------------------------------------
 * Holds what the library would otherwise keep in globals: the internal non-blocking streams and the fork / join events of
 * the multi-lane forward and the two-lane backward (g2s_render_fused_fwd / _bwd), created for the device that is current at
 * creation time, and the tuning
 * read once from the environment (G2S_FWD_LANES, G2S_NO_PIPELINE).  One context per device and per thread of use; a NULL
 * context is accepted everywhere and means "no internal streams" (single lane). */
typedef struct g2s_context g2s_context;
int g2s_context_create(g2s_context **out);
void g2s_context_destroy(g2s_context *ctx);

/* ---- workspace sizes --------------------------------------------------------------------------
 * Every workspace is caller-owned; g2s_workspace_bytes returns the bytes one needs (0 for bad arguments).
 *   G2S_WS_ZBUFFER     forward z-buffer for n views: packed 64-bit keys (zp bits << 32 | face index) per sub-pixel, the work
 *                      list of the rasteriser's second stage and its counters (same as g2s_zbuffer_bytes); must be
 *                      initialised once with g2s_zbuffer_init, every forward leaves it re-initialised;
 *   G2S_WS_RASTER_BWD  backward of the rasteriser for n views: [n,9,S,S] floats (projected vertices | vertex gradients |
 *                      masked quarter gradient); 16-byte aligned; needs no initialisation UNLESS the fused backward is
 *                      given the forward's projected vertices (proj_ws): then its first n*4*S*S floats must be zero on entry
 *                      and are left zero (see g2s_render_fused_bwd);
 *   G2S_WS_TEX_BWD     per-view texture gradient of the fused backward: [n,S,S,4] floats;
 *   G2S_WS_TEXELS      packed texel map of the fused render for n IMAGES: [n,S,S,8] floats;
 *   G2S_WS_GRAD_NORMAL normal-map gradient of the fused backward for n IMAGES: [n,S,S,3] floats;
 *   G2S_WS_RGB_MAP     rgb backward for n views: supersampled colour map [n,2S,2S,4] + quarter gradient [n,S,S,4] floats.
 *   G2S_WS_LOSS        partial sums of the fused photometric loss for n views (g2s_render_fused_loss_fwd). */
enum { G2S_WS_ZBUFFER = 0, G2S_WS_RASTER_BWD = 1, G2S_WS_TEX_BWD = 2, G2S_WS_TEXELS = 3, G2S_WS_GRAD_NORMAL = 4,
       G2S_WS_RGB_MAP = 5, G2S_WS_LOSS = 6 };
size_t g2s_workspace_bytes(int kind, int n, int image_size);
size_t g2s_zbuffer_bytes(int n_views, int image_size);
int g2s_zbuffer_init(void *zbuf, int n_views, int image_size, float far_z, void *stream);

/* ---- warp_canon_depth: renderer.py:116-125 (+ nr.Renderer.render_depth, renderer.py:120) -------
 * depth: [*, S, S] with `depth_view_stride` floats between consecutive views (0 = one depth map
 * shared by all views, the `expand`ed tensor of model.py:260-262); R [n_views,3,3], t [n_views,3].
 * Outputs: recon_depth [n_views,S,S] (clamped), face_idx int32 [n_views,2S,2S]. */
int g2s_warp_depth_fwd(const g2s_camera *cam, const float *depth, long depth_view_stride, const float *R,
                       const float *t, int n_views, void *zbuf, float *recon_depth, int32_t *face_idx,
                       void *stream);

/* Backward of the above: neural_renderer's backward_depth_map (approximate x/y gradient) chained
 * through flip / 2x2 average / clamp on one side and gather / projection / rotation on the other.
 * grad_depth is ACCUMULATED (caller zero-fills) with `grad_depth_view_stride` floats between views
 * (0 = sum over views into one [S,S] map); grad_R [n_views,3,3] and grad_t [n_views,3] are
 * ACCUMULATED too (NULL to skip both).  Workspace grad_sub_ws: n_views * 9 * S * S floats, 16-byte
 * aligned (projected vertices uvz- | vertex gradients uvz- | masked quarter gradient). */
int g2s_warp_depth_bwd(const g2s_camera *cam, const float *depth, long depth_view_stride, const float *R,
                       const float *t, int n_views, const int32_t *face_idx, const float *recon_depth,
                       const float *grad_recon_depth, float *grad_sub_ws, float *grad_depth,
                       long grad_depth_view_stride, float *grad_R, float *grad_t, void *stream);

/* ---- get_warped_2d_grid / get_inv_warped_2d_grid: renderer.py:104-114 ---------------------------
 * depth [B,H,W] (stride as above), R/t [B,..] -> grid [B,H,W,2].  inverse != 0 selects the inverse warp. */
int g2s_warp_grid_fwd(const g2s_camera *cam, const float *depth, long depth_view_stride, const float *R,
                      const float *t, int B, int H, int W, int inverse, float *grid, void *stream);
/* grad_depth [B,H,W] is WRITTEN; grad_R / grad_t ACCUMULATED (NULL to skip both). */
int g2s_warp_grid_bwd(const g2s_camera *cam, const float *depth, long depth_view_stride, const float *R,
                      const float *t, int B, int H, int W, int inverse, const float *grad_grid,
                      float *grad_depth, float *grad_R, float *grad_t, void *stream);

/* ---- get_normal_from_depth: renderer.py:127-139 -------------------------------------------------
 * depth [B,H,W] -> normal [B,H,W,3]; backward WRITES grad_depth [B,H,W]. */
int g2s_normal_fwd(const g2s_camera *cam, const float *depth, int B, int H, int W, float *normal, void *stream);
int g2s_normal_bwd(const g2s_camera *cam, const float *depth, int B, int H, int W, const float *grad_normal,
                   float *grad_depth, int accumulate, void *stream);

/* ---- grid_sample as the reference uses it: model.py:151,270; renderer.py:179,223,241,261,263 -----
 * input [B,C,H,W], grid [B,Ho,Wo,2] -> out [B,C,Ho,Wo]; zeros padding.
 * mode: 0 = bilinear, 1 = nearest.  `input_batch_stride` = floats between batch items of input
 * (0 = one image broadcast over the batch).  Backward: grad_input ACCUMULATED (caller zero-fills; NULL to
 * skip), grad_grid WRITTEN (NULL to skip; zeros for nearest). */
int g2s_sample_fwd(const float *input, long input_batch_stride, const float *grid, int B, int C, int H, int W,
                   int Ho, int Wo, int mode, int align_corners, float *out, void *stream);
int g2s_sample_bwd(const float *input, long input_batch_stride, const float *grid, const float *grad_out, int B,
                   int C, int H, int W, int Ho, int Wo, int mode, int align_corners, float *grad_input,
                   long grad_input_batch_stride, float *grad_grid, void *stream);

/* ---- the fused projected-view render ("chain C", model.py:243-270) -------------------------------
 * For image i (n_images of them) and its views_per_image views b:
 *   normal   = get_normal_from_depth(depth[i])                    (renderer.py:127-139)
 *   texture  = shade(normal, light[b], albedo[i])                 (model.py:355-360)
 *   recon_depth[b] = warp_canon_depth(depth[i]; R[b], t[b])       (renderer.py:116-125)
 *   grid     = get_inv_warped_2d_grid(recon_depth[b])             (renderer.py:110-114)
 *   recon_im[b] = grid_sample(texture, grid).clamp(-1, 1)         (model.py:270)
 * light [n_views,5] = (ambient a, diffuse b, direction dx,dy,dz) as get_lighting_directions returns.
 * The views are processed in chunks that bound the z-buffer workspace; g2s_chunk_views(S) returns the recommended chunk
 * (256 MB of z-buffer keys: 512 views at 128^2, 128 at 256^2 -- measured: longer launches win).  `ws_views` = views
 * the z-buffer workspace holds: with room for L >= 2 recommended chunks (and more views than one chunk) the chunks rotate
 * over L lanes (L <= 4; the Python host asks for 2, the measured optimum) -- one part of the workspace and one stream each:
 * `stream` and the non-blocking streams of `ctx` (NULL ctx: one lane), forked from and joined back into `stream` with events
 * (capturable in a CUDA graph) -- so that one chunk's rasteriser fills the SMs the tail of the previous one leaves idle;
 * otherwise one chunk of min(ws_views, recommended) views at a time on `stream` alone (also while per-kernel timing is on,
 * or with G2S_NO_PIPELINE set).  Workspaces: zbuf (g2s_zbuffer_bytes(ws_views, S), initialised), normal_ws [n_images,S,S,8] (packed texels: normal xyz,
 * albedo rgb, 2 pad; kept for the backward).
 * Outputs: recon_im [n_views,3,S,S], recon_depth [n_views,S,S], face_idx [n_views,2S,2S] (may be NULL).
 * Optional (sample_pseudo_imgs, model.py:291-328 -> render_given_view(..., mask, grid_sample=True), renderer.py:257-264):
 * mask_out [n_views,S,S] = grid_sample(mask, grid, mode='nearest') of mask_in [n_images,S,S] (NULL = all ones); pass
 * mask_out = NULL to skip.
 * proj_ws (may be NULL) [n_views,S,S,4] floats: the rasteriser also stores every vertex it projected (sub-pixel x, y, depth z,
 * pad) for the backward, which then does not project the mesh again (pass the same buffer to g2s_render_fused_bwd). */
int g2s_chunk_views(int image_size);
int g2s_chunk_views_bwd(int image_size);   /* recommended ws_views of g2s_render_fused_bwd (~2 GB of scratch) */
int g2s_render_fused_fwd(g2s_context *ctx, const g2s_camera *cam, const float *depth, const float *albedo, const float *R,
                         const float *t, const float *light, int n_images, int views_per_image,
                         int align_corners, void *zbuf, int ws_views, float *normal_ws, float *recon_im,
                         float *recon_depth, int32_t *face_idx, const float *mask_in, float *mask_out,
                         float *proj_ws, void *stream);

/* Backward of the fused render.  Cotangents: grad_recon_im [n_views,3,S,S] (required),
 * grad_recon_depth [n_views,S,S] (may be NULL).  Workspaces: grad_sub_ws [ws_views,9,S,S] (as above),
 * grad_tex_ws [ws_views,S,S,4] (per-view texture gradient, packed rgb-; 16-byte aligned; chunked like the forward), grad_normal_ws [n_images,S,S,8] (packed texels: normal xyz,
 * albedo rgb, 2 pad; kept for the backward).
 * Outputs, all WRITTEN: grad_depth [n_images,S,S], grad_albedo [n_images,3,S,S], grad_R [n_views,3,3],
 * grad_t [n_views,3], grad_light [n_views,5].
 * ctx (may be NULL = everything on `stream`): the bandwidth-bound kernels of a chunk run on the context's side stream under the
 * issue-bound ones, forked from / joined to `stream` with the context's events.
 * proj_ws (may be NULL): the projected vertices the forward stored; NULL = re-project them here (k_project_verts).
 * WITH proj_ws the first ws_views * 4 * S * S floats of grad_sub_ws (the vertex-gradient scratch) must be ZERO on entry and are
 * zero again on return (the vertex kernel clears what it consumed): clear the workspace once when it is allocated, not per
 * call.  Without proj_ws the workspace needs no initialisation. */
int g2s_render_fused_bwd(g2s_context *ctx, const g2s_camera *cam, const float *depth, const float *albedo, const float *R,
                         const float *t, const float *light, int n_images, int views_per_image,
                         int align_corners, const float *normal_ws, const float *recon_depth,
                         const int32_t *face_idx, const float *grad_recon_im, const float *grad_recon_depth,
                         const float *proj_ws, int ws_views, float *grad_sub_ws, float *grad_tex_ws,
                         float *grad_normal_ws, float *grad_depth, float *grad_albedo, float *grad_R, float *grad_t,
                         float *grad_light, void *stream);

/* ---- fused render + masked photometric loss: model.py:243-274 as one pass ----------------------------------------
 * The step-3 loss of the reference, loss_l1_im = PhotometricLoss(recon_im, projected_samples, mask = (recon_depth <
 * max_depth + margin) * masks) (model.py:265-274, losses.py:39-51 without conf_sigma), taken inside the render:
 * the forward's z-buffer resolve also accumulates sum(|recon_im - target| * m) and sum(m) -- per-CTA pairs in loss_ws
 * (G2S_WS_LOSS for n_views), summed in a fixed order -- and writes out3 = {loss, numerator, denominator} (device); the
 * backward forms the loss's cotangent sign(recon_im - target) * m * grad_loss / denominator inside its pixel stage and ADDS it
 * to grad_recon_im (may be NULL: e.g. no perceptual loss on the render), so neither the loss's own passes over recon_im /
 * target / recon_depth nor the cotangent image exist.  recon_im, recon_depth, face_idx are written as by
 * g2s_render_fused_fwd.  grad_loss is a DEVICE scalar; sums3 = the forward's out3. */
typedef struct {
    const float *target;     /* [n_views,3,S,S] */
    const float *view_mask;  /* [n_views,S,S] or NULL (all ones) */
    float depth_thresh;      /* a pixel counts when recon_depth < depth_thresh */
} g2s_photo_loss;
int g2s_render_fused_loss_fwd(g2s_context *ctx, const g2s_camera *cam, const float *depth, const float *albedo,
                              const float *R, const float *t, const float *light, int n_images, int views_per_image,
                              int align_corners, void *zbuf, int ws_views, float *normal_ws, float *recon_im,
                              float *recon_depth, int32_t *face_idx, const g2s_photo_loss *loss, void *loss_ws,
                              float *out3, float *proj_ws, void *stream);
int g2s_render_fused_loss_bwd(g2s_context *ctx, const g2s_camera *cam, const float *depth, const float *albedo,
                              const float *R, const float *t, const float *light, int n_images, int views_per_image,
                              int align_corners, const float *normal_ws, const float *recon_depth,
                              const int32_t *face_idx, const float *grad_recon_im, const float *grad_recon_depth,
                              const g2s_photo_loss *loss, const float *sums3, const float *grad_loss,
                              const float *proj_ws, int ws_views, float *grad_sub_ws, float *grad_tex_ws,
                              float *grad_normal_ws, float *grad_depth, float *grad_albedo, float *grad_R, float *grad_t,
                              float *grad_light, void *stream);

/* ---- mesh-texture render: nr.Renderer.render_rgb as renderer.py:196,230,248,272,275 call it ------
 * vertices3d [n_views,S*S,3] (already rotated/translated 3-D grid), im [*,C,S,S] per-vertex colours
 * (tex_cube_size 2: perspective-correct barycentric blend of the three vertex colours; 1: one colour
 * per face) with `im_view_stride` floats between views (0 = shared), background colour bg[C] (host).
 * Output rgb [n_views,C,S,S] = 2x2 mean of the 2S x 2S render, then clamp(-1,1) when clamp != 0.
 * face_idx (may be NULL) as above.  C <= 4. */
int g2s_render_rgb_fwd(const g2s_camera *cam, const float *vertices3d, const float *im, long im_view_stride,
                       int n_views, int C, int tex_cube_size, const float *bg, int clamp, void *zbuf,
                       float *rgb, int32_t *face_idx, void *stream);

/* ---- the neural_renderer-level boundary: nr.Renderer.render_depth(vertices, faces) as renderer.py:120 calls it ----
 * vertices3d [n_views,S*S,3]: the 3-D vertices of the S x S grid mesh (faces = utils.py:76-80, fill_back) in camera
 * space.  depth_out [n_views,S,S] = 2x2 mean of the 2S x 2S depth map, background = cam->far_z, NOT clamped (cam->clamp_*
 * are ignored); face_idx may be NULL.  Backward = neural_renderer's backward_depth_map (approximate x/y gradient) +
 * flip / mean + vertices_to_faces gather + projection: WRITES grad_vertices [n_views,S*S,3]; raster_ws as
 * g2s_warp_depth_bwd's grad_sub_ws (n_views * 9 * S * S floats, 16-byte aligned). */
int g2s_render_depth_fwd(const g2s_camera *cam, const float *vertices3d, int n_views, void *zbuf, float *depth_out,
                         int32_t *face_idx, void *stream);
int g2s_render_depth_bwd(const g2s_camera *cam, const float *vertices3d, int n_views, const int32_t *face_idx,
                         const float *grad_depth_out, float *raster_ws, float *grad_vertices, void *stream);

/* Backward of g2s_render_rgb_fwd with respect to the per-vertex colours `im`: neural_renderer's backward_textures chained
 * through get_textures_from_im (utils.py:98-109), the fill_back texture permutation, the 2x2 mean and clamp(-1,1).
 * face_idx [n_views,2S,2S] as written by the forward; grad_rgb [n_views,C,S,S]; grad_im is ACCUMULATED (caller zero-fills)
 * with grad_im_view_stride floats between views (0 = one image shared by all views); NULL to skip.
 * Geometry gradient (NULL grad_vertices to skip): neural_renderer's backward_pixel_map -- the approximate silhouette / colour
 * edge gradient of Kato et al., one thread per face scanning along each edge and each axis -- chained
 * through vertices_to_faces and the projection: WRITES grad_vertices [n_views,S*S,3].  It needs rgb_ws
 * (G2S_WS_RGB_MAP) and raster_ws (G2S_WS_RASTER_BWD); C must be 3. */
int g2s_render_rgb_bwd(const g2s_camera *cam, const float *vertices3d, const float *im, long im_view_stride, int n_views,
                       int C, int tex_cube_size, const float *bg, int clamp, const int32_t *face_idx,
                       const float *grad_rgb, float *grad_im, long grad_im_view_stride, float *rgb_ws, float *raster_ws,
                       float *grad_vertices, void *stream);

/* ---- set_transform_matrices: utils.py:33-73 (view [B, 3|5|6] -> R = Rz Ry Rx [B,3,3], t [B,3]; other widths return
 * G2S_ERR_UNSUPPORTED as utils.py:70-71 raises) and get_lighting_directions: model.py:347-353 (raw light [B,4] ->
 * light5 [B,5] = ambient a, diffuse b, unit direction).  Backward: grad_R / grad_t may be NULL (zeros). */
int g2s_view_fwd(const float *view, int view_width, int B, float *R, float *t, void *stream);
int g2s_view_bwd(const float *view, int view_width, int B, const float *grad_R, const float *grad_t,
                 float *grad_view, void *stream);
int g2s_light_fwd(const float *light, int B, float *light5, void *stream);
int g2s_light_bwd(const float *light, int B, const float *grad_light5, float *grad_light, void *stream);
/* both of the above for the same B views in one launch each way (what Renderer.render_chain uses: the reference's one-image
 * steps are launch-bound).  grad_R / grad_t may be NULL (= zero). */
int g2s_view_light_fwd(const float *view, int view_width, const float *light, int B, float *R, float *t, float *light5,
                       void *stream);
int g2s_view_light_bwd(const float *view, int view_width, const float *light, int B, const float *grad_R,
                       const float *grad_t, const float *grad_light5, float *grad_view, float *grad_light, void *stream);

/* ---- 3-D grid helpers used by render_yaw / render_view / render_given_view ----------------------
 * depth_to_3d_grid (renderer.py:74-80) followed by an optional inverse warp by (R0,t0)
 * (renderer.py:164-167), then a forward rotation R1 (renderer.py:181-183) and optional (R2,t2)
 * (renderer.py:185-192).  Any of the transforms may be NULL.  crop = {top,bottom,left,right} host
 * ints or NULL (renderer.py:145-158).  out [B,H*W,3]. */
int g2s_grid3d_fwd(const g2s_camera *cam, const float *depth, long depth_view_stride, int B, int H, int W,
                   const int *crop, const float *R0, const float *t0, const float *R1, const float *R2,
                   const float *t2, float *out, void *stream);
/* Backward of the three public forms of the above (no crop): depth_to_3d_grid (mode 0, renderer.py:74-80),
 * get_warped_3d_grid (mode 1: R, t of the forward warp, renderer.py:90-95) and get_inv_warped_3d_grid (mode 2,
 * renderer.py:97-102).  grad_out [B,H*W,3]; grad_depth [B,H,W] WRITTEN; grad_R / grad_t ACCUMULATED (NULL to skip both). */
int g2s_grid3d_bwd(const g2s_camera *cam, const float *depth, long depth_view_stride, int B, int H, int W, int mode,
                   const float *R, const float *t, const float *grad_out, float *grad_depth, float *grad_R, float *grad_t,
                   void *stream);
/* grid_3d_to_2d: renderer.py:82-88.  grid3d [B,H*W,3] -> grid [B,H,W,2] in [-1,1] (with cam->K_grid); backward WRITES
 * grad_grid3d [B,H*W,3]. */
int g2s_grid_3d_to_2d_fwd(const g2s_camera *cam, const float *grid3d, int B, int H, int W, float *grid, void *stream);
int g2s_grid_3d_to_2d_bwd(const g2s_camera *cam, const float *grid3d, int B, int H, int W, const float *grad_grid,
                          float *grad_grid3d, void *stream);

/* ---- callers either side of the path -------------------------------------------
 * Every reduction below is deterministic: per-block partial sums go to the caller-owned `reduce_ws`
 * (g2s_reduce_ws_bytes() bytes, 8-byte aligned) and are finished in a fixed order. */
size_t g2s_reduce_ws_bytes(void);

/* get_clamped_depth + rescale_depth: model.py:337-345, 85-86.  depth_raw is n_groups groups of maps_per_group maps [H,W];
 * the mean that is subtracted before tanh is taken per group (the reference's `view(1,-1).mean(1)` is ONE group over the
 * whole tensor).  clamp_border: the 2 leftmost / 2 rightmost columns are blended with the reference's literal F.pad
 * weight 1.02: depth*(1-1.02) + 1.02*border_depth.  mean_out [n_groups] is kept for the backward.  W >= 4;
 * n_groups <= 148 (reduce_ws holds 32 partial sums per group). */
int g2s_clamped_depth_fwd(const float *depth_raw, int n_groups, int maps_per_group, int H, int W, float min_depth,
                          float max_depth, float border_depth, int clamp_border, void *reduce_ws, float *mean_out,
                          float *depth, void *stream);
int g2s_clamped_depth_bwd(const float *depth_raw, const float *mean, const float *grad_depth, int n_groups,
                          int maps_per_group, int H, int W, float min_depth, float max_depth, int clamp_border,
                          void *reduce_ws, float *grad_raw, void *stream);

/* get_shading: model.py:355-360.  normal [*,H*W,3], albedo [*,3,H*W] with `*_view_stride` floats between views (0 = one
 * map shared by all views), light5 [B,5] (g2s_light_fwd) -> diffuse [B,1,H*W] (may be NULL), texture [B,3,H*W].
 * Backward: grad_diffuse / grad_texture (either may be NULL, not both); grad_normal, grad_albedo (strides as above,
 * 0 = summed over the views) and grad_light5 [B,5] are ACCUMULATED (caller zero-fills; any may be NULL). */
int g2s_shading_fwd(const float *normal, long normal_view_stride, const float *light5, const float *albedo,
                    long albedo_view_stride, int B, int HW, float *diffuse, float *texture, void *stream);
int g2s_shading_bwd(const float *normal, long normal_view_stride, const float *light5, const float *albedo,
                    long albedo_view_stride, int B, int HW, const float *grad_diffuse, const float *grad_texture,
                    float *grad_normal, long grad_normal_view_stride, float *grad_light5, float *grad_albedo,
                    long grad_albedo_view_stride, void *stream);

/* Validity mask + PhotometricLoss in one pass: model.py:146-150 / 265-269 and losses.py:39-51.
 *   mask[b,i] = (recon_depth ? recon_depth[b,i] < depth_thresh : 1) * (mask_in ? mask_in[b,i] : 1)
 *   l         = |im1 - im2|;  with conf_sigma (losses.py:44-45): l * 2**0.5 / (sigma + 1e-7) + log(sigma + 1e-7)
 *   loss      = sum(l * mask) / (C * sum(mask))                   (no mask at all: the plain mean of losses.py:50)
 * im1 [B,C,HW]; im2 [*,C,HW] with im2_batch_stride floats between items (0 = one target for all views);
 * recon_depth, mask_in [B,HW]; conf_sigma [B,sigma_channels,HW] with sigma_channels = 1 or C, or NULL.
 * out3 = {loss, numerator, denominator}; the backward reads it back as sums3 and WRITES grad_im1 and/or grad_im2 [B,C,HW]
 * (grad_im2 needs im2_batch_stride == C*HW) and/or grad_sigma [B,sigma_channels,HW]; grad_loss is a DEVICE scalar. */
int g2s_photometric_fwd(const float *im1, const float *im2, long im2_batch_stride, const float *recon_depth,
                        float depth_thresh, const float *mask_in, const float *conf_sigma, int sigma_channels, int B,
                        int C, int HW, void *reduce_ws, float *out3, void *stream);
int g2s_photometric_bwd(const float *im1, const float *im2, long im2_batch_stride, const float *recon_depth,
                        float depth_thresh, const float *mask_in, const float *conf_sigma, int sigma_channels, int B,
                        int C, int HW, const float *sums3, const float *grad_loss, float *grad_im1, float *grad_im2,
                        float *grad_sigma, void *stream);

/* SmoothLoss of ONE map [M,H,W]: losses.py:54-79 (mean|dx2| + mean|dxdy| + mean|dydx| + mean|dy2|, the differences taken
 * in the reference's order).  out5 = {loss, the four means}.  H, W >= 3.  Backward WRITES grad_map [M,H,W]; grad_loss is a
 * DEVICE scalar. */
int g2s_smooth_fwd(const float *map, int M, int H, int W, void *reduce_ws, float *out5, void *stream);
int g2s_smooth_bwd(const float *map, int M, int H, int W, const float *grad_loss, float *grad_map, void *stream);

/* ---- instrumentation (bench.py / tests; not part of the reference surface) ---------------------------
 * g2s_launch_count: kernels launched by this library since it was loaded.
 * g2s_profile_enable(1): record a CUDA event pair around every kernel launch on its stream;
 * g2s_profile_read: waits for the recorded events, returns the number of distinct kernels n (<= max_kernels)
 * and fills names[n], total_ms[n], launches[n]; clears the records. */
long g2s_launch_count(void);
/* g2s_selftest_division: compares the kernels' shared-reciprocal division with IEEE __fdiv_rn on n_pairs pseudo-random
 * operand pairs and ADDS the number of mismatching results to *mismatches_dev (device, caller zero-fills). */
int g2s_selftest_division(unsigned long long n_pairs, unsigned seed, unsigned long long *mismatches_dev, void *stream);
/* g2s_selftest_face_vertices: the kernels' face-index -> vertex-index map (division by S-1 through a reciprocal estimate
 * and one fix-up) against plain integer arithmetic for every face of an image_size^2 grid mesh; ADDS mismatches. */
int g2s_selftest_face_vertices(int image_size, unsigned long long *mismatches_dev, void *stream);
/* g2s_selftest_index_math: the kernels' cheap index arithmetic -- v / S, v % S by a float estimate + fix-up for every vertex of
 * an S x S mesh, view / views_per_image as a multiply-high for every view of a call of n_views views -- against integer
 * division; the mismatch count is added to *mismatches_dev. */
int g2s_selftest_index_math(int image_size, int views_per_image, long n_views, unsigned long long *mismatches_dev,
                            void *stream);
/* g2s_selftest_raster: compares the rasteriser's fast per-face / per-hit arithmetic (shared reciprocals, structural
 * operand-range guards) with the plain IEEE formulation bit for bit on n_triangles pseudo-random triangles (ordinary,
 * degenerate and extreme-magnitude ones, 4 sub-pixels each) and ADDS the number of mismatches to *mismatches_dev. */
int g2s_selftest_raster(unsigned long long n_triangles, unsigned seed, int image_size, unsigned long long *mismatches_dev,
                        void *stream);
int g2s_profile_enable(int on);
int g2s_profile_read(int max_kernels, const char **names, float *total_ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* G2S_B200_H */
