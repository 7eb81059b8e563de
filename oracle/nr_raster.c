/*
 * oracle/nr_raster.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, OpenMP) of the CUDA kernels of the third-party package
 * `neural_renderer` (daniilidis-group fork, https://github.com/daniilidis-group/neural_renderer,
 * UNPINNED HEAD in the reference: /root/reference/README.md:32-37).  The package is not in
 * /root/reference, not installed and not fetchable, so this file restates the published algorithm of
 * neural_renderer/cuda/rasterize_cuda_kernel.cu from knowledge of the public source.
 * Reference call sites it serves: GAN2Shape/renderer/renderer.py:120 (render_depth) and
 * renderer.py:196, 230, 248, 272, 275 (render_rgb).
 *
 * PARITY UNPINNED at this boundary: the reference holds no test, golden image or known-answer vector
 * for the rasteriser.  What pins this file instead: the analytic known-answer tests
 * in tests/test_oracle_known_answers.py (identity view + flat depth, closed-form face-index map,
 * tie rule), finite-difference checks of d(zp)/d(z), and tests/test_raycast_known_answers.py: an independent
 * float64 3-D ray caster (tilted planes, occlusion, fill_back, near / far, rgb blend) and the backward_depth_map
 * formula evaluated in float64.
 *
 * Arithmetic contract: the C expressions below keep the source's literal types (`0.5`, `2.`, `1.`,
 * `0.` are double literals in the CUDA source, everything else is float), source evaluation order, and
 * NO fused multiply-add (compile with -ffp-contract=off; the historical nvcc build's FMA contraction
 * is unknowable).  The CUDA product kernels reproduce exactly this arithmetic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* [nr] forward_face_index_map_cuda_kernel_1: per face, back-face cull and the 3x3 inverse used for
 * barycentric weights.  `face` = 9 floats (v0 xyz, v1 xyz, v2 xyz) in NDC; back faces keep zeros. */
static void face_inv_setup(const float *face, int is, float *face_inv_g)
{
    if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
        return;
    float p[3][2];
    for (int num = 0; num < 3; num++)
        for (int dim = 0; dim < 2; dim++)
            p[num][dim] = 0.5 * (face[3 * num + dim] * is + is - 1);
    float face_inv[9] = {
        p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
        p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
        p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
    float face_inv_denominator = (p[2][0] * (p[0][1] - p[1][1]) +
                                  p[0][0] * (p[1][1] - p[2][1]) +
                                  p[1][0] * (p[2][1] - p[0][1]));
    for (int k = 0; k < 9; k++)
        face_inv[k] /= face_inv_denominator;
    for (int k = 0; k < 9; k++)
        face_inv_g[k] = face_inv[k];
}

/* Body of the face loop of [nr] forward_face_index_map_cuda_kernel_2 for ONE (sub-pixel, face) pair.
 * Returns 1 when the face is a z-buffer candidate at this sub-pixel (front-facing, inside or on an
 * edge, near < zp < far) and then writes zp and the clamped, renormalised weights. */
static inline int eval_pixel_face(const float *face, const float *face_inv, int xi, int yi, float xp,
                                  float yp, float near, float far, float *zp_out, float *w)
{
    if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
        return 0;
    if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
        ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
        ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
        return 0;
    w[0] = face_inv[3 * 0 + 0] * xi + face_inv[3 * 0 + 1] * yi + face_inv[3 * 0 + 2];
    w[1] = face_inv[3 * 1 + 0] * xi + face_inv[3 * 1 + 1] * yi + face_inv[3 * 1 + 2];
    w[2] = face_inv[3 * 2 + 0] * xi + face_inv[3 * 2 + 1] * yi + face_inv[3 * 2 + 2];
    float w_sum = 0;
    for (int k = 0; k < 3; k++) {
        /* device min/max on (float, double) are fmin/fmax: a NaN operand yields the other one */
        w[k] = fmin(fmax(w[k], 0.), 1.);
        w_sum += w[k];
    }
    for (int k = 0; k < 3; k++)
        w[k] /= w_sum;
    const float zp = 1. / (w[0] / face[2] + w[1] / face[5] + w[2] / face[8]);
    if (zp <= near || far <= zp)
        return 0;
    *zp_out = zp;
    return 1;
}

static void prefill(long n, int32_t *face_index_map, float *weight_map, float *depth_map,
                    float *face_inv_map, float far)
{
    /* Python-side pre-fill of [nr] RasterizeFunction.forward */
    for (long i = 0; i < n; i++) {
        face_index_map[i] = -1;
        depth_map[i] = far;
    }
    memset(weight_map, 0, sizeof(float) * 3 * (size_t)n);
    if (face_inv_map)
        memset(face_inv_map, 0, sizeof(float) * 9 * (size_t)n);
}

/*
 * Faithful O(is^2 * nf) rasteriser: [nr] forward_face_index_map (kernel_1 + kernel_2).
 * faces [B, nf, 9]; outputs in nr's native orientation (row yi counts UPWARDS, before the flip):
 * face_index_map [B,is,is] i32, weight_map [B,is,is,3], depth_map [B,is,is], face_inv_map [B,is,is,9]
 * (may be NULL).  stats (may be NULL): [0] += number of sub-pixels where >1 face attains the winning
 * depth (exact depth ties), [1] += candidate hits whose sub-pixel lies outside the face's pixel-space
 * bounding box grown by 1/64 px (the candidate set the product's splat relies on).
 */
EXPORT void nr_forward_face_index_map(const float *faces, int batch_size, int num_faces,
                                      int image_size, float near, float far,
                                      int32_t *face_index_map, float *weight_map, float *depth_map,
                                      float *face_inv_map, int64_t *stats)
{
    const int is = image_size, nf = num_faces;
    const long npix = (long)batch_size * is * is;
    float *faces_inv = (float *)calloc((size_t)batch_size * nf * 9, sizeof(float));
    prefill(npix, face_index_map, weight_map, depth_map, face_inv_map, far);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)batch_size * nf; i++)
        face_inv_setup(&faces[i * 9], is, &faces_inv[i * 9]);

    int64_t n_ties = 0, n_outside = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_ties, n_outside)
    for (long i = 0; i < npix; i++) {
        const int bn = i / (is * is);
        const int pn = i % (is * is);
        const int yi = pn / is;
        const int xi = pn % is;
        const float yp = (2. * yi + 1 - is) / is;
        const float xp = (2. * xi + 1 - is) / is;
        const float *face = &faces[(long)bn * nf * 9] - 9;
        const float *face_inv = &faces_inv[(long)bn * nf * 9] - 9;
        float depth_min = far;
        int face_index_min = -1;
        int n_at_min = 0;
        float weight_min[3];
        float face_inv_min[9];
        for (int fn = 0; fn < nf; fn++) {
            face += 9;
            face_inv += 9;
            float zp, w[3];
            if (!eval_pixel_face(face, face_inv, xi, yi, xp, yp, near, far, &zp, w))
                continue;
            if (stats) {
                float px[3], py[3];
                for (int k = 0; k < 3; k++) {
                    px[k] = 0.5 * (face[3 * k] * is + is - 1);
                    py[k] = 0.5 * (face[3 * k + 1] * is + is - 1);
                }
                const float m = 1.0f / 64;
                if (xi < fminf(px[0], fminf(px[1], px[2])) - m ||
                    xi > fmaxf(px[0], fmaxf(px[1], px[2])) + m ||
                    yi < fminf(py[0], fminf(py[1], py[2])) - m ||
                    yi > fmaxf(py[0], fmaxf(py[1], py[2])) + m)
                    n_outside++;
                if (zp == depth_min && face_index_min >= 0)
                    n_at_min++;
            }
            if (zp < depth_min) {
                depth_min = zp;
                face_index_min = fn;
                n_at_min = 1;
                for (int k = 0; k < 3; k++)
                    weight_min[k] = w[k];
                for (int k = 0; k < 9; k++)
                    face_inv_min[k] = face_inv[k];
            }
        }
        if (0 <= face_index_min) {
            depth_map[i] = depth_min;
            face_index_map[i] = face_index_min;
            for (int k = 0; k < 3; k++)
                weight_map[3 * i + k] = weight_min[k];
            if (face_inv_map)
                for (int k = 0; k < 9; k++)
                    face_inv_map[9 * i + k] = face_inv_min[k];
            if (n_at_min > 1)
                n_ties++;
        }
    }
    if (stats) {
        stats[0] += n_ties;
        stats[1] += n_outside;
    }
    free(faces_inv);
}

/*
 * Bounding-box-culled variant: identical per-(sub-pixel, face) arithmetic and the identical winner
 * rule (smallest zp, then smallest face index), but each face only visits the sub-pixels of its
 * pixel-space bounding box grown by 1 px.  LABELLED "culled": it is NOT the reference's loop; it
 * exists so that parity tests at many views finish in seconds.  tests/test_oracle_raster.py checks
 * it bit-for-bit against nr_forward_face_index_map.
 */
EXPORT void nr_forward_face_index_map_culled(const float *faces, int batch_size, int num_faces,
                                             int image_size, float near, float far,
                                             int32_t *face_index_map, float *weight_map,
                                             float *depth_map, float *face_inv_map)
{
    const int is = image_size, nf = num_faces;
    const long npix = (long)batch_size * is * is;
    prefill(npix, face_index_map, weight_map, depth_map, face_inv_map, far);
    /* parallel over (view, band of rows): bands own disjoint sub-pixels, so the ascending-fn order
     * per sub-pixel is preserved */
    const int band_rows = 16;
    const int nbands = (is + band_rows - 1) / band_rows;
#pragma omp parallel for schedule(dynamic, 1)
    for (int job = 0; job < batch_size * nbands; job++) {
        const int bn = job / nbands;
        const int band_y0 = (job % nbands) * band_rows;
        const int band_y1 = band_y0 + band_rows - 1 < is - 1 ? band_y0 + band_rows - 1 : is - 1;
        int32_t *fim = face_index_map + (long)bn * is * is;
        float *dm = depth_map + (long)bn * is * is;
        float *wm = weight_map + (long)bn * is * is * 3;
        float *fvm = face_inv_map ? face_inv_map + (long)bn * is * is * 9 : NULL;
        for (int fn = 0; fn < nf; fn++) {
            const float *face = &faces[((long)bn * nf + fn) * 9];
            float pxmin = INFINITY, pxmax = -INFINITY, pymin = INFINITY, pymax = -INFINITY;
            int bad = 0;
            for (int k = 0; k < 3; k++) {
                const float px = 0.5 * (face[3 * k] * is + is - 1);
                const float py = 0.5 * (face[3 * k + 1] * is + is - 1);
                if (!(px == px) || !(py == py))
                    bad = 1;
                pxmin = fminf(pxmin, px); pxmax = fmaxf(pxmax, px);
                pymin = fminf(pymin, py); pymax = fmaxf(pymax, py);
            }
            int x0 = 0, x1 = is - 1, y0 = band_y0, y1 = band_y1;
            if (!bad) {
                if (pxmax < -2 || pymax < band_y0 - 2 || pxmin > is + 1 || pymin > band_y1 + 2)
                    continue;
                const float lo_x = floorf(pxmin) - 1, hi_x = ceilf(pxmax) + 1;
                const float lo_y = floorf(pymin) - 1, hi_y = ceilf(pymax) + 1;
                if (lo_x > x0) x0 = (int)lo_x;
                if (hi_x < x1) x1 = (int)hi_x;
                if (lo_y > y0) y0 = (int)lo_y;
                if (hi_y < y1) y1 = (int)hi_y;
            }
            float face_inv[9] = {0};
            face_inv_setup(face, is, face_inv);
            for (int yi = y0; yi <= y1; yi++) {
                const float yp = (2. * yi + 1 - is) / is;
                for (int xi = x0; xi <= x1; xi++) {
                    const float xp = (2. * xi + 1 - is) / is;
                    float zp, w[3];
                    if (!eval_pixel_face(face, face_inv, xi, yi, xp, yp, near, far, &zp, w))
                        continue;
                    const long i = (long)yi * is + xi;
                    /* ascending fn + strict '<' == the reference's loop order */
                    if (zp < dm[i]) {
                        dm[i] = zp;
                        fim[i] = fn;
                        for (int k = 0; k < 3; k++)
                            wm[3 * i + k] = w[k];
                        if (fvm)
                            for (int k = 0; k < 9; k++)
                                fvm[9 * i + k] = face_inv[k];
                    }
                }
            }
        }
    }
}

/* [nr] backward_depth_map_cuda_kernel; grad_faces [B,nf,9] is accumulated in sub-pixel order
 * (the CUDA original uses atomicAdd, i.e. an unspecified order). */
EXPORT void nr_backward_depth_map(const float *faces, const float *depth_map,
                                  const int32_t *face_index_map, const float *face_inv_map,
                                  const float *weight_map, const float *grad_depth_map,
                                  float *grad_faces, int batch_size, int num_faces, int image_size)
{
    const int is = image_size, nf = num_faces;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bn = 0; bn < batch_size; bn++) {
        for (long pn = 0; pn < (long)is * is; pn++) {
            const long i = (long)bn * is * is + pn;
            const int fn = face_index_map[i];
            if (0 <= fn) {
                const float *face = &faces[((long)bn * nf + fn) * 9];
                const float depth = depth_map[i];
                const float depth2 = depth * depth;
                const float *face_inv = &face_inv_map[i * 9];
                const float *weight = &weight_map[i * 3];
                const float grad_depth = grad_depth_map[i];
                float *grad_face = &grad_faces[((long)bn * nf + fn) * 9];
                /* derivative wrt z */
                for (int k = 0; k < 3; k++) {
                    const float z_k = face[3 * k + 2];
                    grad_face[3 * k + 2] += grad_depth * weight[k] * depth2 / (z_k * z_k);
                }
                /* derivative wrt x, y */
                float tmp[3] = {0, 0, 0};
                for (int k = 0; k < 3; k++)
                    for (int l = 0; l < 3; l++)
                        tmp[k] += -face_inv[3 * l + k] / face[3 * l + 2];
                for (int k = 0; k < 3; k++)
                    for (int l = 0; l < 2; l++)
                        grad_face[3 * k + l] += -grad_depth * tmp[l] * weight[k] * depth2 * is / 2;
            }
        }
    }
}

/* [nr] forward_texture_sampling_cuda_kernel + forward_background + forward_alpha_map.
 * textures [B,nf,ts,ts,ts,3]; rgb_map [B,is,is,3] (pre-filled 0 then background where empty),
 * sampling_index_map i32 [B,is,is,8], sampling_weight_map [B,is,is,8], alpha_map [B,is,is]. */
EXPORT void nr_forward_texture_sampling(const float *faces, const float *textures,
                                        const int32_t *face_index_map, const float *weight_map,
                                        const float *depth_map, float *rgb_map,
                                        int32_t *sampling_index_map, float *sampling_weight_map,
                                        float *alpha_map, const float *background_color,
                                        int batch_size, int num_faces, int image_size,
                                        int texture_size, float eps)
{
    const int is = image_size, nf = num_faces, ts = texture_size;
    const long npix = (long)batch_size * is * is;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < npix; i++) {
        float *pixel = &rgb_map[i * 3];
        int32_t *sampling_indices = &sampling_index_map[i * 8];
        float *sampling_weights = &sampling_weight_map[i * 8];
        for (int k = 0; k < 8; k++) {
            sampling_indices[k] = 0;
            sampling_weights[k] = 0;
        }
        const int face_index = face_index_map[i];
        if (face_index >= 0) {
            const int bn = i / ((long)is * is);
            const float *face = &faces[((long)bn * nf + face_index) * 9];
            const float *texture = &textures[((long)bn * nf + face_index) * ts * ts * ts * 3];
            const float *weight = &weight_map[i * 3];
            const float depth = depth_map[i];
            float texture_index_float[3];
            for (int k = 0; k < 3; k++) {
                float tif = weight[k] * (ts - 1) * (depth / (face[3 * k + 2]));
                tif = fmax(tif, 0.);
                tif = fmin(tif, ts - 1 - eps);
                texture_index_float[k] = tif;
            }
            float new_pixel[3] = {0, 0, 0};
            for (int pn = 0; pn < 8; pn++) {
                float w = 1;
                int texture_index_int[3];
                for (int k = 0; k < 3; k++) {
                    if ((pn >> k) % 2 == 0) {
                        w *= 1 - (texture_index_float[k] - (int)texture_index_float[k]);
                        texture_index_int[k] = (int)texture_index_float[k];
                    } else {
                        w *= texture_index_float[k] - (int)texture_index_float[k];
                        texture_index_int[k] = (int)texture_index_float[k] + 1;
                    }
                }
                const int isc = texture_index_int[0] * ts * ts + texture_index_int[1] * ts +
                                texture_index_int[2];
                for (int k = 0; k < 3; k++)
                    new_pixel[k] += w * texture[isc * 3 + k];
                sampling_indices[pn] = isc;
                sampling_weights[pn] = w;
            }
            for (int k = 0; k < 3; k++)
                pixel[k] = new_pixel[k];
            alpha_map[i] = 1;
        } else {
            for (int k = 0; k < 3; k++)
                pixel[k] = background_color[k];
            alpha_map[i] = 0;
        }
    }
}

/* [nr] backward_textures_cuda_kernel (sub-pixel order instead of atomicAdd order). */
EXPORT void nr_backward_textures(const int32_t *face_index_map, const float *sampling_weight_map,
                                 const int32_t *sampling_index_map, const float *grad_rgb_map,
                                 float *grad_textures, int batch_size, int num_faces,
                                 int image_size, int texture_size)
{
    const int is = image_size, nf = num_faces, ts = texture_size;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bn = 0; bn < batch_size; bn++) {
        for (long pn_ = 0; pn_ < (long)is * is; pn_++) {
            const long i = (long)bn * is * is + pn_;
            const int face_index = face_index_map[i];
            if (0 <= face_index) {
                float *grad_texture = &grad_textures[((long)bn * nf + face_index) * ts * ts * ts * 3];
                for (int pn = 0; pn < 8; pn++) {
                    const float w = sampling_weight_map[i * 8 + pn];
                    const int isc = sampling_index_map[i * 8 + pn];
                    for (int k = 0; k < 3; k++)
                        grad_texture[isc * 3 + k] += w * grad_rgb_map[i * 3 + k];
                }
            }
        }
    }
}


/*
 * [nr] backward_pixel_map_cuda_kernel: the approximate silhouette / colour gradient of Kato et al. with respect to the x, y
 * of the projected vertices.  One "thread" per face: for each edge, for each axis, walk the integer
 * positions d0 between the edge's end points, find the crossing d1_cross and the pixel just inside / outside; the "out" pass
 * (only when the inside pixel belongs to this face) walks from the outside pixel to the image border, the "in" pass walks
 * from the inside pixel to the opposite edge over the pixels that belong to this face; a visited pixel whose colour
 * difference to the pixel across the edge correlates positively with the incoming gradient pulls the edge's two vertices
 * by diff_grad / dist.  grad_faces [B,nf,9] is WRITTEN for front faces (x, y components; z stays 0), as the original does.
 * Maps in nr's native orientation (row yi counts upwards).  Restated from knowledge of the public source: PARITY UNPINNED
 * (the reference never differentiates render_rgb: SURVEY.md 8a').
 */
EXPORT void nr_backward_pixel_map(const float *faces, const int32_t *face_index_map, const float *rgb_map,
                                  const float *alpha_map, const float *grad_rgb_map, const float *grad_alpha_map,
                                  float *grad_faces, int batch_size, int num_faces, int image_size, float eps,
                                  int return_rgb, int return_alpha)
{
    const int is = image_size;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)batch_size * num_faces; i++) {
        const int bn = i / num_faces;
        const int fn = i % num_faces;
        const float *face = &faces[i * 9];
        float grad_face[9] = {0};

        /* check backside */
        if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
            continue;

        /* for each edge */
        for (int edge_num = 0; edge_num < 3; edge_num++) {
            int pi[3];
            float pp[3][2];
            for (int num = 0; num < 3; num++)
                pi[num] = (edge_num + num) % 3;
            for (int num = 0; num < 3; num++)
                for (int dim = 0; dim < 2; dim++)
                    pp[num][dim] = 0.5 * (face[3 * pi[num] + dim] * is + is - 1);

            /* for dy, dx */
            for (int axis = 0; axis < 2; axis++) {
                float p[3][2];
                for (int num = 0; num < 3; num++)
                    for (int dim = 0; dim < 2; dim++)
                        p[num][dim] = pp[num][(dim + axis) % 2];

                /* set direction */
                int direction;
                if (axis == 0)
                    direction = (p[0][0] < p[1][0]) ? -1 : 1;
                else
                    direction = (p[0][0] < p[1][0]) ? 1 : -1;

                /* along edge */
                const int d0_from = (int)fmax(ceilf(fminf(p[0][0], p[1][0])), 0.);
                const int d0_to = (int)fmin(fmaxf(p[0][0], p[1][0]), is - 1.);
                for (int d0 = d0_from; d0 <= d0_to; d0++) {
                    /* get cross point */
                    int d1_in, d1_out;
                    const float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                    if (0 < direction)
                        d1_in = (int)floorf(d1_cross);
                    else
                        d1_in = (int)ceilf(d1_cross);
                    d1_out = d1_in + direction;

                    /* continue if cross point is not shown */
                    if (d1_in < 0 || is <= d1_in)
                        continue;
                    if (d1_out < 0 || is <= d1_out)
                        continue;

                    /* get color of in-pixel and out-pixel */
                    float alpha_in = 0, alpha_out = 0;
                    const float *rgb_in = NULL, *rgb_out = NULL;
                    long map_index_in, map_index_out;
                    if (axis == 0) {
                        map_index_in = (long)bn * is * is + (long)d1_in * is + d0;
                        map_index_out = (long)bn * is * is + (long)d1_out * is + d0;
                    } else {
                        map_index_in = (long)bn * is * is + (long)d0 * is + d1_in;
                        map_index_out = (long)bn * is * is + (long)d0 * is + d1_out;
                    }
                    if (return_alpha) {
                        alpha_in = alpha_map[map_index_in];
                        alpha_out = alpha_map[map_index_out];
                    }
                    if (return_rgb) {
                        rgb_in = &rgb_map[map_index_in * 3];
                        rgb_out = &rgb_map[map_index_out * 3];
                    }

                    /* out */
                    const int is_in_fn = (face_index_map[map_index_in] == fn);
                    if (is_in_fn) {
                        const int d1_limit = (0 < direction) ? is - 1 : 0;
                        const int d1_from = d1_out < d1_limit ? (d1_out > 0 ? d1_out : 0) : (d1_limit > 0 ? d1_limit : 0);
                        const int d1_to_ = d1_out > d1_limit ? d1_out : d1_limit;
                        const int d1_to = d1_to_ < is - 1 ? d1_to_ : is - 1;
                        const long map_offset = (axis == 0) ? is : 1;
                        long idx = (axis == 0) ? (long)bn * is * is + (long)d1_from * is + d0
                                               : (long)bn * is * is + (long)d0 * is + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, idx += map_offset) {
                            float diff_grad = 0;
                            if (return_alpha)
                                diff_grad += (alpha_map[idx] - alpha_in) * grad_alpha_map[idx];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[idx * 3 + k] - rgb_in[k]) * grad_rgb_map[idx * 3 + k];
                            if (diff_grad <= 0)
                                continue;
                            if (p[1][0] != d0) {
                                float dist = (p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2. / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0) {
                                float dist = (p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2. / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }

                    /* in */
                    {
                        int d1_limit;
                        float d0_cross2;
                        if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0)
                            d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                        else
                            d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
                        if (0 < direction)
                            d1_limit = (int)ceilf(d0_cross2);
                        else
                            d1_limit = (int)floorf(d0_cross2);
                        const int lo = d1_in < d1_limit ? d1_in : d1_limit, hi = d1_in > d1_limit ? d1_in : d1_limit;
                        const int d1_from = lo > 0 ? lo : 0;
                        const int d1_to = hi < is - 1 ? hi : is - 1;
                        const long map_offset = (axis == 0) ? is : 1;
                        long idx = (axis == 0) ? (long)bn * is * is + (long)d1_from * is + d0
                                               : (long)bn * is * is + (long)d0 * is + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, idx += map_offset) {
                            if (face_index_map[idx] != fn)
                                continue;
                            float diff_grad = 0;
                            if (return_alpha)
                                diff_grad += (alpha_map[idx] - alpha_out) * grad_alpha_map[idx];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[idx * 3 + k] - rgb_out[k]) * grad_rgb_map[idx * 3 + k];
                            if (diff_grad <= 0)
                                continue;
                            if (p[1][0] != d0) {
                                float dist = (p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2. / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0) {
                                float dist = (p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2. / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                }
            }
        }
        for (int k = 0; k < 9; k++)
            grad_faces[i * 9 + k] = grad_face[k];
    }
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* bench.py pins the thread count explicitly (torchrun exports OMP_NUM_THREADS=1) */
EXPORT int nr_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

EXPORT int nr_oracle_version(void) { return 1; }
