/*
 * oracle/fma_mm.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference writes its 3-wide contractions as torch.matmul (GAN2Shape/renderer/renderer.py:67, 79,
 * 85; utils.py:49, 95; neural_renderer projection).  Run on the CPU in the build container (torch 2.11 +
 * oneMKL 2024.2) every one of those evaluates, bit for bit, as the fused chain
 *     acc = a0*b0;  acc = fma(a1, b1, acc);  acc = fma(a2, b2, acc)
 * (measured: 100 % of 12 288 random samples; un-fused left-to-right matches only 65 %).  This file
 * states that chain explicitly with fmaf() so the oracle does not depend on the host BLAS;
 * tests/test_oracle_vs_reference.py pins it against the reference's unmodified code.
 */
#include <math.h>
#define EXPORT __attribute__((visibility("default")))

/* out[b,n,j] = sum_k v[b,n,k] * M[b or 0, j, k]   (v @ M^T), K = 3 */
EXPORT void fma_mm3_nt(const float *v, const float *M, float *out, long B, long N, int m_batched)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < B * N; i++) {
        const float *m = M + (m_batched ? (i / N) * 9 : 0);
        const float *a = v + i * 3;
        for (int j = 0; j < 3; j++) {
            float acc = a[0] * m[3 * j + 0];
            acc = fmaf(a[1], m[3 * j + 1], acc);
            acc = fmaf(a[2], m[3 * j + 2], acc);
            out[i * 3 + j] = acc;
        }
    }
}

/* C[b] = A[b] @ B[b], row-major [R,3] x [3,C] with the same chain; used for Rz@(Ry@Rx) and the
 * texture-cube coefficients. */
EXPORT void fma_mm_k3(const float *A, const float *Bm, float *C, long batch, int R, int Cc,
                      int a_batched, int b_batched)
{
    for (long b = 0; b < batch; b++) {
        const float *a = A + (a_batched ? b * R * 3 : 0);
        const float *bm = Bm + (b_batched ? b * 3 * Cc : 0);
        float *c = C + b * R * Cc;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < Cc; j++) {
                float acc = a[i * 3 + 0] * bm[0 * Cc + j];
                acc = fmaf(a[i * 3 + 1], bm[1 * Cc + j], acc);
                acc = fmaf(a[i * 3 + 2], bm[2 * Cc + j], acc);
                c[i * Cc + j] = acc;
            }
    }
}
