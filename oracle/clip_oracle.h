/*
 * oracle/clip_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic on clip.cpp's encode hot path (see clip_oracle.c for the
 * reference file:line each function follows).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product library
 * (clip.cpp_b200/libclip_b200.so) never links or calls it.
 */
#ifndef CLIP_ORACLE_H
#define CLIP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml type ids used by clip.cpp model files (ggml.h enum ggml_type) */
enum { ORC_F32 = 0, ORC_F16 = 1, ORC_Q4_0 = 2, ORC_Q4_1 = 3, ORC_Q5_0 = 6, ORC_Q5_1 = 7, ORC_Q8_0 = 8 };

size_t orc_row_size(int type, int64_t k);                              /* bytes of one row of k elements   */
int    orc_quantize_row(int type, const float *x, void *y, int64_t k); /* weight quantizer (reference rows) */
int    orc_dequantize_row(int type, const void *x, float *y, int64_t k);

/* Y[m, n] = X[m, k] . W[n, k]^T with the reference's activation rounding for W's type */
int    orc_mul_mat(int type, const void *W, int64_t n, int64_t k, const float *X, int64_t m, float *Y, int n_threads);

void   orc_layer_norm(const float *x, float *y, int64_t rows, int64_t h, float eps);          /* no affine */
void   orc_gelu(const float *x, float *y, int64_t n, int quick);                              /* fp16 LUT semantics */
void   orc_softmax_rows(float *x, int64_t rows, int64_t cols);                                /* in place */
/* full attention for one sequence: q,k,v are [T, H*dh] row-major f32 (q already scaled); out [T, H*dh] */
void   orc_attention(const float *q, const float *k, const float *v, float *out, int T, int H, int dh, int causal, int n_threads);
void   orc_f32_to_f16_to_f32(const float *x, float *y, int64_t n);                            /* round trip */
float  orc_sum_sq_sqrt(const float *x, int64_t n);                                            /* sqrt(sum x^2), double accum */

#ifdef __cplusplus
}
#endif
#endif
