/*
 * oracle/clip_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See clip_oracle.h.
 *
 * A plain-C restatement of the arithmetic the reference (monatis/clip.cpp @3484ffc + ggml @c3ae31e)
 * performs on the image/text encode path, INCLUDING its rounding points, so that it can be pinned
 * against the reference itself (oracle/_ref, see tests/test_oracle_pin.py) before it is trusted as
 * the checker for the CUDA path.  Nothing here is copied from the reference; each function cites
 * the reference lines whose behaviour it restates.  All paths are relative to /root/reference.
 *
 * Rounding points reproduced:
 *   - weight GEMM inputs: fp16 (f16 files) or q8_0 / q8_1 blocks (q* files), x86-AVX2 flavour:
 *     round-to-nearest-even, id = 127/amax                           (ggml/src/ggml.c:1188-1277, 1398-1494)
 *   - block dot = (int32 sum of q_w*q_x) * d_w * d_x in fp32          (ggml/src/ggml.c:2406-2707, 2783-2864,
 *                                                                      3029-3176, 3349-3497, 3564-3621)
 *   - softmax exp and GELU through fp16 in / fp16 out tables          (ggml/src/ggml.c:3752-3815, 4525-4540, 12201-12270)
 *   - LayerNorm / sums with double accumulators                       (ggml/src/ggml.c:10796-10845)
 * Summation ORDER inside a dot product differs from the SIMD reference (8-lane partial sums), which
 * is why the pin test uses a tight tolerance (1 - cos <= 1e-6) instead of bit equality.
 */
#include "clip_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16_t;                      /* IEEE binary16, conversions round-to-nearest-even */
static inline float h2f(uint16_t h) { f16_t v; memcpy(&v, &h, 2); return (float)v; }
static inline uint16_t f2h(float f) { f16_t v = (f16_t)f; uint16_t h; memcpy(&h, &v, 2); return h; }

#define QK 32
/* block sizes: ggml/src/ggml.c:866-911 */
static size_t block_bytes(int type) {
    switch (type) {
    case ORC_Q4_0: return 2 + 16;
    case ORC_Q4_1: return 4 + 16;
    case ORC_Q5_0: return 2 + 4 + 16;
    case ORC_Q5_1: return 4 + 4 + 16;
    case ORC_Q8_0: return 2 + 32;
    default: return 0;
    }
}

size_t orc_row_size(int type, int64_t k) {
    if (type == ORC_F32) return (size_t)k * 4;
    if (type == ORC_F16) return (size_t)k * 2;
    return (size_t)(k / QK) * block_bytes(type);
}

static inline int imin(int a, int b) { return a < b ? a : b; }

/* ---- weight quantizers: the "_reference" rows clip_model_quantize always uses ---------------------
 * ggml/src/ggml.c:914-1116 via ggml_quantize_q* (ggml.c:19382-19403), called from clip.cpp:1771-1786 */
int orc_quantize_row(int type, const float *x, void *vy, int64_t k) {
    if (k % QK) return -1;
    const int64_t nb = k / QK;
    uint8_t *y = (uint8_t *)vy;
    const size_t bs = block_bytes(type);
    if (!bs) return -1;
    for (int64_t i = 0; i < nb; i++) {
        const float *xb = x + i * QK;
        uint8_t *b = y + i * bs;
        if (type == ORC_Q4_0 || type == ORC_Q5_0) {
            /* signed absmax; d = max / -(2^(bits-1)) */
            float amax = 0.0f, max = 0.0f;
            for (int j = 0; j < QK; j++) { float v = xb[j]; if (amax < fabsf(v)) { amax = fabsf(v); max = v; } }
            const int half = (type == ORC_Q4_0) ? 8 : 16;
            const float d = max / (float)(-half);
            const float id = d ? 1.0f / d : 0.0f;
            const float off = (float)half + 0.5f;          /* 8.5f / 16.5f: ONE rounding in x*id + off */
            uint16_t dh = f2h(d); memcpy(b, &dh, 2);
            uint8_t *qs = b + (type == ORC_Q4_0 ? 2 : 6);
            uint32_t qh = 0;
            for (int j = 0; j < QK / 2; j++) {
                const float x0 = xb[j] * id, x1 = xb[QK / 2 + j] * id;
                const uint8_t q0 = (uint8_t)imin(2 * half - 1, (int8_t)(x0 + off));
                const uint8_t q1 = (uint8_t)imin(2 * half - 1, (int8_t)(x1 + off));
                qs[j] = (uint8_t)((q0 & 0x0F) | ((q1 & 0x0F) << 4));
                if (type == ORC_Q5_0) {
                    qh |= (uint32_t)((q0 & 0x10) >> 4) << (j + 0);
                    qh |= (uint32_t)((q1 & 0x10) >> 4) << (j + QK / 2);
                }
            }
            if (type == ORC_Q5_0) memcpy(b + 2, &qh, 4);
        } else if (type == ORC_Q4_1 || type == ORC_Q5_1) {
            float mn = FLT_MAX, mx = -FLT_MAX;
            for (int j = 0; j < QK; j++) { float v = xb[j]; if (v < mn) mn = v; if (v > mx) mx = v; }
            const int levels = (type == ORC_Q4_1) ? 15 : 31;
            const float d = (mx - mn) / (float)levels;
            const float id = d ? 1.0f / d : 0.0f;
            uint16_t dh = f2h(d), mh = f2h(mn); memcpy(b, &dh, 2); memcpy(b + 2, &mh, 2);
            uint8_t *qs = b + (type == ORC_Q4_1 ? 4 : 8);
            uint32_t qh = 0;
            for (int j = 0; j < QK / 2; j++) {
                const float x0 = (xb[j] - mn) * id, x1 = (xb[QK / 2 + j] - mn) * id;
                uint8_t q0, q1;
                if (type == ORC_Q4_1) {
                    q0 = (uint8_t)imin(15, (int8_t)(x0 + 0.5f));
                    q1 = (uint8_t)imin(15, (int8_t)(x1 + 0.5f));
                } else {                       /* q5_1 casts straight to uint8_t, no clamp (ggml.c:1074-1075) */
                    q0 = (uint8_t)(x0 + 0.5f);
                    q1 = (uint8_t)(x1 + 0.5f);
                }
                qs[j] = (uint8_t)((q0 & 0x0F) | ((q1 & 0x0F) << 4));
                if (type == ORC_Q5_1) {
                    qh |= (uint32_t)((q0 & 0x10) >> 4) << (j + 0);
                    qh |= (uint32_t)((q1 & 0x10) >> 4) << (j + QK / 2);
                }
            }
            if (type == ORC_Q5_1) memcpy(b + 4, &qh, 4);
        } else { /* ORC_Q8_0: d = amax/127, roundf (ties away) -- the file quantizer, ggml.c:1097-1114 */
            float amax = 0.0f;
            for (int j = 0; j < QK; j++) { float v = fabsf(xb[j]); if (v > amax) amax = v; }
            const float d = amax / 127.0f;
            const float id = d ? 1.0f / d : 0.0f;
            uint16_t dh = f2h(d); memcpy(b, &dh, 2);
            int8_t *qs = (int8_t *)(b + 2);
            for (int j = 0; j < QK; j++) qs[j] = (int8_t)roundf(xb[j] * id);
        }
    }
    return 0;
}

/* ---- dequantizers: ggml/src/ggml.c:1496-1606 ------------------------------------------------------ */
static inline void unpack_block(int type, const uint8_t *b, int *q /*[32]*/, float *d, float *m) {
    uint16_t dh; memcpy(&dh, b, 2); *d = h2f(dh); *m = 0.0f;
    const uint8_t *qs; uint32_t qh = 0;
    switch (type) {
    case ORC_Q4_0: qs = b + 2;
        for (int j = 0; j < 16; j++) { q[j] = (qs[j] & 0x0F) - 8; q[j + 16] = (qs[j] >> 4) - 8; } break;
    case ORC_Q4_1: { uint16_t mh; memcpy(&mh, b + 2, 2); *m = h2f(mh); qs = b + 4;
        for (int j = 0; j < 16; j++) { q[j] = (qs[j] & 0x0F); q[j + 16] = (qs[j] >> 4); } } break;
    case ORC_Q5_0: memcpy(&qh, b + 2, 4); qs = b + 6;
        for (int j = 0; j < 16; j++) {
            q[j]      = ((qs[j] & 0x0F) | (((qh >> j) & 1) << 4)) - 16;
            q[j + 16] = ((qs[j] >> 4)   | (((qh >> (j + 16)) & 1) << 4)) - 16;
        } break;
    case ORC_Q5_1: { uint16_t mh; memcpy(&mh, b + 2, 2); *m = h2f(mh); memcpy(&qh, b + 4, 4); qs = b + 8;
        for (int j = 0; j < 16; j++) {
            q[j]      = (qs[j] & 0x0F) | (((qh >> j) & 1) << 4);
            q[j + 16] = (qs[j] >> 4)   | (((qh >> (j + 16)) & 1) << 4);
        } } break;
    default: /* ORC_Q8_0 */
        for (int j = 0; j < 32; j++) q[j] = ((const int8_t *)(b + 2))[j];
    }
}

int orc_dequantize_row(int type, const void *vx, float *y, int64_t k) {
    if (type == ORC_F32) { memcpy(y, vx, (size_t)k * 4); return 0; }
    if (type == ORC_F16) { const uint16_t *x = (const uint16_t *)vx; for (int64_t i = 0; i < k; i++) y[i] = h2f(x[i]); return 0; }
    const size_t bs = block_bytes(type);
    if (!bs || k % QK) return -1;
    const uint8_t *x = (const uint8_t *)vx;
    for (int64_t i = 0; i < k / QK; i++) {
        int q[32]; float d, m;
        unpack_block(type, x + i * bs, q, &d, &m);
        for (int j = 0; j < 32; j++) y[i * QK + j] = (float)q[j] * d + m;   /* m == 0 for the symmetric types */
    }
    return 0;
}

/* ---- activation quantization as mul_mat's INIT phase does it on an AVX2 build ----------------------
 * q8_0: ggml.c:1188-1277 (d stored fp16), q8_1: ggml.c:1398-1494 (d kept fp32, s = d * sum q).
 * id = 127/amax, round half to even (_mm256_round_ps NEAREST). */
typedef struct { float d; float s; int8_t q[QK]; } act_block;

static void quantize_act(const float *x, act_block *y, int64_t k, int q81) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float *xb = x + i * QK;
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) { float v = fabsf(xb[j]); if (v > amax) amax = v; }
        const float d = amax / 127.f;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        int sum = 0;
        for (int j = 0; j < QK; j++) { int v = (int)nearbyintf(xb[j] * id); y[i].q[j] = (int8_t)v; sum += v; }
        if (q81) { y[i].d = d; y[i].s = d * (float)sum; }
        else     { y[i].d = h2f(f2h(d)); y[i].s = 0.0f; }
    }
}

/* ---- y = x . W^T : ggml_compute_forward_mul_mat, ggml/src/ggml.c:11223-11437 ----------------------- */
int orc_mul_mat(int type, const void *W, int64_t n, int64_t k, const float *X, int64_t m, float *Y, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    const size_t rs = orc_row_size(type, k);
    const uint8_t *Wb = (const uint8_t *)W;
    if (type == ORC_F32) {
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; i++)
            for (int64_t j = 0; j < n; j++) {
                const float *w = (const float *)(Wb + j * rs), *x = X + i * k;
                float acc = 0.0f;
                for (int64_t t = 0; t < k; t++) acc += x[t] * w[t];
                Y[i * n + j] = acc;
            }
        return 0;
    }
    if (type == ORC_F16) {
        /* activations rounded to fp16 (vec_dot_type f16, ggml.c:1629-1637), fp32 accumulate (ggml.c:2370-2404) */
        float *xr = (float *)malloc((size_t)m * k * 4);
        float *wr = (float *)malloc((size_t)n * k * 4);
        if (!xr || !wr) { free(xr); free(wr); return -1; }
        orc_f32_to_f16_to_f32(X, xr, m * k);
        #pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < n; j++) {
            const uint16_t *w = (const uint16_t *)(Wb + j * rs);
            for (int64_t t = 0; t < k; t++) wr[j * k + t] = h2f(w[t]);
        }
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; i++)
            for (int64_t j = 0; j < n; j++) {
                const float *w = wr + j * k, *x = xr + i * k;
                float acc = 0.0f;
                for (int64_t t = 0; t < k; t++) acc += x[t] * w[t];
                Y[i * n + j] = acc;
            }
        free(xr); free(wr);
        return 0;
    }
    const size_t bs = block_bytes(type);
    if (!bs || k % QK) return -1;
    const int q81 = (type == ORC_Q4_1 || type == ORC_Q5_1);
    const int64_t nb = k / QK;
    act_block *A = (act_block *)malloc((size_t)m * nb * sizeof(act_block));
    /* unpack weights once to int8 + scales so the inner loop is a plain integer dot */
    int8_t *Wq = (int8_t *)malloc((size_t)n * k);
    float *Wd = (float *)malloc((size_t)n * nb * 4), *Wm = (float *)malloc((size_t)n * nb * 4);
    if (!A || !Wq || !Wd || !Wm) { free(A); free(Wq); free(Wd); free(Wm); return -1; }
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; i++) quantize_act(X + i * k, A + i * nb, k, q81);
    #pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; j++)
        for (int64_t b = 0; b < nb; b++) {
            int q[32]; float d, mm;
            unpack_block(type, Wb + j * rs + b * bs, q, &d, &mm);
            for (int t = 0; t < 32; t++) Wq[j * k + b * QK + t] = (int8_t)q[t];
            Wd[j * nb + b] = d; Wm[j * nb + b] = mm;
        }
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++) {
            const act_block *a = A + i * nb;
            const int8_t *wq = Wq + j * k;
            float acc = 0.0f, summs = 0.0f;
            for (int64_t b = 0; b < nb; b++) {
                int s = 0;
                for (int t = 0; t < QK; t++) s += (int)wq[b * QK + t] * (int)a[b].q[t];
                /* q4_0/q5_0/q8_0 . q8_0: sum * d_w * d_x (ggml.c:2695-2707, 3155-3176, 3611-3621)
                 * q4_1/q5_1 . q8_1: (d_w*d_x)*sum + m_w*s_x   (ggml.c:2851-2864, 3476-3497)          */
                acc += (float)s * (Wd[j * nb + b] * a[b].d);
                if (q81) summs += Wm[j * nb + b] * a[b].s;
            }
            Y[i * n + j] = acc + summs;
        }
    free(A); free(Wq); free(Wd); free(Wm);
    return 0;
}

/* ---- ggml_norm (no affine): two passes, double sums -- ggml/src/ggml.c:10796-10845 ------------------ */
void orc_layer_norm(const float *x, float *y, int64_t rows, int64_t h, float eps) {
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const float *xr = x + r * h; float *yr = y + r * h;
        double sum = 0.0;
        for (int64_t i = 0; i < h; i++) sum += (double)xr[i];
        const float mean = (float)(sum / (double)h);
        double sum2 = 0.0;
        for (int64_t i = 0; i < h; i++) { float v = xr[i] - mean; yr[i] = v; sum2 += (double)(v * v); }
        const float variance = (float)(sum2 / (double)h);
        const float scale = 1.0f / sqrtf(variance + eps);
        for (int64_t i = 0; i < h; i++) yr[i] *= scale;
    }
}

/* ---- GELU / quick-GELU through the fp16 tables -- ggml/src/ggml.c:3752-3815, 4525-4540 --------------- */
void orc_gelu(const float *x, float *y, int64_t n, int quick) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float f = h2f(f2h(x[i]));
        float g;
        if (quick) g = f * (1.0f / (1.0f + expf(-1.702f * f)));
        else       g = 0.5f * f * (1.0f + tanhf(0.79788456080286535587989211986876f * f * (1.0f + 0.044715f * f * f)));
        y[i] = h2f(f2h(g));
    }
}

/* ---- soft_max rows: max, exp via fp16 table, double sum -- ggml/src/ggml.c:12201-12270 --------------- */
void orc_softmax_rows(float *x, int64_t rows, int64_t cols) {
    for (int64_t r = 0; r < rows; r++) {
        float *p = x + r * cols;
        float mx = -INFINITY;
        for (int64_t i = 0; i < cols; i++) if (p[i] > mx) mx = p[i];
        double sum = 0.0;
        for (int64_t i = 0; i < cols; i++) {
            if (p[i] == -INFINITY) { p[i] = 0.0f; continue; }
            const float val = h2f(f2h(expf(h2f(f2h(p[i] - mx)))));
            sum += (double)val; p[i] = val;
        }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < cols; i++) p[i] *= inv;
    }
}

/* ---- attention core: KQ = K.Q (fp32), [causal mask], soft_max, KQV = V.KQ (fp32) ----------------------
 * clip.cpp:1100-1108 (text, ggml_diag_mask_inf n_past=0: col > row -> -inf, ggml.c:12117-12165)
 * clip.cpp:1382-1388 (vision, no mask).  q is pre-scaled by 1/sqrt(dh) (clip.cpp:1082,1363). */
void orc_attention(const float *q, const float *k, const float *v, float *out, int T, int H, int dh, int causal, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    const int hs = H * dh;
    #pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        float *s = (float *)malloc((size_t)T * 4);
        for (int i = 0; i < T; i++) {
            for (int j = 0; j < T; j++) {
                if (causal && j > i) { s[j] = -INFINITY; continue; }
                float acc = 0.0f;
                for (int t = 0; t < dh; t++) acc += k[(size_t)j * hs + h * dh + t] * q[(size_t)i * hs + h * dh + t];
                s[j] = acc;
            }
            orc_softmax_rows(s, 1, T);
            for (int t = 0; t < dh; t++) {
                float acc = 0.0f;
                for (int j = 0; j < T; j++) acc += v[(size_t)j * hs + h * dh + t] * s[j];
                out[(size_t)i * hs + h * dh + t] = acc;
            }
        }
        free(s);
    }
}

void orc_f32_to_f16_to_f32(const float *x, float *y, int64_t n) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] = h2f(f2h(x[i]));
}

/* sqrt(sum(sqr(x))) with ggml_sum's double accumulator (ggml.c:9611-9700), clip.cpp:1164, 1451 */
float orc_sum_sq_sqrt(const float *x, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += (double)(x[i] * x[i]);
    return sqrtf((float)s);
}
