// search.cu -- scoring on the device: similarity matrix, softmax_with_sorting arithmetic, top-k selection.
//
// Replaces, for batches, the host loops of clip_zero_shot_label_image (clip.cpp:1624-1659: dot products clip.cpp:1525-1532,
// softmax_with_sorting clip.cpp:1591-1622) and the nearest-neighbour lookup of examples/image-search/search.cpp:114-158 (USearch
// there; brute force here: at <= 1e7 vectors the scan is one HBM pass).  Everything is fp32 on the CUDA cores ON PURPOSE: the whole
// 4096 x 1000 x 768 zero-shot problem is 6.3 GFLOP (0.008 % of the encode that produced the embeddings) and the ranking must not
// move in the 4th digit, which bf16 tensor-core products would do.  Each dot product is ONE sequential fmaf chain over k, the same
// order the reference's scalar loop uses.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace cb {

namespace {

constexpr int ST = 64, SK = 16;      // 64 x 64 output tile per CTA, k-chunks of 16 staged through shared memory

// S[i, j] = sum_k A[i, k] * B[j, k]   (A: [na, d] queries / images, B: [nb, d] labels / database rows), row-major fp32
__global__ void __launch_bounds__(256) sim_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ S,
                                                  int na, int nb, int d, size_t lds) {
    __shared__ float sa[SK][ST + 1], sb[SK][ST + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 x 16 threads, 4 x 4 outputs each
    const int i0 = blockIdx.y * ST, j0 = blockIdx.x * ST;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += SK) {
        for (int e = threadIdx.x; e < ST * SK; e += 256) {
            const int r = e / SK, k = e % SK;
            sa[k][r] = (i0 + r < na && k0 + k < d) ? A[(size_t)(i0 + r) * d + k0 + k] : 0.f;
            sb[k][r] = (j0 + r < nb && k0 + k < d) ? B[(size_t)(j0 + r) * d + k0 + k] : 0.f;
        }
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < SK; k++) {
            float a[4], b[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) { a[u] = sa[k][ty + 16 * u]; b[u] = sb[k][tx + 16 * u]; }
            #pragma unroll
            for (int u = 0; u < 4; u++)
                #pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
    #pragma unroll
    for (int u = 0; u < 4; u++)
        #pragma unroll
        for (int v = 0; v < 4; v++) {
            const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
            if (i < na && j < nb) S[(size_t)i * lds + j] = acc[u][v];
        }
}

CB_DEVINL double warp_sum_d(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// p = (exp(s) + 1e-9) / sum  -- softmax_with_sorting's arithmetic (clip.cpp:1599-1607): exp in double, the element rounded to float,
// the sum of the rounded elements in double, the quotient rounded to float; no max subtraction (as the reference).
__global__ void __launch_bounds__(256) softmax_ref_kernel(float* __restrict__ x, int rows, int cols, size_t ld) {
    __shared__ double part[8];
    const int row = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (row >= rows) return;
    float* r = x + (size_t)row * ld;
    double s = 0.0;
    for (int i = threadIdx.x; i < cols; i += 256) { const float e = (float)(exp((double)r[i]) + 1e-9); r[i] = e; s += (double)e; }
    s = warp_sum_d(s);
    if (lane == 0) part[w] = s;
    __syncthreads();
    double tot = 0.0;
    #pragma unroll
    for (int i = 0; i < 8; i++) tot += part[i];
    for (int i = threadIdx.x; i < cols; i += 256) r[i] = (float)((double)r[i] / tot);
}

// One stage of top-k: the CTA (slice s, row r) sorts its <= SL candidates (value descending, ties by ascending index: a total
// order, unlike the reference's qsort whose tie order is unspecified) with a bitonic network in shared memory and emits the
// best kk.  in_idx == nullptr: candidate j of the row has index idx_base + j.
constexpr int SL = 4096;
struct Cand { float v; int i; };
CB_DEVINL bool before(const Cand& a, const Cand& b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

__global__ void __launch_bounds__(256) topk_stage_kernel(const float* __restrict__ vals, const int* __restrict__ in_idx, size_t ld, int n,
                                                         int idx_base, int kk, float* __restrict__ out_v, int* __restrict__ out_i,
                                                         size_t out_ld, int out_off) {
    __shared__ Cand c[SL];
    const int row = blockIdx.y, s0 = blockIdx.x * SL;
    const float* v = vals + (size_t)row * ld;
    const int* ii = in_idx ? in_idx + (size_t)row * ld : nullptr;
    const int cnt = min(SL, n - s0);
    int m = 1;
    while (m < cnt) m <<= 1;                                   // sort only the power of two that covers this slice
    for (int e = threadIdx.x; e < m; e += 256) {
        Cand x;
        if (e < cnt) {
            const float f = v[s0 + e];
            x.v = (f != f) ? -FLT_MAX : f;                     // NaN scores rank last instead of breaking the network
            x.i = ii ? ii[s0 + e] : idx_base + s0 + e;
        } else { x.v = -INFINITY; x.i = 0x7fffffff; }
        c[e] = x;
    }
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (m >> 1); t += 256) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;            // final order: best first
                const Cand a = c[lo], b = c[hi];
                if (before(b, a) == desc) { c[lo] = b; c[hi] = a; }
            }
            __syncthreads();
        }
    const int take = min(kk, cnt);
    float* ov = out_v + (size_t)row * out_ld + out_off + (size_t)blockIdx.x * kk;
    int* oi = out_i + (size_t)row * out_ld + out_off + (size_t)blockIdx.x * kk;
    for (int e = threadIdx.x; e < kk; e += 256) {
        ov[e] = e < take ? c[e].v : -INFINITY;
        oi[e] = e < take ? c[e].i : 0x7fffffff;
    }
}

}  // namespace

void launch_similarity(const float* a, const float* b, float* s, int na, int nb, int d, size_t lds, cudaStream_t st) {
    if (na <= 0 || nb <= 0) return;
    dim3 grid((nb + ST - 1) / ST, (na + ST - 1) / ST);
    sim_kernel<<<grid, 256, 0, st>>>(a, b, s, na, nb, d, lds);
}

void launch_softmax_ref(float* logits, int rows, int cols, size_t ld, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    softmax_ref_kernel<<<rows, 256, 0, st>>>(logits, rows, cols, ld);
}

int topk_slice() { return SL; }

void launch_topk_stage(const float* vals, const int* in_idx, size_t ld, int rows, int n, int idx_base, int kk, float* out_v, int* out_i,
                       size_t out_ld, int out_off, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return;
    dim3 grid((n + SL - 1) / SL, rows);
    topk_stage_kernel<<<grid, 256, 0, st>>>(vals, in_idx, ld, n, idx_base, kk, out_v, out_i, out_ld, out_off);
}

}  // namespace cb
