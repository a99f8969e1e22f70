// attention_tc.cu -- K3 on tcgen05: S = Q K^T and O = P V as UMMA instructions with both accumulators in TENSOR MEMORY.
//
// Reference semantics: clip.cpp:1082-1108 (text, causal) / 1363-1388 (vision); soft_max ggml.c:12201-12270.
// Persistent CTAs, TWO per SM (256 TMEM columns and 87 KB of shared memory each, so one CTA's softmax runs under the other's
// MMAs / TMA loads), walk work items (sequence, head, 128-query tile):
//   warp 0      TMA: Q box [128 x 64], K and V boxes [256 (+16) keys x 64] straight out of the fused QKV activation matrix
//               (row stride 3*hidden), 128B swizzle
//   warp 1      one elected thread: S[128 x <=256] = Q.K^T (SS form, K=64 -> 4 UMMA k-steps), then
//               O[128 x 64] = P.V (TS form: A = P from TMEM, B = V taken MN-major from the same [key][dh] tile)
//   warps 4-7   softmax in the TMEM lane == query-row mapping (no cross-thread reduction at all): pass 1 row max over
//               tcgen05.ld chunks, pass 2 p = 2^((s-m)*log2e) -> bf16/fp16 pairs written IN PLACE over the consumed low half of
//               S (tcgen05.st), row sum kept in fp32; after the PV commit: O / l -> 128-byte row stores.  Key 256 of
//               ViT-L/14 (257 tokens) does not fit the 256-column S tile: its score and its p*V term are one 64-long dot
//               product / axpy per row on the CUDA cores, straight from the shared-memory tiles.
//   warp 2      TMEM alloc (S: columns [0,256); P: [0,128) in place; O: [128,192) once S is dead)
// HBM traffic: Q, K, V read once per (sequence, head) (re-reads by the other query tiles hit L2), O written once.
// Limits: head_dim 64, T <= 257.  Measured (ViT-L/14, 52 images per launch): 94 us warp-level mma.sync kernel -> 66 us;
// the binding resource is now the double read of S through tcgen05.ld plus MUFU.EX2.
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "kernels.h"

namespace cb {

namespace {

constexpr int AQ = 128, AKMAX = 272, DH = 64;
constexpr uint32_t Q_BYTES = AQ * 128, KV_BYTES = AKMAX * 128;
constexpr uint32_t STAGE = Q_BYTES + 2 * KV_BYTES;     // 86016
constexpr uint32_t BAR_OFF = STAGE;                    // single stage: the co-resident CTA (2 per SM) hides the load latency
constexpr uint32_t ATT_SMEM = BAR_OFF + 128 + 1024;
constexpr uint32_t TMEM_COLS = 256, O_COL = 128;   // TMEM: S fp32 [0,256); P (16-bit pairs) overwrites [0,128); O fp32 [128,192) after S is dead
static_assert(STAGE % 1024 == 0, "stage alignment");

struct AParams {
    CUtensorMap tm_q, tm_kv256, tm_kv16;
    uint16_t* out;
    int T, H, nseq, causal, ntile, nk16;
};

CB_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
CB_DEVINL void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
CB_DEVINL float ex2f(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <bool BF>
CB_DEVINL uint32_t pack2(float lo, float hi) {
    uint32_t r;
    if constexpr (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// instruction descriptor with runtime N; bmn = B operand is MN-major (V tile [key][dh]: dh contiguous)
CB_DEVINL uint32_t idesc_n(bool bf, int N, bool bmn) {
    return (1u << 4) | ((bf ? 1u : 0u) << 7) | ((bf ? 1u : 0u) << 10) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// MN-major operand tile: rows = k (128 B each = 64 MN elements), 8-row swizzle groups 1024 B apart (SBO); one 64-wide MN block (LBO unused)
CB_DEVINL uint64_t umma_desc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// bf16 / fp16 pair -> two floats
template <bool BF>
CB_DEVINL float2 unpack2(uint32_t u) {
    if constexpr (BF) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    else { const __half2 h = *reinterpret_cast<const __half2*>(&u); return __half22float2(h); }
}
CB_DEVINL uint4 lds128a(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

template <bool BF>
__global__ void __launch_bounds__(256, 2) attention_tc_kernel(const __grid_constant__ AParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base + BAR_OFF;
    const uint32_t kv_full = bars, kv_empty = bars + 8, s_full = bars + 16, p_full = bars + 24, o_full = bars + 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1); mbar_init(kv_empty, 4); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_kv256); tma_prefetch_desc(&p.tm_kv16);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();      // qkv is read (and the output written) only after the QKV GEMM grid has completed

    const int hid = p.H * DH, total = p.nseq * p.H * p.ntile;
    const int nmma = p.nk16 < 256 ? p.nk16 : 256;          // keys whose scores come from the tensor core
    const bool extra = p.T > 256;                          // key 256 (ViT-L/14: T = 257) is folded in on the CUDA cores
    const uint32_t kv_tx = 32768u + (extra ? 2048u : 0u);
    const uint32_t q_s = smem_base, k_s = smem_base + Q_BYTES, v_s = k_s + KV_BYTES;

    if (warp == 0) {
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const int qt = it % p.ntile, head = (it / p.ntile) % p.H, seq = it / (p.ntile * p.H);
            mbar_wait(kv_empty, (j & 1) ^ 1);
            if (elect_one()) {
                const int row0 = seq * p.T, c = head * DH;
                mbar_arrive_expect_tx(kv_full, Q_BYTES + 2 * kv_tx);
                tma_load_2d(q_s, &p.tm_q, c, row0 + qt * AQ, kv_full);
                tma_load_2d(k_s, &p.tm_kv256, hid + c, row0, kv_full);
                tma_load_2d(v_s, &p.tm_kv256, 2 * hid + c, row0, kv_full);
                if (extra) {
                    tma_load_2d(k_s + 32768, &p.tm_kv16, hid + c, row0 + 256, kv_full);
                    tma_load_2d(v_s + 32768, &p.tm_kv16, 2 * hid + c, row0 + 256, kv_full);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t id_s = idesc_n(BF, nmma, false), id_pv = idesc_n(BF, DH, true);
        const int npv = nmma >> 4;
        const uint64_t dq = umma_desc_k128(q_s), dk = umma_desc_k128(k_s), dv = umma_desc_mn128(v_s);
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const uint32_t jp = j & 1;
            mbar_wait(kv_full, jp);            // also implies the previous item's epilogue is done (kv_empty gates the loads)
            tc_fence_after();
            if (elect_one()) {
                umma_f16_init(tmem_base, dq, dk, id_s);
                umma_f16_acc(tmem_base, dq + 2, dk + 2, id_s);
                umma_f16_acc(tmem_base, dq + 4, dk + 4, id_s);
                umma_f16_acc(tmem_base, dq + 6, dk + 6, id_s);
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, jp);
            tc_fence_after();
            if (elect_one()) {
                umma_f16_ts_init(tmem_base + O_COL, tmem_base, dv, id_pv);
                for (int ks = 1; ks < npv; ks++)       // 16 keys per step: 8 packed TMEM columns of P, 16 rows (2048 B) of V
                    umma_f16_ts_acc(tmem_base + O_COL, tmem_base + 8 * ks, dv + (uint64_t)ks * (2048 >> 4), id_pv);
                umma_commit(o_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax + epilogue (TMEM lane == query row)
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float LOG2E = 1.4426950408889634f;
        const int nch = (nmma + 31) >> 5;
        const uint32_t q_row = q_s + r * 128, sw = r & 7;
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const int qt = it % p.ntile, head = (it / p.ntile) % p.H, seq = it / (p.ntile * p.H);
            const uint32_t jp = j & 1;
            const int qrow = qt * AQ + r;
            const int klim = p.causal ? min(p.T, qrow + 1) : p.T;      // keys [0, klim) are visible to this row
            const int klim_min = __shfl_sync(0xffffffffu, klim, 0);    // lane 0 holds the smallest row of the warp
            const bool warp_live = (qt * AQ + (warp & 3) * 32) < p.T;  // warp-uniform: any valid query row in this warp?
            mbar_wait(s_full, jp);
            tc_fence_after();
            float l = 0.f, p_x = 0.f;
            if (warp_live) {
                // ---- key 256: one dot product per row on the CUDA cores (Q row and K row 256 are in shared memory)
                float s_x = -INFINITY;
                if (extra && 256 < klim) {
                    float acc = 0.f;
                    #pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const uint4 a = lds128a(q_row + ((c ^ sw) << 4)), b = lds128a(k_s + 32768 + (c << 4));
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
                        #pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float2 fa = unpack2<BF>(aw[i]), fb = unpack2<BF>(bw[i]);
                            acc = fmaf(fa.x, fb.x, acc); acc = fmaf(fa.y, fb.y, acc);
                        }
                    }
                    s_x = acc;
                }
                // ---- pass 1: row max
                float m = s_x;
                for (int c = 0; c < nch; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32(lane_addr + c * 32, v);
                    tmem_ld_wait();
                    if (c * 32 + 32 <= klim_min) {                         // warp-uniform: every key of the chunk is visible
                        #pragma unroll
                        for (int i = 0; i < 32; i++) m = fmaxf(m, __uint_as_float(v[i]));
                    } else {
                        #pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (c * 32 + i < klim) m = fmaxf(m, __uint_as_float(v[i]));
                    }
                }
                if (m == -INFINITY) m = 0.f;                               // padded query rows past T
                const float mb = m * LOG2E;
                p_x = (s_x == -INFINITY) ? 0.f : ex2f(s_x * LOG2E - mb);
                l = p_x;
                // ---- pass 2: p = 2^(s*log2e - m*log2e); the packed 16-bit pairs overwrite the low half of the S columns this
                // thread has already consumed (chunk c -> columns [16c, 16c+16) <= [32c, ..)); only this warp touches these lanes
                for (int c = 0; c < nch; c++) {
                    uint32_t v[32], pk[16];
                    tmem_ld_32x32(lane_addr + c * 32, v);
                    tmem_ld_wait();
                    if (c * 32 + 32 <= klim_min) {
                        #pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float p0 = ex2f(__uint_as_float(v[2 * i]) * LOG2E - mb), p1 = ex2f(__uint_as_float(v[2 * i + 1]) * LOG2E - mb);
                            l += p0 + p1;
                            pk[i] = pack2<BF>(p0, p1);
                        }
                    } else {
                        #pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int k0 = c * 32 + 2 * i;
                            const float p0 = (k0 < klim) ? ex2f(__uint_as_float(v[2 * i]) * LOG2E - mb) : 0.f;
                            const float p1 = (k0 + 1 < klim) ? ex2f(__uint_as_float(v[2 * i + 1]) * LOG2E - mb) : 0.f;
                            l += p0 + p1;
                            pk[i] = pack2<BF>(p0, p1);
                        }
                    }
                    tmem_st_32x16(lane_addr + c * 16, pk);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            // ---- O (+ p_256 * V[256]) / l -> global
            mbar_wait(o_full, jp);
            tc_fence_after();
            if (warp_live) {
                const float inv = 1.0f / l;
                uint16_t* orow = p.out + ((size_t)seq * p.T + qrow) * hid + head * DH;
                #pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    uint32_t v[32];
                    tmem_ld_32x32(lane_addr + O_COL + h2 * 32, v);
                    tmem_ld_wait();
                    float o[32];
                    #pragma unroll
                    for (int i = 0; i < 32; i++) o[i] = __uint_as_float(v[i]);
                    if (extra) {
                        #pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const uint4 b = lds128a(v_s + 32768 + ((h2 * 4 + c) << 4));     // V row 256, broadcast read
                            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
                            #pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const float2 fv = unpack2<BF>(bw[i]);
                                o[8 * c + 2 * i] = fmaf(p_x, fv.x, o[8 * c + 2 * i]);
                                o[8 * c + 2 * i + 1] = fmaf(p_x, fv.y, o[8 * c + 2 * i + 1]);
                            }
                        }
                    }
                    if (qrow < p.T) {
                        #pragma unroll
                        for (int i = 0; i < 4; i++) {
                            uint4 q4;
                            q4.x = pack2<BF>(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
                            q4.y = pack2<BF>(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
                            q4.z = pack2<BF>(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
                            q4.w = pack2<BF>(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
                            *reinterpret_cast<uint4*>(orow + h2 * 32 + 8 * i) = q4;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(kv_empty);      // Q/K/V tile and the TMEM columns are free for the next item
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

bool attention_tc_supported(int T) { return T >= 1 && T <= 257; }   // 256 keys on the tensor core + at most one folded in on the CUDA cores

// query tiles handled by the tcgen05 kernel (all of them since v3: a ragged last tile only keeps its live warps busy)
int attention_tc_tiles(int T) { return (T + AQ - 1) / AQ; }

cudaError_t attention_tc_init() {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
}

cudaError_t launch_attention_tc(const TmaMap* map_q, const TmaMap* map_kv256, const TmaMap* map_kv16, void* out16, int nseq, int T,
                                int H, int causal, int bf16, int num_sms, cudaStream_t st) {
    if (nseq <= 0) return cudaSuccess;
    AParams p;
    memcpy(&p.tm_q, map_q, sizeof(CUtensorMap));
    memcpy(&p.tm_kv256, map_kv256, sizeof(CUtensorMap));
    memcpy(&p.tm_kv16, map_kv16, sizeof(CUtensorMap));
    p.out = (uint16_t*)out16; p.T = T; p.H = H; p.nseq = nseq; p.causal = causal;
    p.ntile = attention_tc_tiles(T);
    p.nk16 = ((T + 15) / 16) * 16;
    const int total = nseq * H * p.ntile;
    const int grid = total < 2 * num_sms ? total : 2 * num_sms;      // two CTAs per SM (256 TMEM columns, 87 KB smem each)
    cudaError_t e = bf16 ? launch_pdl(attention_tc_kernel<true>, (unsigned)grid, 256u, ATT_SMEM, st, 1, p)
                         : launch_pdl(attention_tc_kernel<false>, (unsigned)grid, 256u, ATT_SMEM, st, 1, p);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace cb
