// attention_tc.cu -- K3 on tcgen05: S = Q K^T and O = P V as UMMA instructions with both accumulators in TENSOR MEMORY.
//
// Reference semantics: clip.cpp:1082-1108 (text, causal) / 1363-1388 (vision); soft_max ggml.c:12201-12270.
// One persistent CTA per SM walks work items (sequence, head, 128-query tile):
//   warp 0      TMA: Q box [128 x 64], K and V boxes [<=272 keys x 64] straight out of the fused QKV activation matrix
//               (row stride 3*hidden), 128B swizzle, 2-stage ring -> the next item's loads overlap this item's math
//   warp 1      one elected thread: S[128 x nk] = Q.K^T (SS form, K=64 -> 4 UMMA k-steps; N = 256 (+16) columns), then
//               O[128 x 64] = P.V (TS form: A = P from TMEM, B = V taken MN-major from the same [key][dh] tile, K = nk)
//   warps 4-11  softmax in the TMEM lane == query-row mapping (two warps per lane quarter split the key chunks): pass 1 row max over
//               tcgen05.ld chunks, pass 2 p = 2^((s-m)*log2e) -> bf16/fp16 pairs written to the P columns
//               (tcgen05.st), row sum kept in fp32; after the PV commit: O / l -> 128-byte row stores
//   warp 2      TMEM alloc (S: columns [0,272), P: [272,408), O: [408,472))
// HBM traffic: Q, K, V read once per (sequence, head) (K/V re-reads of the second query tile hit L2), O written once.
// Limits: head_dim 64, T <= 272 keys; query rows beyond the last full 128-tile (e.g. row 256 of ViT-L/14's 257) are left
// to the warp-level kernel in attention.cu, launched with a query offset.
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "kernels.h"

namespace cb {

namespace {

constexpr int AQ = 128, AKMAX = 272, DH = 64;
constexpr uint32_t Q_BYTES = AQ * 128, KV_BYTES = AKMAX * 128;
constexpr uint32_t STAGE = Q_BYTES + 2 * KV_BYTES;     // 86016
constexpr int NS = 2;
constexpr uint32_t BAR_OFF = NS * STAGE;
constexpr uint32_t ATT_SMEM = BAR_OFF + 128 + 2048 /*row max / row sum exchange*/ + 1024;
constexpr uint32_t P_COL = 272, O_COL = 408;   // TMEM columns: S [0,272)  P (16-bit pairs) [272,408)  O [408,472)
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

template <bool BF>
__global__ void __launch_bounds__(384, 1) attention_tc_kernel(const __grid_constant__ AParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base + BAR_OFF;
    const uint32_t kv_full = bars, kv_empty = bars + 16, s_full = bars + 32, p_full = bars + 40, o_full = bars + 48, o_empty = bars + 56;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(kv_full + 8 * i, 1); mbar_init(kv_empty + 8 * i, 1); }
        mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(o_full, 1); mbar_init(o_empty, 8);
        mbar_fence_init();
        tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_kv256); tma_prefetch_desc(&p.tm_kv16);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int hid = p.H * DH, total = p.nseq * p.H * p.ntile;
    const int n0 = p.nk16 < 256 ? p.nk16 : 256, n1 = p.nk16 - n0;       // key columns of the two S MMAs
    const uint32_t kv_tx = 32768u + (n1 > 0 ? 2048u : 0u);

    if (warp == 0) {
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const int qt = it % p.ntile, head = (it / p.ntile) % p.H, seq = it / (p.ntile * p.H);
            const uint32_t s = j & 1, ph = (j >> 1) & 1;
            mbar_wait(kv_empty + 8 * s, ph ^ 1);
            if (elect_one()) {
                const uint32_t dq = smem_base + s * STAGE, dk = dq + Q_BYTES, dv = dk + KV_BYTES, bar = kv_full + 8 * s;
                const int row0 = seq * p.T, c = head * DH;
                mbar_arrive_expect_tx(bar, Q_BYTES + 2 * kv_tx);
                tma_load_2d(dq, &p.tm_q, c, row0 + qt * AQ, bar);
                tma_load_2d(dk, &p.tm_kv256, hid + c, row0, bar);
                tma_load_2d(dv, &p.tm_kv256, 2 * hid + c, row0, bar);
                if (n1 > 0) {
                    tma_load_2d(dk + 32768, &p.tm_kv16, hid + c, row0 + 256, bar);
                    tma_load_2d(dv + 32768, &p.tm_kv16, 2 * hid + c, row0 + 256, bar);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t id_s0 = idesc_n(BF, n0, false), id_s1 = idesc_n(BF, n1 > 0 ? n1 : 16, false), id_pv = idesc_n(BF, DH, true);
        const int npv = p.nk16 >> 4;
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const uint32_t s = j & 1, ph = (j >> 1) & 1, jp = j & 1;
            mbar_wait(kv_full + 8 * s, ph);        // S/P columns are free: PV(j-1) was issued before us and UMMAs run in order
            tc_fence_after();
            const uint32_t qa = smem_base + s * STAGE;
            const uint64_t dq = umma_desc_k128(qa), dk = umma_desc_k128(qa + Q_BYTES), dk1 = umma_desc_k128(qa + Q_BYTES + 32768);
            if (elect_one()) {
                umma_f16_init(tmem_base, dq, dk, id_s0);
                umma_f16_acc(tmem_base, dq + 2, dk + 2, id_s0);
                umma_f16_acc(tmem_base, dq + 4, dk + 4, id_s0);
                umma_f16_acc(tmem_base, dq + 6, dk + 6, id_s0);
                if (n1 > 0) {
                    umma_f16_init(tmem_base + 256, dq, dk1, id_s1);
                    umma_f16_acc(tmem_base + 256, dq + 2, dk1 + 2, id_s1);
                    umma_f16_acc(tmem_base + 256, dq + 4, dk1 + 4, id_s1);
                    umma_f16_acc(tmem_base + 256, dq + 6, dk1 + 6, id_s1);
                }
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, jp);
            mbar_wait(o_empty, jp ^ 1);            // previous item's epilogue has drained the O columns
            tc_fence_after();
            const uint64_t dv = umma_desc_mn128(qa + Q_BYTES + KV_BYTES);
            if (elect_one()) {
                umma_f16_ts_init(tmem_base + O_COL, tmem_base + P_COL, dv, id_pv);
                for (int ks = 1; ks < npv; ks++)       // 16 keys per step: 8 packed TMEM columns of P, 16 rows (2048 B) of V
                    umma_f16_ts_acc(tmem_base + O_COL, tmem_base + P_COL + 8 * ks, dv + (uint64_t)ks * (2048 >> 4), id_pv);
                umma_commit(o_full);
                umma_commit(kv_empty + 8 * s);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax + epilogue: 8 warps, two per TMEM lane quarter.
        // Warp pair (w, w+4) shares the 32 query rows of its quarter and splits the key chunks; row max and row sum are
        // exchanged through shared memory (one named barrier per item), the O columns are split 32 / 32 for the store.
        const int r = (warp & 3) * 32 + lane, half = (warp - 4) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float* xmax = reinterpret_cast<float*>(smem + BAR_OFF + 128);         // [2][128]
        float* xsum = xmax + 256;                                             // [2][128]
        const float LOG2E = 1.4426950408889634f;
        const int nch = (p.nk16 + 31) >> 5;
        const int c_lo = half ? (nch + 1) / 2 : 0, c_hi = half ? nch : (nch + 1) / 2;
        int j = 0;
        for (int it = blockIdx.x; it < total; it += gridDim.x, j++) {
            const int qt = it % p.ntile, head = (it / p.ntile) % p.H, seq = it / (p.ntile * p.H);
            const uint32_t jp = j & 1;
            const int qrow = qt * AQ + r;
            const int klim = p.causal ? min(p.T, qrow + 1) : p.T;      // keys [0, klim) are visible to this row
            const int klim_min = __shfl_sync(0xffffffffu, klim, 0);    // lane 0 holds the smallest row of the warp
            mbar_wait(s_full, jp);
            tc_fence_after();
            // ---- pass 1: row max over this warp's chunks
            float m = -INFINITY;
            for (int c = c_lo; c < c_hi; c++) {
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
            xmax[half * 128 + r] = m;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            m = fmaxf(m, xmax[(half ^ 1) * 128 + r]);
            if (m == -INFINITY) m = 0.f;                               // padded query rows past T
            const float mb = m * LOG2E;
            // ---- pass 2: p = 2^(s*log2e - m*log2e) as packed 16-bit pairs into the P columns
            float l = 0.f;
            for (int c = c_lo; c < c_hi; c++) {
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
                tmem_st_32x16(lane_addr + P_COL + c * 16, pk);     // P has its own columns: the two warps of a quarter never alias
            }
            xsum[half * 128 + r] = l;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            // ---- O / l -> global (this warp stores 32 of the 64 head-dim columns)
            mbar_wait(o_full, jp);
            tc_fence_after();
            const float inv = 1.0f / (l + xsum[(half ^ 1) * 128 + r]);
            uint16_t* orow = p.out + ((size_t)seq * p.T + qrow) * hid + head * DH + half * 32;
            {
                uint32_t v[32];
                tmem_ld_32x32(lane_addr + O_COL + half * 32, v);
                tmem_ld_wait();
                if (qrow < p.T) {
                    #pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint4 q4;
                        q4.x = pack2<BF>(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
                        q4.y = pack2<BF>(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
                        q4.z = pack2<BF>(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
                        q4.w = pack2<BF>(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + 8 * i) = q4;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool attention_tc_supported(int T) { return T >= 1 && ((T + 15) / 16) * 16 <= AKMAX; }

// query tiles handled by the tcgen05 kernel; rows [ntile*128, T) (if any) belong to the warp-level kernel
int attention_tc_tiles(int T) {
    const int full = T / AQ, rem = T % AQ;
    if (full > 0 && rem <= 16) return full;
    return (T + AQ - 1) / AQ;
}

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
    const int grid = total < num_sms ? total : num_sms;
    if (bf16) attention_tc_kernel<true><<<grid, 384, ATT_SMEM, st>>>(p);
    else attention_tc_kernel<false><<<grid, 384, ATT_SMEM, st>>>(p);
    return cudaGetLastError();
}

}  // namespace cb
