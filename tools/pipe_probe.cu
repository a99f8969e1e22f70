// pipe_probe.cu -- the skeleton of K1's pair kernel with everything but the tensor pipe optional, to find what keeps the tensor pipe
// from its 96 cycles per UMMA inside the real kernel (profiles/r02_pipe_probe.md):
//   * warp 0 of both CTAs: TMA tensor loads of the activation half-tile (cta_group::2, 12 KB per k-block per CTA) into a ring,
//   * warp 1 of the leader: wait full[s] -> 4 x tcgen05.mma cta_group::2 (A from TMEM) -> commit empty[s]   (exactly K1's issue loop)
//   * optional: ALU warps spinning on packed 16-bit arithmetic (the unpack math's issue pressure), tcgen05.ld / tcgen05.st side warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/bin/pipe_probe tools/pipe_probe.cu -Iclip.cpp_b200/csrc -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace cb;

struct Cfg {
    int stages;     // ring depth (12 KB per stage per CTA)
    int tma;        // 1: real TMA loads, 0: the producer just arrives on full[s]
    int alu_warps;  // 0..8 busy ALU warps per CTA
    int ld_warps, st_warps;
    int rounds;
    int rows;       // rows of the activation matrix the CTAs walk over
};

__device__ unsigned long long g_cycles[256];

__global__ void __launch_bounds__(640, 1) pipe_kernel(const __grid_constant__ CUtensorMap tm, Cfg c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bars = sb + 200 * 1024;
    const uint32_t full = bars, empty = bars + 128;
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 256);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 200 * 1024 + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        mbar_fence_init();
        *stop = 0;
    }
    if (warp == 2) tmem_alloc_2sm(smem_u32(slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm_base = *slot;
    const int S = c.stages;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    if (warp == 0) {
        uint32_t s = 0, ph = 0;
        int row = pair * 192;
        for (int r = 0; r < c.rounds; r++) {
            mbar_wait(empty + 8 * s, ph ^ 1);
            if (elect_one()) {
                if (!c.tma) { if (rank == 0) mbar_arrive(full + 8 * s); }
                else {
                    if (rank == 0) mbar_arrive_expect_tx(full + 8 * s, 24576);
                    tma_load_2d_2sm(sb + s * 12288, &tm, (r & 15) * 64, row + (int)rank * 96, mapa_rank0(full + 8 * s));
                }
            }
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            if ((r & 15) == 15) { row += npairs * 192; if (row + 192 > c.rows) row = pair * 192; }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = umma_idesc(true, 256, 192);
            uint32_t s = 0, ph = 0;
            const long long t0 = clock64();
            for (int r = 0; r < c.rounds; r++) {
                mbar_wait(full + 8 * s, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t db = umma_desc_k128(sb + s * 12288);
                    const uint32_t a = tm_base + 384 + (r & 3) * 32, d = tm_base + ((r >> 4) & 1) * 192;
                    umma_f16_ts_2sm_acc(d, a, db, idesc);
                    umma_f16_ts_2sm_acc(d, a + 8, db + 2, idesc);
                    umma_f16_ts_2sm_acc(d, a + 16, db + 4, idesc);
                    umma_f16_ts_2sm_acc(d, a + 24, db + 6, idesc);
                    umma_commit_2sm(empty + 8 * s);
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
            // all MMAs done when the last S commits have landed
            for (int r = c.rounds - S; r < c.rounds; r++) mbar_wait(empty + 8 * (r % S), (r / S) & 1);
            const long long t1 = clock64();
            if (lane == 0) {
                g_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
                *stop = 1;
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote) : "r"(sb + 200 * 1024 + 512));
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
            }
        }
    } else if (warp >= 4 && warp < 8) {
        if (warp - 4 < c.ld_warps) {
            const uint32_t la = tm_base + ((uint32_t)((warp & 3) * 32) << 16) + 192;
            uint32_t v[32], sink = 0;
            while (!*stop) { for (int j = 0; j < 6; j++) { tmem_ld_32x32(la + j * 32, v); tmem_ld_wait(); sink += v[0] ^ v[31]; } }
            if (sink == 0xdeadbeefu) g_cycles[255] = sink;
        }
    } else if (warp >= 8 && warp < 12) {
        if (warp - 8 < c.st_warps) {
            const uint32_t sa = tm_base + ((uint32_t)((warp & 3) * 32) << 16) + 480;
            uint32_t v[32];
            for (int j = 0; j < 32; j++) v[j] = 0x3c003c00u;
            while (!*stop) { tmem_st_32x32(sa, v); tmem_st_wait(); }
        }
    } else if (warp >= 12) {
        if (warp - 12 < c.alu_warps) {
            // the unpack warps' instruction mix: shifts, LOP3, packed bf16 add / mul on independent registers
            uint32_t x[16];
            for (int j = 0; j < 16; j++) x[j] = 0x43004300u + lane + j;
            const uint32_t d2 = 0x3c003c00u;
            while (!*stop) {
                #pragma unroll
                for (int it = 0; it < 8; it++)
                    #pragma unroll
                    for (int j = 0; j < 16; j++) {
                        uint32_t t = (x[j] >> 4) & 0x000f000fu;
                        asm volatile("lop3.b32 %0, %0, 0x000f000f, %1, 0xEA;" : "+r"(t) : "r"(x[(j + 1) & 15]));
                        uint32_t u;
                        asm volatile("sub.rn.bf16x2 %0, %1, %2;" : "=r"(u) : "r"(t), "r"(d2));
                        asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(x[j]) : "r"(u), "r"(d2));
                    }
            }
            if (x[0] == 0x12345u) g_cycles[254] = x[3];
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tm_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int rows = 65792, cols = 1024;          // one 256-image ViT-L/14 micro-batch of activations, bf16
    void* d_x = nullptr;
    cudaMalloc(&d_x, (size_t)rows * cols * 2);
    cudaMemset(d_x, 0x3c, (size_t)rows * cols * 2);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, 96}, estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d_x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); return 1; }
    const size_t smem = 200 * 1024 + 2048;
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("%s, %d SMs; cycles per k-block (4 UMMAs M256 N192 K16, ideal 384) on the slowest CTA pair\n", p.name, p.multiProcessorCount);
    printf("%-7s %-4s %-4s %-3s %-3s | cycles/k-block\n", "stages", "tma", "alu", "ld", "st");
    struct Row { int stages, tma, alu, ld, st; };
    const Row rows_[] = {{12, 0, 0, 0, 0}, {12, 1, 0, 0, 0}, {6, 1, 0, 0, 0}, {4, 1, 0, 0, 0}, {12, 0, 2, 0, 0}, {12, 0, 4, 0, 0}, {12, 0, 8, 0, 0},
                         {12, 1, 4, 0, 0}, {12, 1, 8, 0, 0}, {12, 1, 8, 4, 0}, {12, 1, 8, 4, 4}, {12, 1, 0, 4, 4}, {12, 0, 8, 4, 4}};
    for (const Row& r : rows_) {
        Cfg c{r.stages, r.tma, r.alu, r.ld, r.st, 3200, rows};
        double best = 1e30;
        for (int rep = 0; rep < 2; rep++) {
            cudaError_t e = launch_pdl(pipe_kernel, (unsigned)(p.multiProcessorCount & ~1), 640u, smem, 0, 2, tm, c);
            if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<unsigned long long> h(256);
            cudaMemcpyFromSymbol(h.data(), g_cycles, sizeof(unsigned long long) * 256);
            double mx = 0;
            for (int i = 0; i < p.multiProcessorCount; i += 2) mx = mx > (double)h[i] ? mx : (double)h[i];
            if (rep == 1) best = mx / c.rounds;
        }
        printf("%-7d %-4d %-4d %-3d %-3d | %8.1f\n", r.stages, r.tma, r.alu, r.ld, r.st, best);
    }
    return 0;
}
