// umma_probe.cu -- microbenchmark behind profiles/r02_umma_probe.md: what does one tcgen05.mma cost on B200 in the forms K1 uses, and how
// much do concurrent tcgen05.ld (epilogue) / tcgen05.st (unpack) slow the tensor pipe down?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe tools/umma_probe.cu -Iclip.cpp_b200/csrc && ./umma_probe
// Each CTA (or CTA pair) issues ROUNDS x 4 UMMAs (one "k-block") into one or two accumulators, commits after every k-block and
// keeps at most DEPTH k-blocks in flight; optional side warps hammer TMEM with loads / stores of their own columns.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace cb;

struct Cfg {
    int pair;      // cta_group::2
    int N;         // UMMA N
    int ts;        // A from TMEM (1) or shared memory (0)
    int nacc;      // accumulators used round-robin per UMMA (1 = dependent chain, 2 = alternating)
    int ld_warps;  // side warps issuing tcgen05.ld.32x32b.x32 in a loop (0..4)
    int st_warps;  // side warps issuing tcgen05.st.32x32b.x32 in a loop (0..4)
    int rounds;
};

__device__ unsigned long long g_cycles[256];
__device__ unsigned long long g_side[8];

template <int PAIR>
__global__ void __launch_bounds__(384, 1) probe_kernel(Cfg c) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar = sb + 200 * 1024;            // barriers behind 200 KB of operand space
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 256);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 200 * 1024 + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0 / bf16 small
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; i++) mbar_init(bar + 8 * i, 1);
        mbar_fence_init();
        *stop = 0;
    }
    if (warp == 2) { if (PAIR) tmem_alloc_2sm(smem_u32(slot), 512); else tmem_alloc(smem_u32(slot), 512); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = *slot;
    const int DEPTH = 4;
    if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
            const uint32_t a_col = 2 * 192;          // A stages at columns [384, 512)
            long long t0 = clock64();
            for (int r = 0; r < c.rounds; r++) {
                if (r >= DEPTH) mbar_wait(bar + 8 * (r % DEPTH), ((r / DEPTH) - 1) & 1);
                if (elect_one()) {
                    const uint64_t db = umma_desc_k128(sb + (r % 6) * 24576);
                    const uint64_t da = umma_desc_k128(sb + 6 * 24576 + (r % 2) * 16384);
                    #pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t d = tm + ((c.nacc == 2 && (u & 1)) ? 192 : 0);
                        const uint32_t a = tm + a_col + (r % 4) * 32 + u * 8;
                        if (PAIR) { if (c.ts) umma_f16_ts_2sm_acc(d, a, db + 2 * u, idesc); }
                        else if (c.ts) umma_f16_ts_acc(d, a, db + 2 * u, idesc);
                        else umma_f16_acc(d, da + 2 * u, db + 2 * u, idesc);
                    }
                    if (PAIR) umma_commit_2sm(bar + 8 * (r % DEPTH)); else umma_commit(bar + 8 * (r % DEPTH));
                }
                __syncwarp();
            }
            for (int r = c.rounds - DEPTH; r < c.rounds; r++) mbar_wait(bar + 8 * (r % DEPTH), (r / DEPTH) & 1);
            long long t1 = clock64();
            if (lane == 0) {
                g_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
                *stop = 1;
                if (PAIR) {      // the peer's side warps poll the flag in THEIR shared memory: set it through DSMEM
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote) : "r"(sb + 200 * 1024 + 512));
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // side traffic: tcgen05.ld of the (live) accumulator columns, like the epilogue of the previous tile
        if (warp - 4 < c.ld_warps) {
            const uint32_t la = tm + ((uint32_t)((warp & 3) * 32) << 16) + 192;
            uint32_t v[32], sink = 0;
            unsigned long long it = 0;
            while (!*stop) {
                for (int j = 0; j < 6; j++) { tmem_ld_32x32(la + j * 32, v); tmem_ld_wait(); sink += v[0] ^ v[31]; }
                it++;
            }
            if (sink == 0xdeadbeefu) g_cycles[255] = sink;
            if (lane == 0 && blockIdx.x == 0) g_side[warp - 4] = it * 6ull * 4096ull;        // bytes this warp read from TMEM
        }
    } else if (warp >= 8 && warp < 12) {
        if (warp - 8 < c.st_warps) {
            const uint32_t sa = tm + ((uint32_t)((warp & 3) * 32) << 16) + 480;      // last A stage: never read by the probe's MMAs? (r%4 == 3 reads it; data is irrelevant)
            uint32_t v[32];
            for (int j = 0; j < 32; j++) v[j] = 0x3c003c00u;
            unsigned long long it = 0;
            while (!*stop) { tmem_st_32x32(sa, v); tmem_st_wait(); it++; }
            if (lane == 0 && blockIdx.x == 0) g_side[4 + warp - 8] = it * 4096ull;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) {
        // the peer CTA's side warps poll its own `stop`: set it once the leader finished (cluster barrier orders the exit)
        cluster_sync_all();
    }
    if (warp == 2) { if (PAIR) tmem_dealloc_2sm(tm, 512); else tmem_dealloc(tm, 512); }
}

static double g_ld_bpc = 0, g_st_bpc = 0;      // TMEM bytes per cycle moved by the side warps of CTA 0
static double run(Cfg c, int sms) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_side, z, sizeof z);
    const size_t smem = 200 * 1024 + 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e;
    if (c.pair) e = launch_pdl(probe_kernel<1>, (unsigned)(sms & ~1), 384u, smem, 0, 2, c);
    else e = launch_pdl(probe_kernel<0>, (unsigned)sms, 384u, smem, 0, 1, c);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<unsigned long long> h(256);
    cudaMemcpyFromSymbol(h.data(), g_cycles, sizeof(unsigned long long) * 256);
    double mx = 0;
    for (int i = 0; i < sms; i += (c.pair ? 2 : 1)) mx = mx > (double)h[i] ? mx : (double)h[i];
    cudaMemcpyFromSymbol(z, g_side, sizeof z);
    g_ld_bpc = (double)(z[0] + z[1] + z[2] + z[3]) / (double)h[0];
    g_st_bpc = (double)(z[4] + z[5] + z[6] + z[7]) / (double)h[0];
    return mx / ((double)c.rounds * 4.0);      // cycles per UMMA on the slowest CTA
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    printf("%-6s %-4s %-3s %-5s %-3s %-3s | cycles/UMMA  ideal(N/2)  efficiency\n", "group", "N", "A", "nacc", "ld", "st");
    const int R = 4000;
    struct Row { int pair, N, ts, nacc, ld, st; };
    const Row rows[] = {
        {0, 192, 1, 1, 0, 0}, {0, 192, 0, 1, 0, 0}, {0, 128, 1, 1, 0, 0}, {0, 64, 1, 1, 0, 0},
        {1, 192, 1, 1, 0, 0}, {1, 192, 1, 2, 0, 0}, {1, 128, 1, 1, 0, 0}, {1, 128, 1, 2, 0, 0}, {1, 64, 1, 1, 0, 0}, {1, 96, 1, 1, 0, 0},
        {1, 192, 1, 1, 1, 0}, {1, 192, 1, 1, 2, 0}, {1, 192, 1, 1, 4, 0},
        {1, 192, 1, 1, 0, 1}, {1, 192, 1, 1, 0, 2}, {1, 192, 1, 1, 0, 4}, {1, 192, 1, 1, 4, 4},
        {0, 192, 1, 1, 4, 0}, {0, 192, 1, 1, 0, 4},
    };
    for (const Row& r : rows) {
        Cfg c{r.pair, r.N, r.ts, r.nacc, r.ld, r.st, R};
        run(c, p.multiProcessorCount);                       // warm-up
        const double cyc = run(c, p.multiProcessorCount);
        printf("%-6d %-4d %-3s %-5d %-3d %-3d | %10.1f  %9.1f  %9.3f   side: tcgen05.ld %.1f B/clk, tcgen05.st %.1f B/clk\n", r.pair ? 2 : 1, r.N, r.ts ? "tm" : "sm", r.nacc, r.ld, r.st, cyc, r.N / 2.0, (r.N / 2.0) / cyc, g_ld_bpc, g_st_bpc);
    }
    return 0;
}
