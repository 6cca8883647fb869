// Probe (not part of the library): the pass kernel's mainloop with WARP-CONVERGED issue (whole warp runs the loop,
// one elected lane issues; descriptors live in uniform registers, no R2UR waterfall per instruction) and lean
// descriptor arithmetic, for the 3xTF32 schemes:
//   SCHEME 0: 1-pass tf32                        (A, B streamed)
//   SCHEME 1: pre-split, 3 MMAs per k-step      (A, A_lo, B, B_lo streamed)
//   SCHEME 2: pre-split, N-concatenated 2 MMAs  (A.[B|B_lo] N = 2 bn, then A_lo.B N = bn)
// RESIDENT = 1: no TMA, no waits (pure issue rate).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mainloop2 mainloop2.cu
#include <cstdio>
#include <vector>
#include "../tc_common.cuh"

constexpr int KCH = 32;
constexpr int A_BYTES = 128 * KCH * 4;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t mk(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint32_t HI_K = (1024u >> 4) | (1u << 14) | (2u << 29);    // K-major SW128, SBO 1024
constexpr uint32_t LO_K = (16u >> 4) << 16;
constexpr uint32_t HI_MN = (512u >> 4) | (1u << 14) | (1u << 29);    // MN-major SW128 / 32B atoms, SBO 512
constexpr uint32_t LO_MN = (4096u >> 4) << 16;

struct Maps { CUtensorMap a, alo, b, blo; };

template <int BN, int SCHEME, int RESIDENT>
__global__ void __launch_bounds__(192, 1) probe(const __grid_constant__ Maps maps, int nch, int jobs, int stages, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[8], bar_empty[8], bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int B_BYTES = BN * KCH * 4;
    constexpr uint32_t stage_bytes = SCHEME == 0 ? (A_BYTES + B_BYTES) : (2 * A_BYTES + 2 * B_BYTES);
    constexpr uint32_t offB = SCHEME == 0 ? A_BYTES : 2 * A_BYTES;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
        tc::mbar_init(&bar_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc(&tmem_base_s, 256);
    for (int i = tid; i < (int)(stages * stage_bytes) / 4; i += 192) reinterpret_cast<float*>(smem)[i] = 1.0f + 1e-4f * (i & 1023);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int rt = (blockIdx.x / 8) % 8, ntile = blockIdx.x % 8;
    long long t0 = 0;
    if (tid == 0) t0 = clock64();
    __syncthreads();
    int ps = 0, ms = 0;            // producer / MMA stage cursors
    uint32_t pph = 0, mph = 0;     // their phase bits
    uint32_t pfill = 0;            // chunks issued so far (the first `stages` need no empty-wait)
    for (int j = 0; j < jobs; ++j) {
        if (warp == 4 && !RESIDENT) {
            constexpr uint32_t tx = (uint32_t)(A_BYTES + B_BYTES) * (SCHEME == 0 ? 1u : 2u);
            for (int i = 0; i < nch; ++i) {
                const int c = i & 7;
                if (pfill >= (uint32_t)stages) tc::mbar_wait(&bar_empty[ps], pph ^ 1);
                unsigned char* st = smem + (size_t)ps * stage_bytes;
                if (elect_one()) {
                    tc::mbar_arrive_expect_tx(&bar_full[ps], tx);
                    tc::tma_load_2d(st, &maps.a, &bar_full[ps], c * KCH, rt * 128);
#pragma unroll
                    for (int g = 0; g < BN / 32; ++g) tc::tma_load_2d(st + offB + g * 4096, &maps.b, &bar_full[ps], ntile * 32 + g * 32, c * KCH);
                    if (SCHEME != 0) {
                        tc::tma_load_2d(st + A_BYTES, &maps.alo, &bar_full[ps], c * KCH, rt * 128);
#pragma unroll
                        for (int g = 0; g < BN / 32; ++g) tc::tma_load_2d(st + offB + B_BYTES + g * 4096, &maps.blo, &bar_full[ps], ntile * 32 + g * 32, c * KCH);
                    }
                }
                __syncwarp();
                ++pfill;
                if (++ps == stages) { ps = 0; pph ^= 1; }
            }
        } else if (warp == 5) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, 0, 1);
            constexpr uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * BN, 0, 1);
            const uint32_t base = tc::smem_u32(smem);
            uint32_t acc = 0;
            for (int i = 0; i < nch; ++i) {
                if (!RESIDENT) {
                    tc::mbar_wait(&bar_full[ms], mph);
                    tc::tc_fence_after();
                }
                const uint32_t st = (base + ms * stage_bytes) >> 4;
                if (elect_one()) {
                    const uint32_t a = st | LO_K, alo = a + (A_BYTES >> 4);
                    const uint32_t b = (st + (offB >> 4)) | LO_MN, blo = b + (B_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < KCH / 8; ++k) {
                        if (SCHEME == 2) {
                            tc::mma_tf32(tmem, mk(a + k * 2, HI_K), mk(b + k * 64, HI_MN), idesc2, acc);
                            tc::mma_tf32(tmem, mk(alo + k * 2, HI_K), mk(b + k * 64, HI_MN), idesc, 1);
                        } else {
                            tc::mma_tf32(tmem, mk(a + k * 2, HI_K), mk(b + k * 64, HI_MN), idesc, acc);
                            if (SCHEME == 1) {
                                tc::mma_tf32(tmem, mk(a + k * 2, HI_K), mk(blo + k * 64, HI_MN), idesc, 1);
                                tc::mma_tf32(tmem, mk(alo + k * 2, HI_K), mk(b + k * 64, HI_MN), idesc, 1);
                            }
                        }
                        acc = 1;
                    }
                    if (!RESIDENT) tc::mma_commit(&bar_empty[ms]);
                }
                __syncwarp();
                if (++ms == stages) { ms = 0; mph ^= 1; }
            }
            if (elect_one()) tc::mma_commit(&bar_done);
            __syncwarp();
            tc::mbar_wait(&bar_done, j & 1);
        }
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem, 256);
}

static Maps g_maps;
static long long* g_out;

template <int BN, int SCHEME, int RESIDENT>
void run(const char* what, int grid, int nch) {
    constexpr int stage_bytes = SCHEME == 0 ? A_BYTES + BN * 128 : 2 * A_BYTES + 2 * BN * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    const int smem_max = 200 * 1024 + 1024;
    cudaFuncSetAttribute(probe<BN, SCHEME, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    const int jobs = nch == 8 ? 32 : 4;
    for (int rep = 0; rep < 2; ++rep) probe<BN, SCHEME, RESIDENT><<<grid, 192, smem_max>>>(g_maps, nch, jobs, stages, g_out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), g_out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s bn %2d grid %3d nch %2d stages %d: %6.0f cyc/job %5.0f cyc/chunk (%.2f us/job) %s\n", what, BN, grid, nch, stages,
           mx / (double)jobs, mx / (double)(jobs * nch), mx / (double)jobs / 1965.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int rows = 1024, K = 256, N = 256;
    float *X, *Xlo, *W, *Wlo;
    cudaMalloc(&X, (size_t)rows * K * 4); cudaMalloc(&Xlo, (size_t)rows * K * 4);
    cudaMalloc(&W, (size_t)K * N * 4); cudaMalloc(&Wlo, (size_t)K * N * 4);
    cudaMemset(X, 0, (size_t)rows * K * 4); cudaMemset(Xlo, 0, (size_t)rows * K * 4);
    cudaMemset(W, 0, (size_t)K * N * 4); cudaMemset(Wlo, 0, (size_t)K * N * 4);
    cudaMalloc(&g_out, 148 * 8);
    bool ok = tc::make_tmap_2d_f32(&g_maps.a, X, rows, K, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              tc::make_tmap_2d_f32(&g_maps.alo, Xlo, rows, K, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              tc::make_tmap_2d_f32(&g_maps.b, W, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
              tc::make_tmap_2d_f32(&g_maps.blo, Wlo, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (!ok) { printf("tensor map creation failed\n"); return 1; }
    for (int nch : {64, 8}) {
        run<32, 0, 1>("resident 1-pass", 1, nch);
        run<32, 1, 1>("resident 3 MMA", 1, nch);
        run<32, 2, 1>("resident concat 2 MMA", 1, nch);
        run<64, 1, 1>("resident 3 MMA", 1, nch);
        run<64, 2, 1>("resident concat 2 MMA", 1, nch);
        for (int grid : {1, 64, 148}) {
            run<32, 0, 0>("TMA 1-pass", grid, nch);
            run<32, 1, 0>("TMA pre-split 3 MMA", grid, nch);
            run<32, 2, 0>("TMA pre-split concat 2 MMA", grid, nch);
            run<64, 0, 0>("TMA 1-pass", grid, nch);
            run<64, 1, 0>("TMA pre-split 3 MMA", grid, nch);
            run<64, 2, 0>("TMA pre-split concat 2 MMA", grid, nch);
        }
    }
    return 0;
}
