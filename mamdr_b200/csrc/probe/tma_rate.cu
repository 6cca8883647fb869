// Probe (not part of the library): per-SM load rate from an L2-resident buffer for the access patterns the pass kernel
// could use for its 128-row x 32-float K-major A chunks.
//   mode 0: TMA tensor 2D, box [128 rows x 32 floats], SWIZZLE_128B, row pitch 1 KB   (what the kernel does today)
//   mode 1: TMA tensor 2D, same box, rows contiguous (pitch 128 B: chunk-major global layout)
//   mode 2: 1D bulk copy cp.async.bulk of 16 KB contiguous
//   mode 3: LDG.128 by 128 threads (coalesced 4 rows x 128 B per warp instruction) + STS.128
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ -o tma_rate tma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tc_common.cuh"

constexpr int CHUNK = 128 * 32 * 4;   // 16 KB
constexpr int STAGES = 6;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, const float* base, int mode, int chunks_per_cta, int reps,
                                             long long* cycles_out, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) tc::mbar_init(&full[s], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    uint32_t it = 0;
    for (int r = 0; r < reps; ++r) {
        if (mode <= 2) {
            // producer issues STAGES chunks, everybody waits for each, re-issue (no consumer work): pure load rate
            for (int c0 = 0; c0 < chunks_per_cta; c0 += STAGES) {
                if (tid == 0) {
                    for (int s = 0; s < STAGES && c0 + s < chunks_per_cta; ++s) {
                        const int c = c0 + s;
                        tc::mbar_arrive_expect_tx(&full[s], CHUNK);
                        if (mode == 0) tc::tma_load_2d(smem + s * CHUNK, &map, &full[s], (c % 8) * 32, blockIdx.x * 128);
                        else if (mode == 1) tc::tma_load_2d(smem + s * CHUNK, &map, &full[s], 0, (blockIdx.x * 8 + (c % 8)) * 128);
                        else {
                            const float* src = base + ((size_t)blockIdx.x * 8 + (c % 8)) * (CHUNK / 4);
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                             tc::smem_u32(smem + s * CHUNK)),
                                         "l"(src), "r"(CHUNK), "r"(tc::smem_u32(&full[s]))
                                         : "memory");
                        }
                    }
                }
                for (int s = 0; s < STAGES && c0 + s < chunks_per_cta; ++s) tc::mbar_wait(&full[s], (it / 1) & 1);
                ++it;
                __syncthreads();
            }
        } else {
            for (int c = 0; c < chunks_per_cta; ++c) {
                const float* src = base + ((size_t)blockIdx.x * 128) * 256 + (c % 8) * 32;   // row pitch 256 floats
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int f = tid + 128 * i, row = f >> 3, j = f & 7;
                    v[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)row * 256 + j * 4));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int f = tid + 128 * i, row = f >> 3, j = f & 7;
                    *reinterpret_cast<float4*>(smem + (c % STAGES) * CHUNK + row * 128 + ((j ^ (row & 7)) << 4)) = v[i];
                }
            }
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    acc += reinterpret_cast<float*>(smem)[tid];
    if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const int rows = 148 * 128, cols = 256;
    float* buf;
    cudaMalloc(&buf, (size_t)rows * cols * 4);
    cudaMemset(buf, 0, (size_t)rows * cols * 4);
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    float* sink;
    cudaMalloc(&sink, 4);
    CUtensorMap m0, m1;
    tc::make_tmap_2d_f32(&m0, buf, rows, cols, cols, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    tc::make_tmap_2d_f32(&m1, buf, (uint64_t)rows * 8, 32, 32, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK + 1024);
    const int chunks = 48, reps = 20;
    for (int mode = 0; mode < 4; ++mode)
        for (int grid : {1, 8, 32, 64, 148}) {
            probe<<<grid, 128, STAGES * CHUNK + 1024>>>(mode == 1 ? m1 : m0, buf, mode, chunks, reps, cyc, sink);   // warm L2
            probe<<<grid, 128, STAGES * CHUNK + 1024>>>(mode == 1 ? m1 : m0, buf, mode, chunks, reps, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (auto x : h) mx = x > mx ? x : mx;
            const double bytes = (double)chunks * reps * CHUNK;
            printf("mode %d grid %3d: %6.1f B/clk/SM  (%.2f us per 16 KB chunk @1.965 GHz)  %s\n", mode, grid, bytes / mx,
                   mx / (double)(chunks * reps) / 1965.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
