// Probe (not part of the library): rate of the pass kernel's TMA -> (split) -> tcgen05.mma mainloop for the 3xTF32
// schemes under consideration, with the operands streamed from an L2-resident buffer by `grid` CTAs at once.
//   mode 0: r1 scheme     -- TMA A (128x32 K-major) + B (32 x bn MN-major); 128 workers derive the lo halves in shared
//                            memory; 3 MMAs per k-step (A.B, A.B_lo, A_lo.B); 3 stages of 48 KB
//   mode 1: pre-split     -- TMA brings A, A_lo, B, B_lo (lo arrays live in global memory); 3 MMAs per k-step
//   mode 2: pre-split + N-concatenation -- B and B_lo adjacent in smem form ONE operand of N = 2 bn:
//                            2 MMAs per k-step (A.[B|B_lo] -> two accumulator halves, A_lo.B -> first half)
//   mode 3: 1-pass tf32   -- TMA A + B, 1 MMA per k-step, 6 stages
//   mode 4: operands resident (no TMA, no waits): 3 MMAs per k-step  -- pure issue rate
//   mode 5: operands resident: N-concatenated 2 MMAs per k-step
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mainloop_rate mainloop_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tc_common.cuh"

constexpr int KCH = 32;
constexpr int A_BYTES = 128 * KCH * 4;   // 16 KB

__device__ __forceinline__ float tf32_lo(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t u = __float_as_uint(x - hi);
    u += 0x00000FFFu + ((u >> 13) & 1u);
    return __uint_as_float(u & 0xFFFFE000u);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct Maps { CUtensorMap a, alo, b, blo; };

__global__ void __launch_bounds__(192, 1)
probe(const __grid_constant__ Maps maps, int mode, int bn, int nch, int jobs, int stages, int conv, int flags, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[8], bar_empty[8], bar_split[8], bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B_BYTES = bn * KCH * 4;
    const bool split_smem = mode == 0, presplit = mode == 1 || mode == 2, resident = mode >= 4;
    const bool concat = mode == 2 || mode == 5;
    const int passes = mode == 3 ? 1 : 3;
    // stage layout: [A | A_lo | B | B_lo]  (B_lo right behind B: the concatenated operand)
    const uint32_t stage_bytes = mode == 3 ? (uint32_t)(A_BYTES + B_BYTES) : (uint32_t)(2 * A_BYTES + 2 * B_BYTES);
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); tc::mbar_init(&bar_split[s], 128); }
        tc::mbar_init(&bar_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc(&tmem_base_s, 256);
    for (int i = tid; i < (int)(stages * stage_bytes) / 4; i += 192) reinterpret_cast<float*>(smem)[i] = (flags & 1) ? 0.f : 1.0f + 1e-4f * (i & 1023);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int rt = (blockIdx.x / 8) % 8, ntile = blockIdx.x % 8;
    uint32_t it = 0;
    long long t0 = 0, t_first = 0;
    if (tid == 0) t0 = clock64();
    __syncthreads();
    for (int j = 0; j < jobs; ++j) {
        if (warp == 4) {
            if (conv && !resident) {
                // warp-converged producer: every lane runs the loop, one elected lane issues (uniform-register operands)
                const uint32_t tx = (uint32_t)(A_BYTES + B_BYTES) * (presplit ? 2u : 1u);
                for (int i = 0; i < nch; ++i, ++it) {
                    const int s = it % stages, c = i % 8;
                    if (it >= (uint32_t)stages) tc::mbar_wait(&bar_empty[s], ((it / stages) - 1) & 1);
                    unsigned char* st = smem + (size_t)s * stage_bytes;
                    unsigned char* sA = st;
                    unsigned char* sB = mode == 3 ? st + A_BYTES : st + 2 * A_BYTES;
                    if (elect_one()) {
                        tc::mbar_arrive_expect_tx(&bar_full[s], tx);
                        tc::tma_load_2d(sA, &maps.a, &bar_full[s], c * KCH, rt * 128);
                        for (int g = 0; g < bn / 32; ++g) tc::tma_load_2d(sB + g * 4096, &maps.b, &bar_full[s], ntile * 32 + g * 32, c * KCH);
                        if (presplit) {
                            tc::tma_load_2d(sA + A_BYTES, &maps.alo, &bar_full[s], c * KCH, rt * 128);
                            for (int g = 0; g < bn / 32; ++g) tc::tma_load_2d(sB + B_BYTES + g * 4096, &maps.blo, &bar_full[s], ntile * 32 + g * 32, c * KCH);
                        }
                    }
                    __syncwarp();
                }
            } else if (lane == 0 && !resident) {
                const uint32_t tx = (uint32_t)(A_BYTES + B_BYTES) * (presplit ? 2u : 1u);
                for (int i = 0; i < nch; ++i, ++it) {
                    const int s = it % stages, c = i % 8;
                    if (it >= (uint32_t)stages) tc::mbar_wait(&bar_empty[s], ((it / stages) - 1) & 1);
                    unsigned char* st = smem + (size_t)s * stage_bytes;
                    unsigned char* sA = st;
                    unsigned char* sB = mode == 3 ? st + A_BYTES : st + 2 * A_BYTES;
                    tc::mbar_arrive_expect_tx(&bar_full[s], tx);
                    tc::tma_load_2d(sA, &maps.a, &bar_full[s], c * KCH, rt * 128);
                    for (int g = 0; g < bn / 32; ++g) tc::tma_load_2d(sB + g * 4096, &maps.b, &bar_full[s], ntile * 32 + g * 32, c * KCH);
                    if (presplit) {
                        tc::tma_load_2d(sA + A_BYTES, &maps.alo, &bar_full[s], c * KCH, rt * 128);
                        for (int g = 0; g < bn / 32; ++g) tc::tma_load_2d(sB + B_BYTES + g * 4096, &maps.blo, &bar_full[s], ntile * 32 + g * 32, c * KCH);
                    }
                }
            }
        } else if (warp == 5 && conv) {
            // warp-converged MMA issuer
            const int bk = (flags & 2) ? 1 : 0;
            const uint32_t idesc = tc::make_idesc_tf32(128, bn, 0, bk ? 0 : 1);
            const uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * bn, 0, bk ? 0 : 1);
            uint32_t acc = 0;
            for (int i = 0; i < nch; ++i, ++it) {
                const int s = it % stages;
                if (!resident) {
                    tc::mbar_wait(&bar_full[s], (it / stages) & 1);
                    if (split_smem) tc::mbar_wait(&bar_split[s], (it / stages) & 1);
                    tc::tc_fence_after();
                }
                const uint32_t st = tc::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t aA = st, aAlo = st + A_BYTES;
                const uint32_t aB = mode == 3 ? st + A_BYTES : st + 2 * A_BYTES, aBlo = aB + B_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KCH / 8; ++k) {
                        const uint64_t dA = tc::make_smem_desc(aA + k * 32, 16, 1024, tc::kSwizzle128B);
                        const uint64_t dAlo = tc::make_smem_desc(aAlo + k * 32, 16, 1024, tc::kSwizzle128B);
                        const uint64_t dB = bk ? tc::make_smem_desc(aB + k * 32, 16, 1024, tc::kSwizzle128B) : tc::make_smem_desc(aB + k * 1024, 4096, 512, 1);
                        const uint64_t dBlo = bk ? tc::make_smem_desc(aBlo + k * 32, 16, 1024, tc::kSwizzle128B) : tc::make_smem_desc(aBlo + k * 1024, 4096, 512, 1);
                        if (flags & 4) {
                            tc::mma_tf32(tmem, dA, dB, idesc, acc);
                        } else if (flags & 8) {
                            tc::mma_tf32(tmem, dA, dB, idesc, acc);
                            tc::mma_tf32(tmem + 64, dA, dBlo, idesc, acc);
                            tc::mma_tf32(tmem + 128, dAlo, dB, idesc, acc);
                        } else if (concat) {
                            tc::mma_tf32(tmem, dA, dB, idesc2, acc);
                            tc::mma_tf32(tmem, dAlo, dB, idesc, 1);
                        } else {
                            tc::mma_tf32(tmem, dA, dB, idesc, acc);
                            if (passes == 3) { tc::mma_tf32(tmem, dA, dBlo, idesc, 1); tc::mma_tf32(tmem, dAlo, dB, idesc, 1); }
                        }
                        acc = 1;
                    }
                    if (!resident) tc::mma_commit(&bar_empty[s]);
                }
                __syncwarp();
            }
            if (elect_one()) tc::mma_commit(&bar_done);
            __syncwarp();
            tc::mbar_wait(&bar_done, j & 1);
        } else if (warp == 5) {
            if (lane == 0) {
                const uint32_t idesc = tc::make_idesc_tf32(128, bn, 0, 1);
                const uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * bn, 0, 1);
                uint32_t acc = 0;
                for (int i = 0; i < nch; ++i, ++it) {
                    const int s = it % stages;
                    if (!resident) {
                        tc::mbar_wait(&bar_full[s], (it / stages) & 1);
                        if (split_smem) tc::mbar_wait(&bar_split[s], (it / stages) & 1);
                        tc::tc_fence_after();
                    }
                    if (j == 0 && i == 0) t_first = clock64();
                    const uint32_t st = tc::smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t aA = st, aAlo = st + A_BYTES;
                    const uint32_t aB = mode == 3 ? st + A_BYTES : st + 2 * A_BYTES, aBlo = aB + B_BYTES;
#pragma unroll
                    for (int k = 0; k < KCH / 8; ++k) {
                        const uint64_t dA = tc::make_smem_desc(aA + k * 32, 16, 1024, tc::kSwizzle128B);
                        const uint64_t dAlo = tc::make_smem_desc(aAlo + k * 32, 16, 1024, tc::kSwizzle128B);
                        const uint64_t dB = tc::make_smem_desc(aB + k * 1024, 4096, 512, 1);
                        const uint64_t dBlo = tc::make_smem_desc(aBlo + k * 1024, 4096, 512, 1);
                        if (concat) {
                            tc::mma_tf32(tmem, dA, dB, idesc2, acc);      // [A.B | A.B_lo]
                            tc::mma_tf32(tmem, dAlo, dB, idesc, 1);       // += A_lo.B on the first half
                        } else {
                            tc::mma_tf32(tmem, dA, dB, idesc, acc);
                            if (passes == 3) { tc::mma_tf32(tmem, dA, dBlo, idesc, 1); tc::mma_tf32(tmem, dAlo, dB, idesc, 1); }
                        }
                        acc = 1;
                    }
                    if (!resident) tc::mma_commit(&bar_empty[s]);
                }
                tc::mma_commit(&bar_done);
                tc::mbar_wait(&bar_done, j & 1);
            }
        } else if (split_smem) {
            const int nv = (A_BYTES + B_BYTES) / 16;
            for (int i = 0; i < nch; ++i, ++it) {
                const int s = it % stages;
                tc::mbar_wait(&bar_full[s], (it / stages) & 1);
                float4* hiA = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
                float4* hiB = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + 2 * A_BYTES);
                for (int q = tid; q < nv; q += 128) {
                    float4* src = q < A_BYTES / 16 ? hiA + q : hiB + (q - A_BYTES / 16);
                    float4* dst = q < A_BYTES / 16 ? hiA + A_BYTES / 16 + q : hiB + B_BYTES / 16 + (q - A_BYTES / 16);
                    const float4 x = *src;
                    *dst = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bar_split[s]);
            }
        }
        __syncthreads();   // job boundary (the real kernel has an epilogue here)
    }
    if (tid == 160) { out[2 * blockIdx.x + 1] = t_first; }
    __syncthreads();
    if (tid == 0) { out[2 * blockIdx.x] = clock64() - t0; }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem, 256);
}

int main() {
    const int rows = 1024, K = 256, N = 256;
    float *X, *Xlo, *W, *Wlo;
    cudaMalloc(&X, (size_t)rows * K * 4); cudaMalloc(&Xlo, (size_t)rows * K * 4);
    cudaMalloc(&W, (size_t)K * N * 4); cudaMalloc(&Wlo, (size_t)K * N * 4);
    cudaMemset(X, 0, (size_t)rows * K * 4); cudaMemset(Xlo, 0, (size_t)rows * K * 4);
    cudaMemset(W, 0, (size_t)K * N * 4); cudaMemset(Wlo, 0, (size_t)K * N * 4);
    long long* d;
    cudaMalloc(&d, 148 * 16);
    Maps m;
    bool ok = tc::make_tmap_2d_f32(&m.a, X, rows, K, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              tc::make_tmap_2d_f32(&m.alo, Xlo, rows, K, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              tc::make_tmap_2d_f32(&m.b, W, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
              tc::make_tmap_2d_f32(&m.blo, Wlo, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (!ok) { printf("tensor map creation failed\n"); return 1; }
    const int smem_max = 200 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    const char* names[] = {"r1: smem split, 3 MMA", "pre-split, 3 MMA", "pre-split, N-concat 2 MMA", "1-pass tf32", "resident 3 MMA", "resident concat 2 MMA"};
    struct T { int mode, bn, stages, flags; const char* what; };
    const T tests[] = {
        {4, 32, 5, 0, "resident 3 MMA, 5 stages, nonzero data"},
        {4, 32, 5, 1, "resident 3 MMA, 5 stages, ZERO data"},
        {4, 32, 1, 0, "resident 3 MMA, 1 stage (same addresses)"},
        {4, 32, 5, 2, "resident 3 MMA, B K-major"},
        {4, 32, 5, 4, "resident 1 MMA per k-step (fresh A and B each)"},
        {4, 32, 1, 4, "resident 1 MMA per k-step, 1 stage"},
        {4, 32, 5, 6, "resident 1 MMA per k-step, B K-major"},
        {4, 32, 5, 8, "resident 3 MMA into 3 different accumulators"},
        {4, 64, 4, 8, "resident 3 MMA into 3 different accumulators bn 64"},
        {4, 64, 4, 4, "resident 1 MMA per k-step bn 64"},
        {5, 32, 5, 0, "resident concat 2 MMA"},
        {5, 64, 4, 0, "resident concat 2 MMA bn 64"},
    };
    for (const T& t : tests) {
        const int nch = 64, jobs = 4;
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 192, smem_max>>>(m, t.mode, t.bn, nch, jobs, t.stages, 1, t.flags, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-52s: %6.0f cyc/chunk = %5.1f cyc/k-step  %s\n", t.what, h[0] / (double)(jobs * nch), h[0] / (double)(jobs * nch * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
