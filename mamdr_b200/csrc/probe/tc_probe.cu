// Stand-alone probe (not part of libmamdr_b200.so): validates the hand-written tcgen05 / TMA building
// blocks on a real B200 before they are composed into the fused tower kernels.
//   D[128, N] = A . B^T   A: [128, K] K-major or [K, 128] MN-major;  B: [N, K] K-major
//   one CTA, all K chunks resident in smem, kind::tf32, fp32 accumulate in TMEM.
// Prints max errors against (i) exact product, (ii) product of tf32-truncated inputs, (iii) tf32-rounded
// inputs, and the accuracy of the 3xTF32 split.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tc_common.cuh"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int KCH = 32;  // floats per 128-byte swizzle row

// smem: A chunks then B chunks, each chunk 1024-aligned.
//  A K-major chunk : [128 rows][32 k]  (16 KB)      SBO = 1024 (8-row groups)
//  A MN-major chunk: 4 x [32 k][32 m] (4 x 4 KB)    LBO = 4096 (m groups), SBO = 1024 (8-k groups)
//  B K-major chunk : [N rows][32 k]
template <bool A_MN>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                                                    int N, int K, int passes, float* D, uint32_t mn_lt, uint32_t mn_lbo, uint32_t mn_sbo, uint32_t mn_kadv) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nch = K / KCH;
    const uint32_t a_bytes = 128 * KCH * 4, b_bytes = N * KCH * 4;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)nch * a_bytes;
    unsigned char* sA2 = sB + (size_t)nch * b_bytes;   // "lo" operands for the 3xTF32 pass
    unsigned char* sB2 = sA2 + (size_t)nch * a_bytes;
    if (tid == 0) {
        tc::mbar_init(&bar_load, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base, 256);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t total = (uint32_t)nch * (a_bytes + b_bytes) * (passes >= 3 ? 2 : 1);
        tc::mbar_arrive_expect_tx(&bar_load, total);
        for (int c = 0; c < nch; ++c) {
            if (A_MN) {
                for (int g = 0; g < 4; ++g) tc::tma_load_2d(sA + c * a_bytes + g * 4096, &tmA, &bar_load, g * 32, c * KCH);
            } else {
                tc::tma_load_2d(sA + c * a_bytes, &tmA, &bar_load, c * KCH, 0);
            }
            tc::tma_load_2d(sB + c * b_bytes, &tmB, &bar_load, c * KCH, 0);
            if (passes >= 3) {
                if (A_MN) {
                    for (int g = 0; g < 4; ++g) tc::tma_load_2d(sA2 + c * a_bytes + g * 4096, &tmA2, &bar_load, g * 32, c * KCH);
                } else {
                    tc::tma_load_2d(sA2 + c * a_bytes, &tmA2, &bar_load, c * KCH, 0);
                }
                tc::tma_load_2d(sB2 + c * b_bytes, &tmB2, &bar_load, c * KCH, 0);
            }
        }
        tc::mbar_wait(&bar_load, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc_tf32(128, N, A_MN ? 1 : 0, 0);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
            // pass 0: hi*hi ; pass 1: hi*lo ; pass 2: lo*hi
            unsigned char* pa = (pass >= 2) ? sA2 : sA;
            unsigned char* pb = (pass == 1 || pass == 3) ? sB2 : sB;
            for (int c = 0; c < nch; ++c) {
                for (int k = 0; k < KCH / 8; ++k) {
                    uint64_t da, db;
                    if (A_MN)
                        da = tc::make_smem_desc(tc::smem_u32(pa + c * a_bytes) + k * mn_kadv, mn_lbo, mn_sbo, mn_lt);
                    else
                        da = tc::make_smem_desc(tc::smem_u32(pa + c * a_bytes) + k * 32, 16, 1024, tc::kSwizzle128B);
                    db = tc::make_smem_desc(tc::smem_u32(pb + c * b_bytes) + k * 32, 16, 1024, tc::kSwizzle128B);
                    tc::mma_tf32(tmem, da, db, idesc, acc);
                    acc = 1;
                }
            }
        }
        tc::mma_commit(&bar_mma);
    }
    __syncthreads();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 256);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x00000FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}

template <bool A_MN>
int run(int N, int K, int passes, CUtensorMapSwizzle mn_swz = CU_TENSOR_MAP_SWIZZLE_128B, uint32_t mn_lt = 2, uint32_t mn_lbo = 4096,
        uint32_t mn_sbo = 1024, uint32_t mn_kadv = 1024) {
    const int M = 128;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), Alo(A.size()), Blo(B.size()), Ahi(A.size()), Bhi(B.size());
    srand(1234 + N + K);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    const bool rn = passes == 4 || passes == 13;
    if (passes == 13) passes = 3;
    for (size_t i = 0; i < A.size(); ++i) { Ahi[i] = rn ? tf32_rn(A[i]) : tf32_trunc(A[i]); Alo[i] = tf32_rn(A[i] - Ahi[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bhi[i] = rn ? tf32_rn(B[i]) : tf32_trunc(B[i]); Blo[i] = tf32_rn(B[i] - Bhi[i]); }
    const std::vector<float>& Asrc = rn ? Ahi : A;
    const std::vector<float>& Bsrc = rn ? Bhi : B;
    // device layouts: A K-major = [M][K]; A MN-major = [K][M]
    std::vector<float> Adev(A.size()), Alodev(A.size());
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            const size_t dst = A_MN ? (size_t)k * M + m : (size_t)m * K + k;
            Adev[dst] = Asrc[(size_t)m * K + k];
            Alodev[dst] = Alo[(size_t)m * K + k];
        }
    float *dA, *dB, *dA2, *dB2, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dA2, A.size() * 4)); CK(cudaMalloc(&dB2, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, Adev.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA2, Alodev.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bsrc.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, Blo.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tA, tB, tA2, tB2;
    bool ok;
    if (A_MN) {
        ok = tc::make_tmap_2d_f32(&tA, dA, K, M, M, KCH, 32, mn_swz) &&
             tc::make_tmap_2d_f32(&tA2, dA2, K, M, M, KCH, 32, mn_swz);
    } else {
        ok = tc::make_tmap_2d_f32(&tA, dA, M, K, K, 128, KCH, CU_TENSOR_MAP_SWIZZLE_128B) &&
             tc::make_tmap_2d_f32(&tA2, dA2, M, K, K, 128, KCH, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    ok = ok && tc::make_tmap_2d_f32(&tB, dB, N, K, K, N, KCH, CU_TENSOR_MAP_SWIZZLE_128B) &&
         tc::make_tmap_2d_f32(&tB2, dB2, N, K, K, N, KCH, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { printf("tensor map creation failed\n"); return 1; }
    const size_t smem = (size_t)(K / KCH) * (128 * KCH * 4 + N * KCH * 4) * (passes >= 3 ? 2 : 1) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<A_MN><<<1, 128, smem>>>(tA, tB, tA2, tB2, N, K, passes, dD, mn_lt, mn_lbo, mn_sbo, mn_kadv);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double e_exact = 0, e_trunc = 0, e_rn = 0, ref_max = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, st = 0, sr = 0;
            for (int k = 0; k < K; ++k) {
                const float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
                s += (double)a * b;
                st += (double)tf32_trunc(a) * tf32_trunc(b);
                sr += (double)tf32_rn(a) * tf32_rn(b);
            }
            const double d = D[(size_t)m * N + n];
            e_exact = fmax(e_exact, fabs(d - s)); e_trunc = fmax(e_trunc, fabs(d - st)); e_rn = fmax(e_rn, fabs(d - sr));
            ref_max = fmax(ref_max, fabs(s));
        }
    if (A_MN) printf("[swz=%d lt=%u lbo=%u sbo=%u kadv=%u] ", (int)mn_swz, mn_lt, mn_lbo, mn_sbo, mn_kadv);
    printf("A_%s N=%3d K=%3d passes=%d : max|D-exact|=%.3e  max|D-trunc|=%.3e  max|D-rn|=%.3e  (max|ref|=%.2f)\n",
           A_MN ? "MN" : "K ", N, K, passes, e_exact, e_trunc, e_rn, ref_max);
    cudaFree(dA); cudaFree(dB); cudaFree(dA2); cudaFree(dB2); cudaFree(dD);
    return 0;
}

int main() {
    run<false>(64, 32, 1);
    run<false>(64, 128, 1);
    run<false>(256, 128, 1);
    run<false>(16, 64, 1);
    run<false>(32, 128, 1);
    run<false>(64, 128, 3);
    const CUtensorMapSwizzle S32 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    run<true>(64, 32, 1, S32, 1, 4096, 512, 1024);
    run<true>(64, 128, 1, S32, 1, 4096, 512, 1024);
    run<true>(64, 128, 1, S32, 1, 512, 4096, 1024);
    run<true>(64, 128, 1, S32, 1, 4096, 1024, 1024);
    run<true>(64, 128, 1, S32, 2, 4096, 512, 1024);
    run<true>(64, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B, 1, 4096, 512, 1024);
    run<true>(32, 128, 3, S32, 1, 4096, 512, 1024);
    printf("-- K-major, 3-pass trunc-hi / 3-pass rn-hi (13) / 4-pass rn-hi\n");
    run<false>(64, 128, 3);
    run<false>(64, 128, 13);
    run<false>(64, 128, 4);
    run<false>(32, 128, 3);
    run<false>(32, 128, 13);
    run<false>(32, 128, 4);
    return 0;
}
