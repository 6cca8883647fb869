// Probe: phase timeline (globaltimer) of tcg::gemm_kernel on dH- and fwd-like shapes.
#define TC_TIMING
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "../tc_gemm.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct PlainEpi {
    struct State {};
    float* out; int ld;
    __device__ void begin(State&, int, bool) const {}
    __device__ void cols(State&, int row, bool valid, int col0, int, float* v) const {
        if (!valid) return;
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + (size_t)row * ld + col0 + j) = make_float4(v[j], v[j+1], v[j+2], v[j+3]);
    }
    __device__ void end(State&, int, bool, int, int, unsigned char*) const {}
};

template <int BN, bool B_MN>
void run(const char* name, int M, int N, int K, int passes) {
    float *A, *B, *C;
    CK(cudaMalloc(&A, (size_t)M * K * 4)); CK(cudaMalloc(&B, (size_t)N * K * 4)); CK(cudaMalloc(&C, (size_t)M * N * 4));
    CK(cudaMemset(A, 0, (size_t)M * K * 4)); CK(cudaMemset(B, 0, (size_t)N * K * 4));
    tcg::Maps maps;
    bool ok = tc::make_tmap_2d_f32(&maps.a, A, M, K, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (B_MN) ok = ok && tc::make_tmap_2d_f32(&maps.b, B, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    else ok = ok && tc::make_tmap_2d_f32(&maps.b, B, N, K, K, BN, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    maps.a_lo = maps.a; maps.b_lo = maps.b;
    if (!ok) { printf("tmap fail\n"); exit(1); }
    auto kern = tcg::gemm_kernel<BN, false, B_MN, 4, PlainEpi>;
    const size_t smem = tcg::smem_bytes<BN, 4>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(N / BN, (M + 127) / 128, 1);
    const int nct = grid.x * grid.y;
    unsigned long long* tbuf;
    CK(cudaMalloc(&tbuf, (size_t)nct * 8 * 8));
    CK(cudaMemcpyToSymbol(tcg::g_tc_timing, &tbuf, sizeof(tbuf)));
    PlainEpi epi{C, N};
    tcg::SplitK sk{nullptr, nullptr};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) kern<<<grid, 128, smem>>>(maps, M, K / 32, K / 32, passes, sk, epi);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 50; ++i) kern<<<grid, 128, smem>>>(maps, M, K / 32, K / 32, passes, sk, epi);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h((size_t)nct * 8);
    CK(cudaMemcpy(h.data(), tbuf, h.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull, t6 = 0;
    double ph[7] = {0};
    for (int c = 0; c < nct; ++c) {
        t0 = std::min(t0, h[c * 8]); t6 = std::max(t6, h[c * 8 + 6]);
        for (int i = 1; i <= 6; ++i) if (i != 2) ph[i] += (double)(h[c * 8 + i] - h[c * 8 + (i == 3 ? 1 : i - 1)]);
    }
    printf("%-28s grid=%3d passes=%d : back-to-back %.2f us/launch | in-kernel span %.2f us | mean phase us: prologue %.2f mainloop(all warps at sync) %.2f wait-mma %.2f epilogue %.2f teardown %.2f\n",
           name, nct, passes, ms * 1000 / 50, (t6 - t0) / 1000.0, ph[1] / nct / 1000, ph[3] / nct / 1000, ph[4] / nct / 1000, ph[5] / nct / 1000, ph[6] / nct / 1000);
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(tbuf);
}

int main() {
    run<32, false>("dH2  1024x128 K=64", 1024, 128, 64, 3);
    run<32, false>("dH2  1024x128 K=64", 1024, 128, 64, 1);
    run<32, false>("dH1  1024x256 K=128", 1024, 256, 128, 3);
    run<32, true>("fwd0 1024x256 K=384", 1024, 256, 384, 3);
    run<32, true>("fwd0 1024x256 K=384", 1024, 256, 384, 1);
    run<32, true>("fwd1 1024x128 K=256", 1024, 128, 256, 3);
    run<64, true>("fwd2 1024x64  K=128", 1024, 64, 128, 3);
    return 0;
}
