// Philox4x32-10 and the dropout-mask convention (identical to oracle/philox.py):
//   key = (seed + layer, step & 0xffffffff); counter = (e >> 2, 0, 0, 0), e = row * n_cols + col;
//   word = r[e & 3]; keep iff word < keep_threshold; mask = keep ? 1/keep_prob : 0.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct DropoutParams {
    uint32_t seed;       // dropout_seed + layer
    uint32_t step;       // low 32 bits of the global step
    uint32_t threshold;  // keep iff word < threshold
    float    scale;      // 1 / keep_prob
    int      enabled;
    uint32_t row0;       // batch row of local row 0 (a rank's slice of a data-parallel mini-batch; 0 otherwise)
};

// random words for the 4 consecutive elements starting at element index e (e % 4 == 0)
__device__ __forceinline__ uint4 dropout_words4(const DropoutParams& dp, uint32_t e) {
    return philox4x32_10(e >> 2, 0u, 0u, 0u, dp.seed, dp.step);
}

__device__ __forceinline__ float dropout_apply(const DropoutParams& dp, uint32_t word, float v) {
    return word < dp.threshold ? v * dp.scale : 0.0f;
}
