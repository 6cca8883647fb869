/* Plain-C restatement of the integer parts of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py):
 *   - Philox4x32-10 dropout masks (definition: oracle/philox.py; Random123 KAT-pinned)
 *   - streaming AUC confusion counts, following /root/reference/utils/metrics_utils.py:297-354
 *     (strict fp32 `pred > threshold` compare against every threshold, four fp32 accumulators)
 * Built by oracle/build.py into oracle/_build/liboracle.so; oracle/*.py fall back to numpy (same
 * results, bit for bit -- tests/test_oracle_c.py) when it is absent. */
#include <stdint.h>

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

/* out[e] = word(e) < threshold ? scale : 0, e = row * cols + col, rows*cols % 4 == 0 */
void oracle_dropout_mask_f32(int64_t n_elems, uint32_t seed, uint32_t step, uint32_t threshold, float scale,
                             float* out) {
    for (int64_t q = 0; q < n_elems / 4; ++q) {
        uint32_t c[4] = {(uint32_t)q, 0u, 0u, 0u};
        philox4x32_10(c, seed, step);
        for (int j = 0; j < 4; ++j) out[4 * q + j] = c[j] < threshold ? scale : 0.0f;
    }
}

void oracle_philox_words(int64_t n_quads, uint32_t seed, uint32_t step, uint32_t* out) {
    for (int64_t q = 0; q < n_quads; ++q) {
        uint32_t c[4] = {(uint32_t)q, 0u, 0u, 0u};
        philox4x32_10(c, seed, step);
        for (int j = 0; j < 4; ++j) out[4 * q + j] = c[j];
    }
}

/* acc = [tp | fp | fn | tn], each [T] fp32; one fp32 add of the per-batch integer count per entry */
void oracle_auc_update(const float* p, const float* y, int64_t n, const float* thr, int32_t T, float* acc) {
    for (int32_t j = 0; j < T; ++j) {
        int64_t tp = 0, fp = 0, fn = 0, tn = 0;
        const float t = thr[j];
        for (int64_t i = 0; i < n; ++i) {
            const int pos = p[i] > t, lab = y[i] != 0.0f;
            tp += lab & pos; fp += (!lab) & pos; fn += lab & (!pos); tn += (!lab) & (!pos);
        }
        acc[j] += (float)tp; acc[T + j] += (float)fp; acc[2 * T + j] += (float)fn; acc[3 * T + j] += (float)tn;
    }
}
