// Workspace layout of one mlp mini-batch of the fp32 SIMT path (mlp.cu).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace mlpws {

constexpr int kMaxSplit = 16;
constexpr int kMaxTiles = 4096;
// the sigmoid-BCE head (mlp.cu) splits the rows of a mini-batch over up to kHeadCtas CTAs; each parks a partial record
// (floats): dw[kHeadMaxN] | dg, pad | bce (double) | hist[2][1025] (int); a ticket follows the records
constexpr int kHeadMaxN = 512;  // widest last hidden layer supported by the head's smem partials
constexpr int kHeadCtas = 8;
constexpr int kHeadPartDg = kHeadMaxN, kHeadPartBce = kHeadMaxN + 2, kHeadPartHist = kHeadMaxN + 4;
constexpr int kHeadPartStride = kHeadMaxN + 4 + 2 * 1025 + 2;   // 2568 floats, a 16-byte multiple
inline size_t head_part_bytes() { return (size_t)kHeadCtas * kHeadPartStride * 4 + 64; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
    size_t H[MAMDR_MAX_LAYERS + 1];     // H[0] = X
    size_t dZ[MAMDR_MAX_LAYERS];
    size_t y, p, ds, uid_b, pid_b, partials, tickets, colsum, total;
    size_t hist;
    // trainable user / item tables: input gradient of layer 0 and the de-duplicated sparse gradients
    size_t dX, sp_ids[2], sp_rows[2], sp_n[2], sp_ws;
};

inline WsLayout ws_layout(const mamdr_mlp_desc& d, int B) {
    WsLayout w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    const int in_dim = d.emb_dim[0] + d.emb_dim[1] + d.emb_dim[2];
    // regions whose CONTENT persists from one mini-batch to the next (ticket counters and the AUC
    // bins are left zeroed by their consumers) come first: their offsets must not depend on B, which
    // changes on the ragged last batch of a pass
    w.tickets = take((size_t)kMaxTiles * 4);
    w.hist = take(head_part_bytes());   // per-CTA partial records of the head + its ticket (mlp.cu: kHeadCtas x kHeadPartStride floats)
    w.H[0] = take((size_t)B * in_dim * 4);
    for (int l = 0; l < d.n_layers; ++l) w.H[l + 1] = take((size_t)B * d.hidden[l] * 4);
    for (int l = 0; l < d.n_layers; ++l) w.dZ[l] = take((size_t)B * d.hidden[l] * 4);
    w.y = take((size_t)B * 4);
    w.p = take((size_t)B * 4);
    w.ds = take((size_t)B * 4);
    w.uid_b = take((size_t)B * 4);
    w.pid_b = take((size_t)B * 4);
    // split-K partials: sized for the largest dW (32 x 64 tiles)
    size_t max_mn = 0;
    int prev = in_dim;
    for (int l = 0; l < d.n_layers; ++l) {
        const size_t mn = (size_t)((prev + 127) / 128 * 128) * ((d.hidden[l] + 63) / 64 * 64);
        if (mn > max_mn) max_mn = mn;
        prev = d.hidden[l];
    }
    w.partials = take(max_mn * kMaxSplit * 4);
    size_t hsum = 0;
    for (int l = 0; l < d.n_layers; ++l) hsum += d.hidden[l];
    w.colsum = take(hsum * 4);
    w.dX = w.sp_ws = 0;
    w.sp_ids[0] = w.sp_ids[1] = w.sp_rows[0] = w.sp_rows[1] = w.sp_n[0] = w.sp_n[1] = 0;
    if (d.emb_trainable) {
        w.dX = take((size_t)B * (d.emb_dim[0] + d.emb_dim[1]) * 4);
        for (int t = 0; t < 2; ++t) {
            w.sp_ids[t] = take((size_t)B * 4);
            w.sp_rows[t] = take((size_t)B * d.emb_dim[t] * 4);
            w.sp_n[t] = take(16);
        }
        w.sp_ws = take(2 * mamdr_scatter_workspace_bytes(B));   // one per table
    }
    w.total = off;
    return w;
}

}  // namespace mlpws
