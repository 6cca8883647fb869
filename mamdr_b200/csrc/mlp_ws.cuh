// Workspace layout of one mlp mini-batch (shared by the SIMT path mlp.cu and the tcgen05 path mlp_tc.cu).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace mlpws {

constexpr int kMaxSplit = 16;
constexpr int kMaxTiles = 4096;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
    size_t H[MAMDR_MAX_LAYERS + 1];     // H[0] = X
    size_t Hlo[MAMDR_MAX_LAYERS + 1];   // tf32 "lo" parts (3xTF32), same shapes
    size_t dZ[MAMDR_MAX_LAYERS];
    size_t dZlo[MAMDR_MAX_LAYERS];
    size_t params_lo;                   // lo part of the parameter arena
    size_t y, p, ds, uid_b, pid_b, partials, tickets, colsum, total;
    // per-128-row-tile partials of the tcgen05 path
    size_t hp_db[MAMDR_MAX_LAYERS], hp_dw, hp_loss, hp_dg, hist;
    int    mtiles;
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
    w.H[0] = take((size_t)B * in_dim * 4);
    for (int l = 0; l < d.n_layers; ++l) w.H[l + 1] = take((size_t)B * d.hidden[l] * 4);
    w.Hlo[0] = take((size_t)B * in_dim * 4);
    for (int l = 0; l < d.n_layers; ++l) w.Hlo[l + 1] = take((size_t)B * d.hidden[l] * 4);
    for (int l = 0; l < d.n_layers; ++l) w.dZ[l] = take((size_t)B * d.hidden[l] * 4);
    for (int l = 0; l < d.n_layers; ++l) w.dZlo[l] = take((size_t)B * d.hidden[l] * 4);
    // dense span of the arena only, [off_domain_emb, arena_floats): tables are never GEMM operands
    w.params_lo = take((size_t)(d.arena_floats - d.off_domain_emb) * 4);
    w.y = take((size_t)B * 4);
    w.p = take((size_t)B * 4);
    w.ds = take((size_t)B * 4);
    w.uid_b = take((size_t)B * 4);
    w.pid_b = take((size_t)B * 4);
    // split-K partials: sized for the largest dW (tiles of 128 x 64 on the tcgen05 path, 32 x 64 on SIMT)
    size_t max_mn = 0;
    int prev = in_dim;
    for (int l = 0; l < d.n_layers; ++l) {
        const size_t mn = (size_t)((prev + 127) / 128 * 128) * ((d.hidden[l] + 63) / 64 * 64);
        if (mn > max_mn) max_mn = mn;
        prev = d.hidden[l];
    }
    w.partials = take(max_mn * kMaxSplit * 4);
    w.tickets = take((size_t)kMaxTiles * 4);
    size_t hsum = 0;
    for (int l = 0; l < d.n_layers; ++l) hsum += d.hidden[l];
    w.colsum = take(hsum * 4);
    w.mtiles = (B + 127) / 128;
    for (int l = 0; l < d.n_layers; ++l) w.hp_db[l] = take((size_t)w.mtiles * d.hidden[l] * 4);
    w.hp_dw = take((size_t)w.mtiles * d.hidden[d.n_layers - 1] * 4);
    w.hp_loss = take((size_t)w.mtiles * 8);
    w.hp_dg = take((size_t)w.mtiles * 4);
    w.hist = take((size_t)2 * 1025 * 4);
    w.total = off;
    return w;
}

}  // namespace mlpws
