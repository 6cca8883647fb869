// Device-side "program": a recorded sequence of whole domain passes and element-wise meta sweeps that the persistent
// kernel of mlp_pass.cu executes in ONE cooperative launch (mamdr_program_begin / mamdr_program_end).
#pragma once
#include "common.cuh"
#include "meta_ops.cuh"

namespace passk {

// one entry of a device-side program: a whole domain pass, or an element-wise meta sweep over arenas
enum { PROG_PASS = 0, PROG_META = 1 };
struct ProgOp {
    int kind, domain, steps, method;      // PROG_PASS: domain, steps; PROG_META: method = MetaOp | merged_method << 8
    long long n_data, n;
    const int32_t *uid, *pid, *order;
    const float* label;
    float *losses, *probs;
    float *w0, *w1;
    const float *r0, *r1, *r2;
    float f0, f1;
};

}  // namespace passk

// recording hooks (implemented in mlp_pass.cu); return true when the op was recorded instead of launched
bool mamdr_prog_recording(const mamdr_ctx* ctx);
int  mamdr_prog_push_meta(mamdr_ctx* ctx, int meta_op, const MetaArgs& a);
