// Element-wise DN / DR meta operations over flat parameter arenas (K9 / K10), shared by the stand-alone sweep kernels
// (optim.cu) and the in-kernel program executor (mlp_pass.cu).  One float4 per call; `op` is a compile-time constant in
// the templated sweep kernel and a run-time value in the program kernel -- the arithmetic is the same code either way
// (loads are ld.global.cg: inside the persistent program kernel the operands were written by other SMs; explicit
// round-to-nearest mul / add in the order of the reference's numpy expressions: bit-exact vs the oracle).
//
// Replaces the host numpy algebra of /root/reference/model_zoo/domain_negotiation.py:118-123,
// model_zoo/mamdr.py:168-196, model_zoo/specific_base_model.py:164-172 and the SetVarOp / K.batch_get_value round
// trips (utils/tool.py:36-45, model_zoo/maml.py:181-194).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float merge1(float t, float ti, int method) {
    return method == MAMDR_MERGE_PLUS ? __fadd_rn(t, ti) : __fmul_rn(t, ti);
}

enum MetaOp { OP_COPY, OP_MERGE, OP_DN, OP_DR, OP_DR_ACC, OP_DR_APPLY, OP_SUB, OP_AXPY_DIFF };

struct MetaArgs {
    float*       w0;  // primary output / in-out
    float*       w1;  // secondary output (may be NULL)
    const float* r0;
    const float* r1;
    const float* r2;
    float        f0, f1;
    int          method;
    int64_t      n;
};

#define COMP(v, k) (reinterpret_cast<const float*>(&(v))[k])
#define COMPW(v, k) (reinterpret_cast<float*>(&(v))[k])

__device__ __forceinline__ void meta_float4(const int OP, const MetaArgs& a, const int64_t i) {
    float4 W0 = make_float4(0, 0, 0, 0), W1 = make_float4(0, 0, 0, 0);
    float4 R0 = make_float4(0, 0, 0, 0), R1 = R0, R2 = R0;
    if (OP == OP_DN || OP == OP_DR || OP == OP_DR_ACC || OP == OP_DR_APPLY || OP == OP_AXPY_DIFF) W0 = __ldcg(reinterpret_cast<const float4*>(a.w0 + 4 * i));
    if (OP == OP_DR_APPLY) W1 = __ldcg(reinterpret_cast<const float4*>(a.w1 + 4 * i));
    R0 = __ldcg(reinterpret_cast<const float4*>((OP == OP_DR_APPLY ? a.w1 : a.r0) + 4 * i));
    if (OP == OP_MERGE || OP == OP_DR || OP == OP_DR_ACC || OP == OP_SUB || OP == OP_AXPY_DIFF) R1 = __ldcg(reinterpret_cast<const float4*>(a.r1 + 4 * i));
    if (OP == OP_DR_ACC) R2 = __ldcg(reinterpret_cast<const float4*>(a.r2 + 4 * i));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (OP == OP_COPY) {
            COMPW(W0, k) = COMP(R0, k);
        } else if (OP == OP_MERGE) {  // out = theta (+|*) theta_i
            COMPW(W0, k) = merge1(COMP(R0, k), COMP(R1, k), a.method);
        } else if (OP == OP_DN) {  // theta += (model - theta) * beta ; model_out = theta
            const float t = COMP(W0, k);
            const float nt = __fadd_rn(t, __fmul_rn(__fsub_rn(COMP(R0, k), t), a.f0));
            COMPW(W0, k) = nt;
            COMPW(W1, k) = nt;
        } else if (OP == OP_DR) {  // W0 = theta_i, R0 = model, R1 = theta
            const float ti = COMP(W0, k), t = COMP(R1, k);
            const float merged = merge1(t, ti, a.method);
            const float nti = __fadd_rn(ti, __fmul_rn(__fsub_rn(COMP(R0, k), merged), a.f0));
            COMPW(W0, k) = nti;
            COMPW(W1, k) = merge1(t, nti, a.method);
        } else if (OP == OP_DR_ACC) {  // W0 = accum, R0 = model, R1 = theta, R2 = theta_i
            const float t = COMP(R1, k);
            const float merged = merge1(t, COMP(R2, k), a.method);
            float d = __fsub_rn(COMP(R0, k), merged);
            if (a.method == MAMDR_MERGE_TIMES) d = __fmul_rn(d, t);
            COMPW(W0, k) = __fadd_rn(COMP(W0, k), d);
        } else if (OP == OP_DR_APPLY) {  // W0 = theta_i, W1/R0 = accum ; f0 = sample_num, f1 = beta
            COMPW(W0, k) = __fadd_rn(COMP(W0, k), __fmul_rn(__fdiv_rn(COMP(R0, k), a.f0), a.f1));
            COMPW(W1, k) = 0.f;
        } else if (OP == OP_SUB) {
            COMPW(W0, k) = __fsub_rn(COMP(R0, k), COMP(R1, k));
        } else if (OP == OP_AXPY_DIFF) {  // out += (a - b) * alpha
            COMPW(W0, k) = __fadd_rn(COMP(W0, k), __fmul_rn(__fsub_rn(COMP(R0, k), COMP(R1, k)), a.f0));
        }
    }
    *reinterpret_cast<float4*>(a.w0 + 4 * i) = W0;
    if ((OP == OP_DN || OP == OP_DR) && a.w1) *reinterpret_cast<float4*>(a.w1 + 4 * i) = W1;
    if (OP == OP_DR_APPLY) *reinterpret_cast<float4*>(a.w1 + 4 * i) = W1;
}
