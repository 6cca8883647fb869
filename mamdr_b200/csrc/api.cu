// Context management of the C-ABI (include/mamdr_b200.h).
#include <stdlib.h>

#include "common.cuh"

char g_mamdr_create_err[512] = "";

int mamdr_mlp_init_kernels(mamdr_ctx* ctx);
int mamdr_scatter_init_kernels(mamdr_ctx* ctx);
int mamdr_pass_init_kernels(mamdr_ctx* ctx);
void mamdr_pass_free_ctx(mamdr_ctx* ctx);

extern "C" int mamdr_abi_version(void) { return MAMDR_ABI_VERSION; }

extern "C" int mamdr_ctx_create(mamdr_ctx** out, int device) {
    if (!out) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "out is NULL");
        return MAMDR_E_INVALID;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "no CUDA device: %s", cudaGetErrorString(e));
        return MAMDR_E_CUDA;
    }
    if (device < 0 || device >= count) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "device %d out of range [0,%d)", device, count);
        return MAMDR_E_INVALID;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return MAMDR_E_CUDA;
    }
    if (prop.major != 10) {
        // the library is compiled for sm_100a only: no PTX fallback, no other architectures
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err),
                 "device %d is sm_%d%d; libmamdr_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return MAMDR_E_UNSUPPORTED;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "cudaSetDevice: %s", cudaGetErrorString(e));
        return MAMDR_E_CUDA;
    }
    mamdr_ctx* c = (mamdr_ctx*)calloc(1, sizeof(mamdr_ctx));
    if (!c) return MAMDR_E_INVALID;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    int rc = mamdr_mlp_init_kernels(c);
    if (rc == MAMDR_OK) rc = mamdr_scatter_init_kernels(c);
    if (rc == MAMDR_OK) rc = mamdr_pass_init_kernels(c);
    if (rc != MAMDR_OK) {
        snprintf(g_mamdr_create_err, sizeof(g_mamdr_create_err), "%s", c->err);
        free(c);
        return rc;
    }
    *out = c;
    return MAMDR_OK;
}

extern "C" void mamdr_ctx_destroy(mamdr_ctx* ctx) {
    if (!ctx) return;
    mamdr_pass_free_ctx(ctx);
    free(ctx);
}

extern "C" const char* mamdr_last_error(const mamdr_ctx* ctx) { return ctx ? ctx->err : g_mamdr_create_err; }

extern "C" int mamdr_sm_count(const mamdr_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int mamdr_ctx_set_pass_ctas(mamdr_ctx* ctx, int32_t n_ctas) {
    if (!ctx) return MAMDR_E_INVALID;
    MAMDR_REQUIRE(ctx, n_ctas == 0 || (n_ctas >= 32 && n_ctas <= ctx->sm_count), MAMDR_E_INVALID,
                  "pass kernel CTAs must be 0 (one per SM) or in [32, %d] (a CTA runs at most one domain job per mini-batch)", ctx->sm_count);
    ctx->pass_ctas = n_ctas;
    return MAMDR_OK;
}
