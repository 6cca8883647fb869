"""ctypes binding of ``libmamdr_b200.so`` (C-ABI in ``include/mamdr_b200.h``).

The product has NO CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with ``python -m mamdr_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAMDR_B200_LIB") or os.path.join(_HERE, "lib", "libmamdr_b200.so")   # env override: A/B kernel builds

MAX_LAYERS = 8
PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
MERGE_PLUS, MERGE_TIMES = 0, 1
ABI_VERSION = 1

E_NAMES = {0: "MAMDR_OK", -1: "MAMDR_E_INVALID", -2: "MAMDR_E_CUDA", -3: "MAMDR_E_WORKSPACE",
           -4: "MAMDR_E_UNSUPPORTED"}


class MamdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (E_NAMES.get(code, "?"), code, msg))
        self.code = code


class MlpDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("emb_dim", C.c_int32 * 3),
        ("hidden", C.c_int32 * MAX_LAYERS),
        ("n_domain", C.c_int32),
        ("emb_trainable", C.c_int32),
        ("n_uid", C.c_int64),
        ("n_pid", C.c_int64),
        ("dropout_rate", C.c_float),
        ("dropout_seed", C.c_uint32),
        ("l2_emb", C.c_float),
        ("frozen_reg", C.c_float),
        ("off_user_emb", C.c_int64),
        ("off_item_emb", C.c_int64),
        ("off_domain_emb", C.c_int64),
        ("off_kernel", C.c_int64 * MAX_LAYERS),
        ("off_bias", C.c_int64 * MAX_LAYERS),
        ("off_dense_kernel", C.c_int64),
        ("off_global_bias", C.c_int64),
        ("arena_floats", C.c_int64),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("uid_dev", C.c_void_p),
        ("pid_dev", C.c_void_p),
        ("label_dev", C.c_void_p),
        ("order_dev", C.c_void_p),
        ("offset", C.c_int64),
        ("rows", C.c_int32),
        ("domain", C.c_int32),
        ("row0", C.c_int32),
        ("reserved", C.c_int32),
    ]


class StarDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("emb_dim", C.c_int32 * 3), ("hidden", C.c_int32 * MAX_LAYERS),
                ("n_domain", C.c_int32), ("n_uid", C.c_int64), ("n_pid", C.c_int64), ("pn_eps", C.c_float),
                ("pn_momentum", C.c_float), ("off_domain_emb", C.c_int64), ("off_gamma_sp", C.c_int64),
                ("off_beta_sp", C.c_int64), ("off_gamma_sh", C.c_int64), ("off_beta_sh", C.c_int64),
                ("off_ksp", C.c_int64 * MAX_LAYERS), ("off_bsp", C.c_int64 * MAX_LAYERS),
                ("off_ksh", C.c_int64 * MAX_LAYERS), ("off_bsh", C.c_int64 * MAX_LAYERS),
                ("off_out_kernel", C.c_int64), ("off_out_bias", C.c_int64), ("arena_floats", C.c_int64)]


MTL_MAX_K = 8


class MtlDesc(C.Structure):
    _fields_ = [("emb_dim", C.c_int32 * 3), ("n_domain", C.c_int32), ("n_uid", C.c_int64), ("n_pid", C.c_int64),
                ("emb_trainable", C.c_int32), ("has_gate", C.c_int32), ("k", C.c_int32),
                ("n_expert_layers", C.c_int32), ("n_gate_layers", C.c_int32), ("n_tower_layers", C.c_int32),
                ("expert_hidden", C.c_int32 * MAX_LAYERS), ("gate_hidden", C.c_int32 * MAX_LAYERS),
                ("tower_hidden", C.c_int32 * MAX_LAYERS), ("dropout_rate", C.c_float), ("dropout_seed", C.c_uint32),
                ("l2_emb", C.c_float), ("frozen_reg", C.c_float), ("off_user_emb", C.c_int64),
                ("off_item_emb", C.c_int64), ("off_domain_emb", C.c_int64), ("arena_floats", C.c_int64)]


class MtlDomain(C.Structure):
    _fields_ = [("domain", C.c_int32), ("expert_id", C.c_int32 * MTL_MAX_K),
                ("off_expert_kernel", (C.c_int64 * MAX_LAYERS) * MTL_MAX_K),
                ("off_expert_bias", (C.c_int64 * MAX_LAYERS) * MTL_MAX_K),
                ("off_gate_kernel", C.c_int64 * MAX_LAYERS), ("off_gate_bias", C.c_int64 * MAX_LAYERS),
                ("off_gate_out", C.c_int64), ("off_tower_kernel", C.c_int64 * MAX_LAYERS),
                ("off_tower_bias", C.c_int64 * MAX_LAYERS), ("off_tower_out", C.c_int64), ("off_bias", C.c_int64)]


class Pass(C.Structure):
    _fields_ = [
        ("uid_dev", C.c_void_p),
        ("pid_dev", C.c_void_p),
        ("label_dev", C.c_void_p),
        ("order_dev", C.c_void_p),
        ("n_data", C.c_int64),
        ("batch_size", C.c_int32),
        ("steps", C.c_int32),
        ("domain", C.c_int32),
    ]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/mamdr_b200.h declares
SIGNATURES = {
    "mamdr_abi_version": (C.c_int, []),
    "mamdr_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "mamdr_ctx_destroy": (None, [_P]),
    "mamdr_last_error": (C.c_char_p, [_P]),
    "mamdr_sm_count": (C.c_int, [_P]),
    "mamdr_ctx_set_pass_ctas": (C.c_int, [_P, _I32]),
    "mamdr_gather_f32": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _P, _I64, _P]),
    "mamdr_scatter_max_n": (_I64, []),
    "mamdr_scatter_workspace_bytes": (_SZ, [_I64]),
    "mamdr_scatter_dedup_f32": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _SZ, _P]),
    "mamdr_scatter_large_workspace_bytes": (_SZ, [_I64, _I32]),
    "mamdr_scatter_dedup_large_f32": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _SZ, _P]),
    "mamdr_opt_state_bytes": (_SZ, []),
    "mamdr_opt_state_init": (C.c_int, [_P, _P, _F, _F, _P]),
    "mamdr_opt_state_read": (C.c_int, [_P, _P, C.POINTER(_I64), C.POINTER(_F), C.POINTER(_F), _P]),
    "mamdr_mlp_workspace_bytes": (_SZ, [C.POINTER(MlpDesc), _I32]),
    "mamdr_mlp_train_step": (C.c_int, [_P, C.POINTER(MlpDesc), C.POINTER(Batch), _P, _P, _P, _P, _P, _SZ, _P,
                                       _P, _P, _P, _P, _I32, _I32, _P]),
    "mamdr_mlp_eval_step": (C.c_int, [_P, C.POINTER(MlpDesc), C.POINTER(Batch), _P, _P, _P, _P, _SZ, _P, _P, _P,
                                      _P, _I32, _I32, _P]),
    "mamdr_mlp_pass_workspace_bytes": (_SZ, [C.POINTER(MlpDesc), _I32]),
    "mamdr_mlp_pass_supported": (C.c_int, [_P, C.POINTER(MlpDesc), _I32]),
    "mamdr_mlp_train_pass": (C.c_int, [_P, C.POINTER(MlpDesc), C.POINTER(Pass), _P, _P, _P, _P, _P, _P, _P, _SZ, _P,
                                       _P, _P, _P, _I32, _I32, _F, _F, _F, _F, _I32, _P]),
    "mamdr_mlp_eval_pass": (C.c_int, [_P, C.POINTER(MlpDesc), C.POINTER(Pass), _P, _P, _P, _P, _SZ, _P, _P, _P, _P,
                                      _P, _I32, _I32, _P]),
    "mamdr_star_workspace_bytes": (_SZ, [C.POINTER(StarDesc), _I32]),
    "mamdr_star_state_bytes": (_SZ, [C.POINTER(StarDesc)]),
    "mamdr_star_debug_offsets": (C.c_int, [C.POINTER(StarDesc), _I32, C.POINTER(_I64)]),
    "mamdr_star_train_step": (C.c_int, [_P, C.POINTER(StarDesc), C.POINTER(Batch), _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _P,
                                        _I32, _P]),
    "mamdr_star_eval_step": (C.c_int, [_P, C.POINTER(StarDesc), C.POINTER(Batch), _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _P, _I32,
                                       _P]),
    "mamdr_mtl_workspace_bytes": (_SZ, [C.POINTER(MtlDesc), _I32]),
    "mamdr_mtl_train_step": (C.c_int, [_P, C.POINTER(MtlDesc), C.POINTER(MtlDomain), C.POINTER(Batch), _P, _P, _P, _P, _P, _SZ,
                                       _P, _P, _P, _P, _P, _I32, _P]),
    "mamdr_mtl_eval_step": (C.c_int, [_P, C.POINTER(MtlDesc), C.POINTER(MtlDomain), C.POINTER(Batch), _P, _P, _P, _P, _SZ, _P,
                                      _P, _P, _P, _I32, _P]),
    "mamdr_mtl_sparse_grads": (C.c_int, [C.POINTER(MtlDesc), _I32, _P, _I32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mamdr_mtl_input_grads": (C.c_int, [_P, C.POINTER(MtlDesc), C.POINTER(MtlDomain), _I32, _P, _P, _SZ, _P, _P]),
    "mamdr_adam_ranges_step": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), _I32, _P, _F, _F, _F, _F, _P]),
    "mamdr_program_begin": (C.c_int, [_P]),
    "mamdr_program_end": (C.c_int, [_P, _P, _SZ, C.POINTER(_I32), _P]),
    "mamdr_program_abort": (None, [_P]),
    "mamdr_program_op_bytes": (_I64, []),
    "mamdr_debug_pass_timing": (C.c_int, [_P, _P, _I64]),
    "mamdr_mlp_input_grads": (C.c_int, [_P, C.POINTER(MlpDesc), _I32, _P, _P, _SZ, _P, _P]),
    "mamdr_mlp_sparse_grads": (C.c_int, [C.POINTER(MlpDesc), _I32, _P, _I32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mamdr_mlp_pass_sparse_grads": (C.c_int, [C.POINTER(MlpDesc), _I32, _P, _I32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mamdr_adam_table_workspace_bytes": (_SZ, []),
    "mamdr_adam_table_step": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P, _P, _I64, _P, _F, _P, _F, _F, _F, _F, _P, _P,
                                        _SZ, _P]),
    "mamdr_sgd_table_step": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _I64, _P, _F, _F, _P, _P, _SZ, _P]),
    "mamdr_sum_squares_f64": (C.c_int, [_P, _P, _I64, _P, _P, _SZ, _P]),
    "mamdr_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _P]),
    "mamdr_sgd_step": (C.c_int, [_P, _P, _P, _I64, _P, _F, _P]),
    "mamdr_copy": (C.c_int, [_P, _P, _P, _I64, _P]),
    "mamdr_merge": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "mamdr_dn_update": (C.c_int, [_P, _P, _P, _F, _I64, _P, _P]),
    "mamdr_dr_update": (C.c_int, [_P, _P, _P, _P, _F, _I64, _I32, _P, _P]),
    "mamdr_dr_accumulate": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _P]),
    "mamdr_dr_apply_accum": (C.c_int, [_P, _P, _P, _F, _F, _I64, _P]),
    "mamdr_sub": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "mamdr_axpy_diff": (C.c_int, [_P, _P, _P, _P, _F, _I64, _P]),
    "mamdr_pcgrad_project": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "mamdr_route_plan": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "mamdr_route_gather2": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "mamdr_route_pack_rows": (C.c_int, [_P, _P, _I64, _P, _I32, _I32, _F, _P, _P]),
    "mamdr_auc_update": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I32, _P]),
    "mamdr_auc_result": (C.c_int, [_P, _P, _I32, _P, _P]),
}

_lib = None


def load():
    """dlopen the library and bind every declared symbol.  Raises if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmamdr_b200.so not built (%s). Run `python -m mamdr_b200.build`; there is no "
                          "CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mamdr_abi_version() != ABI_VERSION:
        raise ImportError("libmamdr_b200.so ABI %d != expected %d" % (lib.mamdr_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


class Context(object):
    """One ``mamdr_ctx`` bound to a CUDA device.  ``call`` raises MamdrError on a non-zero code."""

    def __init__(self, device=0):
        self.lib = load()
        h = _P()
        rc = self.lib.mamdr_ctx_create(C.byref(h), int(device))
        if rc != 0:
            raise MamdrError(rc, (self.lib.mamdr_last_error(None) or b"").decode())
        self.handle = h
        self.device = int(device)
        self.sm_count = self.lib.mamdr_sm_count(h)
        self.launches = 0  # kernels / memsets enqueued through this ctx (bench 'gpu_launches')

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            raise MamdrError(rc, (self.lib.mamdr_last_error(self.handle) or b"").decode())

    def close(self):
        if self.handle:
            self.lib.mamdr_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
