"""ctypes bindings of libfalcon_r1cs_b200.so (include/falcon_r1cs_b200.h).

The shared library is the product; this module only loads it and declares the
prototypes.  There is no Python or CPU fallback: if the library is missing the
import of `falcon_r1cs_b200.lib.load()` raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRCS_LIB", os.path.join(_HERE, "libfalcon_r1cs_b200.so"))

OK = 0
E_INVALID_ARG, E_CUDA, E_NO_PK, E_ALLOC, E_INVALID_POINT = -1, -2, -3, -4, -5
E_COEFF_RANGE, E_NORM_BOUND = -16, -17
KIND_NTT, KIND_SCHOOLBOOK, KIND_DUAL_NTT = 0, 1, 2

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class Shape(C.Structure):
    _fields_ = [("logn", C.c_uint32), ("kind", C.c_uint32), ("n_instance", C.c_uint32), ("n_witness", C.c_uint32),
                ("n_constraints", C.c_uint32), ("domain_log2", C.c_uint32), ("nnz_a", C.c_uint64),
                ("nnz_b", C.c_uint64), ("nnz_c", C.c_uint64)]


class PkView(C.Structure):
    _fields_ = [("alpha_g1", u64p), ("beta_g1", u64p), ("delta_g1", u64p), ("beta_g2", u64p), ("delta_g2", u64p),
                ("a_query", u64p), ("a_len", C.c_uint64), ("b_g1_query", u64p), ("b_g1_len", C.c_uint64),
                ("b_g2_query", u64p), ("b_g2_len", C.c_uint64), ("h_query", u64p), ("h_len", C.c_uint64),
                ("l_query", u64p), ("l_len", C.c_uint64)]


# every symbol include/falcon_r1cs_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "frcs_ctx_create": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]),
    "frcs_ctx_destroy": (None, [C.c_void_p]),
    "frcs_last_error": (C.c_char_p, []),
    "frcs_shape_get": (C.c_int32, [C.c_void_p, C.POINTER(Shape)]),
    "frcs_get_matrix": (C.c_int32, [C.c_void_p, C.c_int32, u32p, u32p, u64p]),
    "frcs_witness_batch": (C.c_int32, [C.c_void_p, C.c_uint64, u16p, u16p, u16p, u64p, i32p]),
    "frcs_witness_batch_dev": (C.c_int32, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6),
    "frcs_witness_check_batch": (C.c_int32, [C.c_void_p, C.c_uint64, u16p, u16p, u16p, i64p, i32p]),
    "frcs_witness_check_batch_dev": (C.c_int32, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6),
    "frcs_r1cs_eval_batch": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p, u64p, i64p]),
    "frcs_r1cs_eval_batch_dev": (C.c_int32, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6),
    "frcs_witness_map": (C.c_int32, [C.c_void_p, u64p, u64p]),
    "frcs_witness_map_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "frcs_domain_op": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_int32, u64p]),
    "frcs_msm_g1": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p]),
    "frcs_msm_g2": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p]),
    "frcs_load_pk": (C.c_int32, [C.c_void_p, C.POINTER(PkView)]),
    "frcs_verify_proof": (C.c_int32, [u64p, u64p, u64p, C.c_uint64, u64p, u64p]),
    "frcs_g1_validate": (C.c_int32, [u64p]),
    "frcs_g2_validate": (C.c_int32, [u64p]),
    "frcs_vk_validate": (C.c_int32, [u64p, u64p, u64p, C.c_uint64]),
    "frcs_pairing_eq": (C.c_int32, [u64p, u64p, u64p, u64p]),
    "frcs_pairing_is_one": (C.c_int32, [u64p, u64p]),
    "frcs_setup": (C.c_int32, [C.c_void_p, u64p, u64p, u64p, u64p]),
    "frcs_setup_shard": (C.c_int32, [C.c_void_p, u64p, C.c_uint32, C.c_uint32, u64p, u64p, u64p]),
    "frcs_export_pk": (C.c_int32, [C.c_void_p, C.c_int32, u64p]),
    "frcs_load_pk_shard": (C.c_int32, [C.c_void_p, C.POINTER(PkView), C.c_uint32, C.c_uint32]),
    "frcs_prove_partial_dev": (C.c_int32, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 8),
    "frcs_prove_split_begin_dev": (C.c_int32, [C.c_void_p] * 9),
    "frcs_prove_split_finish_dev": (C.c_int32, [C.c_void_p] * 4),
    "frcs_combine_partials": (C.c_int32, [C.c_uint32, C.c_uint64, u64p, u64p, u64p, u64p]),
    "frcs_prove_batch": (C.c_int32, [C.c_void_p, C.c_uint64, u16p, u16p, u16p, u64p, u64p, u64p, i32p]),
    "frcs_prove_from_z": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p, u64p]),
    "frcs_prove_batch_dev": (C.c_int32, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 8),
    "frcs_proof_compress": (C.c_int32, [u64p, u8p]),
    "frcs_launch_count": (C.c_uint64, [C.c_void_p]),
    "frcs_selftest": (C.c_int32, [C.c_int32, C.c_int32, u64p, C.c_uint64, u64p]),
    "frcs_profile_enable": (C.c_int32, [C.c_void_p, C.c_int32]),
    "frcs_profile_get": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64), C.c_int32]),
    "frcs_debug_windows_g1": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p]),
    "frcs_debug_windows_g2": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p]),
    "frcs_gadget_shape": (C.c_int32, [C.c_void_p, C.c_int32, u32p, u32p, u32p]),
    "frcs_gadget_mod_q": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p, i64p, i32p]),
    "frcs_gadget_add_mod": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, u64p, i64p, i32p]),
    "frcs_gadget_less_than_q": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, i64p, i32p]),
    "frcs_gadget_less_than_6144": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, C.c_int32, u64p, i64p, i32p]),
    "frcs_gadget_norm_bound": (C.c_int32, [C.c_void_p, C.c_uint64, u64p, u64p, i64p, i32p]),
    "frcs_gadget_ntt_circuit": (C.c_int32, [C.c_void_p, C.c_uint64, u16p, u16p, u64p, i64p, i32p]),
    "frcs_debug_host_matrix": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_int32, u32p, u32p, u64p, u64p]),
    "frcs_debug_msm_g1": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint64, u64p, u64p, u64p]),
    "frcs_debug_msm_g2": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint64, u64p, u64p, u64p]),
    "frcs_debug_digits": (C.c_int32, [u64p, C.c_uint64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "frcs_imad_peak": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libfalcon_r1cs_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C falcon_r1cs_b200/csrc).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class FrcsError(RuntimeError):
    def __init__(self, code, where):
        msg = load().frcs_last_error()
        super().__init__("%s failed with %d: %s" % (where, code, msg.decode() if msg else ""))
        self.code = code


def check(code, where):
    if code != OK:
        raise FrcsError(code, where)
