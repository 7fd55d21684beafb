"""ctypes binding of libodf.so (include/odf.h).  The CUDA extension is mandatory: if the shared
object is missing or a call fails, this module raises — there is no eager/PyTorch fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libodf.so")

c_fp = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/odf.h declaration by declaration
SIGNATURES = {
    "odf_last_error": (ctypes.c_char_p, []),
    "odf_version": (c_int, []),
    "odf_set_default_kind": (c_int, [c_int]),
    "odf_default_kind": (c_int, []),
    "odf_operand_pitch": (c_i64, [c_i64, c_int]),
    "odf_operand_bytes": (c_i64, [c_i64, c_i64, c_int]),
    "odf_pad_rows": (c_i64, [c_i64]),
    "odf_tpad": (c_int, [c_i64]),
    "odf_tile_splits": (c_int, [c_i64, c_i64, c_i64, c_int]),
    "odf_prepare_points": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "odf_prepare_points_linear": (c_int, [c_fp, c_i64, c_i64, c_i64, c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "odf_gemm_nt_split": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64, c_i64, c_f, c_f, c_fp,
                                  c_i64, c_fp]),
    "odf_zscore": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_f, c_fp]),
    "odf_split_rhs": (c_int, [c_fp, c_i64, c_i64, c_i64, c_f, c_fp, c_fp, c_i64, c_int, c_fp]),
    "odf_gauss_mmv_prepared": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64, c_i64,
                                       c_fp, c_fp, c_i64, c_int, c_int, c_f, c_fp, c_fp]),
    "odf_gauss_mmv_prepared_spill": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64, c_i64,
                                             c_fp, c_fp, c_i64, c_int, c_int, c_f, c_fp, c_fp, c_i64, c_fp]),
    "odf_panel_splits": (c_int, [c_i64, c_i64]),
    "odf_panel_tmm": (c_int, [c_fp, c_i64, c_fp, c_i64, c_i64, c_int, c_int, c_fp, c_fp]),
    "odf_panel16_bytes": (c_sz, [c_i64, c_i64]),
    "odf_tile_pair_eligible": (c_int, [c_i64]),
    "odf_split_rhs16": (c_int, [c_fp, c_i64, c_i64, c_i64, c_f, c_fp, c_fp, c_fp, c_i64, c_int, c_fp]),
    "odf_gauss_mmv_pair": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64, c_i64,
                                   c_fp, c_fp, c_i64, c_fp, c_int, c_int, c_f, c_fp, c_fp, c_fp]),
    "odf_gauss_mmv_prepared_spill16": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64,
                                               c_i64, c_fp, c_fp, c_i64, c_int, c_int, c_f, c_fp, c_fp, c_fp]),
    "odf_finish_w16": (c_int, [c_fp, c_int, c_i64, c_int, c_i64, c_fp, c_i64, c_fp, c_fp, c_fp, c_fp]),
    "odf_panel16_splits": (c_int, [c_i64, c_i64]),
    "odf_panel16_tmm": (c_int, [c_fp, c_i64, c_i64, c_fp, c_fp, c_int, c_int, c_fp, c_fp]),
    "odf_panel16_mmv_splits": (c_int, [c_i64, c_i64]),
    "odf_panel16_mmv": (c_int, [c_fp, c_i64, c_i64, c_fp, c_fp, c_int, c_int, c_fp, c_fp]),
    "odf_panel16_tmm_hi": (c_int, [c_fp, c_i64, c_i64, c_fp, c_fp, c_int, c_int, c_fp, c_fp]),
    "odf_panel16_mmv_hi": (c_int, [c_fp, c_i64, c_i64, c_fp, c_fp, c_int, c_int, c_fp, c_fp]),
    "odf_finish_rows": (c_int, [c_fp, c_int, c_i64, c_int, c_i64, c_f, c_fp, c_i64, c_fp, c_i64, c_fp]),
    "odf_finish_split": (c_int, [c_fp, c_int, c_i64, c_int, c_i64, c_f, c_fp, c_i64, c_fp, c_fp,
                                 c_i64, c_fp]),
    "odf_gauss_kmm_prepared": (c_int, [c_int, c_fp, c_fp, c_fp, c_fp, c_i64, c_i64, c_f, c_fp, c_i64, c_fp]),
    "odf_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_i64, c_i64]),
    "odf_precond_workspace_bytes": (c_sz, [c_i64]),
    "odf_gauss_mmv": (c_int, [c_fp, c_i64, c_i64, c_fp, c_i64, c_i64, c_i64, c_fp, c_i64, c_i64, c_f,
                              c_fp, c_i64, c_fp, c_sz, c_fp]),
    "odf_gauss_dmmv": (c_int, [c_fp, c_i64, c_i64, c_fp, c_i64, c_i64, c_i64, c_fp, c_i64, c_fp, c_i64,
                               c_i64, c_f, c_fp, c_i64, c_fp, c_sz, c_fp]),
    "odf_gauss_kmm": (c_int, [c_fp, c_i64, c_i64, c_i64, c_f, c_fp, c_i64, c_fp, c_sz, c_fp]),
    "odf_precond_init": (c_int, [c_fp, c_fp, c_i64, c_f, c_f, c_fp, c_sz, c_fp]),
    "odf_precond_build_workspace_bytes": (c_sz, [c_i64]),
    "odf_precond_build": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_f, c_f, c_fp, c_sz, c_fp]),
    "odf_precond_solve": (c_int, [c_fp, c_i64, c_fp, c_i64, c_i64, c_int, c_fp]),
    "odf_precond_invert": (c_int, [c_fp, c_fp, c_i64, c_fp]),
    "odf_precond_apply": (c_int, [c_fp, c_i64, c_fp, c_fp, c_i64, c_i64, c_int, c_fp]),
    "odf_gemm": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_f, c_fp, c_i64, c_fp, c_i64, c_f, c_fp, c_i64, c_fp]),
    "odf_precond_apply_rows": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_fp, c_i64, c_i64, c_i64, c_int, c_fp]),
    "odf_rls_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "odf_rls_train": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_fp, c_fp, c_fp, c_i64, ctypes.c_double, c_fp, c_fp, c_fp, c_sz, c_fp]),
    "odf_rls_apply": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_f, c_f, c_f, c_fp, c_f, c_fp, c_fp]),
    "odf_cg_init": (c_int, [c_fp, c_i64, c_i64, c_i64, c_fp, c_fp, c_sz, c_fp]),
    "odf_cg_alpha": (c_int, [c_fp, c_fp, c_i64, c_i64, c_i64, c_f, c_fp, c_fp, c_sz, c_fp]),
    "odf_cg_axpy_a": (c_int, [c_fp, c_fp, c_i64, c_i64, c_i64, c_f, c_fp, c_fp]),
    "odf_cg_residual": (c_int, [c_fp, c_fp, c_fp, c_i64, c_i64, c_i64, c_fp, c_fp]),
    "odf_cg_beta": (c_int, [c_fp, c_i64, c_i64, c_i64, c_f, c_f, c_fp, c_fp, c_sz, c_fp]),
    "odf_cg_xpby_b": (c_int, [c_fp, c_fp, c_i64, c_i64, c_i64, c_fp, c_fp]),
    "odf_axpby": (c_int, [c_fp, c_f, c_fp, c_f, c_fp, c_i64, c_i64, c_i64, c_fp]),
    "odf_cg_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "odf_potrf_upper": (c_int, [c_fp, c_i64, c_fp, c_sz, c_fp]),
    "odf_add_diag": (c_int, [c_fp, c_i64, c_f, c_fp]),
    "odf_zero_strict_lower": (c_int, [c_fp, c_i64, c_fp]),
    "odf_select_workspace_bytes": (c_sz, [c_i64]),
    "odf_select_indices": (c_int, [c_fp, c_i64, c_i64, c_f, c_int, c_fp, c_fp, c_fp, c_sz, c_fp]),
    "odf_gather_rows": (c_int, [c_fp, c_i64, c_fp, c_fp, c_i64, c_i64, c_fp, c_i64, c_fp]),
    "odf_decode_boxes": (c_int, [c_fp, c_fp, c_i64, c_i64, c_f, c_f, c_fp, c_fp]),
    "odf_postprocess_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "odf_detect_postprocess": (c_int, [c_fp, c_fp, c_i64, c_i64, c_f, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                       c_sz, c_fp]),
}

ODF_OP_MMV, ODF_OP_DMMV, ODF_OP_KMM, ODF_OP_PRECOND = 0, 1, 2, 3
ODF_KIND_TF32, ODF_KIND_F16 = 0, 1
KIND_NAMES = {"tf32": ODF_KIND_TF32, "f16": ODF_KIND_F16}
ODF_SOLVE_T, ODF_SOLVE_TT, ODF_SOLVE_A, ODF_SOLVE_AT = 0, 1, 2, 3


class OdfError(RuntimeError):
    pass


_lib = None


def load():
    """Load libodf.so (once).  Raises OdfError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OdfError(
            "libodf.so not found at %s — build it with `python online-detection_b200/build_lib.py` "
            "(there is no CPU/PyTorch fallback for the FALKON hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().odf_last_error()
        raise OdfError("%s failed (%d): %s" % (what or "libodf call", rc, (msg or b"").decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
