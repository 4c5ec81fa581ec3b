"""ctypes binding of include/cerberus_b200.h (the drop-in C ABI).

There is no CPU fallback: if the shared library is missing, `load()` raises and tells the
user to run `python -c "import __graft_entry__ as g; g.build()"`; if no sm_100 device is
present every compute entry point returns CERB_ERR_NO_DEVICE and `check()` raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcerberus_b200.so")

CERB_U8, CERB_F16, CERB_F32, CERB_I32 = 0, 1, 2, 3
CERB_PREC_F16, CERB_PREC_F16X2 = 0, 1
OP_PREP, OP_CONV, OP_MAXPOOL, OP_UPADD, OP_HEAD, OP_PCLASS = 1, 2, 3, 4, 5, 6
HEAD_INST, HEAD_TYPE = 0, 1
TISSUE_GLAND, TISSUE_LUMEN, TISSUE_NUCLEI = 0, 1, 2


class TensorDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("dtype", ctypes.c_int32)]


class Op(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32), ("in0", ctypes.c_int32), ("in1", ctypes.c_int32),
        ("out", ctypes.c_int32), ("in_coff", ctypes.c_int32), ("in_c", ctypes.c_int32),
        ("out_coff", ctypes.c_int32), ("cout", ctypes.c_int32),
        ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("stride", ctypes.c_int32),
        ("pad", ctypes.c_int32), ("relu", ctypes.c_int32), ("stem", ctypes.c_int32),
        ("head_mode", ctypes.c_int32), ("logits_out", ctypes.c_int32),
        ("w_off", ctypes.c_int64), ("w_lo_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
        ("box_w", ctypes.c_int32), ("w_shift", ctypes.c_int32), ("aux_classes", ctypes.c_int32),
        ("up_prev1", ctypes.c_int32), ("aux_w_off", ctypes.c_int64), ("aux_b_off", ctypes.c_int64),
        ("tail_w_off", ctypes.c_int64), ("tail_b_off", ctypes.c_int64),
        ("tail_w_shift", ctypes.c_int32), ("side", ctypes.c_int32),
    ]


MAX_DECODERS = 8
(L_STEM, L_BLOCK_CONV1, L_BLOCK_CONV2, L_BLOCK_DOWN, L_CONV_MAP, L_DEC_FIRST, L_DEC_CONV,
 L_HEAD_HIDDEN, L_HEAD_OUT, L_PCLASS) = range(10)


class Layer(ctypes.Structure):
    _fields_ = [("role", ctypes.c_int32), ("a", ctypes.c_int32), ("b", ctypes.c_int32),
                ("c", ctypes.c_int32), ("cout", ctypes.c_int32), ("cin", ctypes.c_int32),
                ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("w_shift", ctypes.c_int32),
                ("classes", ctypes.c_int32), ("w_off", ctypes.c_int64), ("w_lo_off", ctypes.c_int64),
                ("b_off", ctypes.c_int64)]


class ModelDesc(ctypes.Structure):
    _fields_ = [("n_decoders", ctypes.c_int32), ("head_mode", ctypes.c_int32 * MAX_DECODERS),
                ("classes", ctypes.c_int32 * MAX_DECODERS), ("canvas_coff", ctypes.c_int32 * MAX_DECODERS),
                ("has_pclass", ctypes.c_int32), ("pclass_classes", ctypes.c_int32),
                ("pclass_canvas_coff", ctypes.c_int32), ("canvas_c", ctypes.c_int32)]


# name -> (restype, argtypes); every symbol include/cerberus_b200.h declares.
_SIGNATURES = {
    "cerb_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "cerb_ctx_destroy": (None, [ctypes.c_void_p]),
    "cerb_last_error": (ctypes.c_char_p, []),
    "cerb_ctx_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "cerb_ctx_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "cerb_ctx_read_prof": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                                          ctypes.c_int]),
    "cerb_ctx_stat": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p]),
    "cerb_ctx_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "cerb_ctx_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "cerb_plan_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(TensorDesc), ctypes.c_int,
                                        ctypes.POINTER(Op), ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "cerb_plan_destroy": (None, [ctypes.c_void_p]),
    "cerb_plan_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "cerb_plan_num_ops": (ctypes.c_int, [ctypes.c_void_p]),
    "cerb_plan_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_int32)]),
    "cerb_plan_tensor_ptr": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "cerb_plan_read_tensor": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_size_t]),
    "cerb_plan_write_tensor": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_size_t]),
    "cerb_plan_preview_folding": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(TensorDesc), ctypes.c_int,
                                                 ctypes.POINTER(Op), ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]),
    "cerb_model_spec": (ctypes.c_int, [ctypes.POINTER(ModelDesc), ctypes.POINTER(Layer), ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.POINTER(TensorDesc), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(Op), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "cerb_model_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ModelDesc), ctypes.POINTER(Layer),
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_void_p)]),
    "cerb_model_destroy": (None, [ctypes.c_void_p]),
    "cerb_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int]),
    "cerb_model_plan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32)]),
    "cerb_nccl_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "cerb_nccl_comm_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.POINTER(ctypes.c_void_p)]),
    "cerb_nccl_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "cerb_bcast_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int]),
    "cerb_extract_patches": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "cerb_scatter_patches": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int]),
    "cerb_crop2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_int]),
    "cerb_region_half": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "cerb_nearest_channel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int]),
    "cerb_channel_plane": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "cerb_stitch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "cerb_postproc_nuclei": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "cerb_postproc_gland_lumen": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                 ctypes.c_void_p, ctypes.c_int]),
    "cerb_postproc_eroded_map": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "cerb_mask_lumen": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t]),
    "cerb_inst_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_int32)]),
    "cerb_inst_info_read": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 6),
    "cerb_pickle_instances": (ctypes.c_int64, [ctypes.c_char_p] + [ctypes.c_void_p] * 6 + [ctypes.c_int64] +
                              [ctypes.c_void_p] * 2 + [ctypes.c_int64]),
    "cerb_ellipse_rows": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                                         ctypes.POINTER(ctypes.c_int32)]),
    "cerb_copy_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_int]),
    "cerb_stream_order": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cerb_ctx_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "cerb_ctx_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cerb_ctx_wait_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "cerb_copy_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cerb_copy_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cerb_copy_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "cerb_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "cerb_host_free": (None, [ctypes.c_void_p]),
    "cerb_dev_alloc": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_size_t]),
    "cerb_dev_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "cerb_memcpy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_size_t, ctypes.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load():
    """Loads the shared library and sets prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "cerberus_b200: %s is missing. The CUDA extension is the product; there is no "
            "fallback path. Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise RuntimeError("cerberus_b200: %s lacks the declared symbol %s; rebuild it"
                               % (LIB_PATH, name))
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cerb_last_error().decode("utf-8", "replace")
        raise RuntimeError("cerberus_b200 %s failed (%d): %s" % (what, rc, msg))
