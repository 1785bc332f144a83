"""ctypes binding of the C-ABI in include/vitta_b200.h.

There is deliberately no fallback: if ``libvitta_b200.so`` is missing or a call fails, the caller gets
an exception -- never a silent eager-PyTorch / CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitta_b200.so")

REG_TYPES = {"l1_loss": 0, "mse_loss": 1, "kld": 2}


class VittaChunking(C.Structure):
    _fields_ = [("chunk_rows", C.c_int32), ("chunks_per_frame", C.c_int32), ("frame_rows", C.c_int64),
                ("n_entries", C.c_int32), ("reserved", C.c_int32)]


class VittaLayerDesc(C.Structure):
    _fields_ = [("C", C.c_int32), ("n_entries", C.c_int32), ("chunk_rows", C.c_int32), ("chunks_per_frame", C.c_int32),
                ("frame_rows", C.c_int64), ("part_off", C.c_int64), ("entry_stride", C.c_int64), ("cnt_off", C.c_int64),
                ("cnt_stride", C.c_int64), ("ch_off", C.c_int64), ("reg_type", C.c_int32), ("has_source", C.c_int32),
                ("w_new", C.c_float), ("w_old", C.c_float)]


class VittaBN(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("eps", C.c_float)]


class VittaLnGather(C.Structure):
    _fields_ = [("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32)]


class VittaSgdTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("buf", C.c_void_p), ("n", C.c_int64)]


class VittaSplitTensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p), ("amax", C.c_void_p), ("R", C.c_int32),
                ("T", C.c_int32), ("Cc", C.c_int32), ("mode", C.c_int32), ("src_tap_inner", C.c_int32),
                ("compute_amax", C.c_int32), ("n", C.c_int64), ("fold_w", C.c_void_p), ("fold_rv", C.c_void_p),
                ("fold_eps", C.c_float), ("reserved", C.c_int32)]


class VittaFoldBias(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("rm", C.c_void_p), ("rv", C.c_void_p), ("out", C.c_void_p),
                ("eps", C.c_float), ("C", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    "vitta_version": (C.c_int, []),
    "vitta_last_error": (C.c_char_p, []),
    "vitta_sm_count": (C.c_int, []),
    "vitta_stats_chunking": (C.c_int, [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.POINTER(VittaChunking)]),
    "vitta_stats_partial": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int64, C.c_int64, _P, _P]),
    "vitta_stats_finalize": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_int,
                                       _P]),
    "vitta_stats_finalize_loss_floats": (C.c_int64, [C.c_int, C.c_int]),
    "vitta_stats_inject": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int64, _P]),
    "vitta_bn_act_fwd": (C.c_int, [_P, VittaBN, _P, C.POINTER(VittaBN), C.c_int, _P, _P, _P, _P, _P, C.c_int64,
                                   C.c_int64, C.c_int, _P]),
    "vitta_bn_act_bwd_ws_floats": (C.c_int64, [C.c_int64, C.c_int64, C.c_int]),
    "vitta_bn_act_bwd": (C.c_int, [_P, _P, _P, VittaBN, _P, C.POINTER(VittaBN), C.c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P]),
    "vitta_tam_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P]),
    "vitta_bn_act_fwd_amax": (C.c_int, [_P, VittaBN, _P, C.POINTER(VittaBN), C.c_int, _P, _P, _P, _P, _P, C.c_int64,
                                        C.c_int64, C.c_int, _P, _P]),
    "vitta_bn_act_bwd_amax": (C.c_int, [_P, _P, _P, VittaBN, _P, C.POINTER(VittaBN), C.c_int, _P, _P, _P, _P, _P, _P, _P,
                                        _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P, _P]),
    "vitta_tam_fwd_amax": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P, _P]),
    "vitta_tam_num_chunks": (C.c_int, [C.c_int64, C.c_int]),
    "vitta_tam_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P]),
    "vitta_tam_bwd_finish_tickets": (C.c_int, []),
    "vitta_tam_bwd_finish": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_pred_consis": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "vitta_gemm_set_operand_form": (C.c_int, [C.c_int]),
    "vitta_gemm_set_cta_pair": (C.c_int, [C.c_int]),
    "vitta_split_tf32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_gemm_tf32x3": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, _P, _P,
                                    C.c_int64, C.c_int, C.c_int, _P]),
    "vitta_conv2d_tf32x3": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, _P, _P, C.c_int, _P]),
    "vitta_conv2d_tf32x3_ex": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, _P, _P, _P, C.c_int, _P]),
    "vitta_conv2d_dgrad_tf32x3": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, _P, _P]),
    "vitta_conv2d_wgrad_ws_floats": (C.c_int64, [C.c_int] * 9),
    "vitta_conv2d_wgrad_plan": (C.c_int, [C.c_int] * 10 + [C.POINTER(C.c_int)]),
    "vitta_conv2d_wgrad_tf32x3": (C.c_int, [_P, _P] + [C.c_int] * 9 + [_P, C.c_int, _P, _P]),
    "vitta_gemm_tf32x3_ex": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, _P, _P,
                                       C.c_int64, C.c_int, _P, _P, C.c_int64, C.c_int, _P]),
    "vitta_amax_f32": (C.c_int, [_P, C.c_int64, _P, _P]),
    "vitta_split_f16": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_gemm_f16x3_ex": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      _P, _P, C.c_int64, C.c_int, _P, _P, C.c_int64, C.c_int, _P]),
    "vitta_conv2d_f16x3_ex": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
    "vitta_conv2d_wgrad_f16x3": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 9 + [_P, C.c_int, _P, _P]),
    "vitta_conv2d_wgrad_f16x3_bias": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 9 + [_P, _P, C.c_int, _P, _P]),
    "vitta_conv2d_dgrad_f16x3": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "vitta_ln_chunking": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(VittaChunking)]),
    "vitta_ln_fwd": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, _P, _P, C.c_int64, C.c_int, C.POINTER(VittaLnGather), _P]),
    "vitta_ln_fwd_amax": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, _P, _P, C.c_int64, C.c_int, C.POINTER(VittaLnGather), _P,
                                    _P]),
    "vitta_ln_bwd_amax": (C.c_int, [_P] * 15 + [C.c_int64, C.c_int, C.POINTER(VittaLnGather), _P, _P]),
    "vitta_ln_bwd_ws_floats": (C.c_int64, [C.c_int64, C.c_int]),
    "vitta_ln_bwd": (C.c_int, [_P] * 15 + [C.c_int64, C.c_int, C.POINTER(VittaLnGather), _P]),
    "vitta_colsum_ws_floats": (C.c_int64, [C.c_int64, C.c_int]),
    "vitta_colsum": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, _P, _P]),
    "vitta_frame_mean": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "vitta_frame_mean_bwd": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "vitta_patchify3d": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "vitta_row_scale": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P]),
    "vitta_row_scale_amax": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P, _P]),
    "vitta_wmsa3d_fwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, _P]),
    "vitta_wmsa3d_fwd_amax": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, _P, _P]),
    "vitta_wmsa3d_fwd_trace": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, _P, C.c_int, _P]),
    "vitta_wmsa3d_bwd_amax": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, C.c_int, _P, _P]),
    "vitta_wmsa3d_bwd_ws_floats": (C.c_int64, [C.c_int] * 5),
    "vitta_wmsa3d_bwd_set_trace": (C.c_int, [_P, C.c_int]),
    "vitta_wmsa3d_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, C.c_int, _P]),
    "vitta_gather_normalize_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, _P, _P]),
    "vitta_resample_ksize": (C.c_int, [C.c_int, C.c_int]),
    "vitta_resample_coeffs_u8": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vitta_gather_crop_resize_normalize_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int32),
                                                        C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, _P,
                                                        _P]),
    "vitta_cv_linear_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vitta_cv_resize_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                     _P, _P, C.c_int, C.c_int, _P, _P]),
    "vitta_cv_resize_normalize_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               _P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                               C.c_int, C.c_int, _P, _P]),
    "vitta_tam_gate_fwd": (C.c_int, [_P, _P, VittaBN, _P, _P, VittaBN, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_tam_gate_bwd_ws_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "vitta_tam_gate_bwd": (C.c_int, [_P, _P, VittaBN, _P, _P, VittaBN, _P] + [_P] * 16 + [C.c_int, C.c_int, C.c_int, _P]),
    "vitta_stem_pack": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_stem_pack_weight": (C.c_int, [_P, _P, _P, _P]),
    "vitta_stem_conv_tf32x3": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "vitta_stem_wgrad_ws_floats": (C.c_int64, []),
    "vitta_stem_wgrad": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_bn_relu_pool_fwd": (C.c_int, [_P, VittaBN, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_bn_relu_pool_fwd_amax": (C.c_int, [_P, VittaBN, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "vitta_bn_relu_pool_bwd_ws_floats": (C.c_int64, [C.c_int]),
    "vitta_bn_relu_pool_bwd": (C.c_int, [_P, _P, _P, VittaBN, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_bn_fold_bias_multi": (C.c_int, [_P, C.c_int, _P]),
    "vitta_conv2d_f16x3_infer": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P]),
    "vitta_gemm_f16x3_amax": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        _P, _P, C.c_int64, C.c_int, _P, _P, C.c_int64, _P, _P]),
    "vitta_split_block_elems": (C.c_int, []),
    "vitta_split_multi": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vitta_sgd_block_elems": (C.c_int, []),
    "vitta_sgd_step": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, _P]),
}

_lib = None
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


class VittaError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VittaError("%s not found: build it with `python -m vitta_b200.build` (or __graft_entry__.build()). "
                         "vitta_b200 has no CPU or eager-PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    if os.environ.get("VITTA_GEMM_CTA_PAIR") == "1":      # opt-in cta_group::2 kernels
        lib.vitta_gemm_set_cta_pair(1)
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().vitta_last_error()
        raise VittaError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


profile = None  # set to a list to record (name, start_event, end_event, args) for every call (bench.py attribution)


def call(name, *args):
    global launch_count
    lib = load()
    if profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        profile.append((name, e0, e1, args))
    else:
        rc = getattr(lib, name)(*args)
    launch_count += 1
    check(rc, name)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def chunking(O, Cc, I, frames=1):
    out = VittaChunking()
    check(load().vitta_stats_chunking(O, Cc, I, frames, C.byref(out)), "vitta_stats_chunking")
    return out


def ln_chunking(rows, Cc, want_stats=True):
    out = VittaChunking()
    check(load().vitta_ln_chunking(rows, Cc, 1 if want_stats else 0, C.byref(out)), "vitta_ln_chunking")
    return out


def make_bn(weight, bias, running_mean, running_var, eps):
    return VittaBN(weight.data_ptr(), bias.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(), float(eps))
