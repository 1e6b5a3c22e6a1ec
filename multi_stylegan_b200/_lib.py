"""ctypes binding of the C-ABI in include/msg_b200.h.  Fails loudly when the library is absent."""
import ctypes
import os
import subprocess
import sys
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libmsg_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
INCLUDE_DIR = os.path.join(REPO_ROOT, "include")

MSG_OK, MSG_ERR_BAD_ARG, MSG_ERR_UNSUPPORTED, MSG_ERR_CUDA, MSG_ERR_WORKSPACE = 0, 1, 2, 3, 4
MSG_F32, MSG_F64 = 0, 1
CONV_AUTO, CONV_FORCE_SIMT, CONV_FORCE_TC = 0, 1, 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class ConvDesc(ctypes.Structure):
    """msg_conv_desc (include/msg_b200.h)."""
    _fields_ = [("B", ctypes.c_int), ("C", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("O", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
                ("stride_h", ctypes.c_int), ("stride_w", ctypes.c_int),
                ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int),
                ("OH", ctypes.c_int), ("OW", ctypes.c_int),
                ("w_batch_stride", ctypes.c_int64), ("layout", ctypes.c_int), ("w_transposed", ctypes.c_int)]


class ConvEpilogue(ctypes.Structure):
    """msg_conv_epilogue (include/msg_b200.h)."""
    _fields_ = [("bias", ctypes.c_void_p), ("noise", ctypes.c_void_p), ("noise_w", ctypes.c_void_p),
                ("noise_batch_stride", ctypes.c_int64), ("add", ctypes.c_void_p), ("act", ctypes.c_int),
                ("slope", ctypes.c_float), ("gain", ctypes.c_float),
                ("col_scale", ctypes.c_void_p), ("col_scale_batch_stride", ctypes.c_int64),
                ("y2", ctypes.c_void_p), ("y2_scale", ctypes.c_void_p), ("y2_scale_batch_stride", ctypes.c_int64)]


class LinearItem(ctypes.Structure):
    """msg_linear_item (include/msg_b200.h)."""
    _fields_ = [("W", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("N", ctypes.c_int), ("K", ctypes.c_int),
                ("in_off", ctypes.c_int), ("out_off", ctypes.c_int), ("w_off", ctypes.c_int), ("b_off", ctypes.c_int),
                ("alpha", ctypes.c_float), ("beta", ctypes.c_float)]


class LinearSlot(ctypes.Structure):
    """msg_linear_slot (include/msg_b200.h)."""
    _fields_ = [("in_off", ctypes.c_int), ("K", ctypes.c_int), ("first", ctypes.c_int), ("count", ctypes.c_int)]


class ProfileEntry(ctypes.Structure):
    """msg_profile_entry (include/msg_b200.h)."""
    _fields_ = [("kind", ctypes.c_int), ("taps", ctypes.c_int), ("k_channels", ctypes.c_int),
                ("n_channels", ctypes.c_int), ("pixels", ctypes.c_int64), ("launches", ctypes.c_int64),
                ("ms_total", ctypes.c_double), ("flops_per_launch", ctypes.c_double)]


def sources():
    return sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith(".cu"))


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/libmsg_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith(".cuh")] + \
        [os.path.join(INCLUDE_DIR, "msg_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_c = ctypes
_SIGNATURES = {
    "msg_abi_version": (_c.c_int, []),
    "msg_last_error": (_c.c_char_p, []),
    "msg_launch_count": (_c.c_uint64, []),
    "msg_tensor_core_path_available": (_c.c_int, []),
    "msg_debug_buffer": (_c.POINTER(_c.c_uint32), [_c.POINTER(_c.c_size_t)]),
    "msg_profile_enable": (None, [_c.c_int]),
    "msg_profile_summary": (_c.c_int, [_c.POINTER(ProfileEntry), _c.c_int]),
    "msg_style_mapping_supported": (_c.c_int, [_c.c_int, _c.c_int]),
    "msg_style_mapping_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p),
                                             _c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                             _c.c_float, _c.c_float, _c.c_void_p]),
    "msg_style_mapping_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                              _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_int,
                                              _c.c_float, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p]),
    "msg_linear_group_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.POINTER(LinearItem),
                                            _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "msg_linear_group_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                             _c.c_int64, _c.POINTER(LinearItem), _c.c_int, _c.POINTER(LinearSlot), _c.c_int,
                                             _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "msg_nl_split_pool": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                     _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "msg_nl_merge_unpool": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                       _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "msg_softmax_rows": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p]),
    "msg_softmax_rows_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p]),
    "msg_demod_factors_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                         _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p]),
    "msg_colsum_workspace": (_c.c_size_t, [_c.c_int64, _c.c_int]),
    "msg_colsum_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_void_p, _c.c_size_t,
                                   _c.c_void_p]),
    "msg_dot": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_float, _c.c_void_p, _c.c_void_p]),
    "msg_mbstd_workspace": (_c.c_size_t, [_c.c_int]),
    "msg_mbstd_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_float,
                                     _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "msg_mbstd_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int,
                                      _c.c_float, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "msg_tf32_mma_rate_probe": (_c.c_int, [_c.c_int, _c.c_void_p, _c.POINTER(_c.c_double), _c.c_void_p]),
    "msg_fused_bias_act": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                      _c.c_double, _c.c_double, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int,
                                      _c.c_void_p]),
    "msg_fused_bias_act_bwd_workspace": (_c.c_size_t, [_c.c_int64, _c.c_int64, _c.c_int64, _c.c_int]),
    "msg_fused_bias_act_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_double,
                                          _c.c_double, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_void_p,
                                          _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_upfirdn2d_out_size": (_c.c_int, [_c.c_int] * 6),
    "msg_upfirdn2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64] + [_c.c_int] * 13 +
                      [_c.c_int, _c.c_void_p]),
    "msg_upfirdn2d_bias_act": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64] + [_c.c_int] * 9 +
                               [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_float, _c.c_float,
                                _c.c_void_p]),
    "msg_upfirdn2d_bias_act_mod": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64] +
                                   [_c.c_int] * 9 + [_c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_int64,
                                                     _c.c_void_p, _c.c_int, _c.c_float, _c.c_float, _c.c_void_p,
                                                     _c.c_int64, _c.c_void_p]),
    "msg_demod_factors": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                     _c.c_int, _c.c_float, _c.c_void_p]),
    "msg_styled_act_bwd_workspace": (_c.c_size_t, [_c.c_int, _c.c_int64, _c.c_int]),
    "msg_styled_act_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                      _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int,
                                      _c.c_int64, _c.c_int, _c.c_float, _c.c_float, _c.c_void_p, _c.c_size_t,
                                      _c.c_void_p]),
    "msg_conv2d_workspace": (_c.c_size_t, [_c.POINTER(ConvDesc), _c.c_int, _c.c_int]),
    "msg_conv2d_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc), _c.c_float,
                                      _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_forward_fused": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc), _c.c_float,
                                            _c.POINTER(ConvEpilogue), _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_forward_cat2": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,
                                           _c.POINTER(ConvDesc), _c.c_float, _c.POINTER(ConvEpilogue), _c.c_void_p,
                                           _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_dgrad": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc), _c.c_float,
                                    _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_dgrad_acc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc), _c.c_float,
                                        _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_dgrad_mask_supported": (_c.c_int, [_c.POINTER(ConvDesc), _c.c_int]),
    "msg_conv2d_dgrad_mask": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc),
                                         _c.c_float, _c.c_void_p, _c.c_float, _c.c_float, _c.c_void_p, _c.c_size_t,
                                         _c.c_int, _c.c_void_p]),
    "msg_conv2d_wgrad": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(ConvDesc), _c.c_float,
                                    _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "msg_conv2d_last_engine": (_c.c_int, []),
    "msg_modulate_weights": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                        _c.c_int, _c.c_int, _c.c_float, _c.c_int, _c.c_void_p]),
    "msg_modulate_weights_bwd_workspace": (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "msg_modulate_weights_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                            _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_int, _c.c_void_p,
                                            _c.c_size_t, _c.c_void_p]),
    "msg_noise_bias_act": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                      _c.c_int, _c.c_int64, _c.c_int64, _c.c_float, _c.c_float, _c.c_void_p]),
    "msg_noise_bias_act_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                           _c.c_int64, _c.c_int, _c.c_int64, _c.c_float, _c.c_float, _c.c_void_p]),
    "msg_noise_bias_act_nhwc_bwd_workspace": (_c.c_size_t, [_c.c_int64, _c.c_int]),
    "msg_noise_bias_act_nhwc_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_int64, _c.c_int, _c.c_int64, _c.c_float, _c.c_float, _c.c_void_p,
                                               _c.c_size_t, _c.c_void_p]),
    "msg_affine_warp": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                   _c.c_int, _c.c_void_p]),
    "msg_affine_warp_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_int, _c.c_void_p]),
}


def exported_symbols():
    """Every entry point include/msg_b200.h declares."""
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  No fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "multi_stylegan_b200: CUDA library %s is missing; run `python -c \"import "
                        "__graft_entry__ as g; g.build()\"` (there is no CPU fallback)" % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != MSG_OK:
        msg = lib().msg_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))
