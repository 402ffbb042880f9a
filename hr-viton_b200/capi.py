"""ctypes binding of include/hrviton_sm100.h.  There is NO CPU fallback: if the shared library is missing
or a call fails, this raises."""
import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libhrviton_sm100.so")

BF16, F32 = 0, 1  # hrv_dtype codes: 0 = the flavour's 16-bit storage type (bf16 for hrv_<op>, IEEE fp16 for hrv_<op>_f16), 1 = fp32
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
NHWC, NCHW = 0, 1
EPI_LINEAR, EPI_SPADE = 0, 1
RES_ADD, RES_GATE_RELU, RES_GATE_LRELU = 0, 1, 2

EXPORTS = ["hrv_conv2d_fwd", "hrv_instnorm_stats", "hrv_instnorm_stats2", "hrv_instnorm_apply", "hrv_norm_apply_affine", "hrv_norm_bwd_reduce", "hrv_norm_bwd_apply", "hrv_act_bwd_bias", "hrv_conv2d_wgrad", "hrv_conv2d_wgrad_workspace_bytes", "hrv_nchw_to_nhwc", "hrv_nhwc_to_nchw",
           "hrv_space_to_depth", "hrv_avgpool3s2", "hrv_bilinear_up2_add", "hrv_flow_warp", "hrv_bilinear_up2_bwd", "hrv_flow_warp_bwd", "hrv_pack_conv_weight", "hrv_space_to_depth_bwd", "hrv_maxpool2_fwd", "hrv_maxpool2_bwd",
           "hrv_avgpool3s2_bwd", "hrv_parse_blur_argmax", "hrv_gaussian_blur", "hrv_flow_warp_nchw", "hrv_onehot_u8", "hrv_im2col", "hrv_l1_sum", "hrv_l1_bwd", "hrv_last_error",
           "hrv_version", "hrv_device_sm_count"]


class Tensor(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("pitch", ctypes.c_int32), ("dtype", ctypes.c_int32)]


class ConvParams(ctypes.Structure):
    _fields_ = [("inp", Tensor), ("wpack", ctypes.c_void_p),
                ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("off_y", ctypes.c_int32), ("off_x", ctypes.c_int32),
                ("bk", ctypes.c_int32), ("bn", ctypes.c_int32), ("n_gemm", ctypes.c_int32),
                ("out", Tensor), ("out_layout", ctypes.c_int32), ("epi", ctypes.c_int32), ("act", ctypes.c_int32),
                ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("res", Tensor), ("x0", Tensor), ("x1", Tensor), ("x0_shift", ctypes.c_int32),
                ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p), ("noise", ctypes.c_void_p),
                ("noise_scale", ctypes.c_void_p), ("gamma_out", Tensor), ("res_mode", ctypes.c_int32)]


FLAVOURED = [n for n in EXPORTS if n not in ("hrv_last_error", "hrv_version", "hrv_device_sm_count")]
EXPORTS_F16 = [n + "_f16" for n in FLAVOURED]

_lib = None
_lib_f16 = None


class _F16View:
    """The fp16-storage flavour of the library: attribute hrv_<op> resolves to the symbol hrv_<op>_f16."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        return getattr(self._L, name + "_f16" if name in FLAVOURED else name)


def lib(dtype=None):
    """ctypes handle of libhrviton_sm100.so.  dtype torch.float16 selects the fp16-storage flavour of every entry point."""
    global _lib, _lib_f16
    if dtype is not None and str(dtype) == "torch.float16":
        if _lib_f16 is None:
            _lib_f16 = _F16View(lib())
        return _lib_f16
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("hrviton_b200: %s is missing — run `python __graft_entry__.py build` (no CPU fallback exists)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.hrv_last_error.restype = ctypes.c_char_p
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
    TP = ctypes.POINTER(Tensor)
    L.hrv_conv2d_fwd.argtypes = [ctypes.POINTER(ConvParams), vp]
    L.hrv_instnorm_stats.argtypes = [TP, i32, TP, i32, i32, vp, vp, f32, vp, vp, vp, ctypes.c_size_t, vp]
    L.hrv_instnorm_stats2.argtypes = [TP, i32, TP, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.hrv_instnorm_apply.argtypes = [TP, vp, vp, i32, TP, vp]
    L.hrv_norm_bwd_reduce.argtypes = [TP, TP, TP, TP, i32, TP, i32, i32, vp, vp, vp, vp, vp, i32, TP, TP, vp, vp]
    L.hrv_norm_apply_affine.argtypes = [TP, vp, vp, vp, vp, TP, i32, TP, vp]
    L.hrv_norm_bwd_apply.argtypes = [TP, TP, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, TP, vp, vp]
    L.hrv_act_bwd_bias.argtypes = [TP, TP, i32, TP, vp, vp]
    L.hrv_conv2d_wgrad.argtypes = [TP, TP, i32, i32, i32, vp, vp, ctypes.c_size_t, vp]
    L.hrv_conv2d_wgrad_workspace_bytes.argtypes = [TP, TP, i32, i32]
    L.hrv_nchw_to_nhwc.argtypes = [vp, i32, i32, i32, TP, vp]
    L.hrv_nhwc_to_nchw.argtypes = [TP, vp, vp]
    L.hrv_space_to_depth.argtypes = [TP, TP, vp]
    L.hrv_avgpool3s2.argtypes = [TP, TP, vp]
    L.hrv_bilinear_up2_add.argtypes = [TP, TP, TP, vp]
    L.hrv_flow_warp.argtypes = [vp, vp, vp, TP, TP, vp, vp, vp]
    L.hrv_bilinear_up2_bwd.argtypes = [TP, TP, vp]
    L.hrv_space_to_depth_bwd.argtypes = [TP, TP, vp]
    L.hrv_maxpool2_fwd.argtypes = [TP, TP, vp]
    L.hrv_maxpool2_bwd.argtypes = [TP, TP, TP, i32, vp]
    L.hrv_avgpool3s2_bwd.argtypes = [TP, TP, vp]
    L.hrv_parse_blur_argmax.argtypes = [vp, i32, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_int32), i32, vp, vp, ctypes.c_uint32, vp, vp]
    L.hrv_onehot_u8.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.hrv_gaussian_blur.argtypes = [vp, i32, i32, i32, i32, f32, vp, vp]
    L.hrv_flow_warp_nchw.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp, i32, i32, f32, f32, vp, vp, vp, vp, i32, vp]
    L.hrv_im2col.argtypes = [TP, TP, i32, i32, i32, vp]
    L.hrv_l1_sum.argtypes = [TP, TP, vp, vp]
    L.hrv_l1_bwd.argtypes = [TP, TP, vp, TP, i32, vp]
    L.hrv_pack_conv_weight.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp]
    L.hrv_flow_warp_bwd.argtypes = [vp, vp, vp, TP, TP, vp, vp, vp, vp]
    for name in FLAVOURED:
        twin = getattr(L, name + "_f16")
        twin.argtypes = getattr(L, name).argtypes
    for name in EXPORTS + EXPORTS_F16:
        fn = getattr(L, name)
        if name.startswith("hrv_conv2d_wgrad_workspace_bytes"):
            fn.restype = ctypes.c_size_t
        elif name != "hrv_last_error":
            fn.restype = ctypes.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().hrv_last_error().decode()))
