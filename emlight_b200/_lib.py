"""ctypes binding of libemlight_b200.so (include/emlight_b200.h).  No CPU fallback: a missing library raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemlight_b200.so")
ABI_VERSION = 23

EML_CONV_1x1, EML_CONV_3x3, EML_CONV_POOL2 = 0, 1, 2
EML_PREC_BF16, EML_PREC_BF16X3, EML_PREC_FP32 = 0, 1, 2
PRECISIONS = {"bf16": EML_PREC_BF16, "bf16x3": EML_PREC_BF16X3, "fp32": EML_PREC_FP32}


class ConvParams(Structure):
    _fields_ = [("in_", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("w_oihw", c_void_p),
                ("wpack", c_void_p), ("out", c_void_p), ("stats", c_void_p), ("stats_stride", c_long),
                ("B", c_int), ("H", c_int), ("W", c_int), ("C_in", c_int), ("in_pitch", c_int),
                ("C_out", c_int), ("out_pitch", c_int), ("out_choff", c_int),
                ("mode", c_int), ("relu", c_int), ("precision", c_int), ("plane_pixels", c_long)]


class DenseLayerParams(Structure):
    _fields_ = [("in_", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("wpack", c_void_p), ("bias9", c_void_p),
                ("out", c_void_p), ("B", c_int), ("H", c_int), ("W", c_int), ("C_in", c_int), ("in_pitch", c_int),
                ("growth", c_int), ("out_pitch", c_int), ("out_choff", c_int), ("precision", c_int), ("plane_pixels", c_long)]


# name -> (restype, argtypes); must list every symbol include/emlight_b200.h declares
SIGNATURES = {
    "eml_version": (c_int, []),
    "eml_error_string": (c_char_p, [c_int]),
    "eml_device_ok": (c_int, []),
    "eml_sg_render_fwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "eml_sg_render_params_fwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_void_p, c_long, c_void_p, c_long,
                                         c_void_p, c_long, c_float, c_void_p, c_long, c_void_p, c_int, c_int, c_void_p]),
    "eml_sg_render_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p]),
    "eml_sinkhorn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "eml_sinkhorn_fwdbwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                    c_float, c_void_p, c_size_t, c_void_p]),
    "eml_conv_wpack_bytes": (c_size_t, [c_int, c_int, c_int]),
    "eml_conv_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eml_conv_forward": (c_int, [POINTER(ConvParams), c_void_p]),
    "eml_dense_layer_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "eml_transition_planes_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "eml_dense_layer_forward": (c_int, [POINTER(DenseLayerParams), c_void_p]),
    "eml_dense_layer_wpack_bytes": (c_size_t, [c_int]),
    "eml_dense_layer_compose": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "eml_stem_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_long,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_bn_fold": (c_int, [c_void_p, c_long, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "eml_head_pool": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_linear_fp32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eml_bn_bwd_reduce": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_int, c_long, c_int, c_void_p, c_long, c_void_p]),
    "eml_bn_bwd_apply": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_int, c_int, c_long, c_int, c_void_p, c_long, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eml_wgrad_1x1": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                              c_long, c_int, c_void_p]),
    "eml_wgrad_3x3": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_wgrad_stem": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eml_im2col_lut": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_long, c_long, c_void_p]),
    "eml_im2col_lut_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                    c_long, c_long, c_void_p]),
    "eml_gemm_bf16": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eml_gemm_bf16_slices": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p, c_long, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_spectral_norm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "eml_gemm_pack_slices": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_long, c_void_p]),
    "eml_gemm_bf16_splitk": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_spade_modulate": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_long,
                                   c_int, c_int, c_void_p]),
    "eml_channel_stats": (c_int, [c_void_p, c_int, c_long, c_int, c_void_p, c_void_p]),
    "eml_bias_residual": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_long, c_int, c_void_p]),
    "eml_resize_nearest": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_resize_bilinear_nchw": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_instance_norm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "eml_tanh_to_nchw": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "eml_bias_act": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_long, c_int, c_void_p]),
    "eml_pool2d": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_needlet_basis": (c_int, [c_void_p, c_long, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_long, c_void_p]),
    "eml_split_bf16": (c_int, [c_void_p, c_long, c_int, c_long, c_void_p, c_void_p, c_int, c_void_p]),
    "eml_needlet_sparsify": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p]),
    "eml_extract_params": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "eml_tonemap_hdr": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_long, c_float, c_float, c_float, c_int, c_int, c_int, c_void_p]),
    "eml_col2im_lut": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_long, c_void_p]),
    "eml_col2im_csr": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_long, c_void_p]),
    "eml_act_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_long, c_int, c_void_p, c_void_p]),
    "eml_bias_act_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_long, c_int, c_void_p, c_void_p]),
    "eml_spade_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_void_p]),
    "eml_bn_free_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_int, c_long, c_int,
                                c_void_p]),
    "eml_instance_norm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_long, c_int, c_float, c_int, c_void_p,
                                      c_void_p, c_int, c_void_p]),
    "eml_im2col_lut_bf16_t": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_long, c_int,
                                      c_long, c_long, c_void_p]),
    "eml_upsample2_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_tanh_nchw_bwd": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_long, c_int, c_void_p, c_void_p]),
    "eml_pool2d_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "eml_loss_seed": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_long, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "eml_pool_act": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "eml_dense_bwd1_wpack_bytes": (c_size_t, [c_int]),
    "eml_dense_bwd1_supported": (c_int, [c_int, c_long, c_int]),
    "eml_dense_bwd1_pack": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "eml_dense_bwd1_prep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "eml_dense_bwd1": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_long, c_void_p, c_long, c_int,
                               c_void_p]),
    "eml_dense_bwd1_accum": (c_int, [c_void_p, c_long, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "eml_dense_bwd1_gather": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_long,
                                      c_void_p]),
    "eml_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    "eml_loss_reduce": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_long, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "emlight_b200: %s is missing -- build it with `python -m emlight_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.eml_version() != ABI_VERSION:
        raise RuntimeError("emlight_b200: ABI version mismatch (library %d, binding %d)" % (lib.eml_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().eml_error_string(code)
        raise RuntimeError("%s failed: %s (code %d)" % (what or "emlight_b200 call", msg.decode() if msg else "?", code))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    """The current device's current stream as a raw cudaStream_t (what every launch in the library takes)."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:                                       # no Stream object per call (~1000 calls per GenProjector iteration)
        return c_void_p(raw(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """Every tensor handed to a kernel lives on a CUDA device, and all of them on the same one."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("emlight_b200 runs on CUDA tensors only (sm_100a kernels; no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("emlight_b200: tensors on different devices (%s and %s)" % (dev, t.device))


def on_tensor_device(fn):
    """Decorator for the public entry points: run the call with the first CUDA tensor argument's device current, so that
    `stream_ptr()` (the current device's stream), the workspaces allocated inside and the raw-pointer launches all refer to the
    device the data is on -- a module on cuda:1 called while cuda:0 is current must not launch on cuda:0."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper
