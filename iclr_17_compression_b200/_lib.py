"""ctypes binding of libb2c.so (include/b2c.h).  There is no fallback: if the library has not been
built, or the device is not sm_100, every entry point raises."""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# B2C_LIB: another build of the same library (A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get("B2C_LIB") or os.path.join(_PKG, "libb2c.so")

_c_f = ctypes.c_void_p      # device pointers travel as void*
_c_i = ctypes.c_int
_c_sz = ctypes.c_size_t
_PP = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol include/b2c.h declares
SIGNATURES = {
    "b2c_version": (_c_i, []),
    "b2c_last_error": (ctypes.c_char_p, []),
    "b2c_check_device": (_c_i, []),
    "b2c_launch_count": (ctypes.c_uint64, []),
    "b2c_conv2d_fwd": (_c_i, [_c_f, _c_f, _c_f, _c_f] + [_c_i] * 8 + [_c_f]),
    "b2c_conv_transpose2d_fwd": (_c_i, [_c_f, _c_f, _c_f, _c_f] + [_c_i] * 9 + [_c_f]),
    "b2c_conv_wgrad": (_c_i, [_c_f, _c_f, _c_f] + [_c_i] * 10 + [_c_f]),
    "b2c_channel_sum": (_c_i, [_c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
    "b2c_gdn_fwd": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_i, _c_f]),
    "b2c_gdn_bwd_workspace_bytes": (_c_sz, [_c_i, _c_i, _c_i]),
    "b2c_gdn_bwd": (_c_i, [_c_f] * 7 + [_c_i] * 4 + [_c_f, _c_sz, _c_f]),
    "b2c_quant_entropy_fwd": (_c_i, [_c_f, _c_f, _PP, _c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
    "b2c_quant_entropy_infer_workspace_bytes": (_c_sz, [_c_i, _c_i]),
    "b2c_quant_entropy_infer": (_c_i, [_c_f, _PP, _c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f, _c_sz, _c_f]),
    "b2c_entropy_lut_build": (_c_i, [_PP, _c_f, _c_i, _c_i, _c_f]),
    "b2c_quant_entropy_infer_lut": (_c_i, [_c_f, _PP, _c_f, _c_i, _c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f, _c_sz, _c_f]),
    "b2c_entropy_bwd": (_c_i, [_c_f, _PP, _c_f, _c_f, _PP, _c_i, _c_i, _c_i, _c_f]),
    "b2c_bitparm_fwd": (_c_i, [_c_f] * 5 + [_c_i] * 3 + [_c_f]),
    "b2c_bitparm_bwd_workspace_bytes": (_c_sz, [_c_i, _c_i]),
    "b2c_bitparm_bwd": (_c_i, [_c_f] * 9 + [_c_i] * 3 + [_c_f, _c_sz, _c_f]),
    "b2c_clamp_sse_workspace_bytes": (_c_sz, [_c_i, _c_i]),
    "b2c_clamp_sse_fwd": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_f, _c_sz, _c_f]),
    "b2c_clamp_sse_bwd": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_f, _c_sz, _c_f]),
}



class TcLayer(ctypes.Structure):
    """struct b2c_tc_layer (include/b2c.h)"""
    _fields_ = [("kind", _c_i), ("split3", _c_i), ("gdn", _c_i), ("out_f32_nchw", _c_i),
                ("B", _c_i), ("Hin", _c_i), ("Win", _c_i),
                ("in_hi", _c_f), ("in_lo", _c_f), ("w_hi", _c_f), ("w_lo", _c_f), ("bias", _c_f),
                ("gamma_hi", _c_f), ("gamma_lo", _c_f), ("beta_eff", _c_f), ("out_hi", _c_f), ("out_lo", _c_f),
                ("ref_image", _c_f), ("sse_partial", _c_f), ("saved_y", _c_f), ("saved_n", _c_f),
                ("channels", _c_i), ("grad_out", _c_f), ("grad_scale", ctypes.c_float), ("ksplit", _c_i),
                ("dbias_out", _c_f), ("lut", _c_f), ("lut_half_range", _c_i), ("bit_params", _c_f), ("z_nhwc_out", _c_f),
                ("bits_partial", _c_f), ("range_flag", _c_f)]


TC_CONV1, TC_CONV5S2, TC_DECONV5S2, TC_DECONV3, TC_DECONV3X = 0, 1, 2, 3, 4
TC_PACK_GDN, TC_PACK_GDN_T = 16, 17          # b2c_tc_pack_job.kind beyond the layer kinds


class TcPackJob(ctypes.Structure):
    """struct b2c_tc_pack_job (include/b2c.h)"""
    _fields_ = [("kind", _c_i), ("channels", _c_i), ("src", _c_f), ("src2", _c_f), ("dst_hi", _c_f), ("dst_lo", _c_f),
                ("beta_eff", _c_f)]

SIGNATURES.update({
    "b2c_tc_packed_weight_elems": (_c_sz, [_c_i, _c_i]),
    "b2c_tc_pack_weights": (_c_i, [_c_i, _c_i, _c_f, _c_f, _c_f, _c_f]),
    "b2c_tc_pack_gdn": (_c_i, [_c_i, _c_f, _c_f, _c_f, _c_f, _c_f, _c_f]),
    "b2c_tc_layer_run": (_c_i, [ctypes.POINTER(TcLayer), _c_f]),
    "b2c_tc_pack_gdn_t": (_c_i, [_c_i, _c_f, _c_f, _c_f, _c_f, _c_f]),
    "b2c_tc_tiles_per_image": (_c_i, [_c_i, _c_i, _c_i]),
    "b2c_tc_wgrad": (_c_i, [_c_f, _c_f, _c_f] + [_c_i] * 8 + [_c_f, _c_sz, _c_f]),
    "b2c_tc_wgrad_workspace_bytes": (_c_sz, [_c_i, _c_i, _c_i, _c_i]),
    "b2c_tc_wgrad_untap": (_c_i, [_c_f, _c_f, _c_i, _c_i, _c_f]),
    "b2c_tc_pack_many": (_c_i, [_c_f, _c_i, _c_f]),
    "b2c_tc_wgrad9": (_c_i, [_c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f, _c_sz, _c_f]),
    "b2c_tc_wgrad9_workspace_bytes": (_c_sz, [_c_i, _c_i, _c_i]),
    "b2c_allreduce_symm": (_c_i, [ctypes.POINTER(ctypes.c_void_p), _c_f, _c_i, _c_i, _c_sz, ctypes.c_float, _c_f]),
    "b2c_allreduce_symm_flag_words": (_c_i, []),
    "b2c_allreduce_symm_fused": (_c_i, [ctypes.POINTER(ctypes.c_void_p), _c_f, _c_i, _c_i, _c_sz, _c_sz, _c_sz,
                                        ctypes.c_float, _c_i, _c_f]),
    "b2c_allreduce_symm_range": (_c_i, [ctypes.POINTER(ctypes.c_void_p), _c_f, _c_i, _c_i, _c_sz, _c_sz, ctypes.c_float,
                                        _c_i, _c_f]),
    "b2c_rans_build_tables": (_c_i, [_PP, _c_i, _c_i, _c_i, _c_f, _c_f, _c_f]),
    "b2c_rans_num_blocks": (_c_i, [_c_i, _c_i, _c_i]),
    "b2c_rans_scratch_bytes": (_c_sz, [_c_i, _c_i, _c_i, _c_i]),
    "b2c_rans_encode": (_c_i, [_c_f, _c_f] + [_c_i] * 6 + [_c_f, _c_sz, _c_f, _c_f, _c_f, _c_sz, _c_f, _c_f]),
    "b2c_rans_decode": (_c_i, [_c_f, _c_f, _c_f, _c_f] + [_c_i] * 6 + [_c_f, _c_f, _c_f]),
    "b2c_resized_crop_flip": (_c_i, [_c_f, _c_sz, _c_f, _c_i, _c_i, _c_i, _c_i, _c_f, _c_f, _c_f]),
    "b2c_u8_to_f32": (_c_i, [_c_f, _c_f, _c_sz, _c_f]),
    "b2c_ssim_workspace_bytes": (_c_sz, [_c_i] * 5),
    "b2c_ssim_level": (_c_i, [_c_f, _c_f, _c_f, _c_i, ctypes.c_float, ctypes.c_float, _c_f, _c_i, _c_i, _c_i, _c_i, _c_f, _c_sz, _c_f]),
    "b2c_avgpool2": (_c_i, [_c_f, _c_f, _c_i, _c_i, _c_i, _c_i, _c_i, _c_f]),
    "b2c_msssim_combine": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_f]),
    "b2c_sigmoid_binarize_fwd": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
    "b2c_sigmoid_binarize_bwd": (_c_i, [_c_f, _c_f, _c_f, _c_f, _c_sz, _c_f]),
    "b2c_adam_clamp_step": (_c_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), _c_i, _c_f, _c_f, _c_f,
                                   _c_f, _c_f, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_float, _c_f]),
    "b2c_tc_xsq": (_c_i, [_c_f, _c_f, _c_f, _c_sz, _c_i, _c_f]),
    "b2c_tc_nhwc_colsum": (_c_i, [_c_f, _c_f, _c_sz, _c_f]),
    "b2c_gdn_reparam_bwd": (_c_i, [_c_f] * 6 + [_c_i, _c_f]),
    "b2c_partial_reduce": (_c_i, [_c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
    "b2c_infer_stats": (_c_i, [_c_f, _c_i, _c_i, _c_f, _c_f, _c_i, ctypes.c_double, ctypes.c_double, _c_f, _c_f, _c_f,
                               _c_f]),
    "b2c_tc_nchw_to_nhwc": (_c_i, [_c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
    "b2c_tc_nhwc_to_nchw": (_c_i, [_c_f, _c_f, _c_f, _c_i, _c_i, _c_i, _c_f]),
})

_lib = None
_lock = threading.Lock()


class B2CError(RuntimeError):
    pass


def load():
    """dlopen libb2c.so and declare all prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise B2CError(
                    f"{LIB_PATH} not found: build it with `python -m iclr_17_compression_b200.build` "
                    "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(code, what=""):
    if code != 0:
        msg = load().b2c_last_error().decode(errors="replace")
        raise B2CError(f"libb2c {what} failed ({code}): {msg}")


_device_ok = set()


def require_device(dev_index):
    """libb2c kernels exist for sm_100a only."""
    if dev_index in _device_ok:
        return
    check(load().b2c_check_device(), "device check")
    _device_ok.add(dev_index)


def launch_count():
    return int(load().b2c_launch_count())


# Parameters updated through raw device pointers (the fused clamp + Adam launch, also when replayed inside a CUDA
# graph) do not bump the tensors' autograd ``_version``; every cache keyed on parameter versions (packed bf16 weight
# panels, GDN packs, the entropy table, rANS tables) also keys on this counter, which those updates increment.
_weights_epoch = 0


def weights_epoch():
    return _weights_epoch


def bump_weights_epoch():
    global _weights_epoch
    _weights_epoch += 1
    return _weights_epoch


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return ctypes.cast(arr, _PP), arr     # keep `arr` alive during the call
