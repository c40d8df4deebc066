"""ctypes binding of libhpvg.so (the C ABI declared in include/hpvg.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is raised."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_longlong, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhpvg.so")


class HpvgError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise HpvgError(
            "libhpvg.so not found at %s — build it with `make -C mindspore-hp-vae-gan_b200` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

vp, i, f, ll, u64 = c_void_p, c_int, c_float, c_longlong, c_uint64

HPVG_MAX_LEVELS, HPVG_BLOCK_LAYERS = 16, 8


class HpvgBlock(ctypes.Structure):
    """include/hpvg.h: HpvgBlock."""
    _fields_ = [("n_layers", c_int), ("cin", c_int * HPVG_BLOCK_LAYERS),
                ("wimg", (c_void_p * 2) * HPVG_BLOCK_LAYERS), ("scale", c_void_p * HPVG_BLOCK_LAYERS),
                ("shift", c_void_p * HPVG_BLOCK_LAYERS)]


class HpvgGenerator(ctypes.Structure):
    """include/hpvg.h: HpvgGenerator (description of a prepared generator for hpvg_generator_sample)."""
    _fields_ = [("n_stages", c_int), ("nc_im", c_int), ("latent_dim", c_int),
                ("T", c_int * HPVG_MAX_LEVELS), ("H", c_int * HPVG_MAX_LEVELS), ("W", c_int * HPVG_MAX_LEVELS),
                ("noise_amp", c_float * HPVG_MAX_LEVELS), ("noise_seed", c_uint64 * HPVG_MAX_LEVELS),
                ("decoder", HpvgBlock), ("body", HpvgBlock * HPVG_MAX_LEVELS)]


_SIGS = {
    "hpvg_version": ([], c_int),
    "hpvg_last_error": ([], c_char_p),
    "hpvg_device_count": ([], c_int),
    "hpvg_init": ([i], c_int),
    "hpvg_sm_count": ([], c_int),
    "hpvg_malloc": ([POINTER(vp), c_size_t], c_int),
    "hpvg_free": ([vp], c_int),
    "hpvg_host_alloc": ([POINTER(vp), c_size_t], c_int),
    "hpvg_host_free": ([vp], c_int),
    "hpvg_memset": ([vp, i, c_size_t, vp], c_int),
    "hpvg_h2d": ([vp, vp, c_size_t, vp], c_int),
    "hpvg_d2h": ([vp, vp, c_size_t, vp], c_int),
    "hpvg_d2d": ([vp, vp, c_size_t, vp], c_int),
    "hpvg_stream_create": ([POINTER(vp)], c_int),
    "hpvg_stream_attach": ([vp], c_int),
    "hpvg_stream_detach": ([vp], c_int),
    "hpvg_stream_destroy": ([vp], c_int),
    "hpvg_stream_sync": ([vp], c_int),
    "hpvg_device_sync": ([], c_int),
    "hpvg_event_create": ([POINTER(vp)], c_int),
    "hpvg_event_destroy": ([vp], c_int),
    "hpvg_event_record": ([vp, vp], c_int),
    "hpvg_event_sync": ([vp], c_int),
    "hpvg_stream_wait_event": ([vp, vp], c_int),
    "hpvg_event_elapsed_ms": ([vp, vp, POINTER(f)], c_int),
    "hpvg_graph_begin": ([vp], c_int),
    "hpvg_graph_end": ([vp, POINTER(vp)], c_int),
    "hpvg_graph_launch": ([vp, vp], c_int),
    "hpvg_graph_destroy": ([vp], c_int),
    "hpvg_launch_count": ([], ll),
    "hpvg_set_pdl": ([i], c_int),
    "hpvg_block_fwd_eval_workspace": ([POINTER(HpvgBlock), i, i, i, i], c_size_t),
    "hpvg_block_fwd_eval": ([POINTER(HpvgBlock), i, i, i, i, i, vp, i, vp, vp, vp, c_size_t, vp], c_int),
    "hpvg_generator_sample_workspace": ([POINTER(HpvgGenerator), i], c_size_t),
    "hpvg_generator_sample": ([POINTER(HpvgGenerator), vp, i, u64, vp, vp, vp, vp, c_size_t, vp], c_int),
    "hpvg_pack_cl": ([vp, i, i, i, i, i, vp, i, i, i, vp], c_int),
    "hpvg_unpack_cl": ([vp, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_conv_wimg_bytes": ([i], c_int),
    "hpvg_conv_pack_weights": ([vp, i, i, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_conv_pack_weights_multi": ([i, POINTER(vp), POINTER(i), POINTER(i), POINTER(i), POINTER(i), POINTER(i), POINTER(i),
                                      POINTER(i), POINTER(i), POINTER(i), POINTER(vp), vp], c_int),
    "hpvg_conv_cl": ([i, i, i, i, i, vp, i, vp, vp, vp, i, i, vp, i, i, i, vp, vp, vp, i, vp], c_int),
    "hpvg_linear_taps": ([i, i, i, POINTER(c_int32), POINTER(c_int32), POINTER(f), POINTER(f)], c_int),
    "hpvg_linear_taps_dev": ([i, i, i, vp, vp, vp, vp, vp], c_int),
    "hpvg_resize3d_fwd": ([vp, i, i, i, i, i, vp, i, i, i, i, vp], c_int),
    "hpvg_resize3d_bwd": ([vp, i, i, i, i, i, vp, i, i, i, i, vp], c_int),
    "hpvg_upsample_noise_pack": ([vp, i, i, i, i, i, i, i, i, vp, f, u64, u64, vp, vp, vp, vp], c_int),
    "hpvg_frames_to_clip": ([vp, i, i, i, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_box_muller_inplace": ([vp, ll, vp], c_int),
    "hpvg_randn": ([vp, ll, u64, u64, vp, vp], c_int),
    "hpvg_counter_add": ([vp, u64, vp], c_int),
    "hpvg_bn_stats_cl": ([vp, ll, vp, vp, vp], c_int),
    "hpvg_bn_finalize": ([vp, vp, ll, vp, vp, f, f, vp, vp, vp, vp, vp, vp, vp], c_int),
    "hpvg_bn_apply_lrelu_cl": ([vp, ll, vp, vp, i, vp, vp], c_int),
    "hpvg_bn_train_apply_cl": ([vp, ll, vp, vp, vp, f, f, vp, vp, vp, i, vp, vp, vp], c_int),
    "hpvg_bn_moving_update_multi": ([i, POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), f, f, vp], c_int),
    "hpvg_bn_center_multi": ([i, POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp),
                              f, vp], c_int),
    "hpvg_sn_power_iter": ([vp, i, i, vp, vp, vp, vp, vp], c_int),
    "hpvg_sn_power_iter_multi": ([i, POINTER(vp), POINTER(i), POINTER(i), POINTER(vp), POINTER(vp), POINTER(vp),
                                  POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), vp], c_int),
    "hpvg_bn_fold_eval": ([vp, vp, vp, vp, f, vp, i, vp, vp, vp], c_int),
    "hpvg_affine_from_bias": ([vp, vp, i, vp, vp, vp], c_int),
    "hpvg_mse": ([vp, vp, ll, vp, vp], c_int),
    "hpvg_mean": ([vp, ll, vp, vp], c_int),
    "hpvg_kl": ([vp, vp, ll, vp, vp], c_int),
    "hpvg_reparam": ([vp, vp, vp, ll, vp, vp], c_int),
    "hpvg_reparam_bwd": ([vp, vp, vp, ll, vp, vp, vp], c_int),
    "hpvg_conv_wgrad_cl": ([vp, i, vp, i, i, i, i, i, vp, i, i, i, i, i, i, i, f, vp], c_int),
    "hpvg_lrelu_bwd_cl": ([vp, vp, ll, vp, vp], c_int),
    "hpvg_reflect_pad_cl": ([vp, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_slice_act_cl": ([vp, i, i, i, i, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_bn_bwd_cl": ([vp, vp, ll, vp, i, vp, vp, vp, i, vp], c_int),
    "hpvg_colsum_cl": ([vp, ll, vp, i, vp], c_int),
    "hpvg_mse_grad": ([vp, vp, ll, f, i, vp, vp], c_int),
    "hpvg_tanh_bwd": ([vp, vp, ll, vp, vp], c_int),
    "hpvg_axpby": ([f, vp, f, vp, ll, vp], c_int),
    "hpvg_fill": ([vp, f, ll, vp], c_int),
    "hpvg_gather_strided": ([vp, ll, ll, ll, vp, vp], c_int),
    "hpvg_channel_sum": ([vp, i, i, ll, i, vp, vp], c_int),
    "hpvg_kl_grad": ([vp, vp, ll, f, vp, vp, vp], c_int),
    "hpvg_sn_grad": ([vp, vp, vp, vp, vp, i, i, i, vp, vp], c_int),
    "hpvg_lerp": ([vp, vp, f, ll, vp, vp], c_int),
    "hpvg_gp_grad": ([vp, i, i, ll, f, vp, vp, vp], c_int),
    # fp32 channels-last / kind::tf32 twins (tf32 precision mode)
    "hpvg_pack_cl_f32": ([vp, i, i, i, i, i, vp, i, i, i, vp], c_int),
    "hpvg_unpack_cl_f32": ([vp, i, i, i, i, i, i, i, vp, vp], c_int),
    "hpvg_upsample_noise_pack_f32": ([vp, i, i, i, i, i, i, i, i, vp, f, u64, u64, vp, vp, vp, vp], c_int),
    "hpvg_conv_wgrad_cl_tf32": ([vp, i, vp, i, i, i, i, i, vp, i, i, i, i, i, i, i, f, vp], c_int),
    "hpvg_bn_stats_cl_f32": ([vp, ll, vp, vp, vp], c_int),
    "hpvg_bn_apply_lrelu_cl_f32": ([vp, ll, vp, vp, i, vp, vp], c_int),
    "hpvg_bn_train_apply_cl_f32": ([vp, ll, vp, vp, vp, f, f, vp, vp, vp, i, vp, vp, vp], c_int),
    "hpvg_lrelu_bwd_cl_f32": ([vp, vp, ll, vp, vp], c_int),
    "hpvg_bn_bwd_cl_f32": ([vp, vp, ll, vp, i, vp, vp, vp, i, vp], c_int),
    "hpvg_colsum_cl_f32": ([vp, ll, vp, i, vp], c_int),
    "hpvg_adam_clip_multi": ([i, POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), POINTER(ll), POINTER(f), f, f,
                              f, i, f, vp, vp], c_int),
}
# MindSpore ops.Custom(func_type="aot") entry points
_AOT = ["HpvgUpsampleTrilinear3D", "HpvgUpsampleTrilinear3DGrad", "HpvgConv3dBias", "HpvgConv3dBiasLRelu",
        "HpvgConv3dBiasTanh", "HpvgConv3dBiasLReluGrad", "HpvgBatchNorm3dLReluTrain", "HpvgBatchNorm3dLReluTrainGrad",
        "HpvgSpectralNormIter", "HpvgClipAdam", "HpvgMSELoss", "HpvgKLLoss"]

EXPORTED = list(_SIGS) + _AOT

for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _res
for _name in _AOT:
    _fn = getattr(lib, _name)
    _fn.argtypes = [c_int, POINTER(vp), POINTER(c_int), POINTER(POINTER(ctypes.c_int64)), POINTER(c_char_p), vp, vp]
    _fn.restype = c_int


def check(rc, what=""):
    if rc != 0:
        msg = lib.hpvg_last_error()
        raise HpvgError("%s failed (rc=%d): %s" % (what or "hpvg call", rc, msg.decode() if msg else ""))
