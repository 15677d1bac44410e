"""ctypes binding of liblgd_b200.so (the C ABI declared in include/lgd_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or an entry point
fails, a RuntimeError is raised with lgd_last_error()."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

LGD_MAX_LEVELS = 8
_HERE = os.path.dirname(os.path.abspath(__file__))
# LGD_B200_LIB: another build of the same library (A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get("LGD_B200_LIB") or os.path.join(_HERE, "liblgd_b200.so")


class Pyramid(Structure):
    """lgd_pyramid_t"""
    _fields_ = [("num_levels", c_int32), ("batch", c_int32),
                ("h", c_int32 * LGD_MAX_LEVELS), ("w", c_int32 * LGD_MAX_LEVELS)]

    @classmethod
    def make(cls, batch, hws):
        if not (1 <= len(hws) <= LGD_MAX_LEVELS):
            raise ValueError("between 1 and %d pyramid levels are supported" % LGD_MAX_LEVELS)
        p = cls()
        p.num_levels = len(hws)
        p.batch = batch
        for i, (h, w) in enumerate(hws):
            p.h[i] = int(h)
            p.w[i] = int(w)
        return p


class StepDesc(Structure):
    """lgd_step_desc_t"""
    _fields_ = [("pyr", Pyramid), ("T", c_int32), ("img_h", c_int32), ("img_w", c_int32), ("heads", c_int32),
                ("max_n", c_int32), ("add_context_box", c_int32)]


_P = POINTER(Pyramid)
_D = POINTER(StepDesc)
_vp = c_void_p
_pp = POINTER(c_void_p)

# name -> (restype, argtypes). Must list every symbol of include/lgd_b200.h (tests/test_abi.py checks).
SIGNATURES = {
    "lgd_version": (c_int, []),
    "lgd_last_error": (c_char_p, []),
    "lgd_launch_count": (c_int64, []),
    "lgd_pyramid_elems": (c_int64, [_P]),
    "lgd_encode_descriptors": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "lgd_linear_workspace": (c_size_t, [c_int, c_int, c_int]),
    "lgd_linear_fwd": (c_int, [_vp, c_int, _vp, c_int, _vp, _vp, c_int, c_int, c_int, c_int, _vp, c_size_t, _vp]),
    "lgd_linear_bwd_input": (c_int, [_vp, c_int, _vp, c_int, _vp, c_int, c_int, c_int, c_int, c_int, _vp, c_size_t,
                                     _vp]),
    "lgd_linear_bwd_weight": (c_int, [_vp, c_int, _vp, c_int, _vp, c_int, _vp, c_int, c_int, c_int, c_int, _vp,
                                      c_size_t, _vp]),
    "lgd_layernorm_fwd": (c_int, [_vp, _vp, _vp, _vp, c_int, c_int, c_int, _vp]),
    "lgd_layernorm_bwd": (c_int, [_vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, _vp]),
    "lgd_rowvec_matmul_fwd": (c_int, [_vp, _vp, _vp, c_int, c_int, _vp]),
    "lgd_rowvec_matmul_bwd": (c_int, [_vp, _vp, _vp, _vp, _vp, c_int, c_int, _vp]),
    "lgd_segmax_concat_fwd": (c_int, [_vp, c_int, _vp, c_int, _vp, c_int, _vp, _vp, _vp]),
    "lgd_segmax_concat_bwd": (c_int, [_vp, c_int, c_int, _vp, c_int, _vp, _vp, _vp, _vp]),
    "lgd_attention_fwd": (c_int, [_vp, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, _vp, _vp, c_int, _vp, _vp, _vp]),
    "lgd_attention_bwd": (c_int, [_vp, _vp, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, _vp, _vp, c_int, _vp,
                                  _vp, _vp, _vp, _vp, _vp]),
    "lgd_box_ranges": (c_int, [_vp, c_int, c_int, c_int, _P, _vp, _vp]),
    "lgd_masks_from_ranges": (c_int, [_vp, c_int, _P, _vp, _vp]),
    "lgd_nchw_to_pyramid": (c_int, [POINTER(c_void_p), _P, _vp, c_int, _vp, _vp]),
    "lgd_pyramid_to_nchw": (c_int, [_vp, _P, POINTER(c_void_p), c_int, _vp]),
    "lgd_nhwc_to_pyramid": (c_int, [POINTER(c_void_p), _P, _vp, _vp, _vp]),
    "lgd_pack_conv_weight": (c_int, [_vp, _vp, c_int, _vp]),
    "lgd_unpack_conv_wgrad": (c_int, [_vp, _vp, c_int, _vp]),
    "lgd_conv3x3_num_tiles": (c_int, [_P]),
    "lgd_conv3x3_fwd_workspace": (c_size_t, [_P]),
    "lgd_conv3x3_fwd": (c_int, [_P, _vp, _vp, _vp, c_int, c_int, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_size_t,
                                _vp]),
    "lgd_conv3x3_fwd_addend": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, c_int, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp,
                                       c_size_t, _vp]),
    "lgd_pack_conv_weight_f16": (c_int, [_vp, _vp, c_int, _vp, _vp, c_size_t, _vp]),
    "lgd_pack_conv_weights_f16_multi": (c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), POINTER(c_void_p), _vp, _vp,
                                                c_size_t, _vp]),
    "lgd_conv3x3_dgrad_f16": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, c_size_t,
                                      _vp]),
    "lgd_conv3x3_fwd_f16": (c_int, [_P, _vp, _vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, _vp, _vp]),
    "lgd_conv3x3_wgrad_workspace": (c_size_t, [_P]),
    "lgd_conv3x3_wgrad": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_conv3x3_wgrad_f16": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_gn_finalize": (c_int, [_P, _vp, _vp, _vp]),
    "lgd_gn_apply_workspace": (c_size_t, [_P]),
    "lgd_gn_apply": (c_int, [_P, _vp, _vp, _vp, c_int, c_int, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_gn_bwd": (c_int, [_P, _vp, _vp, _vp, c_int, _vp, c_int, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_gn_bwd_workspace": (c_size_t, [_P]),
    "lgd_gn_bwd_tile_sums": (c_int, [_P, _vp, _vp, _vp, c_int, _vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_conv3x3_dgrad_f16_gnsums": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp]),
    "lgd_conv3x3_dgrad_f16_gnsums_y": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lgd_gn32_workspace": (c_size_t, [_P]),
    "lgd_gn32_stats": (c_int, [_P, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_gn32_apply": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _vp]),
    "lgd_gn32_bwd": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_maskpool_workspace": (c_size_t, [_P, c_int]),
    "lgd_maskpool_fwd": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, _vp, _vp, c_size_t, _vp]),
    "lgd_maskpool_bwd": (c_int, [_P, _vp, _vp, _vp, c_int, _vp, _vp]),
    "lgd_render_fwd": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, _vp, c_int, _vp, _vp]),
    "lgd_render_bwd": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, c_size_t, _vp]),
    "lgd_ctx_bias_table": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "lgd_ctx_bias_table_bwd": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "lgd_pyramid_channel_sums": (c_int, [_P, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_channel_sums_workspace": (c_size_t, [_P]),
    "lgd_in_stats": (c_int, [_P, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_in_mse_fwd": (c_int, [_P, _vp, _vp, _vp, _vp, c_float, _vp, _vp, c_size_t, _vp]),
    "lgd_in_mse_moments_fwd": (c_int, [_P, _vp, _vp, c_float, _vp, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_in_mse_bwd": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_float, _vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                               c_size_t, _vp]),
    "lgd_grad_scale": (c_int, [_vp, c_int, c_int, _vp, _vp, c_float, _vp, _vp]),
    "lgd_in_workspace": (c_size_t, [_P]),
    "lgd_relu_bwd": (c_int, [_vp, _vp, _vp, c_int64, c_int, _vp]),
    "lgd_round_tf32": (c_int, [_vp, _vp, c_int64, _vp]),
    "lgd_tf32_split": (c_int, [_vp, _vp, c_int64, _vp]),
    "lgd_axpy": (c_int, [_vp, _vp, c_int64, _vp]),
    "lgd_store_to_host": (c_int, [_vp, _vp, c_int, _vp]),
    "lgd_upload_from_host": (c_int, [_vp, _vp, c_int64, _vp]),
    "lgd_encode_descriptors_masks": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "lgd_tap_render_workspace": (c_size_t, [_P, c_int, c_int]),
    "lgd_tap_render_fwd": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_int, c_int, _vp, c_int, c_int, _vp, _vp, _vp, c_size_t,
                                   _vp]),
    "lgd_tap_render_bwd": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_encode_descriptors_norm": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp, _vp]),
    "lgd_masks_from_bytes": (c_int, [_vp, c_int64, _vp, _vp]),
    "lgd_dense_mask_workspace": (c_size_t, [_P, c_int]),
    "lgd_mask_gather": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, c_int, c_int, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_mask_paint": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _vp]),
    "lgd_conv3x3_fwd_f16_cols": (c_int, [_P, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "lgd_conv3x3_dgrad_f16_addend": (c_int, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, c_size_t,
                                             _vp]),
    "lgd_pack_conv_weight_f16_rows": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "lgd_unpack_conv_wgrad_rows": (c_int, [_vp, _vp, c_int, c_int, _vp]),
    "lgd_head_grad_workspace": (c_size_t, [_P, c_int]),
    "lgd_head_grad_prepare": (c_int, [_P, POINTER(c_void_p), POINTER(c_int64), c_int, _vp, _vp, _vp, _vp, c_size_t,
                                      _vp]),
    "lgd_mt_chunk_elems": (c_int, []),
    "lgd_mt_sgd": (c_int, [_vp, _vp, c_int, c_float, c_float, c_float, c_int, _vp]),
    "lgd_mt_adamw": (c_int, [_vp, _vp, c_int, c_float, c_float, c_float, c_float, c_float, c_int, _vp]),
    # ---- step runtime (chain.cu)
    "lgd_ctx_create": (c_void_p, []),
    "lgd_ctx_destroy": (None, [_vp]),
    "lgd_ctx_set_side_streams": (c_int, [_vp, c_int]),
    "lgd_ctx_wait_early_grads": (c_int, [_vp, _vp]),
    "lgd_ctx_set_token_programs": (c_int, [_vp, c_int]),
    "lgd_ctx_profile": (c_int, [_vp, c_int]),
    "lgd_ctx_profile_count": (c_int, [_vp]),
    "lgd_ctx_profile_get": (c_int, [_vp, c_int, POINTER(c_char_p), POINTER(c_float)]),
    "lgd_ctx_profile_reset": (None, [_vp]),
    "lgd_teacher_param_count": (c_int, []),
    "lgd_teacher_param_name": (c_char_p, [c_int]),
    "lgd_adapter_param_count": (c_int, []),
    "lgd_adapter_param_name": (c_char_p, [c_int]),
    "lgd_teacher_tape_bytes": (c_size_t, [_D]),
    "lgd_teacher_scratch_bytes": (c_size_t, [_D, c_int]),
    "lgd_distill_tape_bytes": (c_size_t, [_D]),
    "lgd_distill_scratch_bytes": (c_size_t, [_D, c_int]),
    "lgd_teacher_tape_field": (c_int, [_D, c_char_p, POINTER(c_size_t), POINTER(c_size_t)]),
    "lgd_distill_tape_field": (c_int, [_D, c_char_p, POINTER(c_size_t), POINTER(c_size_t)]),
    "lgd_teacher_forward": (c_int, [_vp, _D, _vp, _vp, _pp, _vp, _vp, _vp, c_size_t, _vp, c_size_t, _vp]),
    "lgd_teacher_backward": (c_int, [_vp, _D, _vp, _vp, _pp, _pp, _vp, _vp, c_size_t, _pp, _pp, _vp, c_int, _vp, _vp,
                                     c_size_t, _vp]),
    "lgd_distill_forward": (c_int, [_vp, _D, _vp, _vp, _pp, c_float, _vp, _vp, _vp, c_size_t, _vp, c_size_t, _vp]),
    "lgd_distill_backward": (c_int, [_vp, _D, _vp, _vp, _pp, c_float, _vp, _vp, c_size_t, _pp, _pp, _vp, c_int, _vp,
                                     _vp, c_size_t, _vp]),
}

_lib = None


def load():
    """Load liblgd_b200.so (no CUDA device needed for loading). Fails loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "lgd_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lgd_b200/csrc`. There is no CPU / PyTorch fallback for the distillation hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    """Current stream of the CURRENT device. The autograd functions of the plugin enter torch.cuda.device(<device of
    their tensors>) before they launch anything, so this is the stream of the tensors' device."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# When set to a list, call() brackets every entry point with CUDA events on the launching stream and appends
# (name, start_event, stop_event): bench.py's live per-kernel timing (roofline.achieved). None = no overhead.
profile = None


_CHAIN_CALLS = frozenset(("lgd_teacher_forward", "lgd_teacher_backward", "lgd_distill_forward", "lgd_distill_backward"))


def call(name, *args):
    """Call an int-returning entry point on the current stream; raise on a non-zero status."""
    lib = load()
    if profile is not None and name not in _CHAIN_CALLS:   # the chains time their inner calls themselves
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        profile.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.lgd_last_error().decode("utf-8", "replace")))


def query(name, *args):
    """Call a size / count query (no stream argument)."""
    return getattr(load(), name)(*args)
