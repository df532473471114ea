"""ctypes binding of ``libafter_b200.so`` (declared in ``include/after_b200.h``).

There is no fallback: if the shared library is missing, cannot be loaded, or a call fails, a
``RuntimeError`` is raised -- never a silent PyTorch/CPU path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libafter_b200.so")

ABI_VERSION = 3
MAX_STAGES = 8

OK, IGNORED = 0, 1
PRECISION_FP32, PRECISION_BF16, PRECISION_FP32_SIMT = 0, 1, 2
PRECISIONS = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16, "fp32_simt": PRECISION_FP32_SIMT}
DTYPE_F32, DTYPE_F64, DTYPE_I64 = 0, 1, 2
MODULE_DENOISER, MODULE_AUTOENCODER, MODULE_STRUCTURE_ENCODER, MODULE_TIMBRE_ENCODER, MODULE_UNET = 0, 1, 2, 3, 4
MODULE_LATENT_MAP = 5
CFG_AUDIO, CFG_MIDI = 0, 1
KERNEL_CLASSES = {"tap_gemm_tc": 0, "tap_gemm_simt": 1, "attention": 2, "row_norm": 3, "act_operand": 4, "pqmf": 5,
                  "mlp_fused": 7}


class AfterConfig(C.Structure):
    """Mirror of ``struct after_config`` (field order and types must match the header)."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_channels", C.c_int32),
        ("seq_len", C.c_int32),
        ("embed_dim", C.c_int32),
        ("cond_dim", C.c_int32),
        ("noise_embed_dims", C.c_int32),
        ("n_layers", C.c_int32),
        ("mlp_multiplier", C.c_int32),
        ("tcond_dim", C.c_int32),
        ("local_attention_size", C.c_int32),
        ("attention_chunk_size", C.c_int32),
        ("drop_value", C.c_float),
        ("max_batch", C.c_int32),
        ("max_steps", C.c_int32),
        ("ae_in_channels", C.c_int32),
        ("ae_channels", C.c_int32),
        ("ae_z_channels", C.c_int32),
        ("ae_pqmf_bands", C.c_int32),
        ("ae_n_stages", C.c_int32),
        ("ae_multipliers", C.c_int32 * (MAX_STAGES + 1)),
        ("ae_dec_multipliers", C.c_int32 * (MAX_STAGES + 1)),
        ("ae_factors", C.c_int32 * MAX_STAGES),
        ("ae_dilations", C.c_int32 * MAX_STAGES),
        ("ae_num_blocks", C.c_int32),
        ("ae_kernel_size", C.c_int32),
        ("ae_use_loudness", C.c_int32),
        ("ae_max_samples", C.c_int64),
        ("se_in_size", C.c_int32),
        ("se_n_blocks", C.c_int32),
        ("se_channels", C.c_int32 * MAX_STAGES),
        ("se_kernel_size", C.c_int32),
        ("se_causal", C.c_int32),
        ("se_use_tanh", C.c_int32),
        ("te_in_size", C.c_int32),
        ("te_n_blocks", C.c_int32),
        ("te_channels", C.c_int32 * MAX_STAGES),
        ("te_kernel_sizes", C.c_int32 * MAX_STAGES),
        ("te_dilations", C.c_int32 * MAX_STAGES),
        ("te_res2net_scale", C.c_int32),
        ("te_se_channels", C.c_int32),
        ("te_attention_channels", C.c_int32),
        ("te_out_dim", C.c_int32),
        ("te_global_context", C.c_int32),
        ("te_use_tanh", C.c_int32),
        ("max_cache_size", C.c_int32),
        ("un_in_size", C.c_int32),
        ("un_out_size", C.c_int32),
        ("un_n_levels", C.c_int32),
        ("un_channels", C.c_int32 * MAX_STAGES),
        ("un_ratios", C.c_int32 * MAX_STAGES),
        ("un_kernel_size", C.c_int32),
        ("un_time_channels", C.c_int32),
        ("un_time_cond_in_channels", C.c_int32),
        ("un_time_cond_channels", C.c_int32),
        ("un_cond_channels", C.c_int32),
        ("un_n_attn_layers", C.c_int32),
        ("un_use_res_last", C.c_int32),
        ("stream_slots", C.c_int32),
        ("stream_max_frames", C.c_int32),
        ("stream_gn_frames", C.c_int32),
    ]


_F = C.POINTER(C.c_float)
_H = C.c_void_p

# name -> (restype, argtypes); every symbol include/after_b200.h declares
PROTOTYPES = {
    "after_abi_version": (C.c_int, []),
    "after_build_info": (C.c_char_p, []),
    "after_device_count": (C.c_int, []),
    "after_create": (C.c_int, [C.POINTER(AfterConfig), C.c_int, C.POINTER(_H)]),
    "after_destroy": (C.c_int, [_H]),
    "after_last_error": (C.c_char_p, [_H]),
    "after_load_tensor": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "after_finalize_weights": (C.c_int, [_H, C.c_int]),
    "after_denoiser_forward": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_unet_forward": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_model_forward": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "after_sample": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "after_sample_host": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "after_denoiser_forward_cached": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p]),
    "after_model_forward_cached": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "after_roll_cache": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "after_reset_cache": (C.c_int, [_H, C.c_void_p]),
    "after_sample_stream": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "after_ae_encode": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "after_ae_decode": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_ae_encode_stream": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "after_ae_decode_stream": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_structure_encode_stream": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_stream_reset": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "after_structure_encode": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_timbre_encode": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "after_latent_map": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "after_generate": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float,
                                 C.c_float, C.c_void_p]),
    "after_generate_host": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                      C.c_float, C.c_float, C.c_void_p]),
    "after_launch_count": (C.c_int64, [_H]),
    "after_device_bytes": (C.c_int64, [_H]),
    "after_ae_ratio": (C.c_int, [_H]),
    "after_profile_enable": (C.c_int, [_H, C.c_int]),
    "after_profile_read": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "after_debug_gemm": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it is absent or broken."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m after_b200.build` (needs nvcc). "
            "after_b200 has no CPU or PyTorch fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.after_abi_version() != ABI_VERSION:
        raise RuntimeError("libafter_b200.so ABI version mismatch; rebuild with `python -m after_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    if rc >= 0:
        return rc
    msg = load().after_last_error(handle)
    raise RuntimeError(f"libafter_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")
