"""ctypes binding of the C ABI in include/clipglass_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``make -C clip_glass_b200/csrc``).  There is deliberately no fallback: if the
library is missing, or no sm_100 GPU is present when an engine is created, the
caller gets an exception, never a silent CPU path.
"""
from __future__ import annotations

import ctypes
import os

GLASS_MAX_BLOCKS = 12
_HERE = os.path.dirname(os.path.abspath(__file__))
# CLIPGLASS_LIB: load another build of the same ABI (A/B timing of kernel variants; never set in tests or bench)
LIB_PATH = os.environ.get("CLIPGLASS_LIB") or os.path.join(_HERE, "libclipglass_b200.so")

# glass_config.flags (include/clipglass_b200.h: GLASS_FLAG_*): cross-checked kernel variants
FLAG_FOLDED_RESAMPLE = 1
FLAG_EXACT_RESAMPLE = 2
FLAG_NO_PAIR_PACK = 4
FLAG_NO_I8_LAYOUT = 8
FLAG_SIMT_ATTENTION = 16
FLAG_NO_GRAPH = 32
FLAG_PROJ_FUSION = 64
FLAG_C1_NHWC = 128
FLAG_NO_ZERO_SKIP = 256
FLAG_NO_FUSED_DOWN = 512
FLAG_NO_IMAGE_FUSION = 1024
FLAG_FP32_BLUR = 2048
FLAG_NO_PROJ_ACC = 4096

SYMBOLS = [
    "glass_create", "glass_set_tensor", "glass_finalize", "glass_set_text_features", "glass_destroy",
    "glass_evaluate_host", "glass_evaluate_device", "glass_generate", "glass_clip_similarity",
    "glass_discriminate", "glass_last_error", "glass_launch_count", "glass_debug_read",
    "glass_set_debug", "glass_conv_breakdown", "glass_last_conv_time", "glass_set_batch_size",
    "glass_debug_build", "glass_set_range_check", "glass_range_report",
    "glass_last_images_gather", "glass_image_grid_u8", "glass_biggan_latent",
    "glass_text_create", "glass_text_set_tensor", "glass_text_finalize", "glass_text_set_image_features",
    "glass_text_generate", "glass_text_similarity", "glass_text_launch_count", "glass_text_last_error",
    "glass_text_destroy", "glass_text_set_timing", "glass_text_gemm_time",
    "glass_ga_uniform", "glass_ga_permutations", "glass_ga_tournament", "glass_ga_rand_count", "glass_ga_offspring",
    "glass_ga_dedup_workspace", "glass_ga_dedup_append", "glass_ga_pad", "glass_ga_cast_f32",
    "glass_ga_survive_workspace", "glass_ga_survive", "glass_ga_gather", "glass_ga_last_error",
]


class GlassConfig(ctypes.Structure):
    _fields_ = [
        ("num_blocks", ctypes.c_int32),
        ("channels", ctypes.c_int32 * GLASS_MAX_BLOCKS),
        ("latent_size", ctypes.c_int32),
        ("mapping_layers", ctypes.c_int32),
        ("batch_size", ctypes.c_int32),
        ("mbstd_group_size", ctypes.c_int32),
        ("use_discriminator", ctypes.c_int32),
        ("clip_width", ctypes.c_int32),
        ("clip_layers", ctypes.c_int32),
        ("clip_patch", ctypes.c_int32),
        ("clip_resolution", ctypes.c_int32),
        ("clip_embed_dim", ctypes.c_int32),
        ("max_population", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("conv_impl", ctypes.c_int32),
        ("flags", ctypes.c_int32),
    ]


class GlassNoise(ctypes.Structure):
    _fields_ = [
        ("noise", ctypes.c_void_p),
        ("noise_on_device", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("first_group", ctypes.c_uint64),
    ]


class GlassGaParams(ctypes.Structure):
    _fields_ = [
        ("sbx_eta", ctypes.c_double),
        ("sbx_prob", ctypes.c_double),
        ("sbx_prob_var", ctypes.c_double),
        ("pm_eta", ctypes.c_double),
        ("pm_prob", ctypes.c_double),
        ("n_var", ctypes.c_int32),
        ("integer", ctypes.c_int32),
    ]


class GlassError(RuntimeError):
    pass


class GlassArgError(AssertionError, ValueError):
    """Bad argument.  Subclasses AssertionError because the reference signals
    ``pop % minibatch != 0`` with ``assert`` (models.py:112,124)."""


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GlassError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the fitness path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.glass_create.argtypes = [ctypes.POINTER(GlassConfig), ctypes.POINTER(vp)]
    lib.glass_set_tensor.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t]
    lib.glass_finalize.argtypes = [vp]
    lib.glass_set_text_features.argtypes = [vp, vp, i32]
    lib.glass_destroy.argtypes = [vp]
    lib.glass_evaluate_host.argtypes = [vp, vp, i32, ctypes.POINTER(GlassNoise), vp, vp, vp]
    lib.glass_evaluate_device.argtypes = [vp, vp, i32, ctypes.POINTER(GlassNoise), vp, vp, vp]
    lib.glass_generate.argtypes = [vp, vp, i32, ctypes.POINTER(GlassNoise), vp, vp]
    lib.glass_clip_similarity.argtypes = [vp, vp, i32, vp, vp]
    lib.glass_discriminate.argtypes = [vp, vp, i32, vp, vp]
    lib.glass_last_error.restype = ctypes.c_char_p
    lib.glass_launch_count.argtypes = [vp]
    lib.glass_launch_count.restype = i64
    lib.glass_debug_read.argtypes = [vp, ctypes.c_char_p, vp, i64]
    lib.glass_debug_read.restype = i64
    lib.glass_set_debug.argtypes = [vp, i32, i32]
    lib.glass_set_batch_size.argtypes = [vp, i32]
    lib.glass_debug_build.argtypes = []
    lib.glass_last_images_gather.argtypes = [vp, vp, i32, vp, vp]
    lib.glass_image_grid_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.glass_biggan_latent.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.glass_set_range_check.argtypes = [vp, i32]
    lib.glass_range_report.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_float)]
    lib.glass_conv_breakdown.argtypes = [vp, vp, vp, i32]
    lib.glass_last_conv_time.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    u64, f64 = ctypes.c_uint64, ctypes.c_double
    lib.glass_ga_uniform.argtypes = [u64, u64, vp, i64, vp]
    lib.glass_ga_permutations.argtypes = [vp, i32, i32, vp, vp]
    lib.glass_ga_tournament.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.glass_ga_rand_count.argtypes = [i32, i32]
    lib.glass_ga_rand_count.restype = i64
    lib.glass_ga_offspring.argtypes = [ctypes.POINTER(GlassGaParams), vp, vp, vp, vp, i32, vp, vp]
    lib.glass_ga_dedup_workspace.argtypes = [i32]
    lib.glass_ga_dedup_workspace.restype = i64
    lib.glass_ga_dedup_append.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, f64, i32, vp, vp, vp]
    lib.glass_ga_pad.argtypes = [vp, i32, vp, i32, vp, vp]
    lib.glass_ga_cast_f32.argtypes = [vp, vp, i64, vp]
    lib.glass_ga_survive_workspace.argtypes = [i32]
    lib.glass_ga_survive_workspace.restype = i64
    lib.glass_ga_survive.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.glass_ga_gather.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, i32, vp]
    lib.glass_ga_last_error.restype = ctypes.c_char_p
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int:
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check_ga(rc: int) -> None:
    if rc < 0:
        msg = load_library().glass_ga_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise GlassArgError(msg)
        raise GlassError(f"clipglass_b200 error {rc}: {msg}")


def check(rc: int) -> None:
    if rc < 0:
        msg = load_library().glass_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise GlassArgError(msg)
        raise GlassError(f"clipglass_b200 error {rc}: {msg}")
