"""ctypes binding of libblim_b200.so (the C ABI declared in include/blim_b200.h).

There is deliberately no fallback: if the shared object is missing or cannot be loaded, importing the engine fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLIM_LIB") or os.path.join(_HERE, "csrc", "libblim_b200.so")   # BLIM_LIB: A/B runs against another build

c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_f64 = ctypes.c_double
c_void_p = ctypes.c_void_p
c_char_p = ctypes.c_char_p


class ModelCfg(ctypes.Structure):
    """blim_model_cfg"""
    _fields_ = [
        ("hidden_size", c_i32), ("num_layers", c_i32), ("num_heads", c_i32), ("num_kv_heads", c_i32),
        ("head_dim", c_i32), ("intermediate_size", c_i32), ("vocab_size", c_i32), ("mm_hidden_size", c_i32),
        ("tokens_per_clip", c_i32), ("max_positions", c_i32), ("rms_norm_eps", c_f32),
        ("max_run_tokens", c_i32), ("max_prefix_tokens", c_i32), ("gemm_cta_group", c_i32),
    ]


class FuseCfg(ctypes.Structure):
    """blim_fuse_cfg"""
    _fields_ = [
        ("alpha", c_f64), ("c_query", c_f64), ("c_ens", c_f64),
        ("use_prior", c_i32), ("use_query", c_i32), ("cpn_zero_f64", c_i32),
    ]


class VisionCfg(ctypes.Structure):
    """blim_vision_cfg (include/blim_vision.h)"""
    _fields_ = [
        ("image_size", c_i32), ("patch_size", c_i32), ("frames_per_clip", c_i32), ("hidden_size", c_i32), ("num_layers", c_i32),
        ("num_heads", c_i32), ("mlp_hidden_size", c_i32), ("tome_tokens_per_frame", c_i32), ("max_clips", c_i32),
        ("ln_eps", c_f32), ("final_ln_eps", c_f32),
    ]


# every symbol include/blim_vision.h declares
VISION_SIGNATURES = {
    "blim_vision_create": (c_int, [ctypes.POINTER(VisionCfg), c_int, ctypes.POINTER(c_void_p)]),
    "blim_vision_destroy": (None, [c_void_p]),
    "blim_vision_last_error": (c_char_p, [c_void_p]),
    "blim_vision_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, ctypes.POINTER(c_i64), c_int, c_void_p]),
    "blim_vision_set_pos_embed": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "blim_vision_encode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "blim_vision_merge_tokens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "blim_vision_extract": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "blim_vision_kernel_launches": (c_i64, [c_void_p]),
    "blim_vision_gemm_flops": (c_f64, [c_void_p]),
    "blim_vision_profile": (c_int, [c_void_p, c_int]),
    "blim_vision_profile_read": (c_int, [c_void_p, c_int, ctypes.POINTER(c_f64), ctypes.POINTER(c_i64)]),
}

# name -> (restype, argtypes); every symbol include/blim_b200.h declares
SIGNATURES = {
    "blim_create": (c_int, [ctypes.POINTER(ModelCfg), c_int, ctypes.POINTER(c_void_p)]),
    "blim_destroy": (None, [c_void_p]),
    "blim_last_error": (c_char_p, [c_void_p]),
    "blim_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, ctypes.POINTER(c_i64), c_int, c_void_p]),
    "blim_set_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "blim_set_videos": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "blim_set_texts": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int]),
    "blim_set_video_vocab": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "blim_build_video_vocab": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "blim_set_tvg_prefix_length": (c_int, [c_void_p, c_int]),
    "blim_score_pairs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "blim_forward_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "blim_project_video": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "blim_forward_visual": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "blim_embed_tokens": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "blim_fuse_rerank": (c_int, [c_void_p, ctypes.POINTER(FuseCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "blim_rank_dense": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "blim_topk_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "blim_scatter_scores": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_f32, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "blim_comm_unique_id": (c_int, [c_void_p]),
    "blim_comm_init": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "blim_comm_destroy": (c_int, [c_void_p]),
    "blim_allgather_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "blim_act_dtype": (c_int, []),
    "blim_kernel_launches": (c_i64, [c_void_p]),
    "blim_gemm_flops": (c_f64, [c_void_p]),
    "blim_profile": (c_int, [c_void_p, c_int]),
    "blim_profile_read": (c_int, [c_void_p, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "blim_profile_read_detail": (c_int, [c_void_p, c_int, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), ctypes.POINTER(c_i64)]),
    "blim_debug_plan_batches": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "blim_debug_umma": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_uint32, ctypes.c_uint32,
                                ctypes.c_uint32, c_void_p]),
    "blim_debug_gemm": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_f32, c_int, c_void_p]),
}

_lib = None


def load():
    """Load the shared object (building it first if nvcc is around and it is missing) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("BLIM_LIB"):
        # never load a binary built from other sources: build() compares a digest of csrc/ + include/ + flags with the
        # stamp next to the .so and returns at once when it is fresh (the .so and its stamp travel to the GPU box together)
        from . import build as _build
        try:
            _build.build()
        except Exception as ex:
            if not _build.is_fresh():
                raise ImportError(f"{LIB_PATH} is missing or stale and could not be rebuilt ({ex}); there is no CPU fallback") from ex
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m blim_b200.build` (no CPU fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in list(SIGNATURES.items()) + list(VISION_SIGNATURES.items()):
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
