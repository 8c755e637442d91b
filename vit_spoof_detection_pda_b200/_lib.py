"""ctypes binding of libvitk.so (the C-ABI declared in include/vitk.h).

There is deliberately NO fallback: if the shared library is missing or a symbol cannot be bound this
module raises, and every op in the package fails loudly (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VITK_LIB=dev selects the development build (libvitk_dev.so: device-side tracer, timing experiments); tools only
LIB_PATH = os.path.join(HERE, "libvitk_dev.so" if os.environ.get("VITK_LIB") == "dev" else "libvitk.so")
FLAG_WGRAD_INLINE = 1

# enums (mirror include/vitk.h)
F32, BF16 = 0, 1
PREC_FP32, PREC_BF16 = 0, 1
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_QKV_SCATTER = 0, 1, 2, 3
LAYOUT_ROWMAJOR, LAYOUT_HEADMAJOR = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2

IMG, PATCH, NPATCH, NTOK, DIM, HEADS, HEAD_DIM, MLP, HEAD_HIDDEN = 224, 16, 196, 197, 768, 12, 64, 3072, 512

vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t


class VitkModel(C.Structure):
    _fields_ = [("batch", C.c_int32), ("depth", C.c_int32), ("num_classes", C.c_int32), ("precision", C.c_int32),
                ("training", C.c_int32), ("engine", C.c_int32),
                ("params", vp), ("params16", vp), ("grads", vp), ("workspace", vp), ("images", vp), ("logits", vp),
                ("mask1", vp), ("mask2", vp), ("dlogits", vp), ("frozen_backbone", C.c_int32), ("sm_budget", C.c_int32),
                ("images_u8", vp), ("norm_mean", C.c_float * 3), ("norm_std", C.c_float * 3),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/vitk.h declares
PROTOTYPES = {
    "vitk_version": (i32, []),
    "vitk_last_error_string": (C.c_char_p, []),
    "vitk_device_info": (i32, [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "vitk_is_dev_build": (i32, []),
    "vitk_trace_start": (i32, [vp, sz]),
    "vitk_trace_stop": (i32, []),
    "vitk_layernorm_fwd": (i32, [vp, i64, vp, vp, vp, i32, vp, vp, i32, f32, vp]),
    "vitk_layernorm_bwd": (i32, [vp, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "vitk_linear_fwd": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "vitk_linear_dgrad": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "vitk_linear_wgrad": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "vitk_linear_fwd_ws": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
    "vitk_linear_dgrad_ws": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]),
    "vitk_patch_embed_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "vitk_patch_embed_fwd_u8": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "vitk_u8_to_nchw": (i32, [vp, vp, vp, vp, i32, vp]),
    "vitk_patch_embed_wgrad": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "vitk_attn_fwd": (i32, [vp, vp, vp, i32, i32, vp]),
    "vitk_attn_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "vitk_head_save_floats": (sz, [i32]),
    "vitk_head_fwd": (i32, [vp] * 11 + [i32, i32, vp]),
    "vitk_head_bwd": (i32, [vp] * 14 + [i32, i32, vp]),
    "vitk_focal_fwd_bwd": (i32, [vp, vp, vp, f32, i32, f32, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "vitk_threshold_hist": (i32, [vp, vp, vp, i32, i32, vp, vp]),
    "vitk_threshold_counts": (i32, [vp, i32, vp, vp]),
    "vitk_grad_sumsq_scratch_floats": (sz, []),
    "vitk_grad_sumsq": (i32, [vp, sz, vp, vp, vp]),
    "vitk_adam_step_graph": (i32, [vp, vp, vp, vp, vp, sz, vp, vp, vp, f64, f64, f64, f64, i32, f32, vp, f32, vp]),
    "vitk_grad_scale": (i32, [vp, sz, f32, vp, f32, vp]),
    "vitk_adam_step": (i32, [vp, vp, vp, vp, vp, sz, f64, f64, f64, f64, f64, i32, i32, f32, vp, f32, vp]),
    "vitk_nvls_scratch_floats": (sz, []),
    "vitk_nvls_reduce_sumsq": (i32, [vp, vp, sz, f32, vp, vp, i32, i32, vp]),
    "vitk_nvls_adam_bcast": (i32, [vp, vp, vp, vp, vp, vp, sz, f64, f64, f64, f64, f64, i32, i32, f32, vp, i32, f32,
                             C.POINTER(i64), i32, vp]),
    "vitk_nvls_bcast_f32": (i32, [vp, vp, sz, i32, vp]),
    "vitk_cast_f32_to_bf16": (i32, [vp, vp, sz, vp]),
    "vitk_param_layout": (i64, [i32, i32, C.POINTER(i64), C.POINTER(i64), i32]),
    "vitk_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "vitk_model_fwd": (i32, [C.POINTER(VitkModel), vp]),
    "vitk_model_bwd_stage": (i32, [C.POINTER(VitkModel), i32, vp]),
    "vitk_model_num_bwd_stages": (i32, [i32]),
    "vitk_debug_set": (i32, [i32, i32]),
    "vitk_set_sm_budget": (i32, [i32]),
    "vitk_gemm_plan": (i32, [i32, i32, i32, i32, i32] + [C.POINTER(i32)] * 7),
    "vitk_gemm_plan_items": (i32, [i32, i32, i32, i32, i32, i32, C.POINTER(i32), i32]),
    "vitk_gemm_tail_plan": (i32, [i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "vitk_gemm_tail_scratch_floats": (sz, [i32]),
    "vitk_launch_count": (C.c_longlong, []),
    "vitk_prof_enable": (i32, [i32]),
    "vitk_prof_read": (i32, [C.POINTER(C.c_float), C.POINTER(i32), i32]),
}

# VITK_NVTX=1: NVTX ranges around the forward call, every backward stage, the gradient-norm / clip pass and the optimizer step
# (the ranges an nsys / ncu --nvtx capture filters on).  Off by default: the range calls are host work in a loop
# whose host side otherwise costs ~50 us per step when graphed.
NVTX = os.environ.get("VITK_NVTX") == "1"


class nvtx_range:
    """`with nvtx_range("vitk/forward"):` -- torch.cuda.nvtx push/pop when VITK_NVTX=1, nothing otherwise."""
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


_lib = None
launch_count = 0  # number of C-ABI compute calls made by this process (bench.py reports kernels separately)


def load():
    """Load libvitk.so and bind every prototype. Raises (never falls back) when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the vitk ops.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.vitk_version() != 101:
        raise RuntimeError("libvitk.so version mismatch: rebuild")
    if os.environ.get("VITK_NO_PDL") == "1":   # A/B timing: plain stream order instead of programmatic dependent launch
        lib.vitk_debug_set(6, 1)
    for kv in os.environ.get("VITK_KNOBS", "").split(","):   # A/B timing knobs, e.g. VITK_KNOBS=8:1 (wgrad on the main stream)
        if kv:
            lib.vitk_debug_set(int(kv.split(":")[0]), int(kv.split(":")[1]))
    _lib = lib
    return lib


def check(rc: int, what: str = "vitk"):
    if rc != 0:
        msg = load().vitk_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    global launch_count
    launch_count += 1
    check(getattr(load(), name)(*args), name)


def param_layout(depth: int, num_classes: int):
    lib = load()
    n = 4 + 12 * depth + 8
    offs = (i64 * n)()
    sizes = (i64 * n)()
    total = lib.vitk_param_layout(depth, num_classes, offs, sizes, n)
    if total < 0:
        raise RuntimeError("vitk_param_layout failed")
    return int(total), list(offs), list(sizes)
