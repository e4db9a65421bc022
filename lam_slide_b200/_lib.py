"""ctypes binding of ``liblamslide.so`` (C ABI: ``include/lamslide.h``).

There is NO fallback: if the shared library is missing or a call fails, a ``LamSlideError`` is raised.
The library is built in-tree by ``__graft_entry__.build()`` (``lam_slide_b200/build.py``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, List, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblamslide.so")

# every symbol include/lamslide.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "lamslide_abi_version", "lamslide_last_error", "lamslide_launch_count",
    "lamslide_backbone_create", "lamslide_backbone_destroy", "lamslide_backbone_workspace_bytes",
    "lamslide_backbone_forward", "lamslide_ode_sample", "lamslide_ode_workspace_bytes", "lamslide_euler_step", "lamslide_setup_conditioning", "lamslide_ksample_errors", "lamslide_lincomb3",
    "lamslide_lincomb_n", "lamslide_rk_error_sumsq",
    "lamslide_first_stage_create", "lamslide_first_stage_destroy", "lamslide_first_stage_workspace_bytes",
    "lamslide_encode", "lamslide_decode", "lamslide_debug_gemm", "lamslide_debug_attention",
    "lamslide_debug_linear1", "lamslide_debug_linear2", "lamslide_debug_gemm_mainloop", "lamslide_debug_fused_mlp",
    "lamslide_debug_fs_linear", "lamslide_debug_fused_mlp_ln", "lamslide_debug_fs_linear_ln",
    "lamslide_profile_begin", "lamslide_profile_end", "lamslide_debug_kernel_count", "lamslide_debug_attention_trace",
]


class LamSlideError(RuntimeError):
    pass


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class BackboneConfig(C.Structure):
    _fields_ = [("depth", C.c_int32), ("in_dim", C.c_int32), ("hidden_size", C.c_int32), ("num_heads", C.c_int32),
                ("mlp_hidden", C.c_int32), ("vec_in_dim", C.c_int32), ("normalize", C.c_int32), ("theta", C.c_float)]


MAX_OUTPUTS = 4


class FirstStageConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dim_input", C.c_int32), ("dim_latent", C.c_int32), ("num_entities", C.c_int32),
        ("entity_dim", C.c_int32), ("qk_norm", C.c_int32),
        ("enc_num_latents", C.c_int32), ("enc_heads_cross", C.c_int32), ("enc_dim_head_cross", C.c_int32),
        ("enc_heads_latent", C.c_int32), ("enc_dim_head_latent", C.c_int32), ("enc_blocks_cross", C.c_int32),
        ("enc_blocks_attn", C.c_int32),
        ("dec_query_splitter", C.c_int32), ("dec_num_split", C.c_int32), ("dec_dim_query", C.c_int32),
        ("dec_heads_cross", C.c_int32), ("dec_dim_head_cross", C.c_int32), ("dec_heads_latent", C.c_int32),
        ("dec_dim_head_latent", C.c_int32), ("dec_blocks_cross", C.c_int32), ("dec_blocks_attn", C.c_int32),
        ("n_outputs", C.c_int32), ("output_names", C.c_char_p * MAX_OUTPUTS), ("output_dims", C.c_int32 * MAX_OUTPUTS),
        ("max_res", C.c_int32), ("n_atom_types", C.c_int32),
    ]


class FrameInputs(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("index0", C.c_void_p), ("index1", C.c_void_p), ("entities", C.c_void_p),
                ("mask", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Load liblamslide.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LamSlideError(
            f"{LIB_PATH} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(lam_slide_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.lamslide_abi_version.restype = C.c_int
    lib.lamslide_last_error.restype = C.c_char_p
    lib.lamslide_launch_count.restype = i64
    lib.lamslide_launch_count.argtypes = [i32]
    lib.lamslide_backbone_create.argtypes = [C.POINTER(BackboneConfig), C.POINTER(TensorDesc), i32, C.POINTER(vp)]
    lib.lamslide_backbone_destroy.argtypes = [vp]
    lib.lamslide_backbone_destroy.restype = None
    lib.lamslide_backbone_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.lamslide_backbone_workspace_bytes.restype = sz
    lib.lamslide_ode_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
    lib.lamslide_ode_workspace_bytes.restype = sz
    lib.lamslide_backbone_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.lamslide_ode_sample.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.lamslide_euler_step.argtypes = [vp, vp, i32, i32, C.c_float, C.c_float, vp, i64, vp]
    lib.lamslide_setup_conditioning.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.lamslide_ksample_errors.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.lamslide_lincomb3.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, i64, vp]
    lib.lamslide_lincomb_n.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_float), i32, i64, vp]
    lib.lamslide_rk_error_sumsq.argtypes = [C.POINTER(vp), C.POINTER(C.c_float), i32, vp, vp, C.c_double, C.c_double, i64, vp, vp]
    lib.lamslide_first_stage_create.argtypes = [C.POINTER(FirstStageConfig), C.POINTER(TensorDesc), i32, C.POINTER(vp)]
    lib.lamslide_first_stage_destroy.argtypes = [vp]
    lib.lamslide_first_stage_destroy.restype = None
    lib.lamslide_first_stage_workspace_bytes.argtypes = [vp, i32, i32]
    lib.lamslide_first_stage_workspace_bytes.restype = sz
    lib.lamslide_encode.argtypes = [vp, C.POINTER(FrameInputs), vp, i32, i32, vp, sz, vp]
    lib.lamslide_decode.argtypes = [vp, vp, vp, C.POINTER(vp), i32, i32, vp, sz, vp]
    lib.lamslide_debug_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.lamslide_debug_attention.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.lamslide_debug_linear1.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.c_float, i32, vp]
    lib.lamslide_debug_linear2.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.lamslide_debug_gemm_mainloop.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.lamslide_debug_fused_mlp.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.lamslide_debug_fused_mlp_ln.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.lamslide_debug_fs_linear_ln.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]
    lib.lamslide_debug_fs_linear.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, i32, i32, vp]
    lib.lamslide_profile_end.argtypes = [C.c_char_p, sz]
    lib.lamslide_debug_attention_trace.argtypes = [vp]
    lib.lamslide_debug_attention_trace.restype = None
    lib.lamslide_debug_kernel_count.argtypes = [C.c_char_p, i32]
    lib.lamslide_debug_kernel_count.restype = i64
    if lib.lamslide_abi_version() != 1:
        raise LamSlideError("liblamslide.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().lamslide_last_error().decode("utf-8", "replace")
        if status == -1:
            raise ValueError(msg)  # reference behaviour: ValueError / assert on bad shapes
        if status == -2:
            raise KeyError(msg)  # load_state_dict: missing / mis-shaped key
        raise LamSlideError(f"liblamslide error {status}: {msg}")


def launch_count(reset: bool = False) -> int:
    return int(load().lamslide_launch_count(1 if reset else 0))


def kernel_count(name: str, reset: bool = False) -> int:
    """Launches of one named kernel family since the last reset (test hook)."""
    return int(load().lamslide_debug_kernel_count(name.encode(), 1 if reset else 0))


def pack_state_dict(sd: Dict[str, torch.Tensor]) -> Tuple[C.Array, List[torch.Tensor]]:
    """Named host fp32 tensors for the create calls. Returns (array, keep-alive list)."""
    keep: List[torch.Tensor] = []
    items = []
    for name, t in sd.items():
        if not torch.is_tensor(t) or not t.dtype.is_floating_point:
            continue
        h = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
        if h.dim() > 4:
            h = h.reshape(h.shape[0], -1)
        keep.append(h)
        d = TensorDesc()
        d.name = name.encode()
        d.data = h.data_ptr()
        d.ndim = h.dim()
        for i, s in enumerate(h.shape):
            d.shape[i] = s
        items.append(d)
    arr = (TensorDesc * len(items))(*items)
    return arr, keep


def profile_begin() -> None:
    check(load().lamslide_profile_begin())


def profile_end() -> dict:
    import json
    buf = C.create_string_buffer(8192)
    check(load().lamslide_profile_end(buf, 8192))
    return json.loads(buf.value.decode())


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LamSlideError("lam_slide_b200 runs on CUDA tensors only (no CPU fallback)")
