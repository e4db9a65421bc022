"""Second stage: drop-in ``LatentSIV3`` (reference: ``src/models/components/latent/latent_si_v31.py:66-188``).

Same constructor kwargs, same ``forward(x, t, x_cond, x_cond_mask, y=None)`` signature, same state-dict keys
(SURVEY.md §8(b)), so a Hydra config can point ``backbone._target_`` at
``lam_slide_b200.backbone.LatentSIV3`` and load the reference's Lightning checkpoints unchanged.  The arithmetic
runs in ``liblamslide.so`` (tcgen05/TMEM/TMA GEMMs + fused epilogues + flash attention); there is no PyTorch or
CPU fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib

PATH_TYPES = {"Linear": 0, "GVP": 1}
PREDICTIONS = {"velocity": 0, "data": 1, "noise": 2, "score": 3}


class _Holder(nn.Module):
    """Parameter container (keeps the reference's module / key names); never called."""


def _mlp_embedder(in_dim: int, hidden: int) -> nn.Module:  # mmdit.py:116-124
    m = _Holder()
    m.in_layer = nn.Linear(in_dim, hidden, bias=True)
    m.out_layer = nn.Linear(hidden, hidden, bias=True)
    return m


def _parallel_block(hidden: int, heads: int, mlp_hidden: int) -> nn.Module:  # mmdit.py:215-238
    m = _Holder()
    m.linear1 = nn.Linear(hidden, hidden * 3 + mlp_hidden)
    m.linear2 = nn.Linear(hidden + mlp_hidden, hidden)
    m.norm = _Holder()
    m.norm.query_norm = _Holder()
    m.norm.query_norm.scale = nn.Parameter(torch.ones(hidden // heads))
    m.norm.key_norm = _Holder()
    m.norm.key_norm.scale = nn.Parameter(torch.ones(hidden // heads))
    return m


def _layer(hidden: int, heads: int, mlp_hidden: int) -> nn.Module:  # latent_si_v31.py:19-43
    m = _Holder()
    m.modulation = _Holder()
    m.modulation.lin = nn.Linear(hidden, 6 * hidden, bias=True)
    m.spatial_block = _parallel_block(hidden, heads, mlp_hidden)
    m.temporal_block = _parallel_block(hidden, heads, mlp_hidden)
    return m


class _DeviceWorkspace:
    """Grow-only byte buffer on one device (PyTorch owns the memory; the C ABI never allocates)."""

    ALIGN = 1024

    def __init__(self):
        self.buf: Optional[Tensor] = None

    def get(self, nbytes: int, device: torch.device, align: int = ALIGN):
        """``(aligned pointer, bytes available behind it)`` with at least ``nbytes`` available.  The capacity check is made on what
        is left AFTER alignment (torch's allocations are 512-byte aligned, so up to ``align - 1`` bytes are skipped), and the true
        remaining size goes to the C ABI, which checks it against what it needs."""
        def avail(buf: Tensor) -> int:
            p = buf.data_ptr()
            return buf.numel() - ((p + align - 1) // align * align - p)

        if self.buf is None or self.buf.device != device or avail(self.buf) < nbytes:
            self.buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
        p = self.buf.data_ptr()
        return (p + align - 1) // align * align, avail(self.buf)


class LatentSIV3(nn.Module):
    """B200-native LatentSIV3.  ``checkpointing`` is accepted and ignored (inference path);
    ``attention_mode`` other than ``"scaled_dot_product"`` raises (unused by every shipped config)."""

    def __init__(self, depth: int, in_dim: int, hidden_size: int, num_heads: int, vec_in_dim: Optional[int] = None,
                 mlp_ratio: int = 2, n_timesteps: int = 10, theta: int = 10_000, checkpointing: bool = False,
                 normalize: bool = False, attention_mode: str = "scaled_dot_product", share_weights: bool = False,
                 reset_parameters: bool = True):
        super().__init__()
        if hidden_size % num_heads != 0:  # latent_si_v31.py:92-95
            raise ValueError(f"Hidden size {hidden_size} must be divisible by num_heads {num_heads}")
        if attention_mode != "scaled_dot_product":
            raise NotImplementedError("only attention_mode='scaled_dot_product' is implemented (mmdit.py:58-72 is unused by the configs)")
        self.depth, self.in_dim, self.out_dim = depth, in_dim, in_dim
        self.hidden_size, self.num_heads = hidden_size, num_heads
        self.vec_in_dim = vec_in_dim
        self.mlp_hidden = int(hidden_size * mlp_ratio)
        self.n_timesteps, self.theta, self.normalize = n_timesteps, theta, normalize
        self.checkpointing, self.attention_mode = checkpointing, attention_mode

        self.x_in = nn.Linear(in_dim, hidden_size)
        self.cond_to_emb = nn.Linear(in_dim, hidden_size)
        self.mask_to_emb = nn.Embedding(2, hidden_size)
        self.time_in = _mlp_embedder(256, hidden_size)
        if vec_in_dim is not None:
            self.vec_in = _mlp_embedder(vec_in_dim, hidden_size)
        self.blocks = nn.ModuleList()
        if share_weights:
            blk = _layer(hidden_size, num_heads, self.mlp_hidden)
            for _ in range(depth):
                self.blocks.append(blk)
        else:
            for _ in range(depth):
                self.blocks.append(_layer(hidden_size, num_heads, self.mlp_hidden))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))
        self.linear = nn.Linear(hidden_size, self.out_dim)
        if reset_parameters:
            self.reset_parameters()
        self._handle = None
        self._packed_versions = None
        self._packed_device = None
        self._ws = _DeviceWorkspace()

    # -- latent_si_v31.py:127-156 (same distributions; the zero-initialised layers stay zero)
    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=1.0 / math.sqrt(2))
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        for emb in [self.time_in] + ([self.vec_in] if hasattr(self, "vec_in") else []):
            nn.init.normal_(emb.in_layer.weight, std=0.02)
            nn.init.normal_(emb.out_layer.weight, std=0.02)
        for blk in self.blocks:
            nn.init.constant_(blk.modulation.lin.weight, 0.0)
            nn.init.constant_(blk.modulation.lin.bias, 0.0)
        nn.init.constant_(self.linear.weight, 0.0)
        nn.init.constant_(self.linear.bias, 0.0)

    # -- weight packing ---------------------------------------------------------------------------------------------
    def _versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _release(self):
        if getattr(self, "_handle", None):
            _lib.load().lamslide_backbone_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def pack(self, device: Optional[torch.device] = None) -> None:
        """(Re)pack the current parameters into the library's device layouts (bf16 K-major GEMM operands, fused
        modulation matrix, transposed input embedding).  Called lazily by forward/ode_sample when parameters changed."""
        lib = _lib.load()
        self._release()
        cfg = _lib.BackboneConfig(self.depth, self.in_dim, self.hidden_size, self.num_heads, self.mlp_hidden,
                                  self.vec_in_dim or 0, 1 if self.normalize else 0, float(self.theta))
        arr, keep = _lib.pack_state_dict(self.state_dict())
        handle = C.c_void_p()
        dev = device if device is not None else next(self.parameters()).device
        with torch.cuda.device(dev):
            _lib.check(lib.lamslide_backbone_create(C.byref(cfg), arr, len(arr), C.byref(handle)))
        self._handle = handle
        self._packed_versions = self._versions()
        self._packed_device = torch.device(dev)

    def _ensure(self, device: torch.device):
        if self._handle is None or self._packed_versions != self._versions() or self._packed_device != device:
            self.pack(device)

    def _workspace(self, B: int, T: int, L: int, device: torch.device, num_steps: int = 0):
        lib = _lib.load()
        need = (lib.lamslide_ode_workspace_bytes(self._handle, B, T, L, num_steps) if num_steps else
                lib.lamslide_backbone_workspace_bytes(self._handle, B, T, L))
        return self._ws.get(need, device)

    @staticmethod
    def _prep(x: Tensor, dtype=torch.float32) -> Tensor:
        _lib.require_cuda(x)
        return x.to(dtype).contiguous()

    # -- LatentSIV3.forward (latent_si_v31.py:168-188) ------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: Tensor, t: Tensor, x_cond: Tensor, x_cond_mask: Tensor, y: Tensor = None) -> Tensor:
        B, T, L, D = x.size()
        if D != self.in_dim:
            raise ValueError(f"last dim {D} != in_dim {self.in_dim}")
        x, t, x_cond = self._prep(x), self._prep(t), self._prep(x_cond)
        m = self._prep(x_cond_mask, torch.int64)
        yy = self._prep(y) if y is not None else None
        self._ensure(x.device)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            ws, nbytes = self._workspace(B, T, L, x.device)
            _lib.check(_lib.load().lamslide_backbone_forward(
                self._handle, x.data_ptr(), t.data_ptr(), x_cond.data_ptr(), m.data_ptr(), _lib.ptr(yy), out.data_ptr(),
                B, T, L, ws, nbytes, _lib.current_stream_ptr()))
        return out

    # -- fused Sampler.sample_ode + Transport.get_drift + Euler (transport.py:158-202, 365-411; integrators.py:84-120) --
    @torch.no_grad()
    def ode_sample(self, init: Tensor, x_cond: Tensor, x_cond_mask: Tensor, y: Tensor = None, *, path_type: str = "GVP",
                   prediction: str = "data", num_steps: int = 10, return_states: bool = True,
                   return_velocities: bool = False):
        """Returns ``states`` [num_steps, B, T, L, D] like the reference's sample_fn (or only the final state when
        ``return_states=False``); with ``return_velocities`` also the ``num_steps-1`` drift evaluations."""
        B, T, L, D = init.size()
        x = self._prep(init).clone()
        x_cond = self._prep(x_cond)
        m = self._prep(x_cond_mask, torch.int64)
        yy = self._prep(y) if y is not None else None
        self._ensure(x.device)
        states = torch.empty((num_steps,) + tuple(x.shape), device=x.device, dtype=torch.float32) if return_states else None
        vel = torch.empty((num_steps - 1,) + tuple(x.shape), device=x.device, dtype=torch.float32) if return_velocities else None
        with torch.cuda.device(x.device):
            ws, nbytes = self._workspace(B, T, L, x.device, num_steps)
            _lib.check(_lib.load().lamslide_ode_sample(
                self._handle, x.data_ptr(), x_cond.data_ptr(), m.data_ptr(), _lib.ptr(yy), PATH_TYPES[path_type],
                PREDICTIONS[prediction], num_steps, _lib.ptr(states), _lib.ptr(vel), B, T, L, ws, nbytes,
                _lib.current_stream_ptr()))
        out = states if return_states else x
        return (out, vel) if return_velocities else out
