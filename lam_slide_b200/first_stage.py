"""First stage: UPT-style entity encoder / decoder behind the reference's ``BackboneBase`` interface.

``FirstStage`` mirrors ``BackboneBase`` (``src/models/composites/lightning_base.py:17-48``) composed with the dataset
``Backbone.prepare_inputs`` (``first_stage/{peptide,md17,nba,pedestrian}.py``), ``Encoder`` (``encoder.py:44-103``) and
``Decoder`` / ``DecoderQuerySplitter`` (``decoder.py:12-102, 313-411``):

    encode(batch: Dict[str, Tensor]) -> latents [F, num_latents, dim_latent]
    decode(z, entities)              -> Dict[name, Tensor [F, N, out_dim]]

with the reference's state-dict keys (``encoder.*``, ``decoder.*``, ``quant.0.*``, ``post_quant.1.*``,
``net_merge.{0,2}.*``, ``embedding_res.weight`` …) so ``first_stage_model.backbone`` checkpoints load unchanged.
All arithmetic runs in ``liblamslide.so`` (``lamslide_encode`` / ``lamslide_decode``); no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .backbone import _DeviceWorkspace

KINDS = {"peptide": 0, "md17": 1, "nba": 2, "pedestrian": 3}
# batch key of the coordinates / the two optional index tensors, per dataset
_POS_KEY = {"peptide": "atom14_pos", "md17": "pos", "nba": "pos", "pedestrian": "pos"}
_IDX_KEYS = {"peptide": ("aatype", None), "md17": ("atom", None), "nba": ("team", "group"), "pedestrian": (None, None)}


def first_stage_param_spec(cfg: dict) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, kind) for every entry of the reference's first-stage ``backbone`` state dict.
    kind: w (matrix), b (bias), ln_w / ln_b (LayerNorm affine), one (RMSNorm scale), emb, latents, buffer_*."""
    e, d = cfg["encoder"], cfg["decoder"]
    Din, D, E = cfg["dim_input"], cfg["dim_latent"], cfg["entity_dim"]
    Cd = Din + E
    dq = d["dim_query"]
    sp: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def lin(p, o, i, bias=True):
        sp[p + ".weight"] = ((o, i), "w")
        if bias:
            sp[p + ".bias"] = ((o,), "b")

    def ln(p, dim):
        sp[p + ".weight"] = ((dim,), "ln_w")
        sp[p + ".bias"] = ((dim,), "ln_b")

    def block(p, cross, dim, ctx, heads, dh):
        inner = heads * dh
        if cross:
            lin(p + "attn.fn.to_q", inner, dim, False)
            lin(p + "attn.fn.to_kv", 2 * inner, ctx, False)
        else:
            lin(p + "attn.fn.to_qkv", 3 * inner, dim, False)
        lin(p + "attn.fn.to_out", dim, inner)
        if e["qk_norm"]:
            sp[p + "attn.fn.norm.query_norm.scale"] = ((dh,), "one")
            sp[p + "attn.fn.norm.key_norm.scale"] = ((dh,), "one")
        ln(p + "attn.norm", dim)
        if cross:
            ln(p + "attn.norm_context", ctx)
        lin(p + "ff.fn.net.0.0", dim, dim)
        lin(p + "ff.fn.net.1", dim, dim)
        ln(p + "ff.norm", dim)

    sp["encoder.latents"] = ((e["num_latents"], D), "latents")
    sp["encoder.entity_embedding.embedding.weight"] = ((cfg["num_entities"], E), "entity")
    lin("encoder.mlp.0", D, Cd)
    lin("encoder.mlp.2", Cd, D)
    for i in range(e["num_block_cross"]):
        block(f"encoder.cross_attn_blocks.{i}.", True, D, Cd, e["num_head_cross"], e["dim_head_cross"])
    for i in range(e["num_block_attn"]):
        block(f"encoder.blocks_attn.{i}.", False, D, D, e["num_head_latent"], e["dim_head_latent"])
    sp["decoder.entity_embedding.embedding.weight"] = ((cfg["num_entities"], E), "entity")
    lin("decoder.query_mlp.1", dq, E)
    for i in range(d["num_block_attn"]):
        block(f"decoder.self_attn_blocks.{i}.", False, D, D, d["num_head_latent"], d["dim_head_latent"])
    for i in range(d["num_block_cross"]):
        block(f"decoder.cross_attn_blocks.{i}.", True, D, dq, d["num_head_cross"], d["dim_head_cross"])
    block("decoder.output_block.", True, dq, D, d["num_head_cross"], d["dim_head_cross"])
    for name, od in d["outputs"]:
        lin(f"decoder.output_layers.{name}.0", dq, dq)
        lin(f"decoder.output_layers.{name}.2", od, dq)
    if d["kind"] == "DecoderQuerySplitter":
        sp["decoder.extender.1.weight"] = ((D * d["num_split"], D, 1), "w")
        sp["decoder.extender.1.bias"] = ((D * d["num_split"],), "b")
    lin("quant.0", D, D)
    lin("post_quant.1", D, D)
    kind = cfg["kind"]
    if kind == "peptide":
        sp["embedding_res.weight"] = ((20, 64), "emb")
        sp["embed_res_pos.embeddings"] = ((cfg["max_res"], Din), "buffer_sincos")
        feat = 64 + 42
    elif kind == "md17":
        sp["embed_entity.embedding.weight"] = ((cfg["num_entities"], E), "entity")
        sp["embed_atom.weight"] = ((cfg["n_atom_types"], 64), "emb")
        sp["embed_pos.basis"] = ((3, 63), "buffer_basis")
        lin("embed_pos.mlp", 128, 129)
        feat = 64 + 128
    elif kind == "nba":
        sp["embed_entity.embedding.weight"] = ((cfg["num_entities"], E), "entity")
        sp["embed_team.weight"] = ((3, 32), "emb")
        sp["embed_group.weight"] = ((2, 32), "emb")
        feat = 2 + 32 + 32
    elif kind == "pedestrian":
        feat = 2
    else:
        raise ValueError(f"unknown first-stage kind {kind!r}")
    lin("net_merge.0", Din, feat)
    lin("net_merge.2", Din, Din)
    return sp


class _Node(nn.Module):
    """Nested container so parameters get the reference's dotted names."""


def _register(root: nn.Module, name: str, tensor: Tensor, buffer: bool, trainable: bool = True) -> None:
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=trainable))


class FirstStage(nn.Module):
    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.kind = cfg["kind"]
        self.dim_latent = cfg["dim_latent"]
        self.num_latents = cfg["encoder"]["num_latents"]
        self.output_names = [n for n, _ in cfg["decoder"]["outputs"]]
        self.output_dims = [d for _, d in cfg["decoder"]["outputs"]]
        spec = first_stage_param_spec(cfg)
        ent = None
        for name, (shape, kind) in spec.items():
            if kind == "w":
                t = torch.empty(shape)
                nn.init.xavier_uniform_(t.view(shape[0], -1), gain=1 / math.sqrt(2))
            elif kind in ("b", "ln_b"):
                t = torch.zeros(shape)
            elif kind in ("ln_w", "one"):
                t = torch.ones(shape)
            elif kind in ("emb", "latents"):
                t = torch.randn(shape)
            elif kind == "entity":  # EntityEmbeddingOrthogonal: one frozen orthogonal table shared by encoder & decoder
                if ent is None:
                    ent = torch.empty(shape)
                    nn.init.orthogonal_(ent)
                t = ent
            elif kind == "buffer_sincos":  # embeddings.py:6-25, 39-47
                omega = 1.0 / 10000 ** (torch.arange(shape[1] // 2, dtype=torch.float64) / (shape[1] / 2.0))
                out = torch.arange(shape[0], dtype=torch.float64)[:, None] * omega[None]
                t = torch.cat([out.sin(), out.cos()], dim=1).float()
            elif kind == "buffer_basis":  # embeddings.py:62-78
                k = shape[1] // 3
                f = (2.0 ** torch.arange(k).float()) * math.pi
                z = torch.zeros(k)
                t = torch.stack([torch.cat([f, z, z]), torch.cat([z, f, z]), torch.cat([z, z, f])])
            else:
                raise AssertionError(kind)
            _register(self, name, t, buffer=kind.startswith("buffer"), trainable=(kind != "entity"))
        self._handle = None
        self._packed_versions = None
        self._packed_device = None
        self._ws = _DeviceWorkspace()

    # -- packing --------------------------------------------------------------------------------------------------------
    def _versions(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _release(self):
        if getattr(self, "_handle", None):
            _lib.load().lamslide_first_stage_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def pack(self, device: torch.device) -> None:
        lib = _lib.load()
        self._release()
        c, e, d = self.cfg, self.cfg["encoder"], self.cfg["decoder"]
        fc = _lib.FirstStageConfig()
        fc.kind = KINDS[self.kind]
        fc.dim_input, fc.dim_latent, fc.num_entities, fc.entity_dim = c["dim_input"], c["dim_latent"], c["num_entities"], c["entity_dim"]
        fc.qk_norm = 1 if e["qk_norm"] else 0
        fc.enc_num_latents, fc.enc_heads_cross, fc.enc_dim_head_cross = e["num_latents"], e["num_head_cross"], e["dim_head_cross"]
        fc.enc_heads_latent, fc.enc_dim_head_latent = e["num_head_latent"], e["dim_head_latent"]
        fc.enc_blocks_cross, fc.enc_blocks_attn = e["num_block_cross"], e["num_block_attn"]
        fc.dec_query_splitter = 1 if d["kind"] == "DecoderQuerySplitter" else 0
        fc.dec_num_split = d.get("num_split", 1)
        fc.dec_dim_query = d["dim_query"]
        fc.dec_heads_cross, fc.dec_dim_head_cross = d["num_head_cross"], d["dim_head_cross"]
        fc.dec_heads_latent, fc.dec_dim_head_latent = d["num_head_latent"], d["dim_head_latent"]
        fc.dec_blocks_cross, fc.dec_blocks_attn = d["num_block_cross"], d["num_block_attn"]
        fc.n_outputs = len(self.output_names)
        self._name_bytes = [n.encode() for n in self.output_names]
        for i, (nb, od) in enumerate(zip(self._name_bytes, self.output_dims)):
            fc.output_names[i] = nb
            fc.output_dims[i] = od
        fc.max_res = c.get("max_res", 0) or 0
        fc.n_atom_types = c.get("n_atom_types", 0) or 0
        arr, keep = _lib.pack_state_dict(self.state_dict())
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.lamslide_first_stage_create(C.byref(fc), arr, len(arr), C.byref(handle)))
        self._handle = handle
        self._packed_versions = self._versions()
        self._packed_device = torch.device(device)

    def _ensure(self, device: torch.device):
        if self._handle is None or self._packed_versions != self._versions() or self._packed_device != device:
            self.pack(device)

    def _workspace(self, frames: int, N: int, device: torch.device):
        need = _lib.load().lamslide_first_stage_workspace_bytes(self._handle, frames, N)
        return self._ws.get(need, device, 256)

    # -- BackboneBase.encode (lightning_base.py:37-40) ----------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, batch: Dict[str, Tensor]) -> Tensor:
        pos = batch[_POS_KEY[self.kind]]
        _lib.require_cuda(pos)
        dev = pos.device
        ent = batch["entities"].to(torch.int64).contiguous()
        F_, N = ent.shape
        pos = pos.to(torch.float32).reshape(F_, N, -1).contiguous()
        k0, k1 = _IDX_KEYS[self.kind]
        i0 = batch[k0].to(torch.int64).contiguous() if k0 else None
        i1 = batch[k1].to(torch.int64).contiguous() if k1 else None
        mask = None
        if self.kind != "peptide" and batch.get("attention_mask") is not None:  # peptide.py:79 passes mask=None
            mask = batch["attention_mask"].to(torch.bool).contiguous().view(torch.uint8)
        self._ensure(dev)
        out = torch.empty(F_, self.num_latents, self.dim_latent, device=dev, dtype=torch.float32)
        fi = _lib.FrameInputs(pos.data_ptr(), _lib.ptr(i0), _lib.ptr(i1), ent.data_ptr(), _lib.ptr(mask))
        with torch.cuda.device(dev):
            ws, nbytes = self._workspace(F_, N, dev)
            _lib.check(_lib.load().lamslide_encode(self._handle, C.byref(fi), out.data_ptr(), F_, N, ws, nbytes,
                                                   _lib.current_stream_ptr()))
        return out

    # -- BackboneBase.decode (lightning_base.py:42-44) ----------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: Tensor, entities: Tensor) -> Dict[str, Tensor]:
        _lib.require_cuda(z)
        dev = z.device
        z = z.to(torch.float32).contiguous()
        ent = entities.to(device=dev, dtype=torch.int64).contiguous()
        F_, N = ent.shape
        if tuple(z.shape) != (F_, self.num_latents, self.dim_latent):
            raise ValueError(f"latents shape {tuple(z.shape)} != {(F_, self.num_latents, self.dim_latent)}")
        self._ensure(dev)
        outs = [torch.empty(F_, N, od, device=dev, dtype=torch.float32) for od in self.output_dims]
        arr = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        with torch.cuda.device(dev):
            ws, nbytes = self._workspace(F_, N, dev)
            _lib.check(_lib.load().lamslide_decode(self._handle, z.data_ptr(), ent.data_ptr(), arr, F_, N, ws, nbytes,
                                                   _lib.current_stream_ptr()))
        return dict(zip(self.output_names, outs))

    def forward(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:  # lightning_base.py:33-35
        return self.decode(self.encode(batch), batch["entities"])
