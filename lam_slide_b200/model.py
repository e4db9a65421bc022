"""End-to-end sampling wrapper — mirrors ``SecondStageCondLightningBase`` (``lightning_base.py:167-263``) and the
dataset ``Wrapper`` / ``CondWrapper`` classes (``second_stage/{peptide,md17,nba,pedestrian}.py``) for the calls on the
sampling path: ``forward``, ``encode``, ``decode``, ``prepare_batch``, ``setup_conditioning`` and ``sample``.
Lightning / Hydra / EMA / losses / metrics are out of scope; weights come from ``load_state_dict``
with the reference's key names (``backbone.*``, ``first_stage_model.backbone.*``, ``vec_in_embedding.weight``).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .backbone import LatentSIV3
from .configs import get_config
from .first_stage import FirstStage
from .transport import CreateTransport, Sampler, Transport

_FRAME_KEYS = ("atom14_pos", "aatype", "pos", "atom", "team", "group", "entities", "attention_mask")


class _FirstStageModel(nn.Module):
    """Stands in for the frozen ``FirstStageLightningBase`` (``lightning_base.py:140-164``): ``.backbone`` + encode/decode."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.backbone = FirstStage(cfg)

    def encode(self, batch: Dict[str, Tensor]) -> Tensor:
        return self.backbone.encode(batch)

    def decode(self, latents: Tensor, entities: Tensor) -> Dict[str, Tensor]:
        return self.backbone.decode(z=latents, entities=entities)


class SecondStageSampler(nn.Module):
    def __init__(self, cfg: dict, sampling_method: str = "ODE",
                 sampling_kwargs: Dict[str, Any] = {"sampling_method": "euler", "num_steps": 10}):
        super().__init__()
        self.cfg = cfg
        bb = cfg["backbone"]
        self.backbone = LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"],
                                   num_heads=bb["num_heads"], vec_in_dim=bb.get("vec_in_dim"), mlp_ratio=bb["mlp_ratio"],
                                   theta=bb.get("theta", 10_000), normalize=bb.get("normalize", False), n_timesteps=cfg["T"])
        self.first_stage_model = _FirstStageModel(cfg["first_stage"])
        self.si: Transport = CreateTransport(path_type=cfg["path_type"], prediction=cfg["prediction"])()
        if cfg.get("n_classes"):  # CondWrapper (nba.py:254-263, pedestrian.py:242-251)
            self.vec_in_embedding = nn.Embedding(cfg["n_classes"], bb["vec_in_dim"])
        self.hparams = SimpleNamespace(cond_idx=list(cfg["cond_idx"]), mask_cond_mean=cfg["mask_cond_mean"],
                                       sampling_method=sampling_method, sampling_kwargs=dict(sampling_kwargs),
                                       n_timesteps=cfg["T"])

    @classmethod
    def from_name(cls, name: str, **kw) -> "SecondStageSampler":
        return cls(get_config(name), **kw)

    @property
    def device(self) -> torch.device:
        return next(self.backbone.parameters()).device

    @property
    def transport(self) -> Transport:
        return self.si

    # lightning_base.py:173-174
    def forward(self, xt: Tensor, t: Tensor, **model_kwargs) -> Tensor:
        return self.backbone(x=xt, t=t, **model_kwargs)

    # second_stage/peptide.py:85-95 (and md17 / nba / pedestrian equivalents): "B T ... -> (B T) ..."
    @torch.no_grad()
    def encode(self, batch: Dict[str, Tensor]) -> Tensor:
        B = batch["entities"].shape[0]
        flat = {k: v.flatten(0, 1) for k, v in batch.items() if k in _FRAME_KEYS and torch.is_tensor(v)}
        latents = self.first_stage_model.encode(flat)
        return latents.unflatten(0, (B, -1))

    # second_stage/peptide.py:97-102: "(B T) L (A D) -> B T L A D"; others "(B T) L D -> B T L D"
    def decode(self, latents: Tensor, entities: Tensor, T: Optional[int] = None) -> Dict[str, Tensor]:
        preds = self.first_stage_model.decode(latents=latents, entities=entities)
        T = T or self.hparams.n_timesteps
        main = self.cfg["main_output"]
        pos = preds[main].unflatten(0, (-1, T))
        if main == "atom14_pos":
            pos = pos.unflatten(-1, (14, 3))
        return {main: pos}

    # lightning_base.py:240-263
    @torch.no_grad()
    def setup_conditioning(self, latents: Tensor) -> Tuple[Tensor, Tensor]:
        _lib.require_cuda(latents)
        latents = latents.to(torch.float32).contiguous()
        B, T, L, D = latents.shape
        x_cond = torch.empty_like(latents)
        mask = torch.empty(B, T, L, dtype=torch.int64, device=latents.device)
        c0, c1 = self.hparams.cond_idx
        with torch.cuda.device(latents.device):
            _lib.check(_lib.load().lamslide_setup_conditioning(latents.data_ptr(), x_cond.data_ptr(), mask.data_ptr(), B, T, L, D,
                                                               c0, c1, 1 if self.hparams.mask_cond_mean else 0,
                                                               _lib.current_stream_ptr()))
        return x_cond, mask

    # lightning_base.py:205-215 (+ CondWrapper.prepare_batch)
    @torch.no_grad()
    def prepare_batch(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        latents = self.encode(batch)
        x_cond, x_cond_mask = self.setup_conditioning(latents)
        batch["x1"] = latents
        batch["model_kwargs"] = {"x_cond": x_cond, "x_cond_mask": x_cond_mask}
        if hasattr(self, "vec_in_embedding") and "cond_scene" in batch:
            batch["model_kwargs"]["y"] = self.vec_in_embedding(batch["cond_scene"])
        return batch

    # lightning_base.py:217-238.  ``noise`` (optional) replaces torch.randn_like for reproducible parity tests.
    @torch.no_grad()
    def sample(self, batch: Dict[str, Tensor], noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
        sample_fn = Sampler(self.si).get_sample_fn(self.hparams.sampling_method, self.hparams.sampling_kwargs)
        B, T = batch["entities"].shape[:2]
        dev = self.device
        for key in list(batch.keys()):
            if isinstance(batch[key], torch.Tensor) and batch[key].device != dev:
                batch[key] = batch[key].to(dev, non_blocking=True)
        batch = self.prepare_batch(batch)
        model_kwargs = batch["model_kwargs"]
        x0 = torch.randn_like(model_kwargs["x_cond"]) if noise is None else noise.to(dev, non_blocking=True)
        latents = sample_fn(x0, self.forward, **model_kwargs)[-1]
        return self.decode(latents.flatten(0, 1), batch["entities"].flatten(0, 1), T=T)
