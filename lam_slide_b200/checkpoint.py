"""Checkpoint ingestion (first half): load a Lightning checkpoint of the reference straight into a
``SecondStageSampler`` — the weights the sampling path uses, optionally the EMA copy the reference evaluates with.

Format (``src/models/composites/lightning_base.py:63-70,109-119``; ``src/modules/ema.py:63-74``): ``torch.save``d dict with
``"state_dict"`` (the LightningModule's parameters: ``backbone.*`` = LatentSIV3, ``first_stage_model.backbone.*`` = frozen
first stage, ``vec_in_embedding.weight`` for the conditional wrappers, plus loss / metric state that sampling never reads)
and, when EMA was on, ``"ema": {"params": {<same keys>}, "decay": float}`` — ``on_validation_start`` / ``on_test_start`` swap
the EMA parameters in before every ``sample()``, so ``use_ema=True`` is the reference's evaluation behaviour.
A first-stage-only checkpoint (``FirstStageLightningBase``: keys ``backbone.*``) can be given separately, as the reference
does through ``first_stage_model`` in its Hydra config.  Output formats (XTC / PDB writers) are out of scope.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional, Union

import torch

from .model import SecondStageSampler

_PathOrDict = Union[str, Mapping[str, Any]]


def _read(ckpt: _PathOrDict) -> Mapping[str, Any]:
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        return torch.load(ckpt, map_location="cpu", weights_only=False)
    return ckpt


def _strip_compile_wrappers(sd: Mapping[str, Any]) -> Dict[str, Any]:
    """``torch.compile`` wraps a module in ``OptimizedModule``, whose parameters live under ``_orig_mod.``: with ``compile: true``
    (second_stage/peptide.py:58-60; the md17 / nba experiment configs) the reference's keys read ``backbone._orig_mod.x_in.weight``
    and ``first_stage_model._orig_mod.backbone.*``.  The wrapper adds no parameters of its own, so dropping the path component
    gives the eager names."""
    out: Dict[str, Any] = {}
    for k, v in sd.items():
        k2 = ".".join(part for part in k.split(".") if part != "_orig_mod")
        if k2 in out:
            raise KeyError(f"checkpoint holds '{k2}' both under its eager name and under a torch.compile alias")
        out[k2] = v
    return out


def select_state_dict(ckpt: _PathOrDict, use_ema: bool = True) -> Dict[str, torch.Tensor]:
    """The parameter dict ``sample()`` would run with: ``ckpt["ema"]["params"]`` when present and ``use_ema`` (lightning_base.py:
    63-70), else ``ckpt["state_dict"]``; a bare state dict passes through.  ``_orig_mod.`` path components (checkpoints of
    ``torch.compile``d modules) are removed."""
    c = _read(ckpt)
    if use_ema and isinstance(c.get("ema"), Mapping) and "params" in c["ema"]:
        return _strip_compile_wrappers(c["ema"]["params"])
    if "state_dict" in c:
        return _strip_compile_wrappers(c["state_dict"])
    return _strip_compile_wrappers(c)


def load_checkpoint(model: SecondStageSampler, ckpt: _PathOrDict, first_stage_ckpt: Optional[_PathOrDict] = None,
                    use_ema: bool = True) -> Dict[str, int]:
    """Loads ``backbone.*``, ``first_stage_model.backbone.*`` (or ``backbone.*`` of ``first_stage_ckpt``) and
    ``vec_in_embedding.weight`` strictly — a missing or mis-shaped tensor raises, as ``load_state_dict(strict=True)`` does in the
    reference — and ignores everything else (losses, metrics).  Returns the number of tensors loaded per component."""
    sd = select_state_dict(ckpt, use_ema)
    bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    fs = {k[len("first_stage_model.backbone."):]: v for k, v in sd.items() if k.startswith("first_stage_model.backbone.")}
    if first_stage_ckpt is not None:
        fsd = select_state_dict(first_stage_ckpt, use_ema)
        fs = {k[len("backbone."):]: v for k, v in fsd.items() if k.startswith("backbone.")}
    if not bb:
        raise KeyError("checkpoint has no 'backbone.*' parameters")
    if not fs:
        raise KeyError("no first-stage parameters: neither 'first_stage_model.backbone.*' in the checkpoint nor a first_stage_ckpt")
    model.backbone.load_state_dict(bb, strict=True)
    model.first_stage_model.backbone.load_state_dict(fs, strict=True)
    n_vec = 0
    if hasattr(model, "vec_in_embedding"):
        if "vec_in_embedding.weight" not in sd:
            raise KeyError("conditional model but the checkpoint has no 'vec_in_embedding.weight'")
        model.vec_in_embedding.load_state_dict({"weight": sd["vec_in_embedding.weight"]}, strict=True)
        n_vec = 1
    return {"backbone": len(bb), "first_stage": len(fs), "vec_in_embedding": n_vec}
