"""CPU: checkpoint ingestion (lam_slide_b200/checkpoint.py) on Lightning-format dicts built the way the reference writes them
(lightning_base.py:109-119: ``checkpoint["ema"] = self.ema.state_dict()``; ema.py:69-74: ``{"params", "decay"}``)."""
import os
import tempfile

import pytest
import torch

import lam_slide_b200 as P
from oracle import lamslide_oracle as O


def _lightning_ckpt(cfg, seed, with_ema=True, conditional=False):
    fs_sd = O.init_first_stage_params(cfg["first_stage"], seed)
    bb_sd = O.init_backbone_params(cfg["backbone"], seed + 1)
    sd = {f"backbone.{k}": v for k, v in bb_sd.items()}
    sd.update({f"first_stage_model.backbone.{k}": v for k, v in fs_sd.items()})
    sd["loss.weight"] = torch.ones(3)            # state the sampling path never reads
    sd["train_loss.mean_value"] = torch.zeros(())
    if conditional:
        sd["vec_in_embedding.weight"] = torch.randn(cfg["n_classes"], 256, generator=torch.Generator().manual_seed(seed + 2))
    ckpt = {"epoch": 7, "global_step": 1234, "state_dict": sd, "hyper_parameters": {"n_timesteps": cfg["T"]}}
    if with_ema:
        ckpt["ema"] = {"params": {k: v * 0.5 + 0.01 for k, v in sd.items()}, "decay": 0.999}
    return ckpt


@pytest.mark.parametrize("name,conditional", [("pedestrian", True), ("md17", False)])
def test_load_checkpoint_picks_ema_and_maps_keys(name, conditional):
    cfg = P.get_config(name, depth=1)
    ckpt = _lightning_ckpt(cfg, 11, with_ema=True, conditional=conditional)
    m = P.SecondStageSampler(cfg)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "last.ckpt")
        torch.save(ckpt, path)
        n = P.load_checkpoint(m, path)  # use_ema=True: what on_test_start swaps in
    assert n["backbone"] == len(m.backbone.state_dict()) and n["first_stage"] == len(m.first_stage_model.backbone.state_dict())
    for k, v in m.backbone.state_dict().items():
        assert torch.equal(v, ckpt["ema"]["params"][f"backbone.{k}"]), k
    for k, v in m.first_stage_model.backbone.state_dict().items():
        assert torch.equal(v, ckpt["ema"]["params"][f"first_stage_model.backbone.{k}"]), k
    if conditional:
        assert torch.equal(m.vec_in_embedding.weight.data, ckpt["ema"]["params"]["vec_in_embedding.weight"])
    P.load_checkpoint(m, ckpt, use_ema=False)  # raw weights; a dict works like a path
    k0 = next(iter(m.backbone.state_dict()))
    assert torch.equal(m.backbone.state_dict()[k0], ckpt["state_dict"][f"backbone.{k0}"])


def test_load_checkpoint_without_ema_and_separate_first_stage():
    cfg = P.get_config("nba", depth=1)
    ckpt = _lightning_ckpt(cfg, 21, with_ema=False, conditional=True)
    fs_only = {"state_dict": {k[len("first_stage_model."):]: v + 1.0 for k, v in ckpt["state_dict"].items()
                              if k.startswith("first_stage_model.backbone.")}}
    m = P.SecondStageSampler(cfg)
    P.load_checkpoint(m, ckpt, first_stage_ckpt=fs_only)
    k0 = next(iter(m.first_stage_model.backbone.state_dict()))
    assert torch.equal(m.first_stage_model.backbone.state_dict()[k0], fs_only["state_dict"][f"backbone.{k0}"])


def test_load_checkpoint_is_strict():
    cfg = P.get_config("pedestrian", depth=1)
    ckpt = _lightning_ckpt(cfg, 31, with_ema=False, conditional=True)
    m = P.SecondStageSampler(cfg)
    bad = {"state_dict": {k: v for k, v in ckpt["state_dict"].items() if not k.endswith("time_in.in_layer.weight")}}
    with pytest.raises(RuntimeError):
        P.load_checkpoint(m, bad)
    no_vec = {"state_dict": {k: v for k, v in ckpt["state_dict"].items() if k != "vec_in_embedding.weight"}}
    with pytest.raises(KeyError):
        P.load_checkpoint(m, no_vec)
    with pytest.raises(KeyError):
        P.load_checkpoint(m, {"state_dict": {"loss.weight": torch.ones(1)}})


def test_load_checkpoint_of_compiled_modules():
    """``compile: true`` runs (second_stage/peptide.py:58-60: ``torch.compile`` around the backbone and the first-stage model) store
    their parameters under ``backbone._orig_mod.*`` / ``first_stage_model._orig_mod.backbone.*`` — in ``state_dict`` and in the EMA copy."""
    cfg = P.get_config("nba", depth=1)
    ckpt = _lightning_ckpt(cfg, 41, with_ema=True, conditional=True)

    def compiled(k):
        if k.startswith("backbone."):
            return "backbone._orig_mod." + k[len("backbone."):]
        if k.startswith("first_stage_model."):
            return "first_stage_model._orig_mod." + k[len("first_stage_model."):]
        return k

    cc = {"state_dict": {compiled(k): v for k, v in ckpt["state_dict"].items()},
          "ema": {"params": {compiled(k): v for k, v in ckpt["ema"]["params"].items()}, "decay": 0.999}}
    assert "backbone._orig_mod.x_in.weight" in cc["state_dict"]
    m = P.SecondStageSampler(cfg)
    n = P.load_checkpoint(m, cc)
    assert n["backbone"] == len(m.backbone.state_dict()) and n["first_stage"] == len(m.first_stage_model.backbone.state_dict())
    for k, v in m.backbone.state_dict().items():
        assert torch.equal(v, ckpt["ema"]["params"][f"backbone.{k}"]), k
    for k, v in m.first_stage_model.backbone.state_dict().items():
        assert torch.equal(v, ckpt["ema"]["params"][f"first_stage_model.backbone.{k}"]), k
    # a compiled first-stage-only checkpoint (FirstStageLightningBase with a compiled backbone: ``backbone._orig_mod.*``)
    fs_only = {"state_dict": {"backbone._orig_mod." + k[len("first_stage_model.backbone."):]: v + 2.0
                              for k, v in ckpt["state_dict"].items() if k.startswith("first_stage_model.backbone.")}}
    P.load_checkpoint(m, cc, first_stage_ckpt=fs_only, use_ema=False)
    k0 = next(iter(m.first_stage_model.backbone.state_dict()))
    assert torch.equal(m.first_stage_model.backbone.state_dict()[k0], ckpt["state_dict"][f"first_stage_model.backbone.{k0}"] + 2.0)
    # an eager and a compiled alias of the same tensor in one dict is a malformed checkpoint
    both = dict(cc["state_dict"])
    both["backbone.x_in.weight"] = both["backbone._orig_mod.x_in.weight"]
    with pytest.raises(KeyError):
        P.load_checkpoint(m, {"state_dict": both}, use_ema=False)
