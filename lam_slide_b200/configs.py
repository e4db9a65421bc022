"""Hyper-parameters of the four shipped LaM-SLidE configurations (plain dicts, no Hydra).

Every number is transcribed from the reference YAMLs (SURVEY.md §8 dimension table):
``configs/model/{peptide,md17,nba,pedestrian}/{first,second}-stage*.yaml`` and
``configs/experiment/*/second-stage.yaml``.
"""
from __future__ import annotations

import copy
from typing import Dict

_ENC = dict(dim_head_cross=16, dim_head_latent=16, num_head_latent=2, num_block_cross=1,
            num_block_attn=1, qk_norm=True)
_DEC = dict(dim_head_cross=16, dim_head_latent=16, num_head_latent=2, num_block_cross=0,
            num_block_attn=1, qk_norm=True, dim_query=128)

CONFIGS: Dict[str, dict] = {
    # configs/model/peptide/first-stage.yaml:36-95, second-stage.yaml:7-60, experiment/peptide/second-stage.yaml:22-26
    "peptide": dict(
        first_stage=dict(
            kind="peptide", dim_input=256, dim_latent=96, num_entities=8, entity_dim=128, max_res=10,
            encoder=dict(_ENC, num_latents=2, num_head_cross=2),
            decoder=dict(_DEC, kind="DecoderQuerySplitter", num_head_cross=2, num_split=8,
                         outputs=(("atom14_pos", 42), ("aatype", 20))),
        ),
        backbone=dict(depth=7, in_dim=96, hidden_size=384, num_heads=16, mlp_ratio=4, vec_in_dim=None,
                      theta=10_000, normalize=False),
        T=1000, N=4, cond_idx=(0, 1), mask_cond_mean=True, path_type="GVP", prediction="data",
        n_classes=None, main_output="atom14_pos",
    ),
    # configs/model/md17/first-stage.yaml:27-91, second-stage.yaml:7-22, experiment/md17/second-stage.yaml:12-13
    "md17": dict(
        first_stage=dict(
            kind="md17", dim_input=128, dim_latent=32, num_entities=32, entity_dim=128, n_atom_types=10,
            encoder=dict(_ENC, num_latents=192, num_head_cross=8),
            decoder=dict(_DEC, kind="Decoder", num_head_cross=8, outputs=(("pos", 3), ("atom", 10))),
        ),
        backbone=dict(depth=4, in_dim=32, hidden_size=256, num_heads=16, mlp_ratio=2, vec_in_dim=None,
                      theta=10_000, normalize=False),
        T=30, N=21, cond_idx=(0, 10), mask_cond_mean=True, path_type="GVP", prediction="data",
        n_classes=None, main_output="pos",
    ),
    # configs/model/nba/first-stage.yaml:34-90, second-stage.yaml:6-22, second-stage_cond.yaml, experiment/nba/second-stage.yaml:11
    "nba": dict(
        first_stage=dict(
            kind="nba", dim_input=128, dim_latent=32, num_entities=11, entity_dim=128,
            encoder=dict(_ENC, num_latents=8, num_head_cross=2),
            decoder=dict(_DEC, kind="Decoder", num_head_cross=2,
                         outputs=(("pos", 2), ("team", 3), ("group", 2))),
        ),
        backbone=dict(depth=6, in_dim=32, hidden_size=256, num_heads=16, mlp_ratio=4, vec_in_dim=256,
                      theta=10_000, normalize=True),
        T=20, N=11, cond_idx=(0, 8), mask_cond_mean=True, path_type="GVP", prediction="data",
        n_classes=2, main_output="pos",
    ),
    # configs/model/pedestrian/first-stage.yaml:23-66, second-stage.yaml:10-34, second-stage_cond.yaml
    "pedestrian": dict(
        first_stage=dict(
            kind="pedestrian", dim_input=128, dim_latent=32, num_entities=10, entity_dim=128,
            encoder=dict(_ENC, num_latents=2, num_head_cross=4),
            decoder=dict(_DEC, kind="Decoder", num_head_cross=4, outputs=(("pos", 2),)),
        ),
        backbone=dict(depth=6, in_dim=32, hidden_size=128, num_heads=4, mlp_ratio=2, vec_in_dim=256,
                      theta=10_000, normalize=True),
        T=20, N=10, cond_idx=(0, 8), mask_cond_mean=True, path_type="GVP", prediction="data",
        n_classes=5, main_output="pos",
    ),
}


def get_config(name: str, **backbone_overrides) -> dict:
    """Deep copy of a named configuration; ``backbone_overrides`` patch the second-stage dict."""
    cfg = copy.deepcopy(CONFIGS[name])
    cfg["name"] = name
    cfg["backbone"].update(backbone_overrides)
    return cfg


def flops_per_eval(cfg: dict, T: int | None = None) -> float:
    """ALGORITHMIC FLOPs of one ``LatentSIV3.forward`` per sample — the formula of SURVEY.md §8(d)."""
    bb = cfg["backbone"]
    T = cfg["T"] if T is None else T
    L = cfg["first_stage"]["encoder"]["num_latents"]
    H, D, h, depth = bb["hidden_size"], bb["in_dim"], bb["num_heads"], bb["depth"]
    M = int(bb["mlp_ratio"] * H)
    hd = H // h
    tok = T * L
    gemm = tok * depth * 2 * (2 * H * (3 * H + M) + 2 * (H + M) * H)
    inout = tok * (2 * 2 * D * H + 2 * H * D)
    attn = depth * h * 4 * hd * (T * L * L + L * T * T)
    vec = 2 * 256 * H + 2 * H * H + depth * 2 * H * 6 * H + 2 * H * 2 * H
    if bb.get("vec_in_dim"):
        vec += 2 * bb["vec_in_dim"] * H + 2 * H * H
    return float(gemm + inout + attn + vec)


# GEMM FLOPs per frame of the first stage (FlopCounterMode probe)
FIRST_STAGE_FLOPS_PER_FRAME = {
    "peptide": 1.787e6 + 1.444e6,
    "md17": 18.63e6 + 17.57e6,
    "nba": 1.467e6 + 2.530e6,
    "pedestrian": 1.375e6 + 1.694e6,
}


def flops_per_trajectory(cfg: dict, num_steps: int = 10, T: int | None = None) -> float:
    T = cfg["T"] if T is None else T
    return (num_steps - 1) * flops_per_eval(cfg, T) + FIRST_STAGE_FLOPS_PER_FRAME[cfg["name"]] * T
