"""Synthetic batches of each configuration's shape and random-init helpers for benchmarks.
There is no network for datasets / checkpoints, so throughput is measured on these (bench.py says so in ``data``)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor


def synthetic_batch(cfg: dict, B: int, seed: int, T: Optional[int] = None, pin: bool = False) -> Dict[str, Tensor]:
    """Host batch with the reference's batch schema (keys pos/atom14_pos, aatype/atom/team/group, entities,
    attention_mask, cond_scene); entity ids are ``randperm(num_entities)[:N]`` per sample as the datasets do
    (e.g. src/datasets/nba.py:142) — ``arange(N)`` for peptides (src/modules/sampling.py:36)."""
    g = torch.Generator().manual_seed(seed)
    name = cfg["name"]
    T = cfg["T"] if T is None else T
    N = cfg["N"]
    n_ent = cfg["first_stage"]["num_entities"]
    b: Dict[str, Tensor] = {}

    def entities(n_valid):
        ent = torch.zeros(B, N, dtype=torch.int64)
        for i in range(B):
            nv = int(n_valid[i])
            ent[i, :nv] = torch.randperm(n_ent, generator=g)[:nv]
        return ent[:, None, :].expand(B, T, N).contiguous()

    if name == "peptide":
        b["atom14_pos"] = torch.randn(B, T, N, 14, 3, generator=g)
        b["aatype"] = torch.randint(0, 20, (B, 1, N), generator=g).expand(B, T, N).contiguous()
        b["entities"] = torch.arange(N)[None, None, :].expand(B, T, N).contiguous()
    elif name == "md17":
        b["pos"] = torch.randn(B, T, N, 3, generator=g)
        z = torch.tensor(([6] * 9 + [8] * 4 + [1] * 8)[:N])
        b["atom"] = z[None, None, :].expand(B, T, N).contiguous()
        b["entities"] = entities([N] * B)
        b["attention_mask"] = torch.ones(B, T, N, dtype=torch.bool)
    elif name == "nba":
        b["pos"] = torch.randn(B, T, N, 2, generator=g)
        b["team"] = torch.tensor([0] + [1] * 5 + [2] * 5)[None, None, :].expand(B, T, N).contiguous()
        b["group"] = torch.tensor([0] + [1] * 10)[None, None, :].expand(B, T, N).contiguous()
        b["entities"] = entities([N] * B)
        b["attention_mask"] = torch.ones(B, T, N, dtype=torch.bool)
        b["cond_scene"] = torch.randint(0, cfg["n_classes"], (B,), generator=g)
    elif name == "pedestrian":
        nv = torch.randint(1, N + 1, (B,), generator=g)
        valid = torch.arange(N)[None, :] < nv[:, None]
        b["pos"] = torch.randn(B, T, N, 2, generator=g) * valid[:, None, :, None]
        b["entities"] = entities(nv) * valid[:, None, :]
        b["attention_mask"] = b["pos"][..., 0] != 0
        b["cond_scene"] = torch.randint(0, cfg["n_classes"], (B,), generator=g)
    else:
        raise ValueError(name)
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    return b


@torch.no_grad()
def randomize_zero_init(module: torch.nn.Module, seed: int = 0, std: float = 0.02) -> None:
    """The reference zero-initialises every ``modulation.lin`` and the output ``linear`` (latent_si_v31.py:152-156), so a
    freshly constructed backbone outputs exactly 0.  Benchmarks / parity runs on random-init weights re-draw those
    tensors (and all biases) from N(0, std) so every kernel does real work."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if not p.requires_grad:
            continue
        if float(p.detach().abs().max()) == 0.0:
            p.copy_(torch.randn(p.shape, generator=g) * std)
