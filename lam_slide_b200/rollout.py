"""Autoregressive roll-out driver — mirrors ``SIAtom14SamplingWrapper`` (``src/modules/sampling.py:16-63``): ``create_batch``
and ``sample_rollout`` with the reference's names, arguments and result layout, plus ``sample_rollouts``, the batched
on-device form the B200 path is built for.

The reference rolls ONE peptide at a time: every roll-out step builds a B = 1 batch whose T frames are copies of the current
frame, calls ``sample()`` (which encodes all T identical frames), moves the block to the host side of the loop and continues
from its last frame.  Here many chains advance together (``[B, R, 14, 3]`` conditioning frames), everything stays on the
device between steps, and the conditioning frame is encoded ONCE per chain and step — the first stage treats frames
independently and picks its kernels by layer shape only, so broadcasting its latents over T is identical to encoding T copies
(tests/test_gpu_parity.py checks that bit for bit against ``model.sample(create_batch(...))``).
``sample_traj_files`` is the file-producing half of ``sample_traj`` + ``eval_peptide.sample_trajectory`` (sampling.py:65-142,
eval_peptide.py:329-349) without mdtraj: conditioning frame and residue types in, ``<prefix>.dcd`` / ``<prefix>.pdb`` out
(``lam_slide_b200/formats.py``; XTC's compression codec is not implemented).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from . import formats
from .model import SecondStageSampler


class SIAtom14SamplingWrapper:
    def __init__(self, model: SecondStageSampler, shift: Optional[float] = None, scale: Optional[float] = None):
        self.model = model
        # the Lightning wrapper carries the dataset normalisation as ``model.shift`` / ``model.scale`` (sampling.py:52,63)
        self.shift = getattr(model, "shift", 0.0) if shift is None else shift
        self.scale = getattr(model, "scale", 1.0) if scale is None else scale

    # sampling.py:24-43 (reference-compatible: B = 1, the frame repeated over all T frames)
    def create_batch(self, pos: Tensor, res: Tensor, res_mask: Tensor) -> Dict[str, Tensor]:
        T = self.model.hparams.n_timesteps
        R = res.shape[0]
        pos = pos * res_mask[..., None].to(device=pos.device, dtype=pos.dtype)
        return {
            "atom14_pos": pos[None, None].expand(1, T, R, 14, 3).contiguous(),
            "aatype": res[None, None].expand(1, T, R).contiguous(),
            "attention_mask": torch.ones(1, T, R, dtype=torch.bool, device=res.device),
            "entities": torch.arange(R, device=res.device)[None, None].expand(1, T, R).contiguous(),
        }

    # sampling.py:45-63.  ``noise`` (optional, [num_rollouts, 1, T, L, D]) replaces the randn_like of the sample() calls.
    @torch.no_grad()
    def sample_rollout(self, cond_pos: Tensor, res: Tensor, res_mask: Tensor, num_rollouts: int = 1,
                       noise: Optional[Tensor] = None) -> Tensor:
        return self.sample_rollouts(cond_pos[None], res[None], res_mask[None], num_rollouts, noise=noise)[0]

    @torch.no_grad()
    def sample_rollouts(self, cond_pos: Tensor, res: Tensor, res_mask: Tensor, num_rollouts: int = 1,
                        noise: Optional[Tensor] = None) -> Tensor:
        """B chains at once: ``cond_pos [B, R, 14, 3]``, ``res [B, R]``, ``res_mask [B, R, 14]`` (bool) ->
        ``[B, num_rollouts * T, R, 14, 3]`` in the caller's units; chain b equals ``sample_rollout`` of its own inputs."""
        m = self.model
        dev = m.device
        T = m.hparams.n_timesteps
        cond_pos = cond_pos.to(dev, torch.float32)
        res = res.to(dev)
        mask = res_mask.to(dev)[..., None].to(torch.float32)
        B, R = res.shape
        cond = (cond_pos - self.shift) / self.scale
        pos = cond.clone()
        entities1 = torch.arange(R, device=dev)[None, None].expand(B, 1, R).contiguous()
        entities = entities1.expand(B, T, R).contiguous()
        blocks = []
        for i in range(num_rollouts):
            frame = {"atom14_pos": (pos * mask)[:, None], "aatype": res[:, None], "entities": entities1}
            lat1 = m.encode(frame)  # [B, 1, L, D]: one frame per chain instead of T copies
            latents = lat1.expand(B, T, *lat1.shape[2:]).contiguous()
            nz = None if noise is None else noise[i]
            pred = m.sample_from_latents(latents, entities, noise=nz)["atom14_pos"]  # [B, T, R, 14, 3]
            blocks.append(pred)
            pos = pred[:, -1].clone()
        positions = torch.cat(blocks, dim=1)
        positions[:, 0] = cond
        return positions * self.scale + self.shift

    # sampling.py:65-100 + eval_peptide.py:340-349, from tensors instead of an mdtraj trajectory
    @torch.no_grad()
    def sample_traj_files(self, cond_pos: Tensor, res: Tensor, prefix: str, num_rollouts: int = 1, noise: Optional[Tensor] = None):
        """``cond_pos [R, 14, 3]`` (nm, heavy atoms in atom14 slots, centred as ``sample_traj`` does), ``res [R]`` residue types ->
        roll-out -> padding atom14 slots zeroed (sampling.py:96) -> atom37 heavy-atom trajectory -> ``<prefix>.dcd`` (all frames) and
        ``<prefix>.pdb`` (first frame).  Returns ``(positions [frames, R, 14, 3] on the host, dcd path, pdb path)``."""
        res = torch.as_tensor(res, dtype=torch.long)
        res_mask = torch.as_tensor(formats.RESTYPE_ATOM14_MASK)[res].to(torch.bool)
        pos = self.sample_rollout(cond_pos.to(torch.float32), res, res_mask, num_rollouts, noise=noise).detach().cpu()
        pos = pos * res_mask[None, ..., None]
        dcd, pdb = formats.save_trajectory(prefix, pos, res.tolist())
        return pos, dcd, pdb
