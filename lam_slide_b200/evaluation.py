"""K-sample evaluation loops on the device — mirrors ``Wrapper.test_step`` of ``second_stage/nba.py:161-238`` /
``pedestrian.py:149-226`` (min-ADE / min-FDE over the first ``num_runs`` of K samples, per real agent; unclustered metric) and of
``second_stage/md17.py:139-171`` (mean over K runs, per sample) — SURVEY.md §8(f) rank 2.

The reference calls ``sample(batch)`` K times in a Python loop (K = 60 / 20 / 5) and reduces on the host side of the loop.  Here
the K runs are ONE batched ODE solve: the batch is encoded once (its first-stage latents do not depend on the run), the latents
are repeated K times, K independent noise tensors are drawn, and the errors are reduced by one kernel
(``lamslide_ksample_errors``).  The FPC post-processing (``post_process=True``: torch_kmeans clustering of the final
positions, nba.py:227-238) is out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .model import SecondStageSampler


def ksample_errors(preds: Tensor, target: Tensor, num_runs: int, mode: str) -> Tuple[Tensor, Tensor]:
    """``preds [K, B, T, A, D]``, ``target [B, T, A, D]`` (fp32, CUDA, frames after the conditioning window only) ->
    ``(ades, fdes)``: ``[B * A]`` each for ``mode="min"`` (per agent, min over the first ``num_runs`` samples), ``[B]`` for
    ``mode="mean"`` (per sample, mean over all K)."""
    _lib.require_cuda(preds)
    preds = preds.to(torch.float32).contiguous()
    target = target.to(torch.float32).contiguous()
    K, B, T, A, D = preds.shape
    if T == 0:
        raise ValueError("no frames after the conditioning window")
    if tuple(target.shape) != (B, T, A, D):
        raise ValueError(f"target shape {tuple(target.shape)} does not match preds {tuple(preds.shape)}")
    n_out = B * A if mode == "min" else B
    ades = torch.empty(n_out, dtype=torch.float32, device=preds.device)
    fdes = torch.empty_like(ades)
    with torch.cuda.device(preds.device):
        _lib.check(_lib.load().lamslide_ksample_errors(preds.data_ptr(), target.data_ptr(), ades.data_ptr(), fdes.data_ptr(), K,
                                                       num_runs, B, T, A, D, 0 if mode == "min" else 1, _lib.current_stream_ptr()))
    return ades, fdes


class KSampleEvaluator:
    """``KSampleEvaluator(model, K, num_runs, mode)``; ``test_step(batch) -> (ades, fdes)`` in normalised units (the reference
    multiplies by ``self.scale`` in ``on_test_epoch_end``)."""

    def __init__(self, model: SecondStageSampler, K: int, num_runs: Optional[int] = None, mode: str = "min",
                 max_trajectories: int = 4096):
        if mode not in ("min", "mean"):
            raise ValueError("mode must be 'min' (nba / pedestrian) or 'mean' (md17)")
        self.model, self.K, self.mode = model, K, mode
        self.num_runs = K if num_runs is None else num_runs
        self.max_trajectories = max_trajectories  # runs are batched in chunks of at most this many trajectories per ODE solve

    @torch.no_grad()
    def sample_k(self, batch: Dict[str, Tensor], noise: Optional[Tensor] = None) -> Tensor:
        """K ``sample()`` results for the batch as ``[K, B, T, A, D]``; ``noise [K, B, T, L, D]`` optional."""
        m = self.model
        dev = m.device
        b = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
        B = b["entities"].shape[0]
        latents = m.encode(b)  # identical for all K runs: encoded once
        y = m.vec_in_embedding(b["cond_scene"]) if hasattr(m, "vec_in_embedding") and "cond_scene" in b else None
        per = max(1, self.max_trajectories // B)
        key = m.cfg["main_output"]
        outs = []
        for k0 in range(0, self.K, per):
            kk = min(per, self.K - k0)
            nz = None if noise is None else noise[k0:k0 + kk].flatten(0, 1)
            out = m.sample_from_latents(latents.repeat(kk, 1, 1, 1), b["entities"].repeat(kk, 1, 1), noise=nz,
                                        y=None if y is None else y.repeat(kk, 1))[key]
            outs.append(out.unflatten(0, (kk, B)))
        return torch.cat(outs)

    @torch.no_grad()
    def test_step(self, batch: Dict[str, Tensor], noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        m = self.model
        dev = m.device
        c1 = m.hparams.cond_idx[1]
        batch = {k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
        true_pos = batch["pos"][:, c1:].clone()
        # the targets never reach the model (nba.py:188-189; md17.py:149-151 also blanks the atom types of those frames)
        batch["pos"][:, c1:] = 0
        if self.mode == "mean" and "atom" in batch:
            batch["atom"][:, c1:] = 0
        preds = self.sample_k(batch, noise)[:, :, c1:]
        ades, fdes = ksample_errors(preds, true_pos, self.num_runs, self.mode)
        if self.mode == "min" and "attention_mask" in batch:
            mask = batch["attention_mask"][:, -1].reshape(-1)  # any frame: the mask is constant along a trajectory (nba.py:212-213)
            ades, fdes = ades[mask], fdes[mask]
        return ades, fdes
