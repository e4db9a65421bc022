"""Multi-GPU sampling: independent trajectories are batch-sharded over the ranks (one process per GPU, weights
replicated), and the decoded coordinates are gathered with ONE collective at the end.  There is no
data-path collective: no sample ever needs another sample's data.  The initial noise of global sample ``i`` comes from a
generator seeded with ``(seed, i)``, so a sharded run reproduces the single-GPU run sample for sample."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal chunks: the first ``global_batch % world`` ranks get one extra sample."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, Tensor], rank: int, world: int) -> Dict[str, Tensor]:
    B = batch["entities"].shape[0]
    lo, hi = shard_range(B, rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in batch.items()}


def per_sample_noise(seed: int, lo: int, hi: int, shape: Tuple[int, ...], device: torch.device) -> Tensor:
    """Noise [hi-lo, *shape]; sample i depends on (seed, global index i) only."""
    out = torch.empty((hi - lo,) + tuple(shape), device=device, dtype=torch.float32)
    for j, i in enumerate(range(lo, hi)):
        g = torch.Generator(device=device).manual_seed((seed * 1_000_003 + i) % (2 ** 63 - 1))
        out[j] = torch.randn(shape, device=device, generator=g)
    return out


def gather_samples(local: Tensor, global_batch: int, group=None) -> Tensor:
    """All-gather the per-rank outputs [b_r, ...] into [global_batch, ...] (ragged shards are padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    buf = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return buf
    buf = buf.view((world, mx) + tuple(local.shape[1:]))
    return torch.cat([buf[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)])


def local_noise(model, batch: Dict[str, Tensor], seed: int = 0, group=None) -> Tensor:
    """The initial noise of this rank's shard of ``batch``: rows ``shard_range(B, rank, world)`` of the global noise, whose row ``i``
    depends on ``(seed, i)`` only."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, T = batch["entities"].shape[:2]
    lo, hi = shard_range(B, rank, world)
    cfg = model.cfg
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    return per_sample_noise(seed, lo, hi, (T, L, D), model.device)


@torch.no_grad()
def sample_sharded(model, batch: Dict[str, Tensor], seed: int = 0, gather: bool = True, group=None,
                   noise: Optional[Tensor] = None) -> Tensor:
    """``model.sample`` on this rank's shard of a replicated host/device batch, then one all-gather of the main output.
    ``noise`` (optional): this rank's rows of the initial noise (``local_noise``), for callers that draw it once."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = batch["entities"].shape[0]
    local = shard_batch(batch, rank, world)
    if noise is None:
        noise = local_noise(model, batch, seed, group)
    out = model.sample(dict(local), noise=noise)[model.cfg["main_output"]]
    return gather_samples(out, B, group) if gather else out


def sample_stream_sharded(model, host_batches, seed: int = 0, group=None, noise: Optional[Tensor] = None, keep: Optional[dict] = None):
    """``SecondStageSampler.sample_stream`` over replicated HOST batches, sharded like ``sample_sharded``: every rank uploads only its
    rows of each batch, samples them, all-gathers the decoded coordinates on the device (``keep["gathered"]`` holds the latest
    ``[B, ...]`` device tensor when a dict is given) and yields its OWN rows as a pinned host tensor — one process per GPU, each
    feeding its consumer, with the copies of neighbouring batches overlapped with compute."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sizes = []

    def shards():
        for b in host_batches:
            sizes.append(b["entities"].shape[0])
            yield shard_batch(b, rank, world)

    def on_device(out: Tensor) -> None:
        g = gather_samples(out, sizes.pop(0), group)
        if keep is not None:
            keep["gathered"] = g

    yield from model.sample_stream(shards(), noise=noise, on_device=on_device)


@torch.no_grad()
def rollouts_sharded(wrapper, cond_pos: Tensor, res: Tensor, res_mask: Tensor, num_rollouts: int = 1, seed: int = 0,
                     gather: bool = True, group=None) -> Tensor:
    """``SIAtom14SamplingWrapper.sample_rollouts`` with the B chains sharded over the ranks (chains are independent: no data-path
    collective), one all-gather of the ``[b_r, num_rollouts * T, R, 14, 3]`` blocks at the end.  The noise of roll-out step ``i`` of
    global chain ``c`` is seeded by ``(seed, i, c)``, so the result does not depend on the number of ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = cond_pos.shape[0]
    lo, hi = shard_range(B, rank, world)
    m = wrapper.model
    cfg = m.cfg
    T = m.hparams.n_timesteps
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    noise = torch.stack([per_sample_noise(seed * 7919 + i + 1, lo, hi, (T, L, D), m.device) for i in range(num_rollouts)])
    out = wrapper.sample_rollouts(cond_pos[lo:hi], res[lo:hi], res_mask[lo:hi], num_rollouts=num_rollouts, noise=noise)
    return gather_samples(out, B, group) if gather else out
