"""CPU, world_size 2, gloo: the N>1 path's host logic — shard ranges, per-sample noise reproducibility across world
sizes, and the final all-gather (ragged shards included).  The CUDA sampling itself is covered by the -m gpu tests."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lam_slide_b200.dist import gather_samples, per_sample_noise, shard_batch, shard_range


def test_shard_ranges_cover_batch():
    for B in (1, 5, 8, 64, 67):
        for w in (1, 2, 3, 8):
            r = [shard_range(B, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_per_sample_noise_independent_of_sharding():
    full = per_sample_noise(7, 0, 5, (3, 2, 4), torch.device("cpu"))
    a = per_sample_noise(7, 0, 3, (3, 2, 4), torch.device("cpu"))
    b = per_sample_noise(7, 3, 5, (3, 2, 4), torch.device("cpu"))
    assert torch.equal(full, torch.cat([a, b]))


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = {"entities": torch.arange(B * 3 * 2).reshape(B, 3, 2), "pos": torch.arange(B * 3 * 2 * 2, dtype=torch.float32).reshape(B, 3, 2, 2),
             "cond_scene": torch.arange(B)}
    local = shard_batch(batch, rank, world)
    lo, hi = shard_range(B, rank, world)
    assert local["pos"].shape[0] == hi - lo and local["cond_scene"].shape[0] == hi - lo
    # stand-in for the per-rank sample(): a deterministic function of (sample data, per-sample noise)
    noise = per_sample_noise(3, lo, hi, (3, 2, 2), torch.device("cpu"))
    out_local = local["pos"] * 2 + noise
    out = gather_samples(out_local, B)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_world2_gather_equals_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pos = torch.arange(B * 3 * 2 * 2, dtype=torch.float32).reshape(B, 3, 2, 2)
    expect = pos * 2 + per_sample_noise(3, 0, B, (3, 2, 2), torch.device("cpu"))
    assert torch.equal(out, expect)


class _StubWrapper:
    """Host-side stand-in for SIAtom14SamplingWrapper on CPU: a deterministic function of (chain inputs, per-chain noise)."""

    class _M:
        cfg = {"first_stage": {"encoder": {"num_latents": 2}}, "backbone": {"in_dim": 4}}
        hparams = type("H", (), {"n_timesteps": 3})()
        device = torch.device("cpu")

    model = _M()

    def sample_rollouts(self, cond_pos, res, res_mask, num_rollouts=1, noise=None):
        B, R = res.shape
        T = self.model.hparams.n_timesteps
        blocks = [cond_pos[:, None].expand(B, T, R, 14, 3) * (i + 1) + noise[i].sum(dim=(1, 2, 3))[:, None, None, None, None] for i in range(num_rollouts)]
        return torch.cat(blocks, dim=1)


def _rollout_worker(rank, world, port, B, q):
    from lam_slide_b200.dist import rollouts_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(B, 4, 14, 3, generator=g)
    res = torch.randint(0, 20, (B, 4), generator=g)
    msk = torch.ones(B, 4, 14, dtype=torch.bool)
    out = rollouts_sharded(_StubWrapper(), cond, res, msk, num_rollouts=2, seed=5)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_world2_rollouts_equal_single_process(B):
    """rollouts_sharded on 2 gloo ranks == the unsharded call: chains are independent and the noise is keyed by the global chain index."""
    from lam_slide_b200.dist import rollouts_sharded
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_rollout_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(B, 4, 14, 3, generator=g)
    res = torch.randint(0, 20, (B, 4), generator=g)
    msk = torch.ones(B, 4, 14, dtype=torch.bool)
    expect = rollouts_sharded(_StubWrapper(), cond, res, msk, num_rollouts=2, seed=5)
    assert out.shape == (B, 6, 4, 14, 3) and torch.equal(out, expect)

