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
