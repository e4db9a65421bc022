"""GPU, >= 2 devices (skipped on a 1-GPU box): the PRODUCT multi-GPU entry point — ``lam_slide_b200.dist.sample_sharded`` over NCCL, one
process per GPU — returns on every rank exactly what the single-GPU ``sample()`` returns for the same global batch and the same
``per_sample_noise`` (trajectories are independent, the only collective is the final all-gather).
Also: handles on two devices in ONE process (``include/lamslide.h``: a handle belongs to the device current at create time)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _model_and_batch(name, B, depth=None):
    import lam_slide_b200 as P
    from lam_slide_b200.synthetic import randomize_zero_init, synthetic_batch
    cfg = P.get_config(name, depth=depth) if depth else P.get_config(name)
    torch.manual_seed(0)
    m = P.SecondStageSampler(cfg)
    randomize_zero_init(m, seed=1)
    return m, cfg, synthetic_batch(cfg, B, seed=11)


def _worker(rank, world, port, name, B, q):
    from lam_slide_b200.dist import sample_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m, cfg, batch = _model_and_batch(name, B, depth=2)
    m = m.cuda()
    out = sample_sharded(m, batch, seed=5)
    assert out.shape[0] == B
    # every rank holds the whole result after the all-gather: rank 1's copy must equal rank 0's
    ref = out.clone()
    dist.broadcast(ref, src=0)
    same = torch.equal(ref, out)
    if rank == 0:
        q.put(out.cpu())
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag.item()) == 1
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,B", [("nba", 6), ("pedestrian", 5), ("peptide", 3)])
def test_sample_sharded_nccl_equals_single_gpu(name, B):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from lam_slide_b200.dist import per_sample_noise
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    m, cfg, batch = _model_and_batch(name, B, depth=2)
    m = m.cuda()
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    noise = per_sample_noise(5, 0, B, (cfg["T"], L, D), torch.device("cuda", 0))
    want = m.sample({k: v.clone() for k, v in batch.items()}, noise=noise)[cfg["main_output"]].cpu()
    assert torch.equal(got, want)  # ragged shards (B = 5, 3) included


def test_two_devices_in_one_process():
    """Per-device kernel attributes (opt-in shared memory sizes) and per-device handles: the same model run on cuda:0 and on cuda:1
    from one process gives the same result (function attributes are per device, so a process-wide "configured" flag would make the
    first large-shared-memory launch on the second device fail)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    m, cfg, batch = _model_and_batch("peptide", 2, depth=2)
    cfg_T = 256
    batch = {k: v[:, :cfg_T].contiguous() for k, v in batch.items()}
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    noise = torch.randn(2, cfg_T, L, D, generator=torch.Generator().manual_seed(3))
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        md = m.to(dev)
        with torch.cuda.device(dev):
            outs.append(md.sample({k: v.clone() for k, v in batch.items()}, noise=noise.clone())[cfg["main_output"]].cpu())
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
