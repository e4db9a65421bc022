"""Developer timing aid: the roll-out driver at the 4AA size (T = 1000 frames per block, R = 4 residues) — chains advanced
together on the device (SIAtom14SamplingWrapper.sample_rollouts, B chains) against the reference's calling pattern (one chain per
call, model.sample(create_batch(...)) which encodes T copies of the frame).  Prints blocks (= trajectories of T frames) per second.
Usage: python scripts/gpu_rollout_bench.py [B] [num_rollouts]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lam_slide_b200 as P  # noqa: E402
from lam_slide_b200.synthetic import randomize_zero_init  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_roll = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = P.get_config("peptide")
    torch.manual_seed(0)
    m = P.SecondStageSampler(cfg, sampling_kwargs={"sampling_method": "euler", "num_steps": 10})
    randomize_zero_init(m, seed=1)
    m = m.cuda()
    w = P.SIAtom14SamplingWrapper(m, shift=0.0, scale=1.0)
    R = cfg["N"]
    g = torch.Generator().manual_seed(3)
    cond = torch.randn(B, R, 14, 3, generator=g)
    res = torch.randint(0, 20, (B, R), generator=g)
    msk = torch.ones(B, R, 14, dtype=torch.bool)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    dt, out = timed(lambda: w.sample_rollouts(cond, res, msk, num_rollouts=n_roll))
    print(f"batched on device : B={B} chains x {n_roll} roll-outs -> {tuple(out.shape)}  {dt * 1e3:8.1f} ms  {B * n_roll / dt:8.1f} blocks/s", flush=True)
    nb = min(B, 4)
    dt1, _ = timed(lambda: [w.sample_rollout(cond[b], res[b], msk[b], num_rollouts=n_roll).cpu() for b in range(nb)])
    print(f"chain by chain    : {nb} chains x {n_roll} roll-outs (B = 1 per sample(), result to the host per chain)  {dt1 * 1e3:8.1f} ms  "
          f"{nb * n_roll / dt1:8.1f} blocks/s", flush=True)

    def ref_style():
        outs = []
        for b in range(nb):
            pos = cond[b].cuda()
            for _ in range(n_roll):
                batch = w.create_batch(pos, res[b].cuda(), msk[b].cuda())
                pred = m.sample(batch)["atom14_pos"].squeeze(0)
                outs.append(pred.cpu())
                pos = pred[-1].clone()
        return outs

    dt2, _ = timed(ref_style)
    print(f"reference pattern : {nb} chains x {n_roll} roll-outs (create_batch with T copies + sample(), B = 1)  {dt2 * 1e3:8.1f} ms  "
          f"{nb * n_roll / dt2:8.1f} blocks/s", flush=True)


if __name__ == "__main__":
    main()
