"""Developer timing aid: K-sample evaluation — KSampleEvaluator.test_step (one batched solve of B * K
trajectories from once-encoded latents + the device metric kernel) against the reference's pattern (K sample() calls in a Python loop,
metric from the stacked results).  Usage: python scripts/gpu_ksample_bench.py [config] [B] [K]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lam_slide_b200 as P  # noqa: E402
from lam_slide_b200.synthetic import randomize_zero_init, synthetic_batch  # noqa: E402
from oracle import lamslide_oracle as O  # noqa: E402  (developer script: the oracle is the checker of the loop variant's metric)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "nba"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    cfg = P.get_config(name)
    torch.manual_seed(0)
    m = P.SecondStageSampler(cfg, sampling_kwargs={"sampling_method": "euler", "num_steps": 10})
    randomize_zero_init(m, seed=1)
    m = m.cuda()
    batch = {k: v.cuda() for k, v in synthetic_batch(cfg, B, seed=4).items()}
    mode = "mean" if name == "md17" else "min"
    ev = P.KSampleEvaluator(m, K=K, num_runs=K, mode=mode)
    c1 = cfg["cond_idx"][1]

    def batched():
        return ev.test_step(batch)

    def loop():
        b = {k: v.clone() for k, v in batch.items()}
        true_pos = b["pos"].clone()
        b["pos"][:, c1:] = 0
        preds = [m.sample(dict(b))["pos"] for _ in range(K)]
        if mode == "min":
            mask = b.get("attention_mask", torch.ones(true_pos.shape[:3], dtype=torch.bool, device=true_pos.device))
            return O.ksample_min_ade_fde(preds, true_pos, mask, c1, K)
        return O.ksample_mean_ade_fde(preds, true_pos, c1)

    for label, fn in [("batched evaluator", batched), ("K sample() calls ", loop)]:
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, f = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{name} B={B} K={K} [{label}]: {dt * 1e3:8.1f} ms  {B * K / dt:9.1f} sampled trajectories/s   ADE {float(a.mean()):.4f} FDE {float(f.mean()):.4f}",
              flush=True)


if __name__ == "__main__":
    main()
