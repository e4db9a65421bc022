#!/usr/bin/env python
"""bench.py — headline benchmark of the sampling hot path (BASELINE.json: "4AA-shaped trajectory samples/sec (flow ODE)").

    python bench.py --gpus N --steps K --warmup W            # native arm (this repo's CUDA path)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host CPU (oracle port)

One "step" = one full ``sample()`` pass (first-stage encode -> setup_conditioning -> SiT Euler ODE, num_steps=10 => 9
evaluations of the latent transformer -> first-stage decode [-> NCCL all-gather of the decoded coordinates when N > 1])
over one batch of synthetic 4AA-shaped trajectories (T=1000 frames, 4 residues, L=2 latents, D=96).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the meaning of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "trajectory_samples_per_sec"
UNIT = "trajectories/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="peptide", choices=["peptide", "md17", "nba", "pedestrian"])
    ap.add_argument("--batch-per-gpu", type=int, default=None)
    ap.add_argument("--num-steps", type=int, default=10, help="ODE grid points (num_steps-1 network evaluations)")
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the NBA / pedestrian / sweep lines and the GPU eager comparator")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the headline step from a CUDA graph")
    ap.add_argument("--ref-batch", type=int, default=4, help="trajectories per CPU step of the reference arm / cpu_baseline")
    return ap.parse_args()


DEFAULT_BATCH = {"peptide": 64, "md17": 256, "nba": 1024, "pedestrian": 1024}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(cfg, num_steps: int, B: int, reps: int, warmup: int):
    """The oracle port of the reference's algorithm (fp32 torch CPU ops) on the host cores — a REPORTED baseline."""
    from oracle import lamslide_oracle as O  # the one place bench.py may use oracle/: as the CPU arm
    torch.set_num_threads(os.cpu_count())
    O.USE_SDPA = True  # same attention library call as the reference (mmdit.py:51)
    fs_sd = O.init_first_stage_params(cfg["first_stage"], 1)
    bb_sd = O.init_backbone_params(cfg["backbone"], 2)
    batch = O.synthetic_batch(cfg, B, 3)
    L = cfg["first_stage"]["encoder"]["num_latents"]
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(B, cfg["T"], L, cfg["backbone"]["in_dim"], generator=g)
    y = torch.randn(B, 256, generator=g) if cfg["n_classes"] else None
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            O.sample(fs_sd, bb_sd, cfg, batch, noise, num_steps=num_steps, y=y)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def gpu_eager_baseline(cfg, num_steps: int, B: int, dev):
    """Same-box GPU comparator: the oracle port (plain torch ops = what the reference's eager PyTorch modules execute) on the GPU,
    (a) fp32 with TF32 matmuls (`matmul_precision: high`, src/train.py:48) and (b) under bf16 autocast (the reference's in-training
    validation mode), SDPA attention as in mmdit.py:51.  A REPORTED baseline next to the CUDA path — never on the product path."""
    from oracle import lamslide_oracle as O
    O.USE_SDPA = True
    fs_sd = {k: v.to(dev) for k, v in O.init_first_stage_params(cfg["first_stage"], 1).items()}
    bb_sd = {k: v.to(dev) for k, v in O.init_backbone_params(cfg["backbone"], 2).items()}
    batch = {k: v.to(dev) for k, v in O.synthetic_batch(cfg, B, 3).items()}
    L = cfg["first_stage"]["encoder"]["num_latents"]
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(B, cfg["T"], L, cfg["backbone"]["in_dim"], generator=g).to(dev)
    y = torch.randn(B, 256, generator=g).to(dev) if cfg["n_classes"] else None
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for name, ctx in (("fp32_tf32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            def run():
                with torch.no_grad():
                    if ctx is None:
                        return O.sample(fs_sd, bb_sd, cfg, batch, noise, num_steps=num_steps, y=y)
                    with ctx:
                        return O.sample(fs_sd, bb_sd, cfg, batch, noise, num_steps=num_steps, y=y)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 2
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out["what"] = "oracle port (eager torch ops, SDPA) on this GPU; same workload per trajectory as the headline line"
    return out


def secondary_lines(P, _lib, args, dev):
    """BASELINE.json configs[1], [2], [4] on ONE GPU: NBA B=1024, pedestrian B=1024 (ragged agent counts), and a batch x ODE-steps
    sweep of the 4AA shape.  Each entry: trajectories/s, ms per step, kernel launches per step; the small configurations also replayed
    from a CUDA graph (SecondStageSampler.use_cuda_graphs)."""
    from lam_slide_b200.synthetic import randomize_zero_init, synthetic_batch

    def build(name, num_steps):
        cfg = P.get_config(name)
        torch.manual_seed(0)
        m = P.SecondStageSampler(cfg, sampling_kwargs={"sampling_method": "euler", "num_steps": num_steps})
        randomize_zero_init(m, seed=1)
        return cfg, m.to(dev)

    def time_sample(m, cfg, B, steps, warm, graph, shares=False):
        batch = {k: v.to(dev) for k, v in synthetic_batch(cfg, B, seed=2000 + B).items()}
        L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
        noise = torch.randn(B, cfg["T"], L, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        m.use_cuda_graphs = graph
        for _ in range(warm):
            m.sample(dict(batch), noise=noise)
        torch.cuda.synchronize()
        _lib.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m.sample(dict(batch), noise=noise)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = _lib.launch_count() / steps
        m.use_cuda_graphs = False
        out = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B,
               "launches_per_step": launches if not graph else "1 graph replay"}
        if shares and not graph:  # live CUDA-event time per kernel class over one more step
            _lib.profile_begin()
            m.sample(dict(batch), noise=noise)
            prof = _lib.profile_end()
            tot = sum(v["ms"] for v in prof.values())
            out["kernel_time_shares"] = {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["ms"] > 0}
        return out

    from lam_slide_b200.configs import flops_per_trajectory
    out = {}
    for name in ("nba", "pedestrian"):
        cfg, m = build(name, 10)
        eager = time_sample(m, cfg, 1024, 10, 3, False, shares=True)
        graphed = time_sample(m, cfg, 1024, 10, 3, True)
        eager["tflops"] = flops_per_trajectory(cfg, 10) * eager["value"] / 1e12
        graphed["tflops"] = flops_per_trajectory(cfg, 10) * graphed["value"] / 1e12
        out[f"{name}_b1024"] = {"workload": f"{name} sample(): T={cfg['T']}, N={cfg['N']}, num_steps=10", "eager": eager, "cuda_graph": graphed}
        # a serving-sized batch: the ~400 launches of a step are launch-bound there, which is where replaying the step from a graph pays
        e32 = time_sample(m, cfg, 32, 20, 3, False)
        g32 = time_sample(m, cfg, 32, 20, 3, True)
        out[f"{name}_b32"] = {"workload": f"{name} sample(): T={cfg['T']}, N={cfg['N']}, num_steps=10", "eager": e32, "cuda_graph": g32,
                              "graph_speedup": g32["value"] / e32["value"]}
        del m
    # MD17 (BASELINE configs[0] is its CPU-runnable small batch): 192 latents per frame make it the one configuration whose SPATIAL axis
    # is long (whole-sequence mma.sync attention kernel) while the temporal axis is short
    cfg, m = build("md17", 10)
    md = time_sample(m, cfg, 64, 5, 2, False, shares=True)
    md["tflops"] = flops_per_trajectory(cfg, 10) * md["value"] / 1e12
    out["md17_b64"] = {"workload": f"md17 sample(): T={cfg['T']}, N={cfg['N']}, num_steps=10", "eager": md}
    del m
    sweep = []
    for ns in (5, 10, 20, 50):
        cfg, m = build("peptide", ns)
        for B in ((8, 64, 512) if ns == 10 else (8, 64)):
            r = time_sample(m, cfg, B, 2, 1, False)
            r["num_steps"] = ns
            r["tflops"] = flops_per_trajectory(cfg, ns) * r["value"] / 1e12
            sweep.append(r)
        del m
    out["peptide_sweep"] = sweep
    torch.cuda.empty_cache()
    return out


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU path (oracle port; the Python reference cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    times, cores = cpu_baseline(cfg, args.num_steps, B, args.steps, args.warmup)
    total = sum(times)
    value = B * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the native arm's workload (same trajectories: T, entities, latents, num_steps); a step here is a bounded sample of it
        "config": {"workload": f"{args.config} sample(): encode + Euler ODE (num_steps={args.num_steps} => {args.num_steps - 1} evals) + decode",
                   "batch_per_step": B, "T": cfg["T"], "entities": cfg["N"], "latents": cfg["first_stage"]["encoder"]["num_latents"],
                   "latent_dim": cfg["backbone"]["in_dim"], "parallelism": f"host threads x{cores}",
                   "weights": "random-init (seeded), zero-init layers re-drawn N(0,0.02)",
                   "sample": f"bounded CPU sample: {B} trajectory(ies) per step instead of {DEFAULT_BATCH.get(args.config, B)}"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} x sample() of {B} trajectory(ies), fp32 torch CPU ops, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    import lam_slide_b200 as P
    from lam_slide_b200 import _lib
    from lam_slide_b200.configs import flops_per_eval, flops_per_trajectory
    from lam_slide_b200.synthetic import randomize_zero_init, synthetic_batch

    cfg = P.get_config(args.config)
    if args.T:
        cfg["T"] = args.T
    if args.impl == "reference":
        return run_reference(args, cfg)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch_per_gpu or DEFAULT_BATCH[args.config]
    T, N = cfg["T"], cfg["N"]
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]

    # model: random-init weights of the named architecture (same seed on every rank = replicated weights)
    torch.manual_seed(0)
    model = P.SecondStageSampler(cfg, sampling_kwargs={"sampling_method": "euler", "num_steps": args.num_steps})
    randomize_zero_init(model, seed=1)
    model = model.to(dev)

    # synthetic inputs: ONE global batch of world * B trajectories, replicated on every rank (host: pinned; device: resident in HBM);
    # lam_slide_b200.dist.sample_sharded / sample_stream_sharded — the product's multi-GPU entry points — take this rank's rows,
    # sample them and all-gather the decoded coordinates (world = 1: the same calls, the gather is a no-op)
    from lam_slide_b200 import dist as D_
    GB = world * B
    host_batch = synthetic_batch(cfg, GB, seed=1004, pin=True)
    dev_batch = {k: v.to(dev) for k, v in host_batch.items()}
    noise = D_.local_noise(model, host_batch, seed=77)  # this rank's rows of the global noise (row i depends on (seed, i) only)
    main_key = cfg["main_output"]
    model.use_cuda_graphs = bool(args.cuda_graph)

    def step_device():
        return D_.sample_sharded(model, dev_batch, noise=noise)

    host_out = None
    keep = {}

    def run_e2e(steps):
        """`steps` calls of the public sampling API on HOST (pinned) batches: every step copies this rank's rows host->device and its
        result device->host (and all-gathers on the device when N > 1); the copies overlap the neighbouring steps' compute."""
        nonlocal host_out
        for host_out in D_.sample_stream_sharded(model, (host_batch for _ in range(steps)), keep=keep):  # noise drawn on device like the reference
            pass

    def timed(fn, steps, whole=False):
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn(steps)  # returns after the last result has reached the host (sample_stream synchronises on its copy)
        else:
            for _ in range(steps):
                fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist_on:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _lib.launch_count(reset=True)
    ms = timed(step_device, args.steps)
    launches = _lib.launch_count()
    clock_info = clocks.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # end-to-end through the public API with HOST (pinned) buffers
    run_e2e(2)
    ms_e2e = timed(run_e2e, args.steps, whole=True)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host_batch.values())  # all ranks together: each uploads its own rows
    d2h = host_out.numel() * host_out.element_size() * world

    # per-kernel-class timing (CUDA events on the launching stream) on extra steps after the timed region
    peaks, peaks_src = load_peaks()
    roofline, shares = None, None
    if not args.no_profile:
        _lib.profile_begin()
        pr_steps = 2
        model.use_cuda_graphs = False  # CUDA events cannot be recorded inside a graph replay
        for _ in range(pr_steps):
            D_.sample_sharded(model, dev_batch, noise=noise, gather=False)
        prof = _lib.profile_end()
        tot = sum(v["ms"] for v in prof.values())
        shares = {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["ms"] > 0}
        bb = cfg["backbone"]
        H, M, depth = bb["hidden_size"], int(bb["mlp_ratio"] * bb["hidden_size"]), bb["depth"]
        n_tok = B * T * L
        evals = args.num_steps - 1
        # algorithmic FLOPs per launch (SURVEY.md §8(d) terms), per kernel class.  The product path splits linear1: its q | k | v third
        # runs in the "gemm_linear1" launches, its MLP half inside the fused MLP kernel that the profiler books as "gemm_linear2"
        # (mlp_fused.cuh: u W1m^T -> GELU -> linear2 -> gated residual -> next block's LN + modulate).
        flops = {
            "gemm_linear1": 2.0 * n_tok * H * (3 * H),
            "gemm_linear2": 2.0 * n_tok * (H * M + (H + M) * H),
            "attn_temporal": 4.0 * bb["num_heads"] * (H // bb["num_heads"]) * B * L * T * T,
        }
        dom = max(flops, key=lambda k: prof[k]["ms"])
        n_launch = prof[dom]["launch_groups"]
        ms_per_launch = prof[dom]["ms"] / n_launch
        achieved = flops[dom] / (ms_per_launch * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same config only)
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.config == "peptide" and B == DEFAULT_BATCH["peptide"] and T == 1000:
            traffic = json.load(open(tpath)).get("kernels", {}).get(dom, {}).get("dram_bytes_per_launch")
        kernel_names = {"gemm_linear2": "gemm_linear2 (mlp_fused_kernel: MLP half of linear1 + GELU + linear2 + gated residual + next LN/modulate)",
                        "gemm_linear1": "gemm_linear1 (q | k | v third of linear1 + QK-RMSNorm + RoPE [+ spatial attention])"}
        roofline = {"kernel": kernel_names.get(dom, dom), "class": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)",
                    "peak_source": f"{peaks_src} bf16_tflops_sustained (kernel timed inside a long step)",
                    "ms_per_launch": ms_per_launch, "launches_per_step": n_launch / pr_steps,
                    "share_of_step": shares.get(dom)}
        ftraj = flops_per_trajectory(cfg, args.num_steps)
        whole = {"achieved_tflops_per_gpu": ftraj * value / world / 1e12,
                 "frac_of_burst_peak": ftraj * value / world / 1e12 / peaks["bf16_tflops"],
                 "frac_of_sustained_peak": ftraj * value / world / 1e12 / peaks["bf16_tflops_sustained"]}
    else:
        whole = None

    secondary, gpu_eager = None, None
    if rank == 0 and world == 1 and not args.no_secondary:
        dev_batch.clear()
        torch.cuda.empty_cache()
        secondary = secondary_lines(P, _lib, args, dev)
        gpu_eager = gpu_eager_baseline(cfg, args.num_steps, 4, dev)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        rb = args.ref_batch
        times, cores = cpu_baseline(cfg, args.num_steps, rb, reps=2, warmup=1)
        cpu = {"value": rb * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 x sample() of {rb} trajectories (T={T}, num_steps={args.num_steps}), oracle port of the reference, fp32 torch CPU ops"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.config} sample(): encode + Euler ODE (num_steps={args.num_steps} => {args.num_steps - 1} evals) + decode"
                                   + (" + all_gather" if dist_on else ""),
                       "batch_per_gpu": B, "global_batch": B * world, "T": T, "entities": N, "latents": L, "latent_dim": D,
                       "parallelism": f"batch-sharded x{world} (lam_slide_b200.dist.sample_sharded)", "cuda_graph": bool(args.cuda_graph),
                       "weights": "random-init (seeded), zero-init layers re-drawn N(0,0.02)",
                       "l2": "working set (activations ~GBs per step) is far larger than the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clock_info,
            "roofline": roofline,
            "whole_step": whole,
            "kernel_time_shares": shares,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager,
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
