"""Developer diagnostic (not a test): runs one stage of the CUDA path and PRINTS the error metrics instead of
asserting, so a single gpurun call tells which kernel is off and by how much.  Usage: python scripts/gpu_diag.py <stage>"""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lam_slide_b200 import _lib as L  # noqa: E402
from oracle import lamslide_oracle as O  # noqa: E402
from tests.helpers import CASE_BY_NAME, case_inputs, frame_slice, load_golden, max_rel, rmsd  # noqa: E402


def gemm(shapes):
    lib = L.load()
    for (M, N, K, bn) in shapes:
        g = torch.Generator().manual_seed(M + N)
        a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        b = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        c = torch.full((M, N), float("nan"), device="cuda")
        st = lib.lamslide_debug_gemm(a.data_ptr(), b.data_ptr(), 0, c.data_ptr(), M, N, K, bn, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = a.double() @ b.double().t()
        err = max_rel(c, ref) if torch.isfinite(c).all() else float("nan")
        print(f"gemm M={M} N={N} K={K} bn={bn}: status={st} max_rel={err:.3e} nan={int((~torch.isfinite(c)).sum())}", flush=True)
        if not (err < 1e-4):
            d = (c.double() - ref).abs()
            bad_rows = (d.amax(dim=1) > 1e-3 * ref.abs().max()).nonzero().flatten()[:16].tolist()
            bad_cols = (d.amax(dim=0) > 1e-3 * ref.abs().max()).nonzero().flatten()[:16].tolist()
            print("   bad rows", bad_rows, "bad cols", bad_cols)
            print("   c[0,:8]", c[0, :8].tolist(), "\n   r[0,:8]", ref[0, :8].tolist())


def attn():
    from tests.test_gpu_kernels import _attention_reference
    lib = L.load()
    for (B, T, Lx, H, heads, temporal, flash) in [(2, 1000, 2, 384, 16, 1, 0), (2, 20, 8, 256, 16, 1, 1), (3, 20, 2, 128, 4, 1, 1),
                                                 (2, 7, 192, 256, 16, 0, 0), (4, 30, 2, 384, 16, 0, 0), (2, 20, 8, 256, 16, 0, 0)]:
        n = B * T * Lx
        g = torch.Generator().manual_seed(n)
        qkv = torch.randn(n, 3 * H, generator=g).to(torch.bfloat16).cuda()
        out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
        st = lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, temporal, flash,
                                          torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = _attention_reference(qkv, B, T, Lx, H, heads, bool(temporal))
        print(f"attn B={B} T={T} L={Lx} H={H} heads={heads} temporal={temporal} flash={flash}: status={st} "
              f"max_rel={max_rel(out.float(), ref):.3e} mean_rel={float((out.float() - ref).abs().mean() / ref.abs().mean()):.3e}", flush=True)


def _build(cfg, fs_sd, bb_sd):
    import lam_slide_b200 as P
    m = P.SecondStageSampler(cfg).cuda()
    m.first_stage_model.backbone.load_state_dict(fs_sd, strict=True)
    m.backbone.load_state_dict(bb_sd, strict=True)
    return m


def first_stage():
    from lam_slide_b200.configs import get_config
    for name in ["pedestrian", "nba", "peptide", "md17"]:
        cfg = get_config(name, depth=1)
        fs_sd = O.init_first_stage_params(cfg["first_stage"], 21)
        bb_sd = O.init_backbone_params(cfg["backbone"], 22)
        m = _build(cfg, fs_sd, bb_sd)
        batch = O.synthetic_batch(cfg, 3, 23, T=5)
        flat = {k: v.flatten(0, 1) for k, v in batch.items() if k != "cond_scene"}
        with torch.no_grad():
            lat_ref = O.first_stage_encode(fs_sd, cfg["first_stage"], flat)
            out_ref = O.first_stage_decode(fs_sd, cfg["first_stage"], lat_ref, flat["entities"])
        fs = m.first_stage_model.backbone
        lat = fs.encode({k: v.cuda() for k, v in flat.items()})
        out = fs.decode(lat_ref.cuda(), flat["entities"].cuda())
        torch.cuda.synchronize()
        print(f"first_stage {name}: encode max_rel={max_rel(lat.cpu(), lat_ref):.3e} " +
              " ".join(f"{k}={max_rel(out[k].cpu(), v):.3e}" for k, v in out_ref.items()), flush=True)


def backbone():
    import lam_slide_b200 as P
    from lam_slide_b200.configs import get_config
    for name, depth, B, T in [("pedestrian", 1, 2, 20), ("nba", 1, 2, 20), ("peptide", 1, 1, 64), ("md17", 1, 1, 6), ("peptide", 7, 1, 1000)]:
        cfg = get_config(name, depth=depth)
        bb = cfg["backbone"]
        bb_sd = O.init_backbone_params(bb, 31)
        net = P.LatentSIV3(depth=depth, in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                           vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).cuda()
        net.load_state_dict(bb_sd, strict=True)
        Lx = cfg["first_stage"]["encoder"]["num_latents"]
        g = torch.Generator().manual_seed(32)
        x = torch.randn(B, T, Lx, bb["in_dim"], generator=g)
        xc = torch.randn(B, T, Lx, bb["in_dim"], generator=g)
        mk = (torch.rand(B, T, Lx, generator=g) < 0.3).long()
        t = torch.rand(B, generator=g)
        y = torch.randn(B, bb["vec_in_dim"], generator=g) if bb["vec_in_dim"] else None
        trace = {}
        with torch.no_grad():
            ref = O.backbone_forward(bb_sd, bb, x, t, xc, mk, y, trace=trace)
        t0 = time.time()
        out = net(x.cuda(), t.cuda(), xc.cuda(), mk.cuda(), None if y is None else y.cuda())
        torch.cuda.synchronize()
        print(f"backbone {name} depth={depth} B={B} T={T}: max_rel={max_rel(out.cpu(), ref):.3e} finite={bool(torch.isfinite(out).all())} "
              f"({time.time() - t0:.2f}s first call)", flush=True)


def sample():
    for name in ["pedestrian_full", "nba_full", "peptide_small", "md17_small", "peptide_linear_velocity", "md17_full", "peptide_full"]:
        fx = load_golden(name)
        c = CASE_BY_NAME[name]
        cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
        m = _build(cfg, fs_sd, bb_sd)
        sl = frame_slice(fx)
        B, T = c["B"], c["T"]
        cb = {k: v.cuda() for k, v in batch.items()}
        latents = m.encode(cb)
        x_cond, x_mask = m.setup_conditioning(latents)
        yy = None if y is None else y.cuda()
        t0, _ = O.sample_interval(cfg["path_type"], cfg["prediction"])
        out0 = m.backbone(noise.cuda(), torch.full((B,), t0).cuda(), x_cond, x_mask, yy)
        states, vel = m.backbone.ode_sample(noise.cuda(), x_cond, x_mask, yy, path_type=cfg["path_type"], prediction=cfg["prediction"],
                                            num_steps=c["num_steps"], return_velocities=True)
        out = m.first_stage_model.decode(states[-1].flatten(0, 1), cb["entities"].flatten(0, 1))
        torch.cuda.synchronize()
        main = cfg["main_output"]
        got = out[main].unflatten(0, (B, T)).cpu()[:, sl]
        v = vel.cpu()[fx["velocity_steps"]][:, :, sl]
        verr = [max_rel(v[i], fx["velocities"][i]) for i in range(v.shape[0])]
        print(f"sample {name}: latents={max_rel(latents.cpu()[:, sl], fx['latents']):.2e} net_t0={max_rel(out0.cpu()[:, sl], fx['net_out_t0']):.2e} "
              f"vel_max={max(verr):.2e} final_lat={max_rel(states[-1].cpu()[:, sl], fx['final_latents']):.2e} "
              f"rmsd_final_frame={rmsd(got[:, -1], fx['outputs'][main][:, -1]):.2e} rmsd_all={rmsd(got, fx['outputs'][main]):.2e}", flush=True)


if __name__ == "__main__":
    stage = sys.argv[1]
    torch.cuda.init()
    print(f"== stage {stage} on {torch.cuda.get_device_name(0)}", flush=True)
    if stage == "gemm1":
        gemm([(128, 32, 64, 32)])
    elif stage == "gemm":
        gemm([(128, 128, 128, 128), (300, 192, 384, 192), (1000, 96, 384, 96), (256, 256, 1920, 256), (4000, 384, 1920, 192),
              (77, 64, 256, 64), (515, 2688, 384, 192), (129, 48, 128, 48), (100, 16, 64, 16)])
    elif stage == "attn":
        attn()
    elif stage == "first_stage":
        first_stage()
    elif stage == "backbone":
        backbone()
    elif stage == "sample":
        sample()
    print(f"== stage {stage} done", flush=True)
