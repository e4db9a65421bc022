"""TEST INFRASTRUCTURE — generates ``tests/golden/*.pt`` by running the REAL reference modules.

Run in the dev container only (needs ``/root/reference``):

    python -m oracle.make_golden

For every case the weights (``init_*_params``), the synthetic batch and the initial noise are regenerated from
seeds by ``oracle/lamslide_oracle.py`` (seeded CPU generators), loaded into the reference's own ``nn.Module``s
with ``load_state_dict(strict=True)`` (which also pins the state-dict key names/shapes of SURVEY.md §8(b)), and
the reference's ``sample()`` path (``oracle/ref_loader.reference_sample``: reference Encoder / Decoder /
LatentSIV3 / Transport / Sampler + the torchdiffeq Euler shim) produces the expected tensors.  Only seeds,
checksums and expected outputs are stored, so the fixtures stay small; the checksums make a silent RNG
difference on another box a loud failure instead of a parity mismatch.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lam_slide_b200.configs import get_config  # noqa: E402
from oracle import lamslide_oracle as O  # noqa: E402
from oracle.ref_loader import (RefFirstStage, RefRolloutModel, RefTestStepSelf, load_reference, load_reference_method,  # noqa: E402
                               load_reference_rollout_wrapper, reference_sample)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name, config, backbone overrides, B, T, num_steps, seeds(fs, bb, batch, noise)
CASES = [
    dict(case="peptide_small", cfg="peptide", overrides=dict(depth=2), B=2, T=24, num_steps=10, seeds=(101, 102, 103, 104)),
    dict(case="peptide_full", cfg="peptide", overrides={}, B=1, T=1000, num_steps=10, seeds=(111, 112, 113, 114)),
    dict(case="md17_small", cfg="md17", overrides=dict(depth=2), B=1, T=8, num_steps=5, seeds=(201, 202, 203, 204)),
    dict(case="md17_full", cfg="md17", overrides={}, B=1, T=30, num_steps=10, seeds=(211, 212, 213, 214)),
    dict(case="nba_full", cfg="nba", overrides={}, B=3, T=20, num_steps=10, seeds=(301, 302, 303, 304)),
    dict(case="pedestrian_full", cfg="pedestrian", overrides={}, B=6, T=20, num_steps=10, seeds=(401, 402, 403, 404)),
    dict(case="peptide_linear_velocity", cfg="peptide", overrides=dict(depth=1), B=1, T=16, num_steps=6,
         seeds=(121, 122, 123, 124), path_type="Linear", prediction="velocity"),
    # BASELINE.json configs[4] (batch x ODE-steps sweep): the GVP / data drift multiplies the network output by (pi/2) / cos(pi t / 2),
    # ~19x at t = 0.947 (20 steps) and ~30x at t = 0.979 (50 steps) — the precision risk SURVEY.md §7 names for bf16 operands
    dict(case="peptide_steps20", cfg="peptide", overrides={}, B=1, T=1000, num_steps=20, seeds=(141, 142, 143, 144),
         vel_steps=(0, 1, 9, 17, 18)),
    dict(case="peptide_steps50", cfg="peptide", overrides={}, B=1, T=1000, num_steps=50, seeds=(151, 152, 153, 154),
         vel_steps=(0, 1, 24, 40, 46, 47, 48)),
]


# autoregressive roll-out: the reference's own SIAtom14SamplingWrapper.sample_rollout
ROLLOUT_CASE = dict(case="peptide_rollout", cfg="peptide", overrides=dict(depth=2), T=16, R=4, num_rollouts=3, num_steps=6,
                    shift=0.05, scale=2.0, seeds=(131, 132, 133, 134))


def rollout_case_inputs(c: dict = ROLLOUT_CASE):
    cfg = get_config(c["cfg"], **c["overrides"])
    cfg["T"] = c["T"]
    s_fs, s_bb, s_in, s_noise = c["seeds"]
    fs_sd = O.init_first_stage_params(cfg["first_stage"], s_fs)
    bb_sd = O.init_backbone_params(cfg["backbone"], s_bb)
    cond_pos, res, res_mask = O.rollout_inputs(c["R"], s_in)
    g = torch.Generator().manual_seed(s_noise)
    L = cfg["first_stage"]["encoder"]["num_latents"]
    noises = [torch.randn(1, c["T"], L, cfg["backbone"]["in_dim"], generator=g) for _ in range(c["num_rollouts"])]
    return cfg, fs_sd, bb_sd, cond_pos, res, res_mask, noises


def make_rollout_golden(ref) -> None:
    c = ROLLOUT_CASE
    cfg, fs_sd, bb_sd, cond_pos, res, res_mask, noises = rollout_case_inputs(c)
    bb = cfg["backbone"]
    fs = RefFirstStage(cfg["first_stage"]).eval()
    fs.load_state_dict(fs_sd, strict=True)
    net = ref.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                         vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"], theta=bb["theta"]).eval()
    net.load_state_dict(bb_sd, strict=True)
    model = RefRolloutModel(fs, net, cfg, noises, c["shift"], c["scale"], c["num_steps"])
    wrapper = load_reference_rollout_wrapper()(model)
    with torch.no_grad():
        positions = wrapper.sample_rollout(cond_pos.clone(), res.clone(), res_mask.clone(), num_rollouts=c["num_rollouts"])
    assert model.calls == c["num_rollouts"] and positions.shape == (c["num_rollouts"] * c["T"], c["R"], 14, 3)
    fixture = dict(case=dict(c), checksums=dict(fs=O.state_checksum(fs_sd), bb=O.state_checksum(bb_sd),
                                                 inputs=float(cond_pos.double().sum() + res.double().sum() + res_mask.double().sum()),
                                                 noise=float(sum(n.double().sum() for n in noises))),
                   positions=positions.clone(), torch_version=torch.__version__)
    path = os.path.join(GOLDEN_DIR, c["case"] + ".pt")
    torch.save(fixture, path)
    print(f"{c['case']:28s} -> {os.path.getsize(path) / 1024:8.1f} KiB   |pos| max {float(positions.abs().max()):.3f}")


# K-sample evaluation metrics: the reference's own Wrapper.test_step bodies on preset predictions
KSAMPLE_CASES = [
    dict(case="ksample_nba", file="src/models/composites/second_stage/nba.py", B=3, T=20, A=11, D=2, K=6, num_runs=4, cond_idx=(0, 8),
         pad=False, seed=141, mode="min"),
    dict(case="ksample_pedestrian", file="src/models/composites/second_stage/pedestrian.py", B=5, T=20, A=10, D=2, K=5, num_runs=5,
         cond_idx=(0, 8), pad=True, seed=142, mode="min"),
    dict(case="ksample_md17", file="src/models/composites/second_stage/md17.py", B=2, T=12, A=21, D=3, K=5, num_runs=5, cond_idx=(0, 4),
         pad=False, seed=143, mode="mean"),
]


def make_ksample_golden() -> None:
    fixtures = {}
    for c in KSAMPLE_CASES:
        preds, true_pos, mask = O.ksample_inputs(c["B"], c["T"], c["A"], c["D"], c["K"], c["seed"], c["pad"])
        test_step = load_reference_method(c["file"], "Wrapper", "test_step", {"KMeans": None})
        me = RefTestStepSelf(preds, c["K"], c["cond_idx"], c["num_runs"])
        batch = {"pos": true_pos.clone(), "attention_mask": mask.clone()}
        if c["mode"] == "mean":
            batch["atom"] = torch.ones(c["B"], c["T"], c["A"], dtype=torch.int64)
        with torch.no_grad():
            test_step(me, batch, 0)
        out = me.test_step_outputs["test"]
        assert me.calls == c["K"]
        # what the reference hands to sample(): the frames after the conditioning window are zeroed (nba.py:188-189, md17.py:149-151)
        assert float(me.seen[0]["pos"][:, c["cond_idx"][1]:].abs().sum()) == 0.0
        fixtures[c["case"]] = dict(case=dict(c), ades=out["ades"][0].clone(), fdes=out["fdes"][0].clone(),
                                   checksum=float(sum(p.double().sum() for p in preds) + true_pos.double().sum() + mask.double().sum()))
        print(f"{c['case']:28s} ades {tuple(out['ades'][0].shape)} mean {float(out['ades'][0].mean()):.4f}  fdes mean {float(out['fdes'][0].mean()):.4f}")
    torch.save(fixtures, os.path.join(GOLDEN_DIR, "ksample_metrics.pt"))


# SDE sampler: the reference's own Sampler.sample_sde / integrators.sde around its LatentSIV3
SDE_CASES = [
    dict(case="sde_gvp_data_euler", cfg="pedestrian", overrides=dict(depth=1), B=2, T=6, path_type="GVP", prediction="data",
         kwargs=dict(sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean", last_step_size=0.04, num_steps=8),
         seeds=(151, 152, 153)),
    dict(case="sde_gvp_data_heun", cfg="pedestrian", overrides=dict(depth=1), B=2, T=6, path_type="GVP", prediction="data",
         kwargs=dict(sampling_method="Heun", diffusion_form="sigma", diffusion_norm=0.7, last_step="Euler", last_step_size=0.04, num_steps=6),
         seeds=(161, 162, 163)),
    dict(case="sde_linear_velocity_tweedie", cfg="pedestrian", overrides=dict(depth=1), B=2, T=6, path_type="Linear", prediction="velocity",
         kwargs=dict(sampling_method="Euler", diffusion_form="linear", diffusion_norm=1.0, last_step="Tweedie", last_step_size=0.04, num_steps=7),
         seeds=(171, 172, 173)),
    dict(case="sde_gvp_data_nolast", cfg="pedestrian", overrides=dict(depth=1), B=2, T=6, path_type="GVP", prediction="data",
         kwargs=dict(sampling_method="Euler", diffusion_form="decreasing", diffusion_norm=1.0, last_step=None, last_step_size=0.04, num_steps=5),
         seeds=(181, 182, 183)),
]


def sde_case_inputs(c: dict):
    cfg = get_config(c["cfg"], **c["overrides"])
    s_bb, s_x, s_noise = c["seeds"]
    bb_sd = O.init_backbone_params(cfg["backbone"], s_bb)
    g = torch.Generator().manual_seed(s_x)
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    shape = (c["B"], c["T"], L, D)
    x0 = torch.randn(shape, generator=g)
    x_cond = torch.randn(shape, generator=g)
    mask = torch.zeros(c["B"], c["T"], L, dtype=torch.int64)
    mask[:, :2] = 1
    y = torch.randn(c["B"], cfg["backbone"]["vec_in_dim"], generator=g) if cfg["backbone"]["vec_in_dim"] else None
    return cfg, bb_sd, x0, x_cond, mask, y, s_noise


def make_sde_golden(ref) -> None:
    import src.modules.transport.integrators as integ  # the reference module (torchdiffeq shim installed by load_reference)
    fixtures = {}
    for c in SDE_CASES:
        cfg, bb_sd, x0, x_cond, mask, y, s_noise = sde_case_inputs(c)
        bb = cfg["backbone"]
        net = ref.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                             vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"], theta=bb["theta"]).eval()
        net.load_state_dict(bb_sd, strict=True)
        si = ref.CreateTransport(path_type=c["path_type"], prediction=c["prediction"])()
        fn = ref.Sampler(si).get_sample_fn("SDE", dict(c["kwargs"]))
        noises = []
        real_randn = integ.th.randn

        def recording_randn(*a, **k):  # integrators.py:31,41 draw th.randn(x.size()) per step: keep what was drawn
            w = real_randn(*a, **k)
            noises.append(w.clone())
            return w

        kw = dict(x_cond=x_cond, x_cond_mask=mask)
        if y is not None:
            kw["y"] = y
        torch.manual_seed(s_noise)
        integ.th.randn = recording_randn
        try:
            with torch.no_grad():
                xs = fn(x0, lambda xt, t, **k: net(x=xt, t=t, **k), **kw)
        finally:
            integ.th.randn = real_randn
        assert len(xs) == c["kwargs"]["num_steps"] and len(noises) == c["kwargs"]["num_steps"] - 1
        fixtures[c["case"]] = dict(case=dict(c), checksums=dict(bb=O.state_checksum(bb_sd), x0=float(x0.double().sum())),
                                   noises=torch.stack(noises), states=torch.stack(xs))
        print(f"{c['case']:28s} states {tuple(torch.stack(xs).shape)}  |x| max {float(torch.stack(xs).abs().max()):.3f}")
    torch.save(fixtures, os.path.join(GOLDEN_DIR, "sde_sampler.pt"))


def case_inputs(c: dict):
    """Everything a test needs to re-create the inputs of a golden case (shared with tests/)."""
    cfg = get_config(c["cfg"], **c["overrides"])
    if "path_type" in c:
        cfg["path_type"], cfg["prediction"] = c["path_type"], c["prediction"]
    s_fs, s_bb, s_batch, s_noise = c["seeds"]
    fs_sd = O.init_first_stage_params(cfg["first_stage"], s_fs)
    bb_sd = O.init_backbone_params(cfg["backbone"], s_bb)
    batch = O.synthetic_batch(cfg, c["B"], s_batch, T=c["T"])
    g = torch.Generator().manual_seed(s_noise)
    L = cfg["first_stage"]["encoder"]["num_latents"]
    noise = torch.randn(c["B"], c["T"], L, cfg["backbone"]["in_dim"], generator=g)
    y = None
    if cfg["n_classes"]:
        table = torch.randn(cfg["n_classes"], 256, generator=g)  # CondWrapper.vec_in_embedding (nba.py:254-263)
        y = table[batch["cond_scene"]]
    return cfg, fs_sd, bb_sd, batch, noise, y


def main() -> None:
    ref = load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if not any(a.startswith("--case=") for a in sys.argv):
        make_rollout_golden(ref)
        make_ksample_golden()
        make_sde_golden(ref)
    if "--rollout-only" in sys.argv or "--widened-only" in sys.argv:
        return
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--case=")]
    for c in CASES:
        if only and c["case"] not in only:
            continue
        cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
        bb = cfg["backbone"]
        fs = RefFirstStage(cfg["first_stage"]).eval()
        fs.load_state_dict(fs_sd, strict=True)
        net = ref.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"],
                             num_heads=bb["num_heads"], vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"],
                             normalize=bb["normalize"], theta=bb["theta"]).eval()
        net.load_state_dict(bb_sd, strict=True)
        rec: dict = {}
        out = reference_sample(fs, net, {k: v.clone() for k, v in batch.items()}, cond_idx=cfg["cond_idx"],
                               path_type=cfg["path_type"], prediction=cfg["prediction"], num_steps=c["num_steps"],
                               noise=noise, y=y, mask_cond_mean=cfg["mask_cond_mean"], record=rec)
        # single backbone evaluation at the first grid point (the per-op parity anchor)
        t0, _ = O.sample_interval(cfg["path_type"], cfg["prediction"])
        tt = torch.full((c["B"],), t0)
        kw = dict(x_cond=rec["x_cond"], x_cond_mask=rec["x_cond_mask"])
        if y is not None:
            kw["y"] = y
        with torch.no_grad():
            net_out = net(x=noise, t=tt, **kw)
        big = c["T"] > 100
        sl = slice(None, None, 20) if big else slice(None)
        heavy = rec["latents"][:, sl].numel() > 100_000  # MD17: L=192 ⇒ keep first/last velocity only
        vsel = [0, c["num_steps"] - 2] if heavy else list(range(c["num_steps"] - 1))
        if "vel_steps" in c:
            vsel = list(c["vel_steps"])
        fixture = dict(
            case={k: v for k, v in c.items()},
            checksums=dict(fs=O.state_checksum(fs_sd), bb=O.state_checksum(bb_sd),
                           batch=O.state_checksum({k: v.float() for k, v in batch.items()}),
                           noise=float(noise.double().sum())),
            latents=rec["latents"][:, sl].clone(),
            x_cond=None if heavy else rec["x_cond"][:, sl].clone(),
            net_out_t0=net_out[:, sl].clone(),
            velocities=rec["velocities"][vsel][:, :, sl].clone(),
            velocity_steps=vsel,
            final_latents=rec["states"][-1].clone() if not big else rec["states"][-1][:, sl].clone(),
            outputs={k: v[:, sl].clone() for k, v in out.items()},
            frame_slice=(sl.start, sl.stop, sl.step),
            torch_version=torch.__version__,
        )
        path = os.path.join(GOLDEN_DIR, c["case"] + ".pt")
        torch.save(fixture, path)
        print(f"{c['case']:28s} -> {os.path.getsize(path) / 1024:8.1f} KiB   vel max {float(rec['velocities'].abs().max()):.3f}")


if __name__ == "__main__":
    main()
