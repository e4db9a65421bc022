"""CPU: the oracle restatement reproduces the golden vectors generated from the real reference modules
(oracle/make_golden.py).  fp32 vs fp32, so the tolerance is rounding only."""
import pytest
import torch

from oracle import lamslide_oracle as O
from tests.helpers import CASE_BY_NAME, case_inputs, check_inputs_match_fixture, frame_slice, load_golden, max_rel

FAST = ["peptide_small", "md17_small", "nba_full", "pedestrian_full", "peptide_linear_velocity", "md17_full", "peptide_full", "peptide_steps20",
        "peptide_steps50"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(8)
    fx = load_golden(name)
    c = CASE_BY_NAME[name]
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    check_inputs_match_fixture(fx, fs_sd, bb_sd, batch, noise)
    rec = {}
    with torch.no_grad():
        out = O.sample(fs_sd, bb_sd, cfg, batch, noise, num_steps=c["num_steps"], y=y, record=rec)
    sl = frame_slice(fx)
    assert max_rel(rec["latents"][:, sl], fx["latents"]) < 1e-5
    if fx["x_cond"] is not None:
        assert max_rel(rec["x_cond"][:, sl], fx["x_cond"]) < 1e-5
    vel = rec["velocities"][fx["velocity_steps"]][:, :, sl]
    assert max_rel(vel, fx["velocities"]) < 2e-4  # (π/2)/cos(πt/2) amplifies fp32 rounding ≈ 30×
    assert max_rel(rec["states"][-1][:, sl], fx["final_latents"]) < 1e-4
    for k, v in fx["outputs"].items():
        assert max_rel(out[k][:, sl], v) < 1e-4


def test_oracle_rollout_matches_reference_golden():
    """Autoregressive roll-out: the oracle restatement of SIAtom14SamplingWrapper.sample_rollout against the
    positions produced by the reference's own class (oracle/make_golden.py: make_rollout_golden)."""
    from oracle.make_golden import ROLLOUT_CASE, rollout_case_inputs
    torch.set_num_threads(8)
    fx = load_golden("peptide_rollout")
    c = ROLLOUT_CASE
    cfg, fs_sd, bb_sd, cond_pos, res, res_mask, noises = rollout_case_inputs(c)
    cs = fx["checksums"]
    assert abs(O.state_checksum(fs_sd) - cs["fs"]) <= 1e-6 * max(1.0, abs(cs["fs"]))
    assert abs(O.state_checksum(bb_sd) - cs["bb"]) <= 1e-6 * max(1.0, abs(cs["bb"]))
    assert abs(float(cond_pos.double().sum() + res.double().sum() + res_mask.double().sum()) - cs["inputs"]) < 1e-6
    assert abs(float(sum(n.double().sum() for n in noises)) - cs["noise"]) < 1e-6
    with torch.no_grad():
        pos = O.sample_rollout(fs_sd, bb_sd, cfg, cond_pos, res, res_mask, noises, shift=c["shift"], scale=c["scale"],
                               num_steps=c["num_steps"])
    assert pos.shape == fx["positions"].shape == (c["num_rollouts"] * c["T"], c["R"], 14, 3)
    assert torch.allclose(pos[0], cond_pos, atol=1e-6)  # sampling.py:62: the conditioning frame is put back at index 0 ...
    assert max_rel(pos, fx["positions"]) < 2e-4  # ... and errors feed forward through three chained sample() calls


def test_oracle_ksample_metrics_match_reference_golden():
    """K-sample ADE / FDE: the oracle restatements against the outputs of the reference's own
    Wrapper.test_step bodies (nba / pedestrian: min over runs per agent; md17: mean over runs per sample) on preset predictions."""
    from oracle.make_golden import KSAMPLE_CASES
    fx = load_golden("ksample_metrics")
    for c in KSAMPLE_CASES:
        preds, true_pos, mask = O.ksample_inputs(c["B"], c["T"], c["A"], c["D"], c["K"], c["seed"], c["pad"])
        f = fx[c["case"]]
        assert abs(float(sum(p.double().sum() for p in preds) + true_pos.double().sum() + mask.double().sum()) - f["checksum"]) < 1e-6
        if c["mode"] == "min":
            ades, fdes = O.ksample_min_ade_fde(preds, true_pos, mask, c["cond_idx"][1], c["num_runs"])
        else:
            ades, fdes = O.ksample_mean_ade_fde(preds, true_pos, c["cond_idx"][1])
        assert ades.shape == f["ades"].shape and fdes.shape == f["fdes"].shape
        assert torch.allclose(ades, f["ades"], rtol=1e-6, atol=1e-7), c["case"]
        assert torch.allclose(fdes, f["fdes"], rtol=1e-6, atol=1e-7), c["case"]


def test_oracle_sde_sampler_matches_reference_golden():
    """SDE sampler: the oracle restatement of Sampler.sample_sde + integrators.sde (Euler-Maruyama / Heun, four
    diffusion forms, all last-step variants) against the states the reference's own Sampler produced with the recorded noise."""
    from oracle.make_golden import SDE_CASES, sde_case_inputs
    torch.set_num_threads(8)
    fx = load_golden("sde_sampler")
    for c in SDE_CASES:
        cfg, bb_sd, x0, x_cond, mask, y, _ = sde_case_inputs(c)
        f = fx[c["case"]]
        assert abs(O.state_checksum(bb_sd) - f["checksums"]["bb"]) <= 1e-6 * max(1.0, abs(f["checksums"]["bb"]))
        assert abs(float(x0.double().sum()) - f["checksums"]["x0"]) < 1e-6
        with torch.no_grad():
            xs = O.sde_sample(lambda x, t: O.backbone_forward(bb_sd, cfg["backbone"], x, t, x_cond, mask, y), x0, list(f["noises"]),
                              path_type=c["path_type"], prediction=c["prediction"], **c["kwargs"])
        got = torch.stack(xs)
        assert got.shape == f["states"].shape
        for i in range(got.shape[0]):
            assert max_rel(got[i], f["states"][i]) < 5e-4, (c["case"], i)


def test_backbone_single_eval_matches_golden():
    fx = load_golden("nba_full")
    c = CASE_BY_NAME["nba_full"]
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    with torch.no_grad():
        flat = {k: v.flatten(0, 1) for k, v in batch.items() if k != "cond_scene"}
        lat = O.first_stage_encode(fs_sd, cfg["first_stage"], flat).unflatten(0, (c["B"], c["T"]))
        x_cond, m = O.setup_conditioning(lat, cfg["cond_idx"], True)
        t0, _ = O.sample_interval(cfg["path_type"], cfg["prediction"])
        out = O.backbone_forward(bb_sd, cfg["backbone"], noise, torch.full((c["B"],), t0), x_cond, m, y)
    assert max_rel(out, fx["net_out_t0"]) < 1e-5


def test_euler_grid_and_eval_count():
    """num_steps=n makes n-1 model calls on linspace(t0,t1,n) (integrators.py:98)."""
    calls = []

    def fn(x, t):
        calls.append(float(t[0]))
        return torch.zeros_like(x)

    O.ode_sample(fn, torch.zeros(2, 3, 1, 4), path_type="GVP", prediction="data", num_steps=10)
    assert len(calls) == 9
    assert abs(calls[0] - 0.001) < 1e-7 and abs(calls[-1] - 0.8881111) < 1e-5
    calls.clear()
    O.ode_sample(fn, torch.zeros(2, 3, 1, 4), path_type="Linear", prediction="velocity", num_steps=10)
    assert len(calls) == 9 and calls[0] == 0.0 and abs(calls[-1] - 8 / 9) < 1e-6


def test_gvp_data_drift_closed_form():
    """v = (π/2)(m − sin(a)x)/cos(a), a = πt/2."""
    import math
    g = torch.Generator().manual_seed(0)
    x, m = torch.randn(4, 5, 2, 8, generator=g, dtype=torch.float64), torch.randn(4, 5, 2, 8, generator=g, dtype=torch.float64)
    t = torch.tensor([0.001, 0.3, 0.7, 0.8881], dtype=torch.float64)
    a = (math.pi * t / 2).reshape(-1, 1, 1, 1)
    closed = (math.pi / 2) * (m - torch.sin(a) * x) / torch.cos(a)
    assert max_rel(O.drift("GVP", "data", x, t, m), closed) < 1e-12


def test_masked_padding_is_exact():
    """Encoder output with zero-padded entities + key mask == un-padded call."""
    from lam_slide_b200.configs import get_config
    cfg = get_config("pedestrian")
    sd = O.init_first_stage_params(cfg["first_stage"], 7)
    g = torch.Generator().manual_seed(1)
    pos = torch.randn(5, 4, 2, generator=g)
    ent = torch.stack([torch.randperm(10, generator=g)[:4] for _ in range(5)])
    small = dict(pos=pos, entities=ent, attention_mask=torch.ones(5, 4, dtype=torch.bool))
    pad = dict(pos=torch.cat([pos, torch.zeros(5, 6, 2)], 1), entities=torch.cat([ent, torch.zeros(5, 6, dtype=torch.int64)], 1),
               attention_mask=torch.cat([torch.ones(5, 4, dtype=torch.bool), torch.zeros(5, 6, dtype=torch.bool)], 1))
    a = O.first_stage_encode(sd, cfg["first_stage"], small)
    b = O.first_stage_encode(sd, cfg["first_stage"], pad)
    assert max_rel(a, b) < 1e-6
