"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors that were
generated from the real reference modules.  Tolerances are the ones BASELINE.json's north_star states:
    per-step velocity max relative error  <= 1e-2   (bf16 tensor-core operands, fp32 state)
    final-frame coordinate RMSD           <= 1e-3 of the data scale (synthetic data scale = 1)
The first stage runs in fp32 and is held to 1e-4."""
import numpy as np
import pytest
import torch

from oracle import lamslide_oracle as O
from tests.helpers import CASE_BY_NAME, case_inputs, check_inputs_match_fixture, frame_slice, load_golden, max_rel, mean_rel, rmsd

pytestmark = pytest.mark.gpu

VEL_TOL = 1e-2        # max |err| / max |ref|  (north_star: "per-step velocity max relative error <= 1e-2 in bf16")
VEL_MEAN_TOL = 5e-3   # mean |err| / mean |ref|: the same contract where the reference is small
RMSD_TOL = 1e-3
FS_TOL = 1e-4


def _cuda_batch(batch):
    return {k: v.cuda() for k, v in batch.items()}


def _build(cfg, fs_sd, bb_sd):
    import lam_slide_b200 as P
    m = P.SecondStageSampler(cfg).cuda()
    m.first_stage_model.backbone.load_state_dict(fs_sd, strict=True)
    m.backbone.load_state_dict(bb_sd, strict=True)
    return m


@pytest.mark.parametrize("name", ["peptide", "md17", "nba", "pedestrian"])
def test_first_stage_encode_decode_vs_oracle(name):
    from lam_slide_b200.configs import get_config
    cfg = get_config(name)
    fs_sd = O.init_first_stage_params(cfg["first_stage"], 21)
    bb_sd = O.init_backbone_params(dict(cfg["backbone"], depth=1), 22)
    cfg["backbone"]["depth"] = 1
    m = _build(cfg, fs_sd, bb_sd)
    B, T = 3, 5
    batch = O.synthetic_batch(cfg, B, 23, T=T)
    flat = {k: v.flatten(0, 1) for k, v in batch.items() if k != "cond_scene"}
    with torch.no_grad():
        lat_ref = O.first_stage_encode(fs_sd, cfg["first_stage"], flat)
        out_ref = O.first_stage_decode(fs_sd, cfg["first_stage"], lat_ref, flat["entities"])
    fs = m.first_stage_model.backbone
    lat = fs.encode(_cuda_batch(flat))
    assert max_rel(lat.cpu(), lat_ref) < FS_TOL
    out = fs.decode(lat_ref.cuda(), flat["entities"].cuda())
    for k, v in out_ref.items():
        assert max_rel(out[k].cpu(), v) < FS_TOL, k


def test_setup_conditioning_vs_oracle():
    import lam_slide_b200 as P
    m = P.SecondStageSampler.from_name("nba").cuda()
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(3, 20, 8, 32, generator=g)
    xc_ref, mk_ref = O.setup_conditioning(lat, (0, 8), True)
    xc, mk = m.setup_conditioning(lat.cuda())
    assert torch.equal(mk.cpu(), mk_ref)
    assert max_rel(xc.cpu(), xc_ref) < 1e-6


@pytest.mark.parametrize("name,depth,B,T", [("peptide", 2, 2, 24), ("md17", 1, 1, 6), ("nba", 2, 3, 20), ("pedestrian", 2, 5, 20),
                                            ("peptide", 1, 1, 200)])
def test_backbone_forward_vs_oracle(name, depth, B, T):
    """One LatentSIV3.forward: bf16 operands vs the fp32 oracle."""
    import lam_slide_b200 as P
    from lam_slide_b200.configs import get_config
    cfg = get_config(name, depth=depth)
    bb = cfg["backbone"]
    bb_sd = O.init_backbone_params(bb, 31)
    net = P.LatentSIV3(depth=depth, in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                       vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).cuda()
    net.load_state_dict(bb_sd, strict=True)
    L = cfg["first_stage"]["encoder"]["num_latents"]
    g = torch.Generator().manual_seed(32)
    x = torch.randn(B, T, L, bb["in_dim"], generator=g)
    xc = torch.randn(B, T, L, bb["in_dim"], generator=g)
    mk = (torch.rand(B, T, L, generator=g) < 0.3).long()
    t = torch.rand(B, generator=g)
    y = torch.randn(B, bb["vec_in_dim"], generator=g) if bb["vec_in_dim"] else None
    with torch.no_grad():
        ref = O.backbone_forward(bb_sd, bb, x, t, xc, mk, y)
    out = net(x.cuda(), t.cuda(), xc.cuda(), mk.cuda(), None if y is None else y.cuda())
    assert torch.isfinite(out).all()
    assert max_rel(out.cpu(), ref) < VEL_TOL


@pytest.mark.parametrize("name", ["peptide_small", "md17_small", "nba_full", "pedestrian_full", "peptide_linear_velocity",
                                  "md17_full", "peptide_full", "peptide_steps20", "peptide_steps50"])
def test_sample_vs_reference_golden(name):
    """Full sample(): encode -> conditioning -> Euler ODE -> decode vs the golden vectors of the REAL reference."""
    fx = load_golden(name)
    c = CASE_BY_NAME[name]
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    check_inputs_match_fixture(fx, fs_sd, bb_sd, batch, noise)
    m = _build(cfg, fs_sd, bb_sd)
    sl = frame_slice(fx)
    B, T = c["B"], c["T"]
    cb = _cuda_batch(batch)
    latents = m.encode(cb)
    assert max_rel(latents.cpu()[:, sl], fx["latents"]) < FS_TOL
    x_cond, x_mask = m.setup_conditioning(latents)
    if fx["x_cond"] is not None:
        assert max_rel(x_cond.cpu()[:, sl], fx["x_cond"]) < FS_TOL
    # single network evaluation at t0
    t0, _ = O.sample_interval(cfg["path_type"], cfg["prediction"])
    yy = None if y is None else y.cuda()
    out0 = m.backbone(noise.cuda(), torch.full((B,), t0).cuda(), x_cond, x_mask, yy)
    assert max_rel(out0.cpu()[:, sl], fx["net_out_t0"]) < VEL_TOL
    assert mean_rel(out0.cpu()[:, sl], fx["net_out_t0"]) < VEL_MEAN_TOL
    # fused ODE with velocity recording
    states, vel = m.backbone.ode_sample(noise.cuda(), x_cond, x_mask, yy, path_type=cfg["path_type"], prediction=cfg["prediction"],
                                        num_steps=c["num_steps"], return_velocities=True)
    vel = vel.cpu()[fx["velocity_steps"]][:, :, sl]
    for i in range(vel.shape[0]):
        assert max_rel(vel[i], fx["velocities"][i]) < VEL_TOL, f"velocity step {fx['velocity_steps'][i]}"
        assert mean_rel(vel[i], fx["velocities"][i]) < VEL_MEAN_TOL, f"velocity step {fx['velocity_steps'][i]} (mean)"
    assert max_rel(states[-1].cpu()[:, sl], fx["final_latents"]) < VEL_TOL
    out = m.first_stage_model.decode(states[-1].flatten(0, 1), cb["entities"].flatten(0, 1))
    main = cfg["main_output"]
    got = out[main].unflatten(0, (B, T)).cpu()[:, sl]
    # final-frame coordinate RMSD (data scale 1 for the synthetic N(0,1) coordinates)
    assert rmsd(got[:, -1], fx["outputs"][main][:, -1]) < RMSD_TOL
    assert rmsd(got, fx["outputs"][main]) < RMSD_TOL


def test_public_sample_api_host_batch():
    """model.sample(batch) with a HOST batch (the call a user of the reference makes): H2D inside, dict out."""
    c = CASE_BY_NAME["nba_full"]
    fx = load_golden("nba_full")
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    g = torch.Generator().manual_seed(c["seeds"][3])
    _ = torch.randn(c["B"], c["T"], 8, 32, generator=g)
    table = torch.randn(cfg["n_classes"], 256, generator=g)
    m.vec_in_embedding.weight.data.copy_(table)
    out = m.sample({k: v.clone() for k, v in batch.items()}, noise=noise)
    assert out["pos"].shape == (c["B"], c["T"], cfg["N"], 2)
    assert rmsd(out["pos"].cpu(), fx["outputs"]["pos"]) < RMSD_TOL


def test_sample_stream_matches_sample():
    """sample_stream (host batches in, pinned host results out, copies overlapped with compute on a second stream) returns
    exactly what sample() returns, batch by batch and in order."""
    c = CASE_BY_NAME["nba_full"]
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    batches = []
    for i in range(4):
        b = {k: v.clone() for k, v in batch.items()}
        b["pos"] = b["pos"] + 0.01 * i
        batches.append({k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in b.items()})
    want = [m.sample({k: v.clone() for k, v in b.items()}, noise=noise)["pos"].cpu() for b in batches]
    got = [o.clone() for o in m.sample_stream(({k: v for k, v in b.items()} for b in batches), noise=noise)]
    assert len(got) == len(want)
    for g_, w_ in zip(got, want):
        assert torch.equal(g_, w_)
    assert not torch.equal(want[0], want[1])
    # a yielded (pinned) result stays valid while the consumer works on the next one: keep the previous result WITHOUT cloning
    prev, n = None, 0
    for k, o in enumerate(m.sample_stream(({kk: v for kk, v in b.items()} for b in batches), noise=noise)):
        assert o.is_pinned()
        if prev is not None:
            torch.cuda.synchronize()  # whatever copies are in flight have landed: prev must still hold result k - 1
            assert torch.equal(prev, want[k - 1]), k
        assert torch.equal(o, want[k])
        prev, n = o, n + 1
    assert n == len(want)


def test_rollout_vs_reference_golden():
    """Roll-out driver: SIAtom14SamplingWrapper.sample_rollout on the GPU path against the positions the
    reference's own class produced (tests/golden/peptide_rollout.pt); three chained sample() calls, so the single-call RMSD
    tolerance is applied per block with the error of the earlier blocks feeding forward (x3)."""
    import lam_slide_b200 as P
    from oracle.make_golden import ROLLOUT_CASE, rollout_case_inputs
    c = ROLLOUT_CASE
    fx = load_golden("peptide_rollout")
    cfg, fs_sd, bb_sd, cond_pos, res, res_mask, noises = rollout_case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    m.hparams.sampling_kwargs["num_steps"] = c["num_steps"]
    w = P.SIAtom14SamplingWrapper(m, shift=c["shift"], scale=c["scale"])
    pos = w.sample_rollout(cond_pos, res, res_mask, num_rollouts=c["num_rollouts"], noise=torch.stack(noises)).cpu()
    assert pos.shape == fx["positions"].shape
    assert torch.allclose(pos[0], cond_pos, atol=1e-6)
    T = c["T"]
    for i in range(c["num_rollouts"]):
        blk = slice(max(i * T, 1), (i + 1) * T)
        assert rmsd(pos[blk], fx["positions"][blk]) < RMSD_TOL * (i + 1) * c["scale"], f"block {i}"


def test_rollout_batched_equals_chain_by_chain():
    """sample_rollouts (B chains together, one encoded frame per chain and step, latents broadcast over T) equals the
    reference-style loop — model.sample(create_batch(...)) per chain, which encodes T copies of the frame — bit for bit."""
    import lam_slide_b200 as P
    from oracle.make_golden import ROLLOUT_CASE, rollout_case_inputs
    c = ROLLOUT_CASE
    cfg, fs_sd, bb_sd, _, _, _, _ = rollout_case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    m.hparams.sampling_kwargs["num_steps"] = c["num_steps"]
    w = P.SIAtom14SamplingWrapper(m, shift=c["shift"], scale=c["scale"])
    B, R, T, n_roll = 3, c["R"], c["T"], 2
    ins = [O.rollout_inputs(R, 900 + b) for b in range(B)]
    cond = torch.stack([x[0] for x in ins])
    res = torch.stack([x[1] for x in ins])
    msk = torch.stack([x[2] for x in ins])
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    noise = torch.randn(n_roll, B, T, L, D, generator=torch.Generator().manual_seed(5))
    got = w.sample_rollouts(cond, res, msk, num_rollouts=n_roll, noise=noise).cpu()
    assert got.shape == (B, n_roll * T, R, 14, 3)
    for b in range(B):
        pos = ((cond[b] - c["shift"]) / c["scale"]).cuda()
        blocks = []
        for i in range(n_roll):
            batch = w.create_batch(pos, res[b].cuda(), msk[b].cuda())
            pred = m.sample(batch, noise=noise[i, b:b + 1])["atom14_pos"].squeeze(0)
            blocks.append(pred)
            pos = pred[-1].clone()
        want = torch.cat(blocks)
        want[0] = ((cond[b] - c["shift"]) / c["scale"]).cuda()
        want = (want * c["scale"] + c["shift"]).cpu()
        assert torch.equal(got[b], want), f"chain {b}: max diff {float((got[b] - want).abs().max()):.3e}"


def test_rollout_to_trajectory_files():
    """sample_traj_files: roll-out -> atom37 heavy atoms -> DCD + PDB (sampling.py:65-142, eval_peptide.py:340-349 without mdtraj)."""
    import os
    import tempfile
    import lam_slide_b200 as P
    from lam_slide_b200 import formats as F
    from oracle.make_golden import ROLLOUT_CASE, rollout_case_inputs
    c = ROLLOUT_CASE
    cfg, fs_sd, bb_sd, cond_pos, res, res_mask, noises = rollout_case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    m.hparams.sampling_kwargs["num_steps"] = c["num_steps"]
    w = P.SIAtom14SamplingWrapper(m, shift=c["shift"], scale=c["scale"])
    with tempfile.TemporaryDirectory() as d:
        pos, dcd, pdb = w.sample_traj_files(cond_pos, res, os.path.join(d, "pep"), num_rollouts=2, noise=torch.stack(noises[:2]))
        n_atoms = int(F.RESTYPE_ATOM14_MASK[res].sum())
        xyz = F.read_dcd(dcd)
        assert xyz.shape == (2 * c["T"], n_atoms, 3) and np.isfinite(xyz).all()
        want = F.atom14_to_heavy_atoms(pos, res) * 10.0
        assert np.array_equal(xyz, want.numpy().astype(np.float32))
        px, atoms = F.read_pdb(pdb)
        assert px.shape == (1, n_atoms, 3) and len(atoms) == n_atoms
        assert np.abs(px[0] - want[0].numpy()).max() <= 5.01e-4


def test_ksample_errors_kernel_vs_reference_golden():
    """lamslide_ksample_errors (C-ABI) against the outputs of the reference's own test_step bodies (tests/golden/ksample_metrics.pt)."""
    import lam_slide_b200 as P
    from oracle.make_golden import KSAMPLE_CASES
    fx = load_golden("ksample_metrics")
    for c in KSAMPLE_CASES:
        preds, true_pos, mask = O.ksample_inputs(c["B"], c["T"], c["A"], c["D"], c["K"], c["seed"], c["pad"])
        c1 = c["cond_idx"][1]
        pk = torch.stack(preds)[:, :, c1:].cuda()
        ades, fdes = P.ksample_errors(pk, true_pos[:, c1:].cuda(), c["num_runs"], c["mode"])
        if c["mode"] == "min":
            sel = mask[:, -1].reshape(-1).cuda()
            ades, fdes = ades[sel], fdes[sel]
        f = fx[c["case"]]
        assert torch.allclose(ades.cpu(), f["ades"], rtol=2e-6, atol=1e-6), c["case"]
        assert torch.allclose(fdes.cpu(), f["fdes"], rtol=2e-6, atol=1e-6), c["case"]


@pytest.mark.parametrize("name,mode", [("nba_full", "min"), ("pedestrian_full", "min"), ("md17_full", "mean")])
def test_ksample_evaluator_equals_k_sample_calls(name, mode):
    """KSampleEvaluator.test_step (one batched solve of B*K trajectories from once-encoded latents + the metric kernel) against K
    separate model.sample() calls on the blanked batch reduced by the oracle's restatement of the reference's test_step."""
    import lam_slide_b200 as P
    c = CASE_BY_NAME[name]
    cfg, fs_sd, bb_sd, batch, noise0, y = case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    m.hparams.sampling_kwargs["num_steps"] = c["num_steps"]
    if cfg["n_classes"]:
        g = torch.Generator().manual_seed(c["seeds"][3])
        _ = torch.randn(c["B"], c["T"], cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"], generator=g)
        m.vec_in_embedding.weight.data.copy_(torch.randn(cfg["n_classes"], 256, generator=g))
    K, runs = 4, 3
    noise = torch.randn(K, *noise0.shape, generator=torch.Generator().manual_seed(9))
    ev = P.KSampleEvaluator(m, K=K, num_runs=runs, mode=mode)
    ades, fdes = ev.test_step({k: v.clone() for k, v in batch.items()}, noise=noise)
    c1 = cfg["cond_idx"][1]
    blank = {k: v.clone() for k, v in batch.items()}
    true_pos = blank["pos"].clone()
    blank["pos"][:, c1:] = 0
    if mode == "mean":
        blank["atom"][:, c1:] = 0
    preds = [m.sample({k: v.clone() for k, v in blank.items()}, noise=noise[k])["pos"].cpu() for k in range(K)]
    if mode == "min":
        mask = batch.get("attention_mask", torch.ones(true_pos.shape[:3], dtype=torch.bool))
        want_a, want_f = O.ksample_min_ade_fde(preds, true_pos, mask, c1, runs)
    else:
        want_a, want_f = O.ksample_mean_ade_fde(preds, true_pos, c1)
    assert ades.shape == want_a.shape
    assert torch.allclose(ades.cpu(), want_a, rtol=1e-5, atol=1e-6)
    assert torch.allclose(fdes.cpu(), want_f, rtol=1e-5, atol=1e-6)
    # the runs may be solved in chunks (memory bound on B * K): the same bits (every kernel is chosen by layer shape, not by row count)
    ev1 = P.KSampleEvaluator(m, K=K, num_runs=runs, mode=mode, max_trajectories=c["B"])
    a1, f1 = ev1.test_step({k: v.clone() for k, v in batch.items()}, noise=noise)
    assert torch.equal(a1, ades) and torch.equal(f1, fdes)


def test_sde_sampler_vs_reference_golden():
    """SDE sampler: Sampler.get_sample_fn("SDE", ...) on the GPU path (C-ABI backbone forward + lamslide_lincomb3
    updates, fp64 host coefficients) against the states of the reference's own Sampler.sample_sde, replaying its recorded noise."""
    import lam_slide_b200 as P
    from oracle.make_golden import SDE_CASES, sde_case_inputs
    fx = load_golden("sde_sampler")
    for c in SDE_CASES:
        cfg, bb_sd, x0, x_cond, mask, y, _ = sde_case_inputs(c)
        f = fx[c["case"]]
        bb = cfg["backbone"]
        net = P.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                           vec_in_dim=bb.get("vec_in_dim"), mlp_ratio=bb["mlp_ratio"], theta=bb.get("theta", 10_000),
                           normalize=bb.get("normalize", False), n_timesteps=c["T"])
        net.load_state_dict(bb_sd, strict=True)
        net = net.cuda()
        noises = list(f["noises"])
        si = P.CreateTransport(path_type=c["path_type"], prediction=c["prediction"])()
        fn = P.Sampler(si, noise_fn=lambda shape, device: noises.pop(0)).get_sample_fn("SDE", dict(c["kwargs"]))
        kw = dict(x_cond=x_cond.cuda(), x_cond_mask=mask.cuda())
        if y is not None:
            kw["y"] = y.cuda()
        xs = fn(x0.cuda(), lambda xt, t, **k: net(x=xt, t=t, **k), **kw)
        assert len(xs) == c["kwargs"]["num_steps"] and not noises
        got = torch.stack(xs).cpu()
        # the stochastic integrators amplify the bf16 network error step by step (SBDM near t0 = 1e-3 by ~1/t): same tolerance class as
        # the per-step velocity of the ODE path, relative to the state magnitude of each step
        for i in range(got.shape[0]):
            assert max_rel(got[i], f["states"][i]) < VEL_TOL, (c["case"], i, max_rel(got[i], f["states"][i]))


def test_full_size_batch_properties():
    """BASELINE.json's full configuration (4AA peptides, T = 1000, B = 64 => 128 000 tokens per launch; far beyond what the oracle
    finishes in seconds) through size-independent properties of the path: trajectories are independent, so (1) the run is
    deterministic, (2) permuting the batch permutes the result, (3) a sub-batch sampled alone equals its slice of the full run —
    bit for bit, which exercises every tile that straddles two samples, the persistent kernels' multi-wave loops and the CTA-pair
    tails."""
    import lam_slide_b200 as P
    from lam_slide_b200.synthetic import randomize_zero_init, synthetic_batch
    cfg = P.get_config("peptide")
    torch.manual_seed(0)
    m = P.SecondStageSampler(cfg, sampling_kwargs={"sampling_method": "euler", "num_steps": 10})
    randomize_zero_init(m, seed=1)
    m = m.cuda()
    B, T = 64, cfg["T"]
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    batch = synthetic_batch(cfg, B, seed=5)
    noise = torch.randn(B, T, L, D, generator=torch.Generator().manual_seed(6))
    key = cfg["main_output"]

    def run(idx):
        b = {k: v[idx].clone() for k, v in batch.items()}
        return m.sample(b, noise=noise[idx].clone())[key]

    full_idx = torch.arange(B)
    full = run(full_idx)
    assert torch.isfinite(full).all()
    assert torch.equal(run(full_idx), full)                      # (1)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(7))
    assert torch.equal(run(perm), full[perm.cuda()])             # (2)
    sub = torch.tensor([3, 17, 18, 40, 63])
    assert torch.equal(run(sub), full[sub.cuda()])               # (3)


FEW = 3


def test_full_batch_sample0_equals_golden_run():
    """The headline batch (B = 64 4AA trajectories, T = 1000, full depth) chained to the reference DIRECTLY: sample 0 of the batch is
    the `peptide_full` golden case (same weights, frames and noise), the other 63 are unrelated synthetic trajectories.  Sample 0 of
    the B = 64 run must match what the real reference produced, and the first samples of the B = 64 run must equal a small run
    of the same samples bit for bit."""
    from lam_slide_b200.synthetic import synthetic_batch
    c = CASE_BY_NAME["peptide_full"]
    fx = load_golden("peptide_full")
    cfg, fs_sd, bb_sd, batch1, noise1, _ = case_inputs(c)
    check_inputs_match_fixture(fx, fs_sd, bb_sd, batch1, noise1)
    m = _build(cfg, fs_sd, bb_sd)
    B, T = 64, c["T"]
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    rest = synthetic_batch(cfg, B - 1, seed=77)
    batch = {k: torch.cat([batch1[k], rest[k].to(batch1[k].dtype)]) for k in batch1}
    noise = torch.cat([noise1, torch.randn(B - 1, T, L, D, generator=torch.Generator().manual_seed(78))])
    one = m.sample({k: v.clone() for k, v in batch1.items()}, noise=noise1.clone())["atom14_pos"]
    few = m.sample({k: v[:FEW].clone() for k, v in batch.items()}, noise=noise[:FEW].clone())["atom14_pos"]
    full = m.sample({k: v.clone() for k, v in batch.items()}, noise=noise.clone())["atom14_pos"]
    assert torch.isfinite(full).all()
    assert torch.equal(full[:FEW], few) and torch.equal(full[:1], one)  # the same bits from every batch size
    sl = frame_slice(fx)
    for run in (one, few[:1], full[:1]):  # B = 1 (the golden case itself), B = FEW and B = 64: all within the contract of the reference
        got = run.cpu()[:, sl].flatten(-2)  # the golden file stores the decoder output [.., 42]
        assert rmsd(got[:, -1], fx["outputs"]["atom14_pos"][:, -1]) < RMSD_TOL
        assert rmsd(got, fx["outputs"]["atom14_pos"]) < RMSD_TOL
    assert rmsd(one.cpu(), full[:1].cpu()) < RMSD_TOL / 4  # same trajectory from two batch sizes


@pytest.mark.parametrize("name", ["nba", "pedestrian"])
def test_large_batch_properties_and_golden_slice(name):
    """BASELINE.json configs[1] / [2] at their real batch size (B = 1024: second-stage.sh:12, configs/data/pedestrian.yaml:15; ragged
    agent counts for pedestrians).  The first samples of the batch are the golden case of the REAL reference; trajectories are
    independent, so (1) their slice of the B = 1024 run equals the small run bit for bit and matches the golden outputs, (2) the
    run is deterministic, (3) an arbitrary sub-batch sampled alone equals its slice."""
    from lam_slide_b200.synthetic import synthetic_batch
    c = CASE_BY_NAME[name + "_full"]
    fx = load_golden(name + "_full")
    cfg, fs_sd, bb_sd, batch_g, noise_g, _ = case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    g = torch.Generator().manual_seed(c["seeds"][3])
    L, D = cfg["first_stage"]["encoder"]["num_latents"], cfg["backbone"]["in_dim"]
    _ = torch.randn(c["B"], c["T"], L, D, generator=g)
    m.vec_in_embedding.weight.data.copy_(torch.randn(cfg["n_classes"], 256, generator=g))  # the table the golden case drew y from
    B, T, Bg = 1024, c["T"], c["B"]
    rest = synthetic_batch(cfg, B - Bg, seed=91)
    batch = {}
    for k in batch_g:
        a, b = batch_g[k], rest[k].to(batch_g[k].dtype)
        if a.dim() >= 3 and a.shape[2] != b.shape[2]:  # ragged pedestrians: pad the entity axis to the larger N_max
            n = max(a.shape[2], b.shape[2])
            pad = lambda t: torch.cat([t, t.new_zeros(t.shape[:2] + (n - t.shape[2],) + t.shape[3:])], 2) if t.shape[2] < n else t
            a, b = pad(a), pad(b)
        batch[k] = torch.cat([a, b])
    noise = torch.cat([noise_g, torch.randn(B - Bg, T, L, D, generator=torch.Generator().manual_seed(92))])
    key = cfg["main_output"]

    def run(idx):
        return m.sample({k: v[idx].clone() for k, v in batch.items()}, noise=noise[idx].clone())[key]

    full = run(torch.arange(B))
    assert torch.isfinite(full).all()
    n_g = fx["outputs"][key].shape[2]
    small = run(torch.arange(Bg))  # the golden case alone, and as the first samples of the big batch: both within the contract
    assert rmsd(small.cpu()[:, :, :n_g], fx["outputs"][key]) < RMSD_TOL
    assert rmsd(full[:Bg].cpu()[:, :, :n_g], fx["outputs"][key]) < RMSD_TOL
    assert torch.equal(full[:Bg], small)                                             # (1) the same bits from every batch size
    head = run(torch.arange(128))
    assert torch.equal(full[:128], head)
    assert torch.equal(run(torch.arange(B)), full)                                   # (2)
    sub = torch.cat([torch.tensor([1, 100, 511, 512, 777, 1023]), torch.arange(300, 422)])
    assert torch.equal(run(sub), full[sub.cuda()])                                   # (3)


def test_unbounded_logits_take_the_streaming_softmax_path():
    """The default temporal-attention kernels evaluate softmax WITHOUT a running maximum, which is only valid while the logits are
    bounded: |q.k| hd^-0.5 log2(e) <= sqrt(hd) log2(e) max|gamma_q| max|gamma_k| <= 64 (checked per block at pack time from the
    QK-RMSNorm scales, mmdit.py:132-148).  Trained checkpoints can exceed it: with one large scale entry the bound is ~68, the
    dispatcher must fall back to the streaming (online-max) kernel for those blocks, and the result must still match the oracle."""
    import lam_slide_b200 as P
    from lam_slide_b200 import _lib
    from lam_slide_b200.configs import get_config
    cfg = get_config("peptide", depth=2)
    bb = cfg["backbone"]
    bb_sd = O.init_backbone_params(bb, 41)
    big = "blocks.1.temporal_block.norm.query_norm.scale"
    bb_sd[big] = bb_sd[big].clone()
    bb_sd[big][0] = 9.5  # bound = sqrt(24) * 1.4427 * 9.5 * 1.02 = 68.5 > 64 for this block only
    net = P.LatentSIV3(depth=2, in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                       vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).cuda()
    net.load_state_dict(bb_sd, strict=True)
    B, T, L = 2, 512, 2
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, T, L, bb["in_dim"], generator=g)
    xc = torch.randn(B, T, L, bb["in_dim"], generator=g)
    mk = (torch.rand(B, T, L, generator=g) < 0.3).long()
    t = torch.rand(B, generator=g)
    with torch.no_grad():
        ref = O.backbone_forward(bb_sd, bb, x, t, xc, mk, None)
    net(x.cuda(), t.cuda(), xc.cuda(), mk.cuda())  # packs the weights
    _lib.kernel_count("attn_flash", reset=True)
    out = net(x.cuda(), t.cuda(), xc.cuda(), mk.cuda())
    assert _lib.kernel_count("attn_flash") == 1  # exactly the one block whose bound exceeds the limit
    assert torch.isfinite(out).all()
    assert max_rel(out.cpu(), ref) < VEL_TOL


@pytest.mark.parametrize("name", ["nba_full", "pedestrian_full", "peptide_small"])
def test_cuda_graph_replay_equals_eager(name):
    """include/lamslide.h: "no allocation, no host sync and no default-stream work happens inside forward / sample / encode / decode,
    so calls can be captured in a CUDA graph".  SecondStageSampler.use_cuda_graphs captures the whole sample() (encode ->
    conditioning -> every network evaluation + Euler update -> decode) once per input signature and replays it: the replayed
    results must equal the eager ones bit for bit, for new inputs of the same shape too, and survive a weight update."""
    c = CASE_BY_NAME[name]
    cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
    m = _build(cfg, fs_sd, bb_sd)
    m.hparams.sampling_kwargs["num_steps"] = c["num_steps"]
    key = cfg["main_output"]
    batches = []
    for i in range(3):
        b = {k: v.clone() for k, v in batch.items()}
        b[_pos_key(cfg)] = b[_pos_key(cfg)] * (1.0 + 0.05 * i)
        batches.append(b)
    noises = [noise, noise.flip(0).contiguous(), noise * 0.9]
    want = [m.sample({k: v.clone() for k, v in b.items()}, noise=n)[key].clone() for b, n in zip(batches, noises)]
    m.use_cuda_graphs = True
    got = [m.sample({k: v.clone() for k, v in b.items()}, noise=n)[key].clone() for b, n in zip(batches, noises)]
    assert len(m.__dict__["_graphs"]) == 1
    for g_, w_ in zip(got, want):
        assert torch.equal(g_, w_)
    assert not torch.equal(want[0], want[1])
    # a parameter update invalidates the captured graph (the packed weights it points to are replaced)
    with torch.no_grad():
        m.backbone.linear.bias.add_(0.01)
    g2 = m.sample({k: v.clone() for k, v in batches[0].items()}, noise=noises[0])[key]
    m.use_cuda_graphs = False
    w2 = m.sample({k: v.clone() for k, v in batches[0].items()}, noise=noises[0])[key]
    assert torch.equal(g2, w2) and not torch.equal(g2, want[0])
    # without given noise the replay draws fresh noise every time (graph-safe philox offsets), like torch.randn_like in the reference
    m.use_cuda_graphs = True
    a = m.sample({k: v.clone() for k, v in batches[0].items()})[key].clone()
    b2 = m.sample({k: v.clone() for k, v in batches[0].items()})[key].clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b2)


def _pos_key(cfg):
    return "atom14_pos" if cfg["main_output"] == "atom14_pos" else "pos"


def test_ode_sample_capturable_through_the_c_abi():
    """lamslide_ode_sample itself under stream capture (the raw C-ABI call, not the Python wrapper's graph cache)."""
    import lam_slide_b200 as P
    from lam_slide_b200.configs import get_config
    cfg = get_config("pedestrian", depth=2)
    bb = cfg["backbone"]
    bb_sd = O.init_backbone_params(bb, 51)
    net = P.LatentSIV3(depth=2, in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                       vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).cuda()
    net.load_state_dict(bb_sd, strict=True)
    g = torch.Generator().manual_seed(52)
    B, T, L, D = 7, 20, 2, bb["in_dim"]
    x0 = torch.randn(B, T, L, D, generator=g).cuda()
    xc = torch.randn(B, T, L, D, generator=g).cuda()
    mk = (torch.rand(B, T, L, generator=g) < 0.4).long().cuda()
    y = torch.randn(B, bb["vec_in_dim"], generator=g).cuda()
    want = net.ode_sample(x0, xc, mk, y, num_steps=10, return_states=False).clone()
    static_x = x0.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        net.ode_sample(static_x, xc, mk, y, num_steps=10, return_states=False)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = net.ode_sample(static_x, xc, mk, y, num_steps=10, return_states=False)
    for scale in (1.0, 0.5):
        static_x.copy_(x0 * scale)
        graph.replay()
        ref = net.ode_sample(x0 * scale, xc, mk, y, num_steps=10, return_states=False)
        assert torch.equal(out, ref)
    static_x.copy_(x0)
    graph.replay()
    assert torch.equal(out, want)


def test_errors_are_loud():
    import lam_slide_b200 as P
    with pytest.raises(ValueError):
        P.LatentSIV3(depth=1, in_dim=32, hidden_size=100, num_heads=16)
    net = P.LatentSIV3(depth=1, in_dim=32, hidden_size=128, num_heads=4)
    x = torch.randn(1, 4, 2, 32)
    with pytest.raises(Exception):  # CPU tensors: no fallback
        net(x, torch.zeros(1), x, torch.zeros(1, 4, 2, dtype=torch.long))
