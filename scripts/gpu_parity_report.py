"""Developer aid: the MEASURED parity numbers behind the tolerances of tests/test_gpu_parity.py, per golden case (prints a table).
Usage (GPU box): python scripts/gpu_parity_report.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lamslide_oracle as O  # noqa: E402
from tests.helpers import CASE_BY_NAME, case_inputs, frame_slice, load_golden, max_rel, mean_rel, rmsd  # noqa: E402
import lam_slide_b200 as P  # noqa: E402


def main():
    names = ["peptide_small", "md17_small", "nba_full", "pedestrian_full", "peptide_linear_velocity", "md17_full", "peptide_full",
             "peptide_steps20", "peptide_steps50"]
    print(f"{'case':26s} {'latents':>9s} {'vel max':>9s} {'vel mean':>9s} {'final lat':>9s} {'rmsd last':>10s} {'rmsd all':>9s}")
    worst = [0.0] * 6
    for name in names:
        fx, c = load_golden(name), CASE_BY_NAME[name]
        cfg, fs_sd, bb_sd, batch, noise, y = case_inputs(c)
        m = P.SecondStageSampler(cfg).cuda()
        m.first_stage_model.backbone.load_state_dict(fs_sd, strict=True)
        m.backbone.load_state_dict(bb_sd, strict=True)
        sl, B, T = frame_slice(fx), c["B"], c["T"]
        cb = {k: v.cuda() for k, v in batch.items()}
        latents = m.encode(cb)
        e_lat = max_rel(latents.cpu()[:, sl], fx["latents"])
        x_cond, x_mask = m.setup_conditioning(latents)
        yy = None if y is None else y.cuda()
        states, vel = m.backbone.ode_sample(noise.cuda(), x_cond, x_mask, yy, path_type=cfg["path_type"], prediction=cfg["prediction"],
                                            num_steps=c["num_steps"], return_velocities=True)
        vel = vel.cpu()[fx["velocity_steps"]][:, :, sl]
        e_vmax = max(max_rel(vel[i], fx["velocities"][i]) for i in range(vel.shape[0]))
        e_vmean = max(mean_rel(vel[i], fx["velocities"][i]) for i in range(vel.shape[0]))
        e_fin = max_rel(states[-1].cpu()[:, sl], fx["final_latents"])
        out = m.first_stage_model.decode(states[-1].flatten(0, 1), cb["entities"].flatten(0, 1))
        main_out = cfg["main_output"]
        got = out[main_out].unflatten(0, (B, T)).cpu()[:, sl]
        r_last, r_all = rmsd(got[:, -1], fx["outputs"][main_out][:, -1]), rmsd(got, fx["outputs"][main_out])
        row = [e_lat, e_vmax, e_vmean, e_fin, r_last, r_all]
        worst = [max(a, float(b)) for a, b in zip(worst, row)]
        print(f"{name:26s} " + " ".join(f"{float(v):9.2e}" for v in row))
        del m
        torch.cuda.empty_cache()
    print(f"{'worst':26s} " + " ".join(f"{v:9.2e}" for v in worst))
    print("tolerances: latents (first stage) 1e-4, velocity max 1e-2 / mean 5e-3, final-frame RMSD 1e-3")


if __name__ == "__main__":
    main()
