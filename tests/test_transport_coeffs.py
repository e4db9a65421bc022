"""CPU: the host-side (fp64) coefficient algebra of lam_slide_b200/transport.py — drift, score and diffusion of the Linear / GVP plans
written as linear maps of (network output, state) — against the oracle's tensor restatement of transport.py / path.py."""
import pytest
import torch

from lam_slide_b200.transport import CreateTransport, diffusion_coeff, drift_coeffs, score_coeffs
from oracle import lamslide_oracle as O


@pytest.mark.parametrize("path", ["GVP", "Linear"])
def test_linear_coefficients_match_oracle(path):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 4, generator=g).double()
    m = torch.randn(2, 3, 4, generator=g).double()
    for pred in ("data", "velocity", "noise", "score"):
        for t in (0.013, 0.4, 0.93):
            tt = torch.full((2,), t, dtype=torch.float64)
            cm, cx = drift_coeffs(path, pred, t)
            sm, sx = score_coeffs(path, pred, t)
            d, s = O.drift(path, pred, x, tt, m), O.score(path, pred, x, tt, m)
            assert float((cm * m + cx * x - d).abs().max()) < 1e-9 * max(1.0, float(d.abs().max()))
            assert float((sm * m + sx * x - s).abs().max()) < 1e-9 * max(1.0, float(s.abs().max()))
    for form in ("constant", "SBDM", "sigma", "linear", "decreasing", "inccreasing-decreasing"):
        for t in (0.013, 0.5):
            D = O.diffusion(path, x, torch.full((2,), t, dtype=torch.float64), form, 0.7)
            D = float(D.flatten()[0]) if torch.is_tensor(D) else D
            assert abs(D - diffusion_coeff(path, t, form, 0.7)) < 1e-12


def test_sde_interval_matches_oracle():
    """Transport.check_interval with sde=True (transport.py:69-101)."""
    for path, pred in (("GVP", "data"), ("Linear", "velocity"), ("Linear", "data")):
        tr = CreateTransport(path_type=path, prediction=pred)()
        for form in ("SBDM", "linear"):
            for size in (0.0, 0.04):
                got = tr.check_interval(tr.train_eps, tr.sample_eps, diffusion_form=form, sde=True, eval=True, last_step_size=size)
                want = O.sde_interval(path, pred, form, size)
                assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
