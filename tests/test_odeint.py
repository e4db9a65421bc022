"""ODE integrators behind ``Sampler.sample_ode`` other than fixed-grid Euler: torchdiffeq's ``dopri5`` (the
reference's default and what ``configs/eval_peptide.yaml`` runs), ``bosh3``, ``adaptive_heun`` and the fixed-grid ``midpoint`` / ``rk4`` /
``heun2`` / ``heun3``.  torchdiffeq itself is absent from this image (and not pinned by the reference), so the algorithm is pinned
on what is published about it:

CPU  - the Dormand-Prince / Bogacki-Shampine nodes, stage and solution weights against scipy's independently written RK45 / RK23 tables;
     - the embedded (error) weights and the mid-point interpolant through their order conditions;
     - oracle and product host logic on closed-form ODEs: accuracy, convergence order, interpolated outputs, controller behaviour.
GPU  - the device path (C-ABI ``lamslide_lincomb_n`` / ``lamslide_rk_error_sumsq`` + host controller) against the oracle restatement on
       an analytic vector field: identical accepted / rejected step sequence, states to fp32 rounding;
     - ``Sampler.get_sample_fn("ODE", {"sampling_method": "dopri5"})`` around the CUDA ``LatentSIV3`` against the oracle solver around
       the fp32 oracle network."""
import math

import numpy as np
import pytest
import torch

from lam_slide_b200 import odeint as PO
from oracle import lamslide_oracle as O


# ------------------------------------------------------------------------------------------------ CPU: tableau pins
def test_dopri5_tableau_matches_scipy_rk45():
    from scipy.integrate._ivp import rk
    for tab in (PO.DOPRI5, type("T", (), O.RK_TABLEAUS["dopri5"])):
        assert np.allclose(np.array(tab.alpha[:5]), rk.RK45.C[1:], rtol=0, atol=1e-15)
        for i, b in enumerate(tab.beta[:5]):
            assert np.allclose(np.array(b), rk.RK45.A[i + 1][: i + 1], rtol=0, atol=1e-15)
        assert np.allclose(np.array(tab.c_sol[:6]), rk.RK45.B, rtol=0, atol=1e-15)
        assert np.allclose(np.array(tab.beta[5]), rk.RK45.B, rtol=0, atol=1e-15)  # FSAL: the last stage is the solution


def test_bosh3_tableau_matches_scipy_rk23():
    from scipy.integrate._ivp import rk
    tab = PO.BOSH3
    assert np.allclose(np.array(tab.alpha[:2]), rk.RK23.C[1:], atol=1e-15)
    assert np.allclose(np.array(tab.beta[1]), rk.RK23.A[2][:2], atol=1e-15)
    assert np.allclose(np.array(tab.c_sol[:3]), rk.RK23.B, atol=1e-15)
    assert np.allclose(np.array(tab.c_error), -rk.RK23.E, atol=1e-15)  # scipy stores (lower - higher) order


def _order_conditions(A, b, c, upto):
    """Residuals of the rooted-tree order conditions of an explicit RK method up to order ``upto`` (<= 4)."""
    A, b, c = np.array(A), np.array(b), np.array(c)
    res = {1: [b.sum() - 1]}
    res[2] = [b @ c - 1 / 2]
    res[3] = [b @ c ** 2 - 1 / 3, b @ (A @ c) - 1 / 6]
    res[4] = [b @ c ** 3 - 1 / 4, (b * c) @ (A @ c) - 1 / 8, b @ (A @ c ** 2) - 1 / 12, b @ (A @ (A @ c)) - 1 / 24]
    return [r for o in range(1, upto + 1) for r in res[o]]


def _dense(tab):
    n = len(tab.alpha) + 1
    A = np.zeros((n, n))
    for i, b in enumerate(tab.beta):
        A[i + 1, : len(b)] = b
    c = np.array([0.0] + list(tab.alpha))
    return A, c


def test_embedded_error_weights_satisfy_order_conditions():
    """c_sol is a 5th-order method, c_sol - c_error the 4th-order companion (Shampine's weights in torchdiffeq's dopri5)."""
    A, c = _dense(PO.DOPRI5)
    hi = np.array(PO.DOPRI5.c_sol)
    lo = hi - np.array(PO.DOPRI5.c_error)
    assert max(abs(r) for r in _order_conditions(A, hi, c, 4)) < 1e-14
    assert abs(hi @ c ** 4 - 1 / 5) < 1e-14  # one of the order-5 conditions
    assert max(abs(r) for r in _order_conditions(A, lo, c, 4)) < 1e-14
    assert abs(lo @ c ** 4 - 1 / 5) > 1e-4  # ... which the companion does not satisfy: it is exactly 4th order
    assert abs(sum(PO.DOPRI5.c_error)) < 1e-15
    A, c = _dense(PO.BOSH3)
    hi = np.array(PO.BOSH3.c_sol)
    lo = hi - np.array(PO.BOSH3.c_error)
    assert max(abs(r) for r in _order_conditions(A, hi, c, 3)) < 1e-14
    assert max(abs(r) for r in _order_conditions(A, lo, c, 2)) < 1e-14
    A, c = _dense(PO.ADAPTIVE_HEUN)
    assert max(abs(r) for r in _order_conditions(A, np.array(PO.ADAPTIVE_HEUN.c_sol), c, 2)) < 1e-14


def test_dopri5_midpoint_weights_are_a_fourth_order_dense_output():
    """y_mid = y0 + dt sum_i c_mid_i k_i approximates y(t0 + dt / 2): order conditions with theta = 1/2 up to order 4."""
    A, c = _dense(PO.DOPRI5)
    m = np.array(PO.DOPRI5.c_mid)
    th = 0.5
    assert abs(m.sum() - th) < 1e-12
    assert abs(m @ c - th ** 2 / 2) < 1e-12
    assert abs(m @ c ** 2 - th ** 3 / 3) < 1e-12 and abs(m @ (A @ c) - th ** 3 / 6) < 1e-12
    assert abs(m @ c ** 3 - th ** 4 / 4) < 1e-11 and abs(m @ (A @ (A @ c)) - th ** 4 / 24) < 1e-11


# ------------------------------------------------------------------------------------------------ CPU: the oracle solver on closed forms
def _exact(t):
    return 1.5 * torch.exp(-t) + 0.5 * (torch.sin(t) - torch.cos(t))  # y' = -y + sin t, y(0) = 1


# (bosh3 / adaptive_heun interpolate the outputs with a mid-point that is only first-order accurate — y0 + dt/2 k2 resp. y0 + dt/2 k1,
# torchdiffeq's _BS_C_MID / _AH_C_MID — so their values at the requested times are far less accurate than their steps)
@pytest.mark.parametrize("method,tol", [("dopri5", 2e-6), ("bosh3", 5e-4), ("adaptive_heun", 2e-5)])
def test_oracle_adaptive_solvers_on_a_closed_form(method, tol):
    t = torch.linspace(0, 5, 11, dtype=torch.float64)
    st = {}
    y = O.odeint(lambda tt, y: -y + torch.sin(tt), torch.ones(3, dtype=torch.float64), t, method=method, rtol=1e-6, atol=1e-8, stats=st)
    assert float((y[:, 0] - _exact(t)).abs().max()) < tol
    assert st["accepted"] >= 5 and st["nfe"] == 2 + (len(O.RK_TABLEAUS[method]["alpha"])) * (st["accepted"] + st["rejected"])


def test_oracle_dopri5_convergence_and_controller():
    """Tighter tolerances -> more steps and smaller error; a looser rtol takes fewer steps (the controller reacts to the tolerance)."""
    t = torch.linspace(0, 4, 9, dtype=torch.float64)
    errs, steps = [], []
    for rtol in (1e-3, 1e-5, 1e-7):
        st = {}
        y = O.odeint(lambda tt, y: -y + torch.sin(tt), torch.ones(2, dtype=torch.float64), t, method="dopri5", rtol=rtol, atol=rtol * 1e-2, stats=st)
        errs.append(float((y[:, 0] - _exact(t)).abs().max()))
        steps.append(st["accepted"])
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-7
    assert steps[0] < steps[1] < steps[2]


@pytest.mark.parametrize("method,order", [("midpoint", 2), ("heun2", 2), ("heun3", 3), ("rk4", 4)])
def test_oracle_fixed_grid_convergence_order(method, order):
    errs = []
    for n in (11, 21, 41):
        t = torch.linspace(0, 2, n, dtype=torch.float64)
        y = O.odeint(lambda tt, y: -y + torch.sin(tt), torch.ones(1, dtype=torch.float64), t, method=method)
        errs.append(float((y[-1, 0] - _exact(t[-1])).abs()))
    rate = math.log2(errs[0] / errs[1]), math.log2(errs[1] / errs[2])
    assert abs(rate[0] - order) < 0.35 and abs(rate[1] - order) < 0.35, rate


def test_oracle_ode_solve_euler_equals_ode_sample():
    """The generic solver with method "euler" reproduces the fixed-grid Euler sampler pinned by the reference goldens."""
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(2, 5, 2, 8, generator=g)
    W = torch.randn(8, 8, generator=g) * 0.3
    net = lambda x, t: torch.tanh(x @ W) * (1 + t.view(-1, 1, 1, 1))
    a = O.ode_sample(net, x0, path_type="GVP", prediction="data", num_steps=7)
    b = O.ode_solve(net, x0, path_type="GVP", prediction="data", method="euler", num_steps=7)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------ GPU: the device path
def _field(W):
    return lambda t, y: torch.tanh(y @ W) * (1.0 + 0.5 * math.sin(3.0 * float(t))) - 0.3 * y


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dopri5", "bosh3", "adaptive_heun", "rk4", "midpoint", "heun2", "heun3"])
def test_device_odeint_matches_oracle_on_an_analytic_field(method):
    """lam_slide_b200.odeint on CUDA tensors against the oracle on the CPU for a smooth vector field built from torch ops: the same
    controller decisions (accepted / rejected counts, function evaluations) and the same states to fp32 rounding."""
    g = torch.Generator().manual_seed(7)
    W = torch.randn(16, 16, generator=g) * 0.4
    y0 = torch.randn(6, 50, 16, generator=g)
    t = torch.linspace(0.001, 0.999, 12)
    so, sp = {}, {}
    ref = O.odeint(lambda tt, y: _field(W)(tt, y), y0, t, method=method, rtol=1e-4, atol=1e-6, stats=so)
    Wd = W.cuda()
    got = PO.odeint(lambda tt, y: _field(Wd)(tt, y), y0.cuda(), t, method=method, rtol=1e-4, atol=1e-6, stats=sp)
    assert got.shape == ref.shape
    if method in PO.ADAPTIVE:
        assert (sp["accepted"], sp["rejected"], sp["nfe"]) == (so["accepted"], so["rejected"], so["nfe"])
    diff = float((got.cpu() - ref).abs().max()) / float(ref.abs().max())
    assert diff < 2e-4, f"{method}: max rel diff {diff:.3e}"  # fp32 rounding; the reference-style quartic interpolant amplifies it ~30x


@pytest.mark.gpu
def test_dopri5_sampler_around_the_cuda_backbone_vs_oracle():
    """Sampler.get_sample_fn("ODE", {"sampling_method": "dopri5", ...}) — eval_peptide.yaml's sampling_kwargs with the reference's
    defaults (num_steps 50, atol 1e-6, rtol 1e-3) — around the CUDA LatentSIV3 (bf16 operands) against the oracle's dopri5 around the fp32
    oracle network.  The adaptive controller sees slightly different error estimates, so the step sequences may differ by a step; the
    solutions agree within the solver tolerance class."""
    import lam_slide_b200 as P
    from lam_slide_b200.configs import get_config
    cfg = get_config("pedestrian", depth=2)
    bb = cfg["backbone"]
    bb_sd = O.init_backbone_params(bb, 61)
    net = P.LatentSIV3(depth=2, in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                       vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).cuda()
    net.load_state_dict(bb_sd, strict=True)
    g = torch.Generator().manual_seed(62)
    B, T, L, D = 4, 20, 2, bb["in_dim"]
    x0 = torch.randn(B, T, L, D, generator=g)
    xc = torch.randn(B, T, L, D, generator=g)
    mk = torch.zeros(B, T, L, dtype=torch.int64)
    mk[:, :8] = 1
    y = torch.randn(B, bb["vec_in_dim"], generator=g)
    so = {}
    with torch.no_grad():
        ref = O.ode_solve(lambda x, t: O.backbone_forward(bb_sd, bb, x, t, xc, mk, y), x0, path_type="GVP", prediction="data",
                          method="dopri5", num_steps=50, stats=so)
    si = P.CreateTransport(path_type="GVP", prediction="data")()
    fn = P.Sampler(si).get_sample_fn("ODE", {"sampling_method": "dopri5"})
    got = fn(x0.cuda(), net, x_cond=xc.cuda(), x_cond_mask=mk.cuda(), y=y.cuda())
    assert got.shape == (50, B, T, L, D) and torch.equal(got[0].cpu(), x0)
    assert abs(fn.stats["accepted"] - so["accepted"]) <= 2 and fn.stats["nfe"] <= so["nfe"] + 12
    scale = float(ref[-1].abs().max())
    assert float((got[-1].cpu() - ref[-1]).abs().max()) < 2e-2 * scale
    assert float(((got[-1].cpu() - ref[-1]) ** 2).mean().sqrt()) < 5e-3 * scale
    mid = got[25].cpu()
    assert float((mid - ref[25]).abs().max()) < 2e-2 * float(ref[25].abs().max())  # interpolated outputs along the way
