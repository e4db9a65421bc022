"""ODE integrators behind ``Sampler.sample_ode`` other than the fused fixed-grid Euler loop: the reference hands its drift to
``torchdiffeq.odeint(fn, x, t, method=sampling_method, atol=[atol], rtol=[rtol])`` (``src/modules/transport/integrators.py:103-120``)
with ``dopri5`` as the default (``transport.py:365-372``; ``configs/eval_peptide.yaml:21-23``).  torchdiffeq is a third-party
dependency that is neither vendored nor pinned by the reference (``environment.yaml``); this module restates its published algorithms
(``torchdiffeq/_impl/{rk_common,dopri5,bosh3,adaptive_heun,fixed_grid,interp,misc}.py``, 0.2.x):

* adaptive embedded Runge-Kutta — ``dopri5`` (Dormand-Prince 5(4), Shampine's error coefficients and 4th-order mid-point
  interpolant), ``bosh3`` (Bogacki-Shampine 3(2)), ``adaptive_heun`` (Heun-Euler 2(1)) — with the same controller: initial step of
  Hairer / Norsett / Wanner (``_select_initial_step``), mixed error tolerance ``atol + rtol * max(|y0|, |y1|)`` under an RMS norm,
  accept iff ratio <= 1, next step ``dt * min(10, max(0.9 / ratio^(1/order), 0.2))`` (no shrink limit after an accepted step),
  time in float64, state in float32, outputs at the requested grid by evaluating the step's interpolating polynomial — steps are
  NOT clipped to the output times, so the last step usually reaches past ``t1``;
* fixed grid (one step per interval of ``t``) — ``midpoint``, ``rk4`` (3/8 rule), ``heun2``, ``heun3`` (``euler`` is the fused C loop).

Everything that touches the state runs on the device through the C ABI: stage combinations ``y0 + sum_j (dt beta_ij) k_j`` and the
interpolant are ``lamslide_lincomb_n`` launches; the error norm is ``lamslide_rk_error_sumsq``, a fused reduction in fp64 that never
materialises the error vector.  The step-size controller is a few scalar operations per step and runs on the host in float64, as in
torchdiffeq: one 8-byte read-back (a stream synchronisation) per attempted step — inherent to an adaptive method, whose control flow
depends on the data.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib

_f32 = np.float32


class Tableau:
    def __init__(self, alpha, beta, c_sol, c_error, c_mid, order):
        self.alpha, self.beta, self.c_sol, self.c_error, self.c_mid, self.order = alpha, beta, c_sol, c_error, c_mid, order


# torchdiffeq/_impl/dopri5.py: _DORMAND_PRINCE_SHAMPINE_TABLEAU, DPS_C_MID
DOPRI5 = Tableau(
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    beta=[
        [1 / 5],
        [3 / 40, 9 / 40],
        [44 / 45, -56 / 15, 32 / 9],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
        [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
    ],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_error=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
             11 / 84 - 649 / 6300, -1.0 / 60.0],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
    order=5,
)
# torchdiffeq/_impl/bosh3.py
BOSH3 = Tableau(
    alpha=[1 / 2, 3 / 4, 1.0],
    beta=[[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
    c_sol=[2 / 9, 1 / 3, 4 / 9, 0.0],
    c_error=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
    c_mid=[0.0, 0.5, 0.0, 0.0],
    order=3,
)
# torchdiffeq/_impl/adaptive_heun.py
ADAPTIVE_HEUN = Tableau(alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.0], order=2)

ADAPTIVE = {"dopri5": DOPRI5, "bosh3": BOSH3, "adaptive_heun": ADAPTIVE_HEUN}
FIXED = ("midpoint", "rk4", "heun2", "heun3")
METHODS = ("euler",) + FIXED + tuple(ADAPTIVE)


# ------------------------------------------------------------------------------------------------ device building blocks
def lincomb_n(out: Optional[Tensor], srcs: Sequence[Tensor], coefs: Sequence[float]) -> Tensor:
    """out = sum_j coefs[j] * srcs[j] (fp32 coefficients; ``out`` may be one of the sources)."""
    n = len(srcs)
    assert 1 <= n == len(coefs) <= 8
    if out is None:
        out = torch.empty_like(srcs[0])
    ptrs = (C.c_void_p * n)(*[s.data_ptr() for s in srcs])
    cf = (C.c_float * n)(*[float(c) for c in coefs])
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().lamslide_lincomb_n(out.data_ptr(), ptrs, cf, n, out.numel(), _lib.current_stream_ptr()))
    return out


class _ErrNorm:
    """rms( (sum_j coefs[j] srcs[j]) / (atol + rtol * max(|a|, |b|)) ) on the device, accumulated in fp64."""

    def __init__(self, device: torch.device):
        self.acc = torch.zeros(1, dtype=torch.float64, device=device)

    def __call__(self, srcs: Sequence[Tensor], coefs: Sequence[float], a: Tensor, b: Tensor, rtol: float, atol: float) -> float:
        n = len(srcs)
        ptrs = (C.c_void_p * n)(*[s.data_ptr() for s in srcs])
        cf = (C.c_float * n)(*[float(c) for c in coefs])
        self.acc.zero_()
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().lamslide_rk_error_sumsq(ptrs, cf, n, a.data_ptr(), b.data_ptr(), float(rtol), float(atol), a.numel(),
                                                           self.acc.data_ptr(), _lib.current_stream_ptr()))
        return math.sqrt(float(self.acc.item()) / a.numel())  # the one host read-back of a step


def _scaled(coefs: Sequence[float], dt: float) -> List[float]:
    """tableau coefficients and dt are cast to the state dtype before they are multiplied (rk_common.py: _runge_kutta_step)."""
    return [float(_f32(c) * _f32(dt)) for c in coefs]


# ------------------------------------------------------------------------------------------------ adaptive solvers
class AdaptiveRK:
    """torchdiffeq RKAdaptiveStepsizeODESolver (rk_common.py) on device tensors.  ``func(t: float, y) -> dy/dt`` (fp32 tensors; t has
    already been rounded to fp32, as torchdiffeq casts the time to the state dtype before it calls the user function)."""

    def __init__(self, func: Callable[[float, Tensor], Tensor], y0: Tensor, method: str, rtol: float, atol: float,
                 safety: float = 0.9, ifactor: float = 10.0, dfactor: float = 0.2, max_num_steps: int = 2 ** 31 - 1):
        self.tab = ADAPTIVE[method]
        self.func = lambda t, y: func(float(_f32(t)), y)
        self.y0 = y0
        self.rtol, self.atol = float(rtol), float(atol)
        self.safety, self.ifactor, self.dfactor, self.max_num_steps = safety, ifactor, dfactor, max_num_steps
        self.norm = _ErrNorm(y0.device)
        self.stats = {"nfe": 0, "accepted": 0, "rejected": 0}

    def _f(self, t: float, y: Tensor) -> Tensor:
        self.stats["nfe"] += 1
        return self.func(t, y)

    def _select_initial_step(self, t0: float, y0: Tensor, f0: Tensor) -> float:
        """misc.py: _select_initial_step (Hairer, Norsett, Wanner: Solving ODEs I, II.4), called with order - 1."""
        order = self.tab.order - 1
        d0 = self.norm([y0], [1.0], y0, y0, self.rtol, self.atol)
        d1 = self.norm([f0], [1.0], y0, y0, self.rtol, self.atol)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        h0 = abs(h0)
        y1 = lincomb_n(None, [y0, f0], [1.0, float(_f32(h0))])
        f1 = self._f(t0 + h0, y1)
        d2 = abs(self.norm([f1, f0], [1.0, -1.0], y0, y0, self.rtol, self.atol) / h0)
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
        return min(100 * h0, abs(h1))

    def _optimal_step_size(self, last_step: float, error_ratio: float) -> float:
        """misc.py: _optimal_step_size."""
        if error_ratio == 0:
            return last_step * self.ifactor
        dfactor = 1.0 if error_ratio < 1 else self.dfactor
        factor = min(self.ifactor, max(self.safety / error_ratio ** (1.0 / self.tab.order), dfactor))
        return last_step * factor

    def _step(self, y0: Tensor, f0: Tensor, t0: float, dt: float):
        """rk_common.py: _runge_kutta_step + _compute_error_ratio.  Returns (y1, f1, ratio, k)."""
        tab = self.tab
        t1 = t0 + dt
        k = [f0]
        yi = y0
        for alpha_i, beta_i in zip(tab.alpha, tab.beta):
            ti = t1 if alpha_i == 1.0 else t0 + float(_f32(alpha_i)) * dt
            yi = lincomb_n(None, [y0] + k, [1.0] + _scaled(beta_i, dt))
            k.append(self._f(ti, yi))
        if not (tab.c_sol[-1] == 0 and list(tab.c_sol[:-1]) == list(tab.beta[-1])):  # not FSAL: a separate solution combination
            yi = lincomb_n(None, [y0] + k, [1.0] + _scaled(tab.c_sol, dt))
        y1, f1 = yi, k[-1]
        ratio = self.norm(k, _scaled(tab.c_error, dt), y0, y1, self.rtol, self.atol)
        return y1, f1, ratio, k

    def integrate(self, t: Sequence[float]) -> List[Tensor]:
        """Solution at every time of ``t`` (increasing; float64 values of the caller's grid)."""
        t = [float(v) for v in t]
        y0 = self.y0
        f0 = self._f(t[0], y0)
        dt = self._select_initial_step(t[0], y0, f0)
        rk_t0, rk_t1 = t[0], t[0]
        interp = None  # (t0, t1, y0, y1, y_mid, f0, f1, dt) of the last accepted step
        out = [y0]
        for next_t in t[1:]:
            n_steps = 0
            while next_t > rk_t1:
                if n_steps >= self.max_num_steps:
                    raise RuntimeError(f"max_num_steps exceeded ({n_steps}>={self.max_num_steps})")
                if not (rk_t1 + dt > rk_t1):
                    raise RuntimeError(f"underflow in dt {dt}")
                t0 = rk_t1
                y1, f1, ratio, k = self._step(y0, f0, t0, dt)
                if not math.isfinite(ratio):
                    raise RuntimeError("non-finite values in the ODE state")
                if ratio <= 1:  # accept
                    y_mid = lincomb_n(None, [y0] + k, [1.0] + _scaled(self.tab.c_mid, dt))
                    interp = (t0, t0 + dt, y0, y1, y_mid, k[0], k[-1], dt)
                    rk_t0, rk_t1 = t0, t0 + dt
                    y0, f0 = y1, f1
                    self.stats["accepted"] += 1
                else:
                    self.stats["rejected"] += 1
                dt = self._optimal_step_size(dt, ratio)
                n_steps += 1
            out.append(self._interp_evaluate(interp, next_t))
        return out

    @staticmethod
    def _interp_evaluate(interp, t: float) -> Tensor:
        """interp.py: _interp_fit + _interp_evaluate — the quartic through (y0, f0), (y_mid), (y1, f1), written as weights of the five
        tensors:  y(x) = y0 + x d + x^2 c + x^3 b + x^4 a  with a, b, c, d linear in (y0, y1, y_mid, dt f0, dt f1)."""
        t0, t1, y0, y1, y_mid, f0, f1, dt = interp
        x = float(_f32((t - t0) / (t1 - t0)))
        x2, x3, x4 = x * x, x * x * x, x * x * x * x
        dtf = float(_f32(dt))
        w_y0 = 1 - 11 * x2 + 18 * x3 - 8 * x4
        w_y1 = -5 * x2 + 14 * x3 - 8 * x4
        w_ym = 16 * x2 - 32 * x3 + 16 * x4
        w_f0 = dtf * (x - 4 * x2 + 5 * x3 - 2 * x4)
        w_f1 = dtf * (x2 - 3 * x3 + 2 * x4)
        return lincomb_n(None, [y0, y1, y_mid, f0, f1], [w_y0, w_y1, w_ym, w_f0, w_f1])


# ------------------------------------------------------------------------------------------------ fixed-grid solvers
def fixed_grid(func: Callable[[float, Tensor], Tensor], y0: Tensor, t: Sequence[float], method: str) -> List[Tensor]:
    """torchdiffeq FixedGridODESolver with grid = t (fixed_grid.py, rk_common.py: rk4_alt_step_func, rk3_step_func, rk2_step_func)."""
    f = lambda tt, y: func(float(_f32(tt)), y)
    out = [y0]
    y = y0
    for a, b in zip(t[:-1], t[1:]):
        a, b = float(a), float(b)
        dt = float(_f32(b) - _f32(a))  # the grid is fp32 (th.linspace), and so is its difference
        if method == "midpoint":
            half = float(_f32(0.5) * _f32(dt))
            y_mid = lincomb_n(None, [y, f(a, y)], [1.0, half])
            y = lincomb_n(None, [y, f(a + half, y_mid)], [1.0, dt])
        elif method == "rk4":  # 3/8 rule
            k1 = f(a, y)
            k2 = f(a + dt / 3, lincomb_n(None, [y, k1], [1.0, dt / 3]))
            k3 = f(a + dt * 2 / 3, lincomb_n(None, [y, k2, k1], [1.0, dt, -dt / 3]))
            k4 = f(b, lincomb_n(None, [y, k1, k2, k3], [1.0, dt, -dt, dt]))
            y = lincomb_n(None, [y, k1, k2, k3, k4], [1.0, dt * 0.125, dt * 0.375, dt * 0.375, dt * 0.125])
        elif method == "heun2":
            k1 = f(a, y)
            k2 = f(b, lincomb_n(None, [y, k1], [1.0, dt]))
            y = lincomb_n(None, [y, k1, k2], [1.0, dt * 0.5, dt * 0.5])
        elif method == "heun3":
            k1 = f(a, y)
            k2 = f(a + dt / 3, lincomb_n(None, [y, k1], [1.0, dt / 3]))
            k3 = f(a + dt * 2 / 3, lincomb_n(None, [y, k2], [1.0, dt * 2 / 3]))
            y = lincomb_n(None, [y, k1, k3], [1.0, dt * 0.25, dt * 0.75])
        else:
            raise NotImplementedError(method)
        out.append(y)
    return out


def odeint(func: Callable[[float, Tensor], Tensor], y0: Tensor, t: Tensor, *, method: str = "dopri5", rtol: float = 1e-3,
           atol: float = 1e-6, stats: Optional[dict] = None) -> Tensor:
    """``torchdiffeq.odeint(func, y0, t, method=..., rtol=..., atol=...)`` for a CUDA fp32 state: stacked solution ``[len(t), *y0.shape]``."""
    _lib.require_cuda(y0)
    y0 = y0.to(torch.float32).contiguous()
    grid = [float(v) for v in t.detach().cpu().to(torch.float64)]
    if any(b <= a for a, b in zip(grid[:-1], grid[1:])):
        raise ValueError("t must be strictly increasing")  # torchdiffeq: _assert_increasing (the samplers integrate forward in time)
    if method in ADAPTIVE:
        solver = AdaptiveRK(func, y0, method, rtol, atol)
        sol = solver.integrate(grid)
        if stats is not None:
            stats.update(solver.stats)
    elif method in FIXED:
        sol = fixed_grid(func, y0, grid, method)
    else:
        raise NotImplementedError(f"ODE method '{method}' (have: {', '.join(METHODS)})")
    return torch.stack(sol)
