"""SiT stochastic-interpolant transport — the sampling half of ``src/modules/transport`` behind the same API.

``CreateTransport(path_type, prediction)()`` -> ``Transport``; ``Sampler(transport).get_sample_fn("ODE",
{"sampling_method": "euler", "num_steps": n})`` -> ``fn(init, model, **model_kwargs)`` returning the stacked states
``[num_steps, *init.shape]`` exactly like the reference (transport/__init__.py:7-79, transport.py:39-101, 229-503,
integrators.py:84-120).  When ``model`` is (a bound method of an object whose ``backbone`` is) a
``lam_slide_b200.LatentSIV3`` the whole loop — network, drift, Euler update — runs inside one C-ABI call
(``lamslide_ode_sample``); for any other callable the loop runs here and each step's drift + update is one
``lamslide_euler_step`` launch.  ``get_sample_fn("SDE", {...})`` -> ``sample_sde`` (transport.py:301-363; integrators.py:7-78):
Euler-Maruyama / Heun with every diffusion form and last-step variant on the Linear and GVP plans; the network runs through
the same C-ABI forward, each update is one or two ``lamslide_lincomb3`` launches with coefficients evaluated here in fp64 (all of
the reference's update rules are linear in state, network output and noise with time-only coefficients).
``get_sample_fn("ODE", {...})`` with any other torchdiffeq method the reference can name — ``dopri5`` (its default; what
``configs/eval_peptide.yaml`` runs), ``bosh3``, ``adaptive_heun``, ``midpoint``, ``rk4``, ``heun2``, ``heun3`` — goes through
``lam_slide_b200/odeint.py`` (torchdiffeq's adaptive Runge-Kutta controller restated on device tensors).
Out of scope: likelihood, the VP path, reverse time and the training losses — requesting them raises
``NotImplementedError``.
"""
from __future__ import annotations

import enum
import math
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .backbone import PATH_TYPES, PREDICTIONS, LatentSIV3


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()
    DATA = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


_PRED_NAME = {ModelType.VELOCITY: "velocity", ModelType.DATA: "data", ModelType.NOISE: "noise", ModelType.SCORE: "score"}
_PATH_NAME = {PathType.LINEAR: "Linear", PathType.GVP: "GVP"}


class Transport:
    def __init__(self, *, model_type: ModelType, path_type: PathType, train_eps: float, sample_eps: float, loss_type=None):
        self.model_type, self.path_type = model_type, path_type
        self.train_eps, self.sample_eps, self.loss_type = train_eps, sample_eps, loss_type

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """transport.py:69-101."""
        t0, t1 = 0, 1
        eps = train_eps if not eval else sample_eps
        if self.path_type == PathType.VP:
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        elif self.model_type != ModelType.VELOCITY or sde:
            t0 = eps if (diffusion_form == "SBDM" and sde) or self.model_type != ModelType.VELOCITY else 0
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1


class CreateTransport:
    """transport/__init__.py:7-79."""

    def __init__(self, path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None):
        self.path_type, self.prediction, self.loss_weight = path_type, prediction, loss_weight
        self.train_eps, self.sample_eps = train_eps, sample_eps

    def __call__(self) -> Transport:
        model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE, "data": ModelType.DATA}.get(self.prediction, ModelType.VELOCITY)
        path_type = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[self.path_type]
        if path_type == PathType.VP:
            train_eps = 1e-5 if self.train_eps is None else self.train_eps
            sample_eps = 1e-3 if self.sample_eps is None else self.sample_eps
        elif model_type != ModelType.VELOCITY:
            train_eps = 1e-3 if self.train_eps is None else self.train_eps
            sample_eps = 1e-3 if self.sample_eps is None else self.sample_eps
        else:
            train_eps, sample_eps = 0, 0
        return Transport(model_type=model_type, path_type=path_type, train_eps=train_eps, sample_eps=sample_eps,
                         loss_type=self.loss_weight)


def _find_backbone(model: Callable):
    if isinstance(model, LatentSIV3):
        return model
    owner = getattr(model, "__self__", None)
    bb = getattr(owner, "backbone", None) if owner is not None else None
    return bb if isinstance(bb, LatentSIV3) else None


# ---- time-only coefficients of the plans (path.py:21-47 ICPlan, :188-206 GVPCPlan), evaluated on the host in fp64 ----------------
def _plan(path: str, t: float) -> Tuple[float, float, float, float, float]:
    """(alpha, d_alpha, sigma, d_sigma, d_alpha / alpha)."""
    if path == "GVP":
        a = t * math.pi / 2
        ratio = math.pi / (2 * math.tan(a)) if t != 0 else math.inf  # t = 0 only occurs for velocity models, which never use it
        return math.sin(a), math.pi / 2 * math.cos(a), math.cos(a), -math.pi / 2 * math.sin(a), ratio
    return t, 1.0, 1.0 - t, -1.0, (1.0 / t if t != 0 else math.inf)


def drift_coeffs(path: str, pred: str, t: float) -> Tuple[float, float]:
    """Transport.get_drift (transport.py:158-202) as ``v = cm * m + cx * x``."""
    if pred == "velocity":
        return 1.0, 0.0
    alpha, _, sigma, d_sigma, ratio = _plan(path, t)
    var = ratio * sigma ** 2 - sigma * d_sigma      # drift_var of compute_drift (path.py:39-47); -drift_mean = ratio * x
    sm, sx = score_coeffs(path, pred, t)
    return var * sm, ratio + var * sx


def score_coeffs(path: str, pred: str, t: float) -> Tuple[float, float]:
    """Transport.get_score (transport.py:204-226, path.py:73-95) as ``score = sm * m + sx * x``."""
    alpha, d_alpha, sigma, d_sigma, _ = _plan(path, t)
    if pred == "noise":
        return -1.0 / sigma, 0.0
    if pred == "score":
        return 1.0, 0.0
    if pred == "velocity":
        rar = alpha / d_alpha
        var = sigma ** 2 - rar * d_sigma * sigma
        return rar / var, -1.0 / var
    return alpha / sigma ** 2, -1.0 / sigma ** 2  # data


def diffusion_coeff(path: str, t: float, form: str, norm: float) -> float:
    """ICPlan.compute_diffusion (path.py:49-71)."""
    _, _, sigma, d_sigma, ratio = _plan(path, t) if form in ("SBDM", "sigma") else (0, 0, 0, 0, 0)
    if form == "constant":
        return norm
    if form == "SBDM":
        return norm * (ratio * sigma ** 2 - sigma * d_sigma)
    if form == "sigma":
        return norm * sigma
    if form == "linear":
        return norm * (1 - t)
    if form == "decreasing":
        return 0.25 * (norm * math.cos(math.pi * t) + 1) ** 2
    if form == "inccreasing-decreasing":  # (sic) the reference's key
        return norm * math.sin(math.pi * t) ** 2
    raise NotImplementedError(f"Diffusion form {form} not implemented")


def _lincomb(out: Tensor, x: Tensor, m: Optional[Tensor], w: Optional[Tensor], px: float, pm: float, pw: float) -> Tensor:
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().lamslide_lincomb3(out.data_ptr(), x.data_ptr(), 0 if m is None else m.data_ptr(),
                                                 0 if w is None else w.data_ptr(), px, pm, pw, x.numel(), _lib.current_stream_ptr()))
    return out


class Sampler:
    """transport.py:229-503: fixed-grid Euler ODE and the Euler-Maruyama / Heun SDE samplers.  ``noise_fn(shape, device)``
    (optional) replaces the ``th.randn`` of the SDE steps (integrators.py:31,41) for reproducible tests."""

    def __init__(self, transport: Transport, noise_fn: Optional[Callable] = None):
        self.transport = transport
        self.noise_fn = noise_fn

    def get_sample_fn(self, sampling_method: str = "ODE", sampling_kwargs: Dict[str, Any] = {}):
        if sampling_method == "SDE":
            kw = {"sampling_method": "Euler", "diffusion_form": "linear", "diffusion_norm": 1.0, "last_step": "Mean",
                  "last_step_size": 0.04, "num_steps": 250}  # transport.py:480-487
            kw.update(sampling_kwargs)
            return self.sample_sde(**kw)
        if sampling_method != "ODE":
            raise NotImplementedError(f"sampling method '{sampling_method}'")
        kw = {"sampling_method": "dopri5", "num_steps": 50, "atol": 1e-6, "rtol": 1e-3, "reverse": False}
        kw.update(sampling_kwargs)
        return self.sample_ode(**kw)

    def sample_sde(self, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean", last_step_size=0.04,
                   num_steps=250):
        """transport.py:301-363.  Returns ``fn(init, model, **model_kwargs) -> list of num_steps states`` (the reference's list)."""
        tr = self.transport
        if tr.path_type not in _PATH_NAME:
            raise NotImplementedError("VP path is not implemented")
        if sampling_method not in ("Euler", "Heun"):
            raise NotImplementedError("Smapler type not implemented.")  # (sic) integrators.py:63
        if last_step not in (None, "Mean", "Tweedie", "Euler"):
            raise NotImplementedError()
        path, pred = _PATH_NAME[tr.path_type], _PRED_NAME[tr.model_type]
        if last_step is None:
            last_step_size = 0.0
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, diffusion_form=diffusion_form, sde=True, eval=True, reverse=False,
                                   last_step_size=last_step_size)
        assert t0 < t1, "SDE sampler has to be in forward time"
        grid = torch.linspace(t0, t1, num_steps)  # fp32 on the host, as integrators.py:24
        dt = float(grid[1] - grid[0])
        noise_fn = self.noise_fn or (lambda shape, device: torch.randn(shape, device=device, dtype=torch.float32))

        def sde_ab(t: float) -> Tuple[float, float, float]:
            """sde_drift = a * m + b * x (transport.py:252-264) and the diffusion D at time t."""
            cm, cx = drift_coeffs(path, pred, t)
            sm, sx = score_coeffs(path, pred, t)
            D = diffusion_coeff(path, t, diffusion_form, diffusion_norm)
            return cm + D * sm, cx + D * sx, D

        @torch.no_grad()
        def _sample(init: Tensor, model: Callable, **model_kwargs) -> List[Tensor]:
            _lib.require_cuda(init)
            x = init.to(torch.float32).contiguous()
            B = x.shape[0]

            def net(xx: Tensor, t: float) -> Tensor:
                tv = torch.full((B,), t, device=xx.device, dtype=torch.float32)
                m = model(xx, tv, **model_kwargs).to(torch.float32).contiguous()
                assert m.shape == xx.shape
                return m

            xs: List[Tensor] = []
            for i in range(num_steps - 1):
                t = float(grid[i])
                w = noise_fn(tuple(x.shape), x.device).to(x.device, torch.float32).contiguous()
                a, b, D = sde_ab(t)
                c = math.sqrt(2 * D) * math.sqrt(dt)
                if sampling_method == "Euler":  # integrators.py:29-38
                    m = net(x, t)
                    x = _lincomb(torch.empty_like(x), x, m, w, 1.0 + dt * b, dt * a, c)
                else:  # Heun, integrators.py:40-52
                    xhat = _lincomb(torch.empty_like(x), x, None, w, 1.0, 0.0, c)
                    m1 = net(xhat, t)
                    xp = _lincomb(torch.empty_like(x), xhat, m1, None, 1.0 + dt * b, dt * a, 0.0)
                    a2, b2, _ = sde_ab(t + dt)
                    m2 = net(xp, t + dt)
                    tmp = _lincomb(torch.empty_like(x), xhat, m1, xp, 1.0 + 0.5 * dt * b, 0.5 * dt * a, 0.5 * dt * b2)
                    x = _lincomb(tmp, tmp, m2, None, 1.0, 0.5 * dt * a2, 0.0)
                xs.append(x)
            t_end = float(t1)
            if last_step is None:
                xl = xs[-1]
            elif last_step == "Mean":  # transport.py:276-280
                a, b, _ = sde_ab(t_end)
                xl = _lincomb(torch.empty_like(x), xs[-1], net(xs[-1], t_end), None, 1.0 + last_step_size * b, last_step_size * a, 0.0)
            elif last_step == "Tweedie":  # transport.py:281-289
                alpha, _, sigma, _, _ = _plan(path, t_end)
                sm, sx = score_coeffs(path, pred, t_end)
                xl = _lincomb(torch.empty_like(x), xs[-1], net(xs[-1], t_end), None, 1.0 / alpha + sigma ** 2 / alpha * sx,
                              sigma ** 2 / alpha * sm, 0.0)
            else:  # "Euler", transport.py:290-294
                cm, cx = drift_coeffs(path, pred, t_end)
                xl = _lincomb(torch.empty_like(x), xs[-1], net(xs[-1], t_end), None, 1.0 + last_step_size * cx, last_step_size * cm, 0.0)
            xs.append(xl)
            assert len(xs) == num_steps, "Samples does not match the number of steps"
            return xs

        return _sample

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False):
        """transport.py:365-411 + integrators.py:84-120.  ``euler`` with a ``LatentSIV3`` model is one fused C-ABI call; every other
        torchdiffeq method the reference can name here (``dopri5`` — its default —, ``bosh3``, ``adaptive_heun``, ``midpoint``, ``rk4``,
        ``heun2``, ``heun3``) goes through ``lam_slide_b200.odeint``.  ``fn.stats`` holds the function-evaluation / step counts of the
        last call of an adaptive method."""
        from . import odeint as _ode
        if sampling_method not in _ode.METHODS:
            raise NotImplementedError(f"ODE method '{sampling_method}' is not implemented (have: {', '.join(_ode.METHODS)})")
        if reverse:
            raise NotImplementedError("reverse-time sampling is not implemented")
        tr = self.transport
        if tr.path_type not in _PATH_NAME:
            raise NotImplementedError("VP path is not implemented")
        path, pred = _PATH_NAME[tr.path_type], _PRED_NAME[tr.model_type]
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, sde=False, eval=True, reverse=False, last_step_size=0.0)
        if (float(t0), float(t1)) not in ((0.0, 1.0), (1e-3, 1 - 1e-3)):
            raise NotImplementedError("custom sample_eps is not implemented")

        @torch.no_grad()
        def _sample(init: Tensor, model: Callable, **model_kwargs) -> Tensor:
            bb = _find_backbone(model)
            if bb is not None and sampling_method == "euler":
                return bb.ode_sample(init, model_kwargs["x_cond"], model_kwargs["x_cond_mask"], model_kwargs.get("y"),
                                     path_type=path, prediction=pred, num_steps=num_steps)
            _lib.require_cuda(init)
            lib = _lib.load()
            grid = torch.linspace(t0, t1, num_steps)  # fp32 on the host, as integrators.py:98
            if sampling_method != "euler":
                B = init.shape[0]

                def drift(t: float, x: Tensor) -> Tensor:
                    """Transport.get_drift (transport.py:158-202) around one network evaluation: v = cm(t) m + cx(t) x."""
                    tv = torch.full((B,), t, device=x.device, dtype=torch.float32)
                    m = model(x, tv, **model_kwargs).to(torch.float32).contiguous()
                    assert m.shape == x.shape, "Output shape from ODE solver must match input shape"  # transport.py:197-199
                    cm, cx = drift_coeffs(path, pred, t)
                    if cx == 0.0 and cm == 1.0:
                        return m
                    return _lincomb(torch.empty_like(x), x, m, None, cx, cm, 0.0)

                stats: Dict[str, Any] = {}
                out = _ode.odeint(drift, init, grid, method=sampling_method, rtol=rtol, atol=atol, stats=stats)
                _sample.stats = stats
                return out
            # fixed-grid Euler around a generic callable: Python loop, one fused drift + update launch per step
            x = init.to(torch.float32).contiguous().clone()
            states = [x.clone()]
            for i in range(num_steps - 1):
                tv = torch.full((x.shape[0],), float(grid[i]), device=x.device, dtype=torch.float32)
                m = model(x, tv, **model_kwargs).to(torch.float32).contiguous()
                assert m.shape == x.shape, "Output shape from ODE solver must match input shape"  # transport.py:197-199
                with torch.cuda.device(x.device):
                    _lib.check(lib.lamslide_euler_step(x.data_ptr(), m.data_ptr(), PATH_TYPES[path], PREDICTIONS[pred],
                                                       float(grid[i]), float(grid[i + 1]), 0, x.numel(),
                                                       _lib.current_stream_ptr()))
                states.append(x.clone())
            return torch.stack(states)

        _sample.stats = {}
        return _sample
