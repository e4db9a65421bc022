"""SiT stochastic-interpolant transport — the sampling half of ``src/modules/transport`` behind the same API.

``CreateTransport(path_type, prediction)()`` -> ``Transport``; ``Sampler(transport).get_sample_fn("ODE",
{"sampling_method": "euler", "num_steps": n})`` -> ``fn(init, model, **model_kwargs)`` returning the stacked states
``[num_steps, *init.shape]`` exactly like the reference (transport/__init__.py:7-79, transport.py:39-101, 229-503,
integrators.py:84-120).  When ``model`` is (a bound method of an object whose ``backbone`` is) a
``lam_slide_b200.LatentSIV3`` the whole loop — network, drift, Euler update — runs inside one C-ABI call
(``lamslide_ode_sample``); for any other callable the loop runs here and each step's drift + update is one
``lamslide_euler_step`` launch.  Out of scope: SDE samplers, likelihood, adaptive dopri5, the
VP path and the training losses — requesting them raises ``NotImplementedError``.
"""
from __future__ import annotations

import enum
from typing import Any, Callable, Dict

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
        """transport.py:69-101 (ODE branch)."""
        if sde:
            raise NotImplementedError("SDE sampling is out of scope of the B200 hot path")
        t0, t1 = 0, 1
        eps = train_eps if not eval else sample_eps
        if self.path_type == PathType.VP:
            t1 = 1 - eps
        elif self.model_type != ModelType.VELOCITY:
            t0, t1 = eps, 1 - eps
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


class Sampler:
    """transport.py:229-503 (ODE / Euler only)."""

    def __init__(self, transport: Transport):
        self.transport = transport

    def get_sample_fn(self, sampling_method: str = "ODE", sampling_kwargs: Dict[str, Any] = {}):
        if sampling_method != "ODE":
            raise NotImplementedError("only the ODE sampler is implemented on the B200 hot path")
        kw = {"sampling_method": "dopri5", "num_steps": 50, "atol": 1e-6, "rtol": 1e-3, "reverse": False}
        kw.update(sampling_kwargs)
        return self.sample_ode(**kw)

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False):
        if sampling_method != "euler":
            raise NotImplementedError(f"ODE method '{sampling_method}' is not implemented (fixed-grid 'euler' only)")
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
            if bb is not None:
                return bb.ode_sample(init, model_kwargs["x_cond"], model_kwargs["x_cond_mask"], model_kwargs.get("y"),
                                     path_type=path, prediction=pred, num_steps=num_steps)
            # generic callable: Python loop, one fused drift+Euler launch per step
            _lib.require_cuda(init)
            lib = _lib.load()
            grid = torch.linspace(t0, t1, num_steps)  # fp32 on the host, as integrators.py:98
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

        return _sample
