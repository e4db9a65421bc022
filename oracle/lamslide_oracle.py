"""TEST INFRASTRUCTURE — CPU oracle for the LaM-SLidE sampling hot path.  NOT part of the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package (``lam_slide_b200``) never does and has no CPU fallback.

It is a functional (state-dict driven) fp32/fp64 restatement of the reference's floating-point algorithm,
written against plain torch CPU ops (matmul / exp / erf …; no nn.Module, no SDPA, no reference code).
Every function cites the reference lines it follows.  Parity status: **pinned** — this restatement is
checked (``tests/test_oracle_vs_reference.py``, run in the dev container) against the reference's own
modules imported from ``/root/reference`` and (``tests/test_oracle_golden.py``, run everywhere) against the
golden vectors in ``tests/golden/`` that ``oracle/make_golden.py`` generated from those reference modules.
The reference itself ships no tests / golden vectors for this path.

Third-party arithmetic that is not under /root/reference: ``torchdiffeq.odeint(method="euler")``
(call site ``src/modules/transport/integrators.py:119``; version un-pinned by ``environment.yaml``).  Its
published fixed-grid semantics are restated in ``ode_sample``: y_{i+1} = y_i + (t_{i+1}-t_i)·f(t_i, y_i) on
``t = linspace(t0, t1, num_steps)``, output at every grid point.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------------------
def gelu_erf(x: Tensor) -> Tensor:
    """mmdit.py:11-18 / torch_modules.py:36-50 — exact erf GELU."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def silu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def layer_norm(x: Tensor, eps: float, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None) -> Tensor:
    """nn.LayerNorm semantics (biased variance) — latent_si_v31.py:30,118,174; torch_modules.py:112-113."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mu) * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * weight + bias
    return y


def rms_norm(x: Tensor, scale: Tensor) -> Tensor:
    """mmdit.py:132-136 / torch_modules.py:89-93 — x·rsqrt(mean(x²)+1e-6)·scale over the head dim."""
    return x * torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + 1e-6) * scale


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


# When True the attention goes through the same library call the reference uses (F.scaled_dot_product_attention) instead
# of the explicit softmax below.  Only bench.py's CPU-baseline / --impl reference legs set it, so that the reported CPU
# number is representative of the reference's own CPU path (the explicit form materialises S x S scores and is ~2x slower).
USE_SDPA = False


def softmax_attention(q: Tensor, k: Tensor, v: Tensor, scale: float, key_mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q kᵀ · scale [+ −inf on masked keys]) v — what F.scaled_dot_product_attention computes at
    mmdit.py:51 (no mask, scale = hd^-0.5) and torch_modules.py:184,251 (bool key mask, True = keep).
    q: [..., Sq, d], k/v: [..., Sk, d], key_mask: broadcastable to [..., 1, Sk]."""
    if USE_SDPA:
        return torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=key_mask, scale=scale)
    s = (q @ k.transpose(-1, -2)) * scale
    if key_mask is not None:
        s = s.masked_fill(~key_mask, float("-inf"))
    s = s - s.amax(dim=-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    return p @ v


def timestep_embedding(t: Tensor, dim: int = 256, max_period: float = 10000.0, time_factor: float = 1000.0) -> Tensor:
    """mmdit.py:93-113 — cat[cos, sin](1000·t·exp(−ln(1e4)·i/half)); frequencies built in fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = (time_factor * t)[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(t.dtype)


def rope_tables(S: int, hd: int, theta: float, device=None) -> Tuple[Tensor, Tensor]:
    """mmdit.py:75-82 — angle = p·theta^(−2i/hd) in fp64, cos/sin cast to fp32. Returns [S, hd/2] each."""
    scale = torch.arange(0, hd, 2, dtype=torch.float64, device=device) / hd
    omega = 1.0 / (float(theta) ** scale)
    ang = torch.arange(S, dtype=torch.float64, device=device)[:, None] * omega[None]
    return torch.cos(ang).float(), torch.sin(ang).float()


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """mmdit.py:85-90 — interleaved pairs: out[2i] = cos·x[2i] − sin·x[2i+1]; out[2i+1] = sin·x[2i] + cos·x[2i+1].
    x: [..., S, hd]; cos/sin: [S, hd/2]."""
    xe, xo = x[..., 0::2], x[..., 1::2]
    cos, sin = cos.to(x.dtype), sin.to(x.dtype)
    out = torch.stack([cos * xe - sin * xo, sin * xe + cos * xo], dim=-1)
    return out.flatten(-2)


# --------------------------------------------------------------------------------------------------
# second stage: LatentSIV3
# --------------------------------------------------------------------------------------------------
def _parallel_block(sd: SD, pfx: str, u: Tensor, H: int, heads: int, cos: Tensor, sin: Tensor) -> Tensor:
    """ParallelMLPAttentionV2.forward — mmdit.py:240-249.  u: [S_batch, S, H]."""
    hd = H // heads
    z = linear(u, sd[pfx + "linear1.weight"], sd[pfx + "linear1.bias"])
    qkv, mlp = z[..., : 3 * H], z[..., 3 * H:]
    Sb, S, _ = u.shape
    qkv = qkv.reshape(Sb, S, 3, heads, hd).permute(2, 0, 3, 1, 4)  # "B L (K H D) -> K B H L D"
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rms_norm(q, sd[pfx + "norm.query_norm.scale"])
    k = rms_norm(k, sd[pfx + "norm.key_norm.scale"])
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    a = softmax_attention(q, k, v, hd ** -0.5)  # [Sb, heads, S, hd]
    a = a.permute(0, 2, 1, 3).reshape(Sb, S, H)  # "B H L D -> B L (H D)"
    return linear(torch.cat([a, gelu_erf(mlp)], dim=-1), sd[pfx + "linear2.weight"], sd[pfx + "linear2.bias"])


def backbone_vec(sd: SD, cfg: dict, t: Tensor, y: Optional[Tensor]) -> Tensor:
    """latent_si_v31.py:176-178 — vec = time_in(timestep_embedding(t)) [+ vec_in(y)]; MLPEmbedder mmdit.py:116-124."""
    e = timestep_embedding(t, 256)
    vec = linear(silu(linear(e, sd["time_in.in_layer.weight"], sd["time_in.in_layer.bias"])),
                 sd["time_in.out_layer.weight"], sd["time_in.out_layer.bias"])
    if y is not None:
        vec = vec + linear(silu(linear(y, sd["vec_in.in_layer.weight"], sd["vec_in.in_layer.bias"])),
                           sd["vec_in.out_layer.weight"], sd["vec_in.out_layer.bias"])
    return vec


def backbone_forward(sd: SD, cfg: dict, x: Tensor, t: Tensor, x_cond: Tensor, x_cond_mask: Tensor,
                     y: Optional[Tensor] = None, trace: Optional[dict] = None) -> Tensor:
    """LatentSIV3.forward — latent_si_v31.py:168-188 with LatentSIV3Layer.forward :45-63."""
    B, T, L, _ = x.shape
    H, heads, depth = cfg["hidden_size"], cfg["num_heads"], cfg["depth"]
    hd = H // heads
    h = (linear(x, sd["x_in.weight"], sd["x_in.bias"])
         + linear(x_cond, sd["cond_to_emb.weight"], sd["cond_to_emb.bias"])
         + sd["mask_to_emb.weight"][x_cond_mask])
    if cfg.get("normalize", False):
        h = layer_norm(h, 1e-5)  # F.layer_norm default eps (:174)
    vec = backbone_vec(sd, cfg, t, y)
    svec = silu(vec)
    cs_s, sn_s = rope_tables(L, hd, cfg.get("theta", 10_000), x.device)
    cs_t, sn_t = rope_tables(T, hd, cfg.get("theta", 10_000), x.device)
    cs_s, sn_s, cs_t, sn_t = (a.to(x.dtype) for a in (cs_s, sn_s, cs_t, sn_t))
    if trace is not None:
        trace["h0"] = h.clone()
        trace["vec"] = vec.clone()
    for i in range(depth):
        p = f"blocks.{i}."
        mod = linear(svec, sd[p + "modulation.lin.weight"], sd[p + "modulation.lin.bias"])
        s1, c1, g1, s2, c2, g2 = (m[:, None, None, :] for m in mod.chunk(6, dim=-1))  # shift, scale, gate ×2
        u = layer_norm(h, 1e-6) * (1 + c1) + s1
        o = _parallel_block(sd, p + "spatial_block.", u.reshape(B * T, L, H), H, heads, cs_s, sn_s)
        h = h + g1 * o.reshape(B, T, L, H)
        u = layer_norm(h, 1e-6) * (1 + c2) + s2
        u = u.permute(0, 2, 1, 3).reshape(B * L, T, H)  # "B T L D -> (B L) T D"
        o = _parallel_block(sd, p + "temporal_block.", u, H, heads, cs_t, sn_t)
        h = h + g2 * o.reshape(B, L, T, H).permute(0, 2, 1, 3)
        if trace is not None:
            trace[f"h{i + 1}"] = h.clone()
    ada = linear(svec, sd["adaLN_modulation.1.weight"], sd["adaLN_modulation.1.bias"])
    shift, scale = (m[:, None, None, :] for m in ada.chunk(2, dim=-1))
    u = layer_norm(h, 1e-6) * (1 + scale) + shift
    return linear(u, sd["linear.weight"], sd["linear.bias"])


# --------------------------------------------------------------------------------------------------
# transport / sampler
# --------------------------------------------------------------------------------------------------
def sample_interval(path_type: str, prediction: str) -> Tuple[float, float]:
    """CreateTransport (transport/__init__.py:57-68) + Transport.check_interval (transport.py:69-101) for the
    ODE sampler (sde=False, eval=True, reverse=False): velocity models on Linear/GVP integrate t∈[0,1];
    every other model type on Linear/GVP uses sample_eps=1e-3 ⇒ [0.001, 0.999]; VP ⇒ [0, 0.999]."""
    if path_type == "VP":
        return 0.0, 1.0 - 1e-3
    if prediction == "velocity":
        return 0.0, 1.0
    return 1e-3, 1.0 - 1e-3


def drift(path_type: str, prediction: str, x: Tensor, t: Tensor, m: Tensor) -> Tensor:
    """Transport.get_drift — transport.py:158-202 with the coupling plans of path.py (ICPlan :21-47,
    GVPCPlan :188-206).  ``m`` is the network output, ``t``: [B].  Written out term by term as the reference
    does (−drift_mean + drift_var·score), not in closed form, so rounding matches."""
    if prediction == "velocity":
        return m
    tt = t.reshape(-1, *([1] * (x.dim() - 1)))
    if path_type == "GVP":
        alpha, d_alpha = torch.sin(tt * math.pi / 2), math.pi / 2 * torch.cos(tt * math.pi / 2)
        sigma, d_sigma = torch.cos(tt * math.pi / 2), -math.pi / 2 * torch.sin(tt * math.pi / 2)
        ratio = math.pi / (2 * torch.tan(tt * math.pi / 2))
    elif path_type == "Linear":
        alpha, d_alpha = tt, 1.0
        sigma, d_sigma = 1 - tt, -1.0
        ratio = 1 / tt
    else:
        raise NotImplementedError(path_type)
    drift_mean = -(ratio * x)
    drift_var = ratio * sigma ** 2 - sigma * d_sigma
    if prediction == "data":
        score = -(1 / sigma ** 2) * (x - alpha * m)
    elif prediction == "noise":
        score = m / -sigma
    elif prediction == "score":
        score = m
    else:
        raise NotImplementedError(prediction)
    return -drift_mean + drift_var * score


def ode_sample(model_fn, x0: Tensor, *, path_type: str = "GVP", prediction: str = "data", num_steps: int = 10,
               record_velocity: Optional[List[Tensor]] = None) -> Tensor:
    """Sampler.sample_ode (transport.py:365-411) → ode.sample (integrators.py:103-120) →
    torchdiffeq fixed-grid Euler.  ``model_fn(x, t[B]) -> net output``.  Returns all ``num_steps`` states."""
    t0, t1 = sample_interval(path_type, prediction)
    grid = torch.linspace(t0, t1, num_steps)  # fp32, as integrators.py:98
    states = [x0]
    x = x0
    for i in range(num_steps - 1):
        t = torch.ones(x.shape[0], device=x.device) * grid[i]
        v = drift(path_type, prediction, x, t.to(x.dtype), model_fn(x, t.to(x.dtype)))
        if record_velocity is not None:
            record_velocity.append(v)
        x = x + (grid[i + 1] - grid[i]).to(x.dtype) * v
        states.append(x)
    return torch.stack(states)


# ---- SDE sampler: Sampler.sample_sde + integrators.sde --------------------------------------
# --------------------------------------------------------------------------------------------------
# torchdiffeq integrators other than fixed-grid Euler
# --------------------------------------------------------------------------------------------------
# The reference calls ``torchdiffeq.odeint(fn, x, t, method=sampling_method, atol=[atol], rtol=[rtol])`` (integrators.py:103-120) with
# ``dopri5`` as default (transport.py:365-372; configs/eval_peptide.yaml:21-23).  torchdiffeq is not vendored and not pinned by the
# reference (environment.yaml), and is absent here: PARITY OF THIS SECTION IS PINNED ON PUBLISHED MATHEMATICS, NOT ON THE PACKAGE —
# the Butcher tableau against scipy's independent RK45 / RK23 tables, the embedded error weights and the mid-point interpolant
# against their order conditions, the solvers against closed-form ODE solutions (tests/test_odeint.py).  The step-size controller is
# restated from torchdiffeq 0.2.x (rk_common.py, misc.py, interp.py) as tensor code, separately from lam_slide_b200/odeint.py.
_DOPRI5 = dict(
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    beta=[[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9], [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
          [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656], [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_error=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
             11 / 84 - 649 / 6300, -1.0 / 60.0],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
    order=5)
_BOSH3 = dict(alpha=[1 / 2, 3 / 4, 1.0], beta=[[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]], c_sol=[2 / 9, 1 / 3, 4 / 9, 0.0],
              c_error=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8], c_mid=[0.0, 0.5, 0.0, 0.0], order=3)
_ADAPTIVE_HEUN = dict(alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.0], order=2)
RK_TABLEAUS = {"dopri5": _DOPRI5, "bosh3": _BOSH3, "adaptive_heun": _ADAPTIVE_HEUN}


def _rms(x: Tensor) -> Tensor:
    return x.abs().pow(2).mean().sqrt()


def odeint(func, y0: Tensor, t: Tensor, *, method: str = "dopri5", rtol: float = 1e-3, atol: float = 1e-6,
           stats: Optional[dict] = None) -> Tensor:
    """``torchdiffeq.odeint`` for the methods the product implements: state in ``y0.dtype``, time in float64 for the adaptive
    solvers (rk_common.py: "all 'time'-like objects use float64"), the user function sees the time cast to the state dtype."""
    nfe = [0]

    def f(tt, y):
        nfe[0] += 1
        return func(torch.as_tensor(tt, dtype=torch.float64).to(y.dtype), y)

    sol = [y0]
    if method in ("euler", "midpoint", "rk4", "heun2", "heun3"):  # fixed_grid.py / rk_common.py step functions, grid = t
        y = y0
        for i in range(len(t) - 1):
            t0, t1 = t[i], t[i + 1]
            dt = t1 - t0
            if method == "euler":
                dy = dt * f(t0, y)
            elif method == "midpoint":
                half = 0.5 * dt
                dy = dt * f(t0 + half, y + f(t0, y) * half)
            elif method == "rk4":
                k1 = f(t0, y)
                k2 = f(t0 + dt / 3, y + dt * k1 / 3)
                k3 = f(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
                k4 = f(t1, y + dt * (k1 - k2 + k3))
                dy = (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
            elif method == "heun2":
                k1 = f(t0, y)
                dy = dt * 0.5 * (k1 + f(t1, y + dt * k1))
            else:  # heun3
                k1 = f(t0, y)
                k2 = f(t0 + dt / 3, y + dt * k1 / 3)
                k3 = f(t0 + dt * 2 / 3, y + dt * 2 / 3 * k2)
                dy = dt * (0.25 * k1 + 0.75 * k3)
            y = y + dy.to(y.dtype)
            sol.append(y)
        if stats is not None:
            stats.update(nfe=nfe[0])
        return torch.stack(sol)

    tab = RK_TABLEAUS[method]
    dt_ = y0.dtype
    alpha = torch.tensor(tab["alpha"], dtype=torch.float64).to(dt_)
    beta = [torch.tensor(b, dtype=torch.float64).to(dt_) for b in tab["beta"]]
    c_sol = torch.tensor(tab["c_sol"], dtype=torch.float64).to(dt_)
    c_error = torch.tensor(tab["c_error"], dtype=torch.float64).to(dt_)
    c_mid = torch.tensor(tab["c_mid"], dtype=torch.float64).to(dt_)
    order = tab["order"]
    rt = torch.as_tensor([rtol], dtype=torch.float64)  # the reference passes one-element lists (integrators.py:116-117)
    at = torch.as_tensor([atol], dtype=torch.float64)
    tt = t.to(torch.float64)

    def rk_step(y, f0, t0, dt):
        t0c, dtc = t0.to(dt_), dt.to(dt_)
        k = [f0]
        yi = y
        for a_i, b_i in zip(alpha, beta):
            ti = t0 + dt if float(a_i) == 1.0 else t0 + a_i * dt
            yi = y + torch.stack(k, -1).matmul(b_i * dtc)
            k.append(f(ti, yi))
        kk = torch.stack(k, -1)
        if not (float(c_sol[-1]) == 0 and bool((c_sol[:-1] == beta[-1]).all())):
            yi = y + kk.matmul(dtc * c_sol)
        return yi, k[-1], kk.matmul(dtc * c_error), kk

    # misc.py: _select_initial_step, called with order - 1
    f0 = f(tt[0], y0)
    scale = at + y0.abs() * rt
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = torch.tensor(1e-6, dtype=torch.float64) if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = h0.abs()
    f1 = f(tt[0] + h0, y0 + h0.to(dt_) * f0)
    d2 = (_rms((f1 - f0) / scale) / h0).abs()
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=torch.float64), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order))
    dt = torch.min(100 * h0, h1.abs()).to(torch.float64)

    y, t0s, t1s, interp = y0, tt[0], tt[0], None
    acc = rej = 0
    for nt in tt[1:]:
        while nt > t1s:
            assert t1s + dt > t1s, "underflow in dt"
            t0 = t1s
            y1, f1, err, kk = rk_step(y, f0, t0, dt)
            ratio = _rms(err / (at + rt * torch.max(y.abs(), y1.abs()))).abs()
            assert torch.isfinite(ratio)
            if ratio <= 1:
                dtc = dt.to(dt_)
                y_mid = y + kk.matmul(dtc * c_mid)
                fa, fb = kk[..., 0], kk[..., -1]
                # interp.py: _interp_fit
                a = 2 * dtc * (fb - fa) - 8 * (y1 + y) + 16 * y_mid
                b = dtc * (5 * fa - 3 * fb) + 18 * y + 14 * y1 - 32 * y_mid
                c = dtc * (fb - 4 * fa) - 11 * y - 5 * y1 + 16 * y_mid
                interp = ([y, dtc * fa, c, b, a], t0, t0 + dt)
                t0s, t1s = t0, t0 + dt
                y, f0 = y1, f1
                acc += 1
            else:
                rej += 1
            # misc.py: _optimal_step_size
            if ratio == 0:
                dt = dt * 10.0
            else:
                dfac = 1.0 if ratio < 1 else 0.2
                dt = dt * min(10.0, max(0.9 / float(ratio) ** (1.0 / order), dfac))
        coef, ta, tb = interp
        x = ((nt - ta) / (tb - ta)).to(dt_)  # interp.py: _interp_evaluate
        total = coef[0] + x * coef[1]
        xp = x
        for cf in coef[2:]:
            xp = xp * x
            total = total + xp * cf
        sol.append(total)
    if stats is not None:
        stats.update(nfe=nfe[0], accepted=acc, rejected=rej)
    return torch.stack(sol)


def ode_solve(model_fn, x0: Tensor, *, path_type: str = "GVP", prediction: str = "data", method: str = "dopri5", num_steps: int = 50,
              atol: float = 1e-6, rtol: float = 1e-3, stats: Optional[dict] = None) -> Tensor:
    """Sampler.sample_ode (transport.py:365-411) + ode.sample (integrators.py:103-120) for any torchdiffeq method restated above."""
    t0, t1 = sample_interval(path_type, prediction)
    grid = torch.linspace(t0, t1, num_steps)

    def fn(t, x):
        tv = torch.ones(x.shape[0], device=x.device) * t
        return drift(path_type, prediction, x, tv.to(x.dtype), model_fn(x, tv.to(x.dtype)))

    return odeint(fn, x0, grid, method=method, rtol=rtol, atol=atol, stats=stats)


def _plan(path_type: str, tt: Tensor):
    """(alpha, d_alpha, sigma, d_sigma, d_alpha / alpha) of ICPlan (path.py:27-37) / GVPCPlan (path.py:192-206) at ``tt``."""
    if path_type == "GVP":
        return (torch.sin(tt * math.pi / 2), math.pi / 2 * torch.cos(tt * math.pi / 2), torch.cos(tt * math.pi / 2),
                -math.pi / 2 * torch.sin(tt * math.pi / 2), math.pi / (2 * torch.tan(tt * math.pi / 2)))
    if path_type == "Linear":
        return tt, 1.0, 1 - tt, -1.0, 1 / tt
    raise NotImplementedError(path_type)


def score(path_type: str, prediction: str, x: Tensor, t: Tensor, m: Tensor) -> Tensor:
    """Transport.get_score — transport.py:204-226 with path.py:73-95 (score from velocity / data), ``m`` = network output."""
    tt = t.reshape(-1, *([1] * (x.dim() - 1)))
    alpha, d_alpha, sigma, d_sigma, _ = _plan(path_type, tt)
    if prediction == "noise":
        return m / -sigma
    if prediction == "score":
        return m
    if prediction == "velocity":
        rar = alpha / d_alpha
        var = sigma ** 2 - rar * d_sigma * sigma
        return (rar * m - x) / var
    if prediction == "data":
        return -(1 / sigma ** 2) * (x - alpha * m)
    raise NotImplementedError(prediction)


def diffusion(path_type: str, x: Tensor, t: Tensor, form: str, norm: float) -> Tensor:
    """ICPlan.compute_diffusion — path.py:49-71 (the SBDM entry is compute_drift(x, t)[1], path.py:39-47)."""
    tt = t.reshape(-1, *([1] * (x.dim() - 1)))
    _, _, sigma, d_sigma, ratio = _plan(path_type, tt)
    choices = {
        "constant": lambda: norm,
        "SBDM": lambda: norm * (ratio * sigma ** 2 - sigma * d_sigma),
        "sigma": lambda: norm * sigma,
        "linear": lambda: norm * (1 - tt),
        "decreasing": lambda: 0.25 * (norm * torch.cos(math.pi * tt) + 1) ** 2,
        "inccreasing-decreasing": lambda: norm * torch.sin(math.pi * tt) ** 2,
    }
    return choices[form]()


def sde_interval(path_type: str, prediction: str, diffusion_form: str, last_step_size: float) -> Tuple[float, float]:
    """Transport.check_interval (transport.py:69-101) for sde=True, eval=True, reverse=False on the Linear / GVP plans with the
    default eps of CreateTransport (transport/__init__.py:57-68: 1e-3 unless velocity, then 0)."""
    eps = 0.0 if prediction == "velocity" else 1e-3
    t0 = eps if (diffusion_form == "SBDM") or prediction != "velocity" else 0
    t1 = 1 - eps if last_step_size == 0 else 1 - last_step_size
    return t0, t1


def sde_sample(model_fn, x0: Tensor, noises: Sequence[Tensor], *, path_type: str = "GVP", prediction: str = "data",
               sampling_method: str = "Euler", diffusion_form: str = "SBDM", diffusion_norm: float = 1.0,
               last_step: Optional[str] = "Mean", last_step_size: float = 0.04, num_steps: int = 250) -> List[Tensor]:
    """Sampler.sample_sde (transport.py:301-363) around integrators.sde (integrators.py:7-78): Euler-Maruyama or Heun over
    ``linspace(t0, t1, num_steps)`` (num_steps - 1 steps), then the last step (None / Mean / Tweedie / Euler).  ``noises[i]``
    stands in for the ``th.randn`` of step i.  Returns the ``num_steps`` states the reference returns."""
    if last_step is None:
        last_step_size = 0.0
    t0, t1 = sde_interval(path_type, prediction, diffusion_form, last_step_size)
    grid = torch.linspace(t0, t1, num_steps)
    dt = grid[1] - grid[0]

    def ode_drift(x, t):
        return drift(path_type, prediction, x, t, model_fn(x, t))

    def sde_drift(x, t):
        m = model_fn(x, t)
        return drift(path_type, prediction, x, t, m) + diffusion(path_type, x, t, diffusion_form, diffusion_norm) * score(
            path_type, prediction, x, t, m)

    x = x0
    xs: List[Tensor] = []
    for i, ti in enumerate(grid[:-1]):
        w = noises[i]
        dw = w * torch.sqrt(dt)
        t = torch.ones(x.shape[0]) * ti
        if sampling_method == "Euler":
            d = sde_drift(x, t)
            mean_x = x + d * dt
            x = mean_x + torch.sqrt(2 * torch.as_tensor(diffusion(path_type, x, t, diffusion_form, diffusion_norm))) * dw
        elif sampling_method == "Heun":
            dif = diffusion(path_type, x, t, diffusion_form, diffusion_norm)
            xhat = x + torch.sqrt(2 * torch.as_tensor(dif)) * dw
            k1 = sde_drift(xhat, t)
            xp = xhat + dt * k1
            k2 = sde_drift(xp, t + dt)
            x = xhat + 0.5 * dt * (k1 + k2)
        else:
            raise NotImplementedError(sampling_method)
        xs.append(x)
    ts = torch.ones(x0.shape[0]) * t1
    if last_step is None:
        x = xs[-1]
    elif last_step == "Mean":
        x = xs[-1] + sde_drift(xs[-1], ts) * last_step_size
    elif last_step == "Tweedie":
        alpha, _, sigma, _, _ = _plan(path_type, ts)
        x = xs[-1] / alpha[0] + (sigma[0] ** 2) / alpha[0] * score(path_type, prediction, xs[-1], ts, model_fn(xs[-1], ts))
    elif last_step == "Euler":
        x = xs[-1] + ode_drift(xs[-1], ts) * last_step_size
    else:
        raise NotImplementedError(last_step)
    xs.append(x)
    return xs


def setup_conditioning(latents: Tensor, cond_idx: Sequence[int], mask_cond_mean: bool = True) -> Tuple[Tensor, Tensor]:
    """SecondStageCondLightningBase.setup_conditioning — lightning_base.py:240-263."""
    B, T, L, _ = latents.shape
    mask = torch.zeros(B, T, L, dtype=torch.int64, device=latents.device)
    mask[:, cond_idx[0]: cond_idx[1]] = 1
    if mask_cond_mean:
        fill = latents[:, cond_idx[0]: cond_idx[1]].mean(dim=1, keepdim=True).expand_as(latents)
    else:
        fill = torch.zeros_like(latents)
    return torch.where(mask[..., None].bool(), latents, fill), mask


# --------------------------------------------------------------------------------------------------
# first stage: UPT-style encoder / decoder
# --------------------------------------------------------------------------------------------------
def embedding_max_norm(table: Tensor, max_norm: float = 1.0) -> Tensor:
    """nn.Embedding(max_norm=…) renormalises looked-up rows in place: rows with ‖row‖₂ > max_norm are scaled by
    max_norm/(norm+1e-7) (entity_embeddings.py:25; first-stage.yaml ``max_norm: 1``).  Idempotent, so applying
    it to the whole table once equals the reference's lazy per-lookup renorm."""
    n = table.norm(dim=-1, keepdim=True)
    return torch.where(n > max_norm, table * (max_norm / (n + 1e-7)), table)


def _heads(x: Tensor, h: int) -> Tensor:
    b, n, _ = x.shape
    return x.reshape(b, n, h, -1).permute(0, 2, 1, 3)  # "b n (h d) -> b h n d"


def _ff(sd: SD, p: str, x: Tensor) -> Tensor:
    """PreNorm(FeedForward) — torch_modules.py:108-144: LN(affine, eps 1e-5) → Linear → GELU → Linear."""
    u = layer_norm(x, 1e-5, sd[p + "ff.norm.weight"], sd[p + "ff.norm.bias"])
    u = gelu_erf(linear(u, sd[p + "ff.fn.net.0.0.weight"], sd[p + "ff.fn.net.0.0.bias"]))
    return linear(u, sd[p + "ff.fn.net.1.weight"], sd[p + "ff.fn.net.1.bias"])


def cross_attention_block(sd: SD, p: str, x: Tensor, ctx: Tensor, heads: int, qk_norm: bool,
                          mask: Optional[Tensor] = None) -> Tensor:
    """CrossAttentionBlock — torch_modules.py:189-218 (+ PreNorm :108-122, Attention :147-186)."""
    xq = layer_norm(x, 1e-5, sd[p + "attn.norm.weight"], sd[p + "attn.norm.bias"])
    xc = layer_norm(ctx, 1e-5, sd[p + "attn.norm_context.weight"], sd[p + "attn.norm_context.bias"])
    q = _heads(linear(xq, sd[p + "attn.fn.to_q.weight"]), heads)
    k, v = linear(xc, sd[p + "attn.fn.to_kv.weight"]).chunk(2, dim=-1)  # k first, then v
    k, v = _heads(k, heads), _heads(v, heads)
    dh = q.shape[-1]
    if qk_norm:
        q = rms_norm(q, sd[p + "attn.fn.norm.query_norm.scale"])
        k = rms_norm(k, sd[p + "attn.fn.norm.key_norm.scale"])
    km = None if mask is None else mask[:, None, None, :]
    a = softmax_attention(q, k, v, dh ** -0.5, km)
    a = a.permute(0, 2, 1, 3).flatten(-2)
    x = linear(a, sd[p + "attn.fn.to_out.weight"], sd[p + "attn.fn.to_out.bias"]) + x
    return _ff(sd, p, x) + x


def self_attention_block(sd: SD, p: str, x: Tensor, heads: int, qk_norm: bool) -> Tensor:
    """SelfAttentionBlock — torch_modules.py:256-273 (SelfAttention :221-253; the mask is dropped by PreNorm)."""
    u = layer_norm(x, 1e-5, sd[p + "attn.norm.weight"], sd[p + "attn.norm.bias"])
    q, k, v = (_heads(c, heads) for c in linear(u, sd[p + "attn.fn.to_qkv.weight"]).chunk(3, dim=-1))
    dh = q.shape[-1]
    if qk_norm:
        q = rms_norm(q, sd[p + "attn.fn.norm.query_norm.scale"])
        k = rms_norm(k, sd[p + "attn.fn.norm.key_norm.scale"])
    a = softmax_attention(q, k, v, dh ** -0.5)
    a = a.permute(0, 2, 1, 3).flatten(-2)
    x = linear(a, sd[p + "attn.fn.to_out.weight"], sd[p + "attn.fn.to_out.bias"]) + x
    return _ff(sd, p, x) + x


def point_embed(sd: SD, p: str, pos: Tensor) -> Tensor:
    """PointEmbed — embeddings.py:50-88: Linear(cat[sin(pos·basis), cos(pos·basis), pos])."""
    proj = pos @ sd[p + "basis"]
    return linear(torch.cat([proj.sin(), proj.cos(), pos], dim=-1), sd[p + "mlp.weight"], sd[p + "mlp.bias"])


def first_stage_features(sd: SD, cfg: dict, batch: Dict[str, Tensor]) -> Tensor:
    """Backbone.prepare_inputs of the four datasets — first_stage/peptide.py:96-103, md17.py:52-58,
    nba.py:54-59, pedestrian.py:39-42.  Tensors are per frame: [F, N, …]."""
    kind = cfg["kind"]
    if kind == "peptide":
        res = embedding_max_norm(sd["embedding_res.weight"])[batch["aatype"]]
        x = torch.cat([res, batch["atom14_pos"].flatten(-2)], dim=-1)
    elif kind == "md17":
        atom = embedding_max_norm(sd["embed_atom.weight"])[batch["atom"]]
        x = torch.cat([atom, point_embed(sd, "embed_pos.", batch["pos"])], dim=-1)
    elif kind == "nba":
        x = torch.cat([batch["pos"], sd["embed_team.weight"][batch["team"]],
                       sd["embed_group.weight"][batch["group"]]], dim=-1)
    elif kind == "pedestrian":
        x = batch["pos"]
    else:
        raise ValueError(kind)
    x = linear(gelu_erf(linear(x, sd["net_merge.0.weight"], sd["net_merge.0.bias"])),
               sd["net_merge.2.weight"], sd["net_merge.2.bias"])
    if kind == "peptide":  # SinCosPositionalEmbedding1D — embeddings.py:39-47
        x = x + sd["embed_res_pos.embeddings"][: x.shape[1]][None]
    return x


def first_stage_encode(sd: SD, cfg: dict, batch: Dict[str, Tensor]) -> Tensor:
    """BackboneBase.encode (lightning_base.py:37-40) → Encoder.forward (encoder.py:35-41, 96-103) → quant (:24-27).
    Returns latents [F, L, D]."""
    e = cfg["encoder"]
    x = first_stage_features(sd, cfg, batch)
    ent = embedding_max_norm(sd["encoder.entity_embedding.embedding.weight"])[batch["entities"]]
    ctx = torch.cat([x, ent], dim=-1)
    ctx = linear(gelu_erf(linear(ctx, sd["encoder.mlp.0.weight"], sd["encoder.mlp.0.bias"])),
                 sd["encoder.mlp.2.weight"], sd["encoder.mlp.2.bias"])
    z = sd["encoder.latents"][None].expand(ctx.shape[0], -1, -1)
    mask = None if cfg["kind"] == "peptide" else batch.get("attention_mask")  # peptide.py:79 passes mask=None
    for i in range(e["num_block_cross"]):
        z = cross_attention_block(sd, f"encoder.cross_attn_blocks.{i}.", z, ctx, e["num_head_cross"], e["qk_norm"], mask)
    for i in range(e["num_block_attn"]):
        z = self_attention_block(sd, f"encoder.blocks_attn.{i}.", z, e["num_head_latent"], e["qk_norm"])
    return layer_norm(linear(z, sd["quant.0.weight"], sd["quant.0.bias"]), 1e-5)


def first_stage_decode(sd: SD, cfg: dict, latents: Tensor, entities: Tensor) -> Dict[str, Tensor]:
    """BackboneBase.decode (lightning_base.py:42-44) → Decoder.forward (decoder.py:83-102) /
    DecoderQuerySplitter.forward (:391-411, extender :385-389)."""
    d = cfg["decoder"]
    z = linear(layer_norm(latents, 1e-5), sd["post_quant.1.weight"], sd["post_quant.1.bias"])
    ent = embedding_max_norm(sd["decoder.entity_embedding.embedding.weight"])[entities]
    q = linear(ent, sd["decoder.query_mlp.1.weight"], sd["decoder.query_mlp.1.bias"])
    for i in range(d["num_block_attn"]):
        z = self_attention_block(sd, f"decoder.self_attn_blocks.{i}.", z, d["num_head_latent"], d["qk_norm"])
    for i in range(d["num_block_cross"]):
        z = cross_attention_block(sd, f"decoder.cross_attn_blocks.{i}.", z, q, d["num_head_cross"], d["qk_norm"])
    if d["kind"] == "DecoderQuerySplitter":
        # Conv1d(D → D·n, k=1) on "B D L", then "B (D N) L -> B (L N) D"
        n = d["num_split"]
        F_, L, D = z.shape
        w = sd["decoder.extender.1.weight"].reshape(D * n, D)
        e = linear(z, w, sd["decoder.extender.1.bias"])  # [F, L, D·n], channel c = d·n + j
        z = e.reshape(F_, L, D, n).permute(0, 1, 3, 2).reshape(F_, L * n, D)
    o = cross_attention_block(sd, "decoder.output_block.", q, z, d["num_head_cross"], d["qk_norm"])
    out = {}
    for name, _ in d["outputs"]:
        p = f"decoder.output_layers.{name}."
        out[name] = linear(gelu_erf(linear(o, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])
    return out


# --------------------------------------------------------------------------------------------------
# end-to-end sample()
# --------------------------------------------------------------------------------------------------
_FRAME_KEYS = ("atom14_pos", "aatype", "pos", "atom", "team", "group", "entities", "attention_mask")


def sample(fs_sd: SD, bb_sd: SD, cfg: dict, batch: Dict[str, Tensor], noise: Tensor, *, num_steps: int = 10,
           y: Optional[Tensor] = None, record: Optional[dict] = None) -> Dict[str, Tensor]:
    """SecondStageCondLightningBase.sample — lightning_base.py:205-238 with Wrapper.encode/decode
    (second_stage/peptide.py:85-102, md17.py:115-130, nba.py:133-148, pedestrian.py:121-136).
    ``noise`` stands in for ``torch.randn_like(x_cond)``; ``y`` for ``vec_in_embedding(cond_scene)``."""
    fcfg, bcfg = cfg["first_stage"], cfg["backbone"]
    B, T = batch["entities"].shape[:2]
    flat = {k: v.flatten(0, 1) for k, v in batch.items() if k in _FRAME_KEYS}
    latents = first_stage_encode(fs_sd, fcfg, flat).unflatten(0, (B, T))
    x_cond, x_mask = setup_conditioning(latents, cfg["cond_idx"], cfg["mask_cond_mean"])
    vel: List[Tensor] = []
    states = ode_sample(lambda x, t: backbone_forward(bb_sd, bcfg, x, t, x_cond, x_mask, y), noise,
                        path_type=cfg["path_type"], prediction=cfg["prediction"], num_steps=num_steps,
                        record_velocity=vel)
    out = first_stage_decode(fs_sd, fcfg, states[-1].flatten(0, 1), flat["entities"])
    out = {k: v.unflatten(0, (B, T)) for k, v in out.items()}
    if record is not None:
        record.update(latents=latents, x_cond=x_cond, x_cond_mask=x_mask, states=states, velocities=torch.stack(vel))
    return out


# --------------------------------------------------------------------------------------------------
# autoregressive roll-out driver
# --------------------------------------------------------------------------------------------------
def rollout_create_batch(pos: Tensor, res: Tensor, res_mask: Tensor, T: int) -> Dict[str, Tensor]:
    """SIAtom14SamplingWrapper.create_batch — src/modules/sampling.py:24-43: one conditioning frame, masked by the
    residue-type atom mask and repeated over all T frames of a B = 1 batch."""
    pos = pos * res_mask[..., None].to(pos.dtype)
    R = res.shape[0]
    return {
        "atom14_pos": pos[None, None].expand(1, T, R, 14, 3).contiguous(),
        "aatype": res[None, None].expand(1, T, R).contiguous(),
        "attention_mask": torch.ones(1, T, R, dtype=torch.bool),
        "entities": torch.arange(R)[None, None].expand(1, T, R).contiguous(),
    }


def sample_rollout(fs_sd: SD, bb_sd: SD, cfg: dict, cond_pos: Tensor, res: Tensor, res_mask: Tensor, noises: Sequence[Tensor],
                   *, shift: float = 0.0, scale: float = 1.0, num_steps: int = 10) -> Tensor:
    """SIAtom14SamplingWrapper.sample_rollout — src/modules/sampling.py:45-63: normalise the conditioning frame, then
    ``len(noises)`` times: build the batch from the current frame, ``sample()`` a block of T frames, continue from its last
    frame; concatenate the blocks, put the conditioning frame back at index 0, de-normalise.  ``noises[i]`` ([1,T,L,D])
    stands in for the ``randn_like`` of the i-th ``sample()`` call.  Returns [len(noises) * T, R, 14, 3]."""
    T = cfg["T"]
    cond = (cond_pos - shift) / scale
    pos = cond.clone()
    blocks = []
    for nz in noises:
        batch = rollout_create_batch(pos, res, res_mask, T)
        pred = sample(fs_sd, bb_sd, cfg, batch, nz, num_steps=num_steps)["atom14_pos"]  # [1, T, R, 42]
        pred = pred.unflatten(-1, (14, 3)).squeeze(0)  # Wrapper.decode: "(B T) L (A D) -> B T L A D" (second_stage/peptide.py:97-102)
        blocks.append(pred)
        pos = pred[-1].clone()
    positions = torch.cat(blocks)
    positions[0] = cond
    return positions * scale + shift


def rollout_inputs(R: int, seed: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Seeded synthetic conditioning frame for the roll-out tests: positions [R,14,3], residue types [R], atom mask [R,14]."""
    g = torch.Generator().manual_seed(seed)
    cond_pos = torch.randn(R, 14, 3, generator=g)
    res = torch.randint(0, 20, (R,), generator=g)
    res_mask = torch.rand(R, 14, generator=g) < 0.7
    res_mask[:, :4] = True  # backbone atoms always present
    return cond_pos, res, res_mask


# --------------------------------------------------------------------------------------------------
# K-sample evaluation metrics
# --------------------------------------------------------------------------------------------------
def ksample_min_ade_fde(preds: Sequence[Tensor], true_pos: Tensor, attention_mask: Tensor, cond_end: int,
                        num_runs: int) -> Tuple[Tensor, Tensor]:
    """The unclustered metric of ``Wrapper.test_step`` — second_stage/nba.py:184-225 (pedestrian.py:172-213 is identical):
    ``preds[k]`` is the k-th ``sample()`` result ``[B, T, A, D]``, ``true_pos [B, T, A, D]`` the ground truth, ``attention_mask
    [B, T, A]`` marks real agents.  Frames before ``cond_end`` are dropped, padded agents removed (mask of the last frame),
    and per remaining agent the minimum over the first ``num_runs`` samples of the time-averaged (ADE) and final-frame (FDE)
    displacement is returned."""
    B, T, A, D = true_pos.shape
    mask = attention_mask[:, -1].reshape(B * A)
    true = true_pos[:, cond_end:].permute(0, 2, 1, 3).reshape(B * A, T - cond_end, D)[mask]
    all_traj = []
    for p in preds:
        p = p[:, cond_end:].permute(0, 2, 1, 3).reshape(B * A, T - cond_end, D)[mask]
        all_traj.append(p)
    all_traj = torch.stack(all_traj, dim=1)  # [n, K, T', D]
    selected = all_traj[:, :num_runs]
    error = torch.norm(selected - true[:, None], dim=-1)  # [n, runs, T']
    return error.mean(dim=-1).min(dim=1).values, error[..., -1].min(dim=1).values


def ksample_mean_ade_fde(preds: Sequence[Tensor], true_pos: Tensor, cond_end: int) -> Tuple[Tensor, Tensor]:
    """``Wrapper.test_step`` of second_stage/md17.py:148-168: per sample, the mean over the K runs of the displacement averaged
    over frames and atoms (ADE) and over the atoms of the last frame (FDE)."""
    true = true_pos[:, cond_end:]
    ades, fdes = [], []
    for p in preds:
        p = p[:, cond_end:]
        ades.append(torch.norm(true - p, dim=-1).mean(dim=(1, 2)))
        fdes.append(torch.norm(true[:, -1] - p[:, -1], dim=-1).mean(dim=1))
    return torch.stack(ades).mean(dim=0), torch.stack(fdes).mean(dim=0)


def ksample_inputs(B: int, T: int, A: int, D: int, K: int, seed: int, pad_agents: bool):
    """Seeded synthetic inputs of the K-sample metric tests: K predictions, the ground truth, the agent mask."""
    g = torch.Generator().manual_seed(seed)
    true_pos = torch.randn(B, T, A, D, generator=g)
    preds = [true_pos + 0.3 * torch.randn(B, T, A, D, generator=g) for _ in range(K)]
    mask = torch.ones(B, T, A, dtype=torch.bool)
    if pad_agents:
        for b in range(B):
            n_valid = int(torch.randint(max(1, A // 2), A + 1, (1,), generator=g))
            mask[b, :, n_valid:] = False
    return preds, true_pos, mask


# --------------------------------------------------------------------------------------------------
# deterministic parameters and synthetic batches (shared by the tests, smoke() and bench.py)
# --------------------------------------------------------------------------------------------------
def _rand(gen: torch.Generator, shape, std: float) -> Tensor:
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def init_backbone_params(cfg: dict, seed: int) -> SD:
    """Random second-stage weights under the reference's state-dict key names.
    Matrices ~ N(0, 1/fan_in); the layers the reference zero-initialises (modulation.lin, final linear —
    latent_si_v31.py:152-156) get N(0, 0.02) so parity is not vacuous; mask_to_emb N(0,1); RMSNorm scales
    1 + N(0, 0.1).  Generated with a seeded CPU generator ⇒ identical on every box with this torch build."""
    g = torch.Generator().manual_seed(seed)
    H, D, depth = cfg["hidden_size"], cfg["in_dim"], cfg["depth"]
    M = int(cfg["mlp_ratio"] * H)
    hd = H // cfg["num_heads"]
    sd: SD = {}

    def lin(name, out_f, in_f, std=None, bias_std=0.02):
        sd[name + ".weight"] = _rand(g, (out_f, in_f), std if std is not None else in_f ** -0.5)
        sd[name + ".bias"] = _rand(g, (out_f,), bias_std)

    lin("x_in", H, D)
    lin("cond_to_emb", H, D)
    sd["mask_to_emb.weight"] = _rand(g, (2, H), 1.0)
    lin("time_in.in_layer", H, 256)
    lin("time_in.out_layer", H, H)
    if cfg.get("vec_in_dim"):
        lin("vec_in.in_layer", H, cfg["vec_in_dim"])
        lin("vec_in.out_layer", H, H)
    for i in range(depth):
        p = f"blocks.{i}."
        lin(p + "modulation.lin", 6 * H, H, std=0.02)
        for blk in ("spatial_block.", "temporal_block."):
            lin(p + blk + "linear1", 3 * H + M, H)
            lin(p + blk + "linear2", H, H + M)
            sd[p + blk + "norm.query_norm.scale"] = 1.0 + _rand(g, (hd,), 0.1)
            sd[p + blk + "norm.key_norm.scale"] = 1.0 + _rand(g, (hd,), 0.1)
    lin("adaLN_modulation.1", 2 * H, H, std=0.02)
    lin("linear", D, H, std=0.02)
    return sd


def _sincos_1d(n_positions: int, dim: int) -> Tensor:
    """get_1d_sincos_pos_embed_from_grid — embeddings.py:6-25 (fp64 → fp32, cat[sin, cos])."""
    omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
    out = torch.arange(n_positions, dtype=torch.float64)[:, None] * omega[None]
    return torch.cat([out.sin(), out.cos()], dim=1).float()


def _point_basis(hidden_dim: int) -> Tensor:
    """PointEmbed.basis — embeddings.py:62-78: block-diagonal 2^k·π frequencies, [3, hidden_dim/2]."""
    k = hidden_dim // 6
    e = (2.0 ** torch.arange(k).float()) * math.pi
    z = torch.zeros(k)
    return torch.stack([torch.cat([e, z, z]), torch.cat([z, e, z]), torch.cat([z, z, e])])


def init_first_stage_params(cfg: dict, seed: int) -> SD:
    """Random first-stage weights under the reference's key names (``first_stage_model.backbone.*``)."""
    g = torch.Generator().manual_seed(seed)
    e, d = cfg["encoder"], cfg["decoder"]
    Din, D, E = cfg["dim_input"], cfg["dim_latent"], cfg["entity_dim"]
    C = Din + E
    sd: SD = {}

    def lin(name, out_f, in_f, bias=True):
        sd[name + ".weight"] = _rand(g, (out_f, in_f), in_f ** -0.5)
        if bias:
            sd[name + ".bias"] = _rand(g, (out_f,), 0.02)

    def ln(name, dim):
        sd[name + ".weight"] = 1.0 + _rand(g, (dim,), 0.1)
        sd[name + ".bias"] = _rand(g, (dim,), 0.05)

    def attn_common(p, dim, inner, dh):
        lin(p + "attn.fn.to_out", dim, inner)
        sd[p + "attn.fn.norm.query_norm.scale"] = 1.0 + _rand(g, (dh,), 0.1)
        sd[p + "attn.fn.norm.key_norm.scale"] = 1.0 + _rand(g, (dh,), 0.1)
        ln(p + "attn.norm", dim)
        lin(p + "ff.fn.net.0.0", dim, dim)
        lin(p + "ff.fn.net.1", dim, dim)
        ln(p + "ff.norm", dim)

    def cross(p, dim, ctx_dim, heads, dh):
        lin(p + "attn.fn.to_q", heads * dh, dim, bias=False)
        lin(p + "attn.fn.to_kv", 2 * heads * dh, ctx_dim, bias=False)
        ln(p + "attn.norm_context", ctx_dim)
        attn_common(p, dim, heads * dh, dh)

    def selfb(p, dim, heads, dh):
        lin(p + "attn.fn.to_qkv", 3 * heads * dh, dim, bias=False)
        attn_common(p, dim, heads * dh, dh)

    # frozen orthogonal entity table (entity_embeddings.py:24-27): QR of a Gaussian, rows orthonormal
    n_ent = cfg["num_entities"]
    q, _ = torch.linalg.qr(_rand(g, (E, n_ent), 1.0))
    ent = q.t().contiguous()
    sd["encoder.entity_embedding.embedding.weight"] = ent
    sd["decoder.entity_embedding.embedding.weight"] = ent
    sd["encoder.latents"] = _rand(g, (e["num_latents"], D), 1.0)
    lin("encoder.mlp.0", D, C)
    lin("encoder.mlp.2", C, D)
    for i in range(e["num_block_cross"]):
        cross(f"encoder.cross_attn_blocks.{i}.", D, C, e["num_head_cross"], e["dim_head_cross"])
    for i in range(e["num_block_attn"]):
        selfb(f"encoder.blocks_attn.{i}.", D, e["num_head_latent"], e["dim_head_latent"])
    dq = d["dim_query"]
    lin("decoder.query_mlp.1", dq, E)
    for i in range(d["num_block_attn"]):
        selfb(f"decoder.self_attn_blocks.{i}.", D, d["num_head_latent"], d["dim_head_latent"])
    for i in range(d["num_block_cross"]):
        cross(f"decoder.cross_attn_blocks.{i}.", D, dq, d["num_head_cross"], d["dim_head_cross"])
    cross("decoder.output_block.", dq, D, d["num_head_cross"], d["dim_head_cross"])
    for name, out_dim in d["outputs"]:
        lin(f"decoder.output_layers.{name}.0", dq, dq)
        lin(f"decoder.output_layers.{name}.2", out_dim, dq)
    if d["kind"] == "DecoderQuerySplitter":
        n = d["num_split"]
        sd["decoder.extender.1.weight"] = _rand(g, (D * n, D, 1), D ** -0.5)
        sd["decoder.extender.1.bias"] = _rand(g, (D * n,), 0.02)
    lin("quant.0", D, D)
    lin("post_quant.1", D, D)
    kind = cfg["kind"]
    if kind == "peptide":
        sd["embedding_res.weight"] = _rand(g, (20, 64), 1.0)  # norms ≈ 8 ⇒ max_norm renorm is exercised
        sd["embed_res_pos.embeddings"] = _sincos_1d(cfg["max_res"], Din)
        feat = 64 + 42
    elif kind == "md17":
        sd["embed_entity.embedding.weight"] = ent
        sd["embed_atom.weight"] = _rand(g, (cfg["n_atom_types"], 64), 1.0)
        sd["embed_pos.basis"] = _point_basis(126)
        lin("embed_pos.mlp", 128, 126 + 3)
        feat = 64 + 128
    elif kind == "nba":
        sd["embed_entity.embedding.weight"] = ent
        sd["embed_team.weight"] = _rand(g, (3, 32), 1.0)
        sd["embed_group.weight"] = _rand(g, (2, 32), 1.0)
        feat = 2 + 32 + 32
    else:
        feat = 2
    lin("net_merge.0", Din, feat)
    lin("net_merge.2", Din, Din)
    return sd


def synthetic_batch(cfg: dict, B: int, seed: int, T: Optional[int] = None) -> Dict[str, Tensor]:
    """Synthetic batches of each config's shape — SURVEY.md §8(d) table (C1..C4).  All on CPU, seeded."""
    g = torch.Generator().manual_seed(seed)
    name = cfg["name"]
    T = cfg["T"] if T is None else T
    N = cfg["N"]
    n_ent = cfg["first_stage"]["num_entities"]
    batch: Dict[str, Tensor] = {}

    def perm_entities(n_valid_per_sample):
        ent = torch.zeros(B, N, dtype=torch.int64)
        for b in range(B):
            nv = int(n_valid_per_sample[b])
            ent[b, :nv] = torch.randperm(n_ent, generator=g)[:nv]
        return ent[:, None, :].expand(B, T, N).contiguous()

    if name == "peptide":
        batch["atom14_pos"] = torch.randn(B, T, N, 14, 3, generator=g)
        batch["aatype"] = torch.randint(0, 20, (B, 1, N), generator=g).expand(B, T, N).contiguous()
        batch["entities"] = torch.arange(N)[None, None, :].expand(B, T, N).contiguous()  # sampling.py:36
    elif name == "md17":
        batch["pos"] = torch.randn(B, T, N, 3, generator=g)
        z = torch.tensor(([6] * 9 + [8] * 4 + [1] * 8)[:N])  # aspirin C9H8O4
        batch["atom"] = z[None, None, :].expand(B, T, N).contiguous()
        batch["entities"] = perm_entities([N] * B)
        batch["attention_mask"] = torch.ones(B, T, N, dtype=torch.bool)
    elif name == "nba":
        batch["pos"] = torch.randn(B, T, N, 2, generator=g)
        batch["team"] = torch.tensor([0] + [1] * 5 + [2] * 5)[None, None, :].expand(B, T, N).contiguous()
        batch["group"] = torch.tensor([0] + [1] * 10)[None, None, :].expand(B, T, N).contiguous()
        batch["entities"] = perm_entities([N] * B)
        batch["attention_mask"] = torch.ones(B, T, N, dtype=torch.bool)
        batch["cond_scene"] = torch.randint(0, cfg["n_classes"], (B,), generator=g)
    elif name == "pedestrian":
        nv = torch.randint(1, N + 1, (B,), generator=g)
        pos = torch.randn(B, T, N, 2, generator=g)
        valid = torch.arange(N)[None, :] < nv[:, None]
        pos = pos * valid[:, None, :, None]
        batch["pos"] = pos
        batch["entities"] = perm_entities(nv) * valid[:, None, :]
        batch["attention_mask"] = pos[..., 0] != 0  # collate_functions.py:71-77
        batch["cond_scene"] = torch.randint(0, cfg["n_classes"], (B,), generator=g)
    else:
        raise ValueError(name)
    return batch


def state_checksum(sd: SD) -> float:
    """Order-independent fp64 checksum used to make sure seeded weights are identical across boxes."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double().flatten()
        w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 97 + 1.0
        tot += float((v * w).sum())
    return tot
