"""TEST INFRASTRUCTURE — loader for the *real* reference (``/root/reference``) in the dev container.

This file is only used (a) by ``oracle/make_golden.py`` to generate the committed golden vectors in
``tests/golden/`` and (b) by ``tests/test_oracle_vs_reference.py`` (skipped when the reference tree
is absent, e.g. on the GPU box).  Nothing in the product package imports it.

What it does
------------
* puts ``/root/reference`` on ``sys.path`` so ``src.models.components.*`` / ``src.modules.*`` import as-is
  (mmdit, latent_si_v31, torch_modules, encoder, decoder, entity_embeddings, embeddings);
* installs a ~10 line ``torchdiffeq.odeint`` shim (fixed-grid explicit Euler; the only method any shipped
  ``sampling_kwargs`` default uses: ``second_stage/*.py`` ``{"sampling_method": "euler", "num_steps": 10}``)
  so ``src.modules.transport`` imports (call site ``src/modules/transport/integrators.py:4,119``);
* composes the reference components into the four first-stage backbones following
  ``src/models/composites/lightning_base.py:17-48`` and ``first_stage/{peptide.py:23-103, md17.py:21-58,
  nba.py:23-59, pedestrian.py:16-42}`` — those files import lightning/hydra/torchmetrics, which are not in
  this image, so the ~15 lines of glue per dataset are re-stated here around the reference's own modules;
* re-states ``SecondStageCondLightningBase.setup_conditioning / sample`` (``lightning_base.py:217-263``) as
  ``reference_sample`` around the reference ``Sampler``/``Transport``.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, Optional

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("LAMSLIDE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models", "components"))


def _install_torchdiffeq_shim() -> None:
    if "torchdiffeq" in sys.modules:
        return
    mod = types.ModuleType("torchdiffeq")

    def odeint(func, y0, t, *, method="euler", atol=None, rtol=None, **_):
        # torchdiffeq fixed-grid solver semantics for method="euler": grid == t,
        # y_{i+1} = y_i + (t_{i+1}-t_i) f(t_i, y_i); returns the state at every grid point.
        if method != "euler":
            raise NotImplementedError(f"shim only provides fixed-grid euler, got {method}")
        ys = [y0]
        y = y0
        for i in range(len(t) - 1):
            y = y + (t[i + 1] - t[i]) * func(t[i], y)
            ys.append(y)
        return torch.stack(ys)

    mod.odeint = odeint
    sys.modules["torchdiffeq"] = mod


def load_reference():
    """Import the reference modules; returns a namespace with the classes used on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _install_torchdiffeq_shim()
    ns = types.SimpleNamespace()
    from src.models.components.latent.latent_si_v31 import LatentSIV3  # noqa
    from src.models.components.encoder import Encoder  # noqa
    from src.models.components.decoder import Decoder, DecoderQuerySplitter  # noqa
    from src.modules.entity_embeddings import EntityEmbeddingOrthogonal  # noqa
    from src.modules.embeddings import PointEmbed, SinCosPositionalEmbedding1D  # noqa
    from src.modules.torch_modules import GELU  # noqa
    from src.modules.transport import CreateTransport  # noqa
    from src.modules.transport.transport import Sampler  # noqa

    ns.LatentSIV3 = LatentSIV3
    ns.Encoder = Encoder
    ns.Decoder = Decoder
    ns.DecoderQuerySplitter = DecoderQuerySplitter
    ns.EntityEmbeddingOrthogonal = EntityEmbeddingOrthogonal
    ns.PointEmbed = PointEmbed
    ns.SinCosPositionalEmbedding1D = SinCosPositionalEmbedding1D
    ns.GELU = GELU
    ns.CreateTransport = CreateTransport
    ns.Sampler = Sampler
    return ns


class RefFirstStage(nn.Module):
    """Reference components wired as ``BackboneBase`` + the dataset ``Backbone`` (state-dict keys identical
    to ``first_stage_model.backbone.*`` of the reference checkpoints)."""

    def __init__(self, cfg: dict):
        super().__init__()
        ref = load_reference()
        self.cfg = cfg
        kind = cfg["kind"]
        act = ref.GELU
        ent = ref.EntityEmbeddingOrthogonal(cfg["num_entities"], cfg["entity_dim"], max_norm=1)
        e = cfg["encoder"]
        d = cfg["decoder"]
        self.encoder = ref.Encoder(
            dim_input=cfg["dim_input"], dim_latent=cfg["dim_latent"],
            dim_head_cross=e["dim_head_cross"], dim_head_latent=e["dim_head_latent"],
            num_latents=e["num_latents"], num_head_cross=e["num_head_cross"],
            num_head_latent=e["num_head_latent"], num_block_cross=e["num_block_cross"],
            num_block_attn=e["num_block_attn"], qk_norm=e["qk_norm"], entity_embedding=ent, act=act)
        dec_cls = ref.DecoderQuerySplitter if d["kind"] == "DecoderQuerySplitter" else ref.Decoder
        kw = dict(
            outputs=dict(d["outputs"]), dim_query=d["dim_query"], dim_latent=cfg["dim_latent"],
            entity_embedding=ent, dim_head_cross=d["dim_head_cross"], dim_head_latent=d["dim_head_latent"],
            num_head_cross=d["num_head_cross"], num_head_latent=d["num_head_latent"],
            num_block_cross=d["num_block_cross"], num_block_attn=d["num_block_attn"],
            dropout_query=0.1, dropout_latent=0.0, qk_norm=d["qk_norm"], act=act)
        if d["kind"] == "DecoderQuerySplitter":
            kw["num_split"] = d["num_split"]
        self.decoder = dec_cls(**kw)
        D = cfg["dim_latent"]
        # lightning_base.py:24-31
        self.quant = nn.Sequential(nn.Linear(D, D), nn.LayerNorm(D, elementwise_affine=False))
        self.post_quant = nn.Sequential(nn.LayerNorm(D, elementwise_affine=False), nn.Linear(D, D))
        Din = cfg["dim_input"]
        if kind == "peptide":  # first_stage/peptide.py:36-56
            self.embedding_res = nn.Embedding(20, 64, max_norm=1)
            self.embed_res_pos = ref.SinCosPositionalEmbedding1D(n_positions=cfg["max_res"], embed_dim=Din)
            feat = 64 + 42
        elif kind == "md17":  # first_stage/md17.py:34-50
            self.embed_entity = ent
            self.embed_atom = nn.Embedding(cfg["n_atom_types"], 64, max_norm=1)
            self.embed_pos = ref.PointEmbed(embedding_dim=128, hidden_dim=126)
            feat = 64 + 128
        elif kind == "nba":  # first_stage/nba.py:36-52
            self.embed_entity = ent
            self.embed_team = nn.Embedding(3, 32)
            self.embed_group = nn.Embedding(2, 32)
            feat = 2 + 32 + 32
        elif kind == "pedestrian":  # first_stage/pedestrian.py:27-37
            feat = 2
        else:
            raise ValueError(kind)
        self.net_merge = nn.Sequential(nn.Linear(feat, Din), act(), nn.Linear(Din, Din))

    def prepare_inputs(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        kind = self.cfg["kind"]
        if kind == "peptide":  # peptide.py:96-103
            pos = batch["atom14_pos"].flatten(-2)
            x = torch.cat([self.embedding_res(batch["aatype"]), pos], dim=-1)
            return self.embed_res_pos(self.net_merge(x))
        if kind == "md17":  # md17.py:52-58
            x = torch.cat([self.embed_atom(batch["atom"]), self.embed_pos(batch["pos"])], dim=-1)
            return self.net_merge(x)
        if kind == "nba":  # nba.py:54-59
            x = torch.cat([batch["pos"], self.embed_team(batch["team"]), self.embed_group(batch["group"])], dim=-1)
            return self.net_merge(x)
        return self.net_merge(batch["pos"])  # pedestrian.py:39-42

    def encode(self, batch):  # lightning_base.py:37-40 (peptide.py:77-80 passes mask=None)
        x = self.prepare_inputs(batch)
        mask = None if self.cfg["kind"] == "peptide" else batch.get("attention_mask")
        return self.quant(self.encoder(x=x, entities=batch["entities"], mask=mask))

    def decode(self, z, entities):  # lightning_base.py:42-44
        return self.decoder(self.post_quant(z), entities)


def reference_setup_conditioning(latents, cond_idx, mask_cond_mean=True):
    """lightning_base.py:240-263."""
    B, T, L, _ = latents.shape
    m = torch.zeros(B, T, L, dtype=torch.int64)
    m[:, cond_idx[0]:cond_idx[1]] = 1
    if mask_cond_mean:
        x_cond = torch.where(m.unsqueeze(-1).bool(), latents,
                             latents[:, cond_idx[0]:cond_idx[1]].mean(dim=1).unsqueeze(1))
    else:
        x_cond = torch.where(m.unsqueeze(-1).bool(), latents, torch.zeros(()))
    return x_cond, m


@torch.no_grad()
def reference_sample(first_stage: RefFirstStage, backbone: nn.Module, batch: Dict[str, torch.Tensor],
                     *, cond_idx, path_type="GVP", prediction="data", num_steps=10, noise: torch.Tensor,
                     y: Optional[torch.Tensor] = None, mask_cond_mean=True, record: Optional[dict] = None):
    """lightning_base.py:205-238 around the reference Sampler; ``noise`` replaces ``randn_like`` so the
    CUDA path can be driven with identical initial noise. ``record`` (optional dict) receives the latents,
    conditioning, every ODE state and every drift ("velocity") evaluation."""
    ref = load_reference()
    B, T = batch["entities"].shape[:2]
    flat = {k: v.flatten(0, 1) for k, v in batch.items()
            if torch.is_tensor(v) and v.dim() >= 2 and k != "cond_scene"}
    latents = first_stage.encode(flat).unflatten(0, (B, T))
    x_cond, x_mask = reference_setup_conditioning(latents, cond_idx, mask_cond_mean)
    si = ref.CreateTransport(path_type=path_type, prediction=prediction)()
    sampler = ref.Sampler(si)
    vel = []
    if record is not None:
        inner = sampler.drift

        def spy(x, t, model, **kw):
            v = inner(x, t, model, **kw)
            vel.append(v.clone())
            return v

        sampler.drift = spy
    fn = sampler.get_sample_fn("ODE", {"sampling_method": "euler", "num_steps": num_steps})
    kw = dict(x_cond=x_cond, x_cond_mask=x_mask)
    if y is not None:
        kw["y"] = y
    states = fn(noise, lambda xt, t, **k: backbone(x=xt, t=t, **k), **kw)
    out = first_stage.decode(states[-1].flatten(0, 1), batch["entities"].flatten(0, 1))
    out = {k: v.unflatten(0, (B, T)) for k, v in out.items()}
    if record is not None:
        record.update(latents=latents, x_cond=x_cond, x_cond_mask=x_mask, states=states,
                      velocities=torch.stack(vel))
    return out


def load_reference_rollout_wrapper():
    """The reference's OWN ``SIAtom14SamplingWrapper`` (``src/modules/sampling.py:16-63``).  Its module cannot be imported here
    (mdtraj, MDAnalysis, lightning ... are not in the image), so the class is compiled straight from the reference file —
    only ``__init__``, ``create_batch`` and ``sample_rollout`` (the methods with no trajectory-file dependency) — with the
    three names they use (torch, einops.repeat, tqdm) in scope.  Nothing is copied into this repository."""
    import ast

    from einops import repeat
    from tqdm import tqdm
    path = os.path.join(REFERENCE_ROOT, "src", "modules", "sampling.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SIAtom14SamplingWrapper")
    cls.body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("__init__", "create_batch", "sample_rollout")]
    mod = ast.Module(body=[cls], type_ignores=[])
    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": Dict, "repeat": repeat, "tqdm": tqdm}
    exec(compile(mod, path, "exec"), ns)  # nosec B102 - the reference's own source, dev container only
    return ns["SIAtom14SamplingWrapper"]


class RefRolloutModel:
    """What ``SIAtom14SamplingWrapper`` needs from the Lightning wrapper: ``hparams.n_timesteps``, ``shift``, ``scale`` and
    ``sample(batch) -> {"atom14_pos": [B, T, R, 14, 3]}`` (``second_stage/peptide.py:97-102`` reshapes the decoder's 42 columns).
    ``noises`` replaces the ``randn_like`` of successive ``sample()`` calls."""

    def __init__(self, first_stage, backbone, cfg, noises, shift, scale, num_steps):
        self.fs, self.net, self.cfg, self.noises = first_stage, backbone, cfg, list(noises)
        self.hparams = types.SimpleNamespace(n_timesteps=cfg["T"])
        self.shift, self.scale, self.num_steps, self.calls = shift, scale, num_steps, 0

    def sample(self, batch):
        out = reference_sample(self.fs, self.net, batch, cond_idx=self.cfg["cond_idx"], path_type=self.cfg["path_type"],
                               prediction=self.cfg["prediction"], num_steps=self.num_steps, noise=self.noises[self.calls],
                               mask_cond_mean=self.cfg["mask_cond_mean"])
        self.calls += 1
        return {"atom14_pos": out["atom14_pos"].unflatten(-1, (14, 3))}


def load_reference_method(relpath: str, class_name: str, method_name: str, extra_ns: Optional[dict] = None):
    """One method of a reference class as a plain function ``f(self, ...)``, compiled from the reference file in place — for
    classes whose modules cannot be imported here (Lightning / hydra / torchmetrics / torch_kmeans are not in the image)."""
    import ast

    from einops import rearrange
    path = os.path.join(REFERENCE_ROOT, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == method_name)
    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": Dict, "rearrange": rearrange}
    ns.update(extra_ns or {})
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)  # nosec B102 - the reference's own source
    return ns[method_name]


class RefTestStepSelf:
    """The attributes ``Wrapper.test_step`` touches (second_stage/nba.py:161-225, md17.py:139-171): ``hparams``, the datamodule's
    dataloader name, the output dict, and ``sample(batch)`` — which here hands out preset predictions ``[B, T, A, D]`` in
    the reference's ``(B T) A D`` layout."""

    def __init__(self, preds, K, cond_idx, num_runs):
        self.hparams = types.SimpleNamespace(K=K, cond_idx=list(cond_idx), num_runs=num_runs, post_process=False)
        self.trainer = types.SimpleNamespace(datamodule=types.SimpleNamespace(dataloader_names=lambda idx: "test"))
        self.test_step_outputs: dict = {}
        self._preds, self.calls, self.seen = list(preds), 0, []

    def sample(self, batch):
        self.seen.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
        out = self._preds[self.calls].flatten(0, 1)
        self.calls += 1
        return {"pos": out}

