"""End-to-end sampling wrapper — mirrors ``SecondStageCondLightningBase`` (``lightning_base.py:167-263``) and the
dataset ``Wrapper`` / ``CondWrapper`` classes (``second_stage/{peptide,md17,nba,pedestrian}.py``) for the calls on the
sampling path: ``forward``, ``encode``, ``decode``, ``prepare_batch``, ``setup_conditioning`` and ``sample``.
Lightning / Hydra / EMA / losses / metrics are out of scope; weights come from ``load_state_dict``
with the reference's key names (``backbone.*``, ``first_stage_model.backbone.*``, ``vec_in_embedding.weight``).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .backbone import LatentSIV3
from .configs import get_config
from .first_stage import FirstStage
from .transport import CreateTransport, Sampler, Transport

_FRAME_KEYS = ("atom14_pos", "aatype", "pos", "atom", "team", "group", "entities", "attention_mask")


class _FirstStageModel(nn.Module):
    """Stands in for the frozen ``FirstStageLightningBase`` (``lightning_base.py:140-164``): ``.backbone`` + encode/decode."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.backbone = FirstStage(cfg)

    def encode(self, batch: Dict[str, Tensor]) -> Tensor:
        return self.backbone.encode(batch)

    def decode(self, latents: Tensor, entities: Tensor) -> Dict[str, Tensor]:
        return self.backbone.decode(z=latents, entities=entities)


class SecondStageSampler(nn.Module):
    def __init__(self, cfg: dict, sampling_method: str = "ODE",
                 sampling_kwargs: Dict[str, Any] = {"sampling_method": "euler", "num_steps": 10}):
        super().__init__()
        self.cfg = cfg
        bb = cfg["backbone"]
        self.backbone = LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"],
                                   num_heads=bb["num_heads"], vec_in_dim=bb.get("vec_in_dim"), mlp_ratio=bb["mlp_ratio"],
                                   theta=bb.get("theta", 10_000), normalize=bb.get("normalize", False), n_timesteps=cfg["T"])
        self.first_stage_model = _FirstStageModel(cfg["first_stage"])
        self.si: Transport = CreateTransport(path_type=cfg["path_type"], prediction=cfg["prediction"])()
        if cfg.get("n_classes"):  # CondWrapper (nba.py:254-263, pedestrian.py:242-251)
            self.vec_in_embedding = nn.Embedding(cfg["n_classes"], bb["vec_in_dim"])
        self.hparams = SimpleNamespace(cond_idx=list(cfg["cond_idx"]), mask_cond_mean=cfg["mask_cond_mean"],
                                       sampling_method=sampling_method, sampling_kwargs=dict(sampling_kwargs),
                                       n_timesteps=cfg["T"])
        # CUDA graphs (opt-in, ``use_cuda_graphs``): one captured graph of the whole sample() per input signature; the small
        # configurations (pedestrian: ~700 launches of a few microseconds each) are launch-bound without it
        self.use_cuda_graphs = False
        self.__dict__["_graphs"] = {}

    @classmethod
    def from_name(cls, name: str, **kw) -> "SecondStageSampler":
        return cls(get_config(name), **kw)

    @property
    def device(self) -> torch.device:
        return next(self.backbone.parameters()).device

    @property
    def transport(self) -> Transport:
        return self.si

    # lightning_base.py:173-174
    def forward(self, xt: Tensor, t: Tensor, **model_kwargs) -> Tensor:
        return self.backbone(x=xt, t=t, **model_kwargs)

    # second_stage/peptide.py:85-95 (and md17 / nba / pedestrian equivalents): "B T ... -> (B T) ..."
    @torch.no_grad()
    def encode(self, batch: Dict[str, Tensor]) -> Tensor:
        B = batch["entities"].shape[0]
        flat = {k: v.flatten(0, 1) for k, v in batch.items() if k in _FRAME_KEYS and torch.is_tensor(v)}
        latents = self.first_stage_model.encode(flat)
        return latents.unflatten(0, (B, -1))

    # second_stage/peptide.py:97-102: "(B T) L (A D) -> B T L A D"; others "(B T) L D -> B T L D"
    def decode(self, latents: Tensor, entities: Tensor, T: Optional[int] = None) -> Dict[str, Tensor]:
        preds = self.first_stage_model.decode(latents=latents, entities=entities)
        T = T or self.hparams.n_timesteps
        main = self.cfg["main_output"]
        pos = preds[main].unflatten(0, (-1, T))
        if main == "atom14_pos":
            pos = pos.unflatten(-1, (14, 3))
        return {main: pos}

    # lightning_base.py:240-263
    @torch.no_grad()
    def setup_conditioning(self, latents: Tensor) -> Tuple[Tensor, Tensor]:
        _lib.require_cuda(latents)
        latents = latents.to(torch.float32).contiguous()
        B, T, L, D = latents.shape
        x_cond = torch.empty_like(latents)
        mask = torch.empty(B, T, L, dtype=torch.int64, device=latents.device)
        c0, c1 = self.hparams.cond_idx
        with torch.cuda.device(latents.device):
            _lib.check(_lib.load().lamslide_setup_conditioning(latents.data_ptr(), x_cond.data_ptr(), mask.data_ptr(), B, T, L, D,
                                                               c0, c1, 1 if self.hparams.mask_cond_mean else 0,
                                                               _lib.current_stream_ptr()))
        return x_cond, mask

    # lightning_base.py:205-215 (+ CondWrapper.prepare_batch)
    @torch.no_grad()
    def prepare_batch(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        latents = self.encode(batch)
        x_cond, x_cond_mask = self.setup_conditioning(latents)
        batch["x1"] = latents
        batch["model_kwargs"] = {"x_cond": x_cond, "x_cond_mask": x_cond_mask}
        if hasattr(self, "vec_in_embedding") and "cond_scene" in batch:
            batch["model_kwargs"]["y"] = self.vec_in_embedding(batch["cond_scene"])
        return batch

    # lightning_base.py:217-238.  ``noise`` (optional) replaces torch.randn_like for reproducible parity tests.
    @torch.no_grad()
    def sample(self, batch: Dict[str, Tensor], noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
        if self.use_cuda_graphs:
            return self._sample_graphed(batch, noise)
        return self._sample_eager(batch, noise)

    def _graph_signature(self, batch: Dict[str, Tensor], noise: Optional[Tensor]):
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in batch.items() if isinstance(v, torch.Tensor)))
        hp = self.hparams
        return (sig, None if noise is None else tuple(noise.shape), str(self.device), hp.sampling_method,
                tuple(sorted((k, str(v)) for k, v in hp.sampling_kwargs.items())), tuple(hp.cond_idx), hp.mask_cond_mean)

    def _param_versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    @torch.no_grad()
    def _sample_graphed(self, batch: Dict[str, Tensor], noise: Optional[Tensor]) -> Dict[str, Tensor]:
        """``sample()`` replayed from a CUDA graph: the first call with a new input signature (shapes, dtypes, sampler settings)
        runs once eagerly (packs weights, sizes workspaces, sets kernel attributes), captures the whole call — encode, conditioning,
        every network evaluation and Euler update, decode — on static input buffers, and later calls copy their inputs into those
        buffers and replay.  The C ABI never allocates, synchronises or touches the default stream (include/lamslide.h), so the
        capture holds exactly the kernels of the eager call; the results are bit-identical to it.  Graphs are dropped when a
        parameter changes; the workspaces a graph was captured with are kept alive with it."""
        dev = self.device
        key = self._graph_signature(batch, noise)
        graphs = self.__dict__["_graphs"]
        versions = self._param_versions()
        ent = graphs.get(key)
        if ent is not None and ent["versions"] != versions:
            graphs.clear()
            ent = None
        tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
        if ent is None:
            with torch.cuda.device(dev):
                static_in = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in tensors.items()}
                static_noise = None if noise is None else torch.empty(noise.shape, dtype=torch.float32, device=dev)
                for k, v in tensors.items():
                    static_in[k].copy_(v, non_blocking=True)
                if noise is not None:
                    static_noise.copy_(noise, non_blocking=True)
                cur = torch.cuda.current_stream(dev)
                side = torch.cuda.Stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):  # warm-up outside the capture (torch's capture rules; also packs and sizes everything)
                    self._sample_eager(dict(static_in), static_noise)
                cur.wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._sample_eager(dict(static_in), static_noise)
                ent = {"graph": graph, "in": static_in, "noise": static_noise, "out": out, "versions": versions,
                       "keep": [self.backbone._ws.buf, self.first_stage_model.backbone._ws.buf]}
                graphs[key] = ent
        with torch.cuda.device(dev):
            for k, v in tensors.items():
                ent["in"][k].copy_(v, non_blocking=True)
            if noise is not None:
                ent["noise"].copy_(noise, non_blocking=True)
            ent["graph"].replay()
            return {k: v.clone() for k, v in ent["out"].items()}

    @torch.no_grad()
    def _sample_eager(self, batch: Dict[str, Tensor], noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
        sample_fn = Sampler(self.si).get_sample_fn(self.hparams.sampling_method, self.hparams.sampling_kwargs)
        B, T = batch["entities"].shape[:2]
        dev = self.device
        for key in list(batch.keys()):
            if isinstance(batch[key], torch.Tensor) and batch[key].device != dev:
                batch[key] = batch[key].to(dev, non_blocking=True)
        batch = self.prepare_batch(batch)
        model_kwargs = batch["model_kwargs"]
        x0 = torch.randn_like(model_kwargs["x_cond"]) if noise is None else noise.to(dev, non_blocking=True)
        latents = sample_fn(x0, self.forward, **model_kwargs)[-1]
        return self.decode(latents.flatten(0, 1), batch["entities"].flatten(0, 1), T=T)

    @torch.no_grad()
    def sample_from_latents(self, latents: Tensor, entities: Tensor, noise: Optional[Tensor] = None,
                            y: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """The part of ``sample()`` after the encoder (lightning_base.py:209-238): conditioning from the given first-stage
        latents ``[B, T, L, D]`` (device), ODE, decode with ``entities [B, T, N]``.  Used by callers that can build the
        latents cheaper than by encoding B·T frames — e.g. the roll-out driver, whose T frames are copies of one frame."""
        sample_fn = Sampler(self.si).get_sample_fn(self.hparams.sampling_method, self.hparams.sampling_kwargs)
        B, T = latents.shape[:2]
        x_cond, x_cond_mask = self.setup_conditioning(latents)
        model_kwargs = {"x_cond": x_cond, "x_cond_mask": x_cond_mask}
        if y is not None:
            model_kwargs["y"] = y
        x0 = torch.randn_like(x_cond) if noise is None else noise.to(self.device, non_blocking=True)
        out = sample_fn(x0, self.forward, **model_kwargs)[-1]
        return self.decode(out.flatten(0, 1), entities.flatten(0, 1), T=T)

    def sample_stream(self, batches, noise: Optional[Tensor] = None, on_device=None):
        """``sample()`` over an iterable of HOST batches (pinned memory), yielding pinned host tensors of the main output.
        Same per-batch work as ``sample()`` (lightning_base.py:217-238), but the host->device copy of batch k + 1 and the
        device->host copy of result k - 1 run on a second CUDA stream while batch k is being computed, so at steady state the
        PCIe transfers cost no device time.  Results come out in order.  Three pinned result buffers rotate: the copy of result
        k + 1 is already in flight when result k is yielded, so a yielded tensor stays untouched until the consumer has asked for
        two more results (``next()`` twice) — keep the previous result while working on the current one, clone anything kept longer.  ``on_device(out)`` (optional) runs on the
        compute stream right after each ``sample()`` with the device result (e.g. the all-gather of ``dist.sample_stream_sharded``)."""
        dev = self.device
        main_key = self.cfg["main_output"]
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            # copy stream, two rotating sets of device input buffers and pinned result buffers, kept on the module between calls:
            # no allocation at steady state (the caching allocator would otherwise cudaMalloc while blocks shared between the
            # two streams are still in flight, and a pinned allocation costs milliseconds)
            pipe = self.__dict__.setdefault("_pipe", {"side": torch.cuda.Stream(dev), "in": [None, None], "host": [None, None, None]})
            side, in_bufs, host_bufs = pipe["side"], pipe["in"], pipe["host"]
            in_free = [None, None]
            side.wait_stream(main)

            def upload(batch, slot):
                with torch.cuda.stream(side):
                    if in_free[slot] is not None:
                        side.wait_event(in_free[slot])  # the compute that read this buffer set has finished
                    tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
                    cur = in_bufs[slot]
                    if cur is None or any(k not in cur or cur[k].shape != v.shape or cur[k].dtype != v.dtype for k, v in tensors.items()):
                        cur = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in tensors.items()}
                        for t in cur.values():
                            t.record_stream(main)
                        in_bufs[slot] = cur
                    for k, v in tensors.items():
                        cur[k].copy_(v, non_blocking=True)
                    d = dict(batch)
                    d.update({k: cur[k] for k in tensors})
                    return d, side.record_event()

            it = iter(batches)
            try:
                nxt = upload(next(it), 0)
            except StopIteration:
                return
            pending, k = None, 0
            while nxt is not None:
                cur, ready = nxt
                try:
                    nxt = upload(next(it), (k + 1) & 1)
                except StopIteration:
                    nxt = None
                main.wait_event(ready)
                out = self.sample(cur, noise=noise)[main_key]
                if on_device is not None:
                    on_device(out)
                done = main.record_event()
                in_free[k & 1] = done
                hb = k % 3
                if host_bufs[hb] is None or host_bufs[hb].shape != out.shape or host_bufs[hb].dtype != out.dtype:
                    host_bufs[hb] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                with torch.cuda.stream(side):
                    side.wait_event(done)
                    host_bufs[hb].copy_(out, non_blocking=True)
                    copied = side.record_event()
                if pending is not None:
                    pending[1].synchronize()
                    yield pending[0]
                pending = (host_bufs[hb], copied, out)  # `out` stays referenced until its copy has completed
                k += 1
            pending[1].synchronize()
            yield pending[0]
