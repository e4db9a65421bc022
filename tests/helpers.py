"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import os

import torch

from oracle import lamslide_oracle as O
from oracle.make_golden import CASES, GOLDEN_DIR, case_inputs

CASE_BY_NAME = {c["case"]: c for c in CASES}


def load_golden(name: str) -> dict:
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def check_inputs_match_fixture(fx: dict, fs_sd, bb_sd, batch, noise) -> None:
    """The golden files store seeds, not tensors: make an RNG difference a loud failure."""
    cs = fx["checksums"]
    assert abs(O.state_checksum(fs_sd) - cs["fs"]) <= 1e-6 * max(1.0, abs(cs["fs"]))
    assert abs(O.state_checksum(bb_sd) - cs["bb"]) <= 1e-6 * max(1.0, abs(cs["bb"]))
    got = O.state_checksum({k: v.float() for k, v in batch.items()})
    assert abs(got - cs["batch"]) <= 1e-6 * max(1.0, abs(cs["batch"]))
    assert abs(float(noise.double().sum()) - cs["noise"]) <= 1e-6 * max(1.0, abs(cs["noise"]))


def frame_slice(fx: dict) -> slice:
    return slice(*fx["frame_slice"])


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| — the "max relative error" of BASELINE.json's north_star tolerance."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def mean_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """mean |a-b| / mean |b| — bounds the error where the reference is small too (a localized error in a low-magnitude region does
    not move max_rel, which is normalised by the largest reference value)."""
    return float((a.double() - b.double()).abs().mean() / b.double().abs().mean().clamp_min(1e-30))


def rmsd(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a.double() - b.double()) ** 2).mean().sqrt())
