"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/lamslide.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from lam_slide_b200.build import build
    return build()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lamslide.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lamslide_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = _header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lamslide.h but not exported"


def test_python_binding_lists_the_same_symbols(lib_path):
    from lam_slide_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _header_symbols()
    assert _lib.load().lamslide_abi_version() == 1


def test_sass_contains_blackwell_tensor_and_tma_instructions(lib_path):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma / TMA load / tcgen05.ld (B200_PROFILING.md)
        assert mnemonic in sass, mnemonic


def test_state_dict_keys_match_reference_names():
    import lam_slide_b200 as P
    from oracle import lamslide_oracle as O
    for name in ("peptide", "md17", "nba", "pedestrian"):
        cfg = P.get_config(name)
        m = P.SecondStageSampler(cfg)
        fs_sd = O.init_first_stage_params(cfg["first_stage"], 0)
        bb_sd = O.init_backbone_params(cfg["backbone"], 0)
        m.first_stage_model.backbone.load_state_dict(fs_sd, strict=True)
        m.backbone.load_state_dict(bb_sd, strict=True)


def test_no_cpu_fallback():
    import torch
    import lam_slide_b200 as P
    from lam_slide_b200._lib import LamSlideError
    net = P.LatentSIV3(depth=1, in_dim=32, hidden_size=128, num_heads=4)
    x = torch.randn(1, 4, 2, 32)
    with pytest.raises(LamSlideError):
        net(x, torch.zeros(1), x, torch.zeros(1, 4, 2, dtype=torch.long))


def test_transport_interval_and_api():
    import lam_slide_b200 as P
    tr = P.CreateTransport(path_type="GVP", prediction="data")()
    assert tr.check_interval(tr.train_eps, tr.sample_eps, eval=True) == (1e-3, 1 - 1e-3)
    tr = P.CreateTransport()()
    assert tr.check_interval(tr.train_eps, tr.sample_eps, eval=True) == (0, 1)
    assert callable(P.Sampler(tr).get_sample_fn("SDE", {}))  # Euler-Maruyama with the reference's defaults (transport.py:480-487)
    with pytest.raises(NotImplementedError):
        P.Sampler(tr).get_sample_fn("SDE", {"sampling_method": "Milstein"})
    assert callable(P.Sampler(tr).get_sample_fn("ODE", {}))  # the reference's default: dopri5, num_steps 50, atol 1e-6, rtol 1e-3 (transport.py:365-372)
    assert callable(P.Sampler(tr).get_sample_fn("ODE", {"sampling_method": "dopri5"}))  # configs/eval_peptide.yaml:21-23
    with pytest.raises(NotImplementedError):
        P.Sampler(tr).get_sample_fn("ODE", {"sampling_method": "dopri8"})
    fn = P.Sampler(tr).get_sample_fn("ODE", {"sampling_method": "euler", "num_steps": 10})
    assert callable(fn)
