"""Dev-container only: the oracle restatement against the REAL reference modules imported from /root/reference
(skipped where the reference tree does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import lamslide_oracle as O
from oracle.ref_loader import reference_available
from tests.helpers import max_rel

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("name,B,T", [("peptide", 2, 12), ("md17", 1, 5), ("nba", 2, 20), ("pedestrian", 4, 20)])
def test_sample_matches_reference_modules(name, B, T):
    from lam_slide_b200.configs import get_config
    from oracle.ref_loader import RefFirstStage, load_reference, reference_sample
    ref = load_reference()
    cfg = get_config(name, depth=2)
    bb = cfg["backbone"]
    fs_sd = O.init_first_stage_params(cfg["first_stage"], 11)
    bb_sd = O.init_backbone_params(bb, 12)
    fs = RefFirstStage(cfg["first_stage"]).eval()
    fs.load_state_dict(fs_sd, strict=True)  # pins the state-dict key names of SURVEY §8(b)
    net = ref.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                         vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"]).eval()
    net.load_state_dict(bb_sd, strict=True)
    batch = O.synthetic_batch(cfg, B, 5, T=T)
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(B, T, cfg["first_stage"]["encoder"]["num_latents"], bb["in_dim"], generator=g)
    y = torch.randn(cfg["n_classes"], 256, generator=g)[batch["cond_scene"]] if cfg["n_classes"] else None
    r1, r2 = {}, {}
    with torch.no_grad():
        out_r = reference_sample(fs, net, {k: v.clone() for k, v in batch.items()}, cond_idx=cfg["cond_idx"], noise=noise, y=y, record=r1)
        out_o = O.sample(fs_sd, bb_sd, cfg, batch, noise, y=y, record=r2)
    assert max_rel(r2["latents"], r1["latents"]) < 1e-5
    assert max_rel(r2["velocities"], r1["velocities"]) < 2e-4
    assert max_rel(r2["states"], r1["states"]) < 1e-4
    for k in out_r:
        assert max_rel(out_o[k], out_r[k]) < 1e-4


def test_reference_default_init_is_zero_output():
    """latent_si_v31.py:152-156 zero-inits modulation.lin and the head ⇒ parity on default init would be vacuous."""
    from oracle.ref_loader import load_reference
    ref = load_reference()
    net = ref.LatentSIV3(depth=1, in_dim=8, hidden_size=32, num_heads=2).eval()
    x = torch.randn(1, 3, 2, 8)
    with torch.no_grad():
        out = net(x=x, t=torch.tensor([0.5]), x_cond=x, x_cond_mask=torch.zeros(1, 3, 2, dtype=torch.int64))
    assert float(out.abs().max()) == 0.0


def test_checkpoint_ingestion_against_reference_ema_class():
    """lam_slide_b200.load_checkpoint on a checkpoint assembled the way the reference's LightningModule writes it
    (lightning_base.py:114-119: ``checkpoint["ema"] = self.ema.state_dict()``) with the reference's OWN ExponentialMovingAverage
    (src/modules/ema.py, compiled from the reference file: its module imports lightning) around the reference's own modules."""
    import ast
    import os
    from collections import OrderedDict

    import torch.nn as nn

    import lam_slide_b200 as P
    from oracle.ref_loader import REFERENCE_ROOT, RefFirstStage, load_reference
    ref = load_reference()
    cfg = P.get_config("pedestrian", depth=1)
    bb = cfg["backbone"]

    class FirstStageModel(nn.Module):  # FirstStageLightningBase: .backbone
        def __init__(self):
            super().__init__()
            self.backbone = RefFirstStage(cfg["first_stage"])

    class Lightning(nn.Module):  # the attribute names of SecondStageCondLightningBase / CondWrapper (pedestrian.py:242-251)
        def __init__(self):
            super().__init__()
            self.backbone = ref.LatentSIV3(depth=bb["depth"], in_dim=bb["in_dim"], hidden_size=bb["hidden_size"], num_heads=bb["num_heads"],
                                           vec_in_dim=bb["vec_in_dim"], mlp_ratio=bb["mlp_ratio"], normalize=bb["normalize"])
            self.first_stage_model = FirstStageModel()
            self.vec_in_embedding = nn.Embedding(cfg["n_classes"], bb["vec_in_dim"])

    torch.manual_seed(3)
    lm = Lightning()
    path = os.path.join(REFERENCE_ROOT, "src", "modules", "ema.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExponentialMovingAverage")
    ns = {"torch": torch, "nn": nn, "OrderedDict": OrderedDict,
          "tensor_tree_map": lambda fn, tree_: OrderedDict((k, fn(v)) for k, v in tree_.items())}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)  # nosec B102 - the reference's own source
    ema = ns["ExponentialMovingAverage"](model=lm, decay=0.9)
    with torch.no_grad():
        for p_ in lm.parameters():
            p_.add_(0.1 * torch.randn_like(p_))
    ema.update(lm)  # EMA parameters now differ from the raw ones
    ckpt = {"epoch": 3, "state_dict": lm.state_dict(), "ema": ema.state_dict()}

    m = P.SecondStageSampler(cfg)
    P.load_checkpoint(m, ckpt, use_ema=True)
    for k, v in m.backbone.state_dict().items():
        assert torch.equal(v, ema.params[f"backbone.{k}"]), k
        assert not torch.equal(v, lm.state_dict()[f"backbone.{k}"]) or v.numel() == 0 or float(v.abs().sum()) == 0.0
    for k, v in m.first_stage_model.backbone.state_dict().items():
        assert torch.equal(v, ema.params[f"first_stage_model.backbone.{k}"]), k
    assert torch.equal(m.vec_in_embedding.weight.data, ema.params["vec_in_embedding.weight"])
    P.load_checkpoint(m, ckpt, use_ema=False)
    for k, v in m.backbone.state_dict().items():
        assert torch.equal(v, lm.state_dict()[f"backbone.{k}"]), k
