"""lam_slide_b200 — B200-native (sm_100a) implementation of LaM-SLidE's sampling hot path.

Public surface (mirrors the reference's interfaces for this path; see DESIGN.md / INTEGRATION.md):
    LatentSIV3            second-stage latent transformer          (latent_si_v31.py)
    FirstStage            BackboneBase + Encoder/Decoder           (lightning_base.py, encoder.py, decoder.py)
    CreateTransport, Sampler                                       (src/modules/transport)
    SecondStageSampler    encode -> conditioning -> Euler ODE -> decode  (SecondStageCondLightningBase.sample)
    SIAtom14SamplingWrapper  autoregressive roll-out driver, batched on device  (src/modules/sampling.py)
    KSampleEvaluator      K-sample min/mean ADE-FDE evaluation, one batched solve  (second_stage/{nba,pedestrian,md17}.py test_step)
    odeint                dopri5 / bosh3 / adaptive_heun / midpoint / rk4 / heun behind Sampler.sample_ode  (torchdiffeq call of integrators.py)
    formats               atom14 -> atom37 -> heavy-atom topology -> PDB / DCD  (sampling.py:64-142, geometry.py:14-33)
"""
from .backbone import LatentSIV3  # noqa: F401
from .configs import CONFIGS, get_config  # noqa: F401
from .first_stage import FirstStage  # noqa: F401
from .model import SecondStageSampler  # noqa: F401
from .rollout import SIAtom14SamplingWrapper  # noqa: F401
from .evaluation import KSampleEvaluator, ksample_errors  # noqa: F401
from .checkpoint import load_checkpoint, select_state_dict  # noqa: F401
from .transport import CreateTransport, Sampler, Transport  # noqa: F401
from . import formats, odeint  # noqa: F401
