"""Subset of the reference ``BaseOptions`` (src/nodes/options/base_options.py:10-60) that the FMT sampling path reads.
Field names and defaults are the reference's, so an options object from either side is interchangeable here."""
from dataclasses import dataclass

TORCHDIFFEQ_FIXED_STEP_SOLVERS = ["euler", "midpoint", "rk4", "heun2", "heun3"]   # src/nodes/__init__.py:15-23


@dataclass
class BaseOptions:
    seed: int = 15
    fix_noise_seed: bool = True
    fps: float = 25.0
    sampling_rate: int = 16000
    wav2vec_sec: float = 2.0
    attention_window: int = 2
    audio_dropout_prob: float = 0.1
    ref_dropout_prob: float = 0.1
    emotion_dropout_prob: float = 0.1
    dim_a: int = 512
    dim_w: int = 512
    dim_h: int = 1024
    dim_e: int = 7
    fmt_depth: int = 8
    num_heads: int = 8
    mlp_ratio: float = 4.0
    num_prev_frames: int = 10
    ode_atol: float = 1e-5
    ode_rtol: float = 1e-5
    nfe: int = 10
    torchdiffeq_ode_method: str = "euler"
    a_cfg_scale: float = 2.0
    e_cfg_scale: float = 1.0
    r_cfg_scale: float = 1.0
    cudnn_benchmark_enabled: bool = False


class FmtModel:
    """Minimal stand-in for a loaded reference ``FlowMatchingTransformer`` as ``LoadFMTModel`` hands it to the sampler
    (nodes_vadv_loader.py:840-866): ``state_dict()``, ``.opt``, ``.final_construction_options``, ``.target_device``.
    Lets the sampler node run from a bare state dict (tests, benchmarks, non-ComfyUI hosts)."""

    def __init__(self, state_dict, opt: BaseOptions = None, target_device="cuda"):
        self._sd = dict(state_dict)
        self.opt = opt or BaseOptions()
        self.final_construction_options = {k: getattr(self.opt, k) for k in vars(self.opt)}
        self.target_device = target_device
        self.cudnn_benchmark_setting = self.opt.cudnn_benchmark_enabled

    def state_dict(self):
        return self._sd
