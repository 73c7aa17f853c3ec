"""Import the UNMODIFIED reference (``src/nodes`` of set-soft/ComfyUI-FLOAT_Optimized).

Test infrastructure only (tests/, ``bench.py --impl reference`` / ``cpu_baseline``, tools/psnr_*).  The reference tree is looked
for in ``$FLOAT_REFERENCE_ROOT``, then ``/root/reference`` (the build container), then ``oracle/_ref/reference`` - the
git-ignored copy ``oracle/make_ref.py`` makes at build time so that the GPU box, where ``/root/reference`` does not exist,
can time and decode with the real reference code.  The reference is pure Python but depends on packages that are
not installed here (``seconohe``, ``comfy``, ``folder_paths``, ``timm``, ``torchdiffeq``,
``librosa``, ``face_alignment``).  This module injects minimal stand-ins for those into
``sys.modules`` and then loads the reference package under the name
``refnodes``.  The reference files are executed unmodified.

The only arithmetic the stand-ins carry is

* ``timm.models.vision_transformer.Mlp`` - fc1 -> act -> fc2 (timm>=1.0.9, used at
  ``FMT.py:162``), and ``timm.layers.use_fused_attn`` -> True (``FMT.py:60``);
* ``torchdiffeq.odeint`` - the fixed-grid solvers the node exposes
  (``src/nodes/__init__.py:15-23``): for every consecutive pair of the given time grid
  ``y1 = y0 + step(f, t0, dt, y0)``; results are exact at the grid points, ``rtol/atol``
  are ignored by fixed-grid solvers (torchdiffeq ``fixed_grid.py`` / ``rk_common.py``).

It is used by ``tests/golden/make_golden.py`` to produce the committed fixtures, by the ``refimpl``-marked tests, by the
reference arm of ``bench.py`` and by the PSNR gate (reference decoder).
"""
import contextlib
import importlib.machinery
import importlib.util
import logging
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref_root() -> str:
    cands = [os.environ.get("FLOAT_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref", "reference")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "nodes", "models", "float", "FMT.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_ref_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "nodes", "models", "float", "FMT.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Mlp(nn.Module):
    """timm Mlp semantics: fc1 -> act -> drop -> norm(Identity) -> fc2 -> drop."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0., use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


def _fixed_grid_odeint(func, y0, t, *, method="euler", rtol=None, atol=None, options=None, **_):
    """Stand-in for torchdiffeq.odeint restricted to fixed-grid solvers (grid == t)."""
    def euler(f, t0, dt, t1, y):
        return dt * f(t0, y)

    def midpoint(f, t0, dt, t1, y):
        half = 0.5 * dt
        return dt * f(t0 + half, y + f(t0, y) * half)

    def rk4(f, t0, dt, t1, y):   # torchdiffeq's fixed-grid rk4 is the 3/8-rule variant
        k1 = f(t0, y)
        k2 = f(t0 + dt / 3, y + dt * k1 / 3)
        k3 = f(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
        k4 = f(t1, y + dt * (k1 - k2 + k3))
        return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125

    def heun2(f, t0, dt, t1, y):  # tableau alpha=[1], beta=[[1]], c_sol=[1/2, 1/2]
        k1 = f(t0, y)
        k2 = f(t0 + dt, y + dt * k1)
        return dt * (0.5 * k1 + 0.5 * k2)

    def heun3(f, t0, dt, t1, y):  # alpha=[1/3, 2/3], beta=[[1/3],[0,2/3]], c_sol=[1/4,0,3/4]
        k1 = f(t0, y)
        k2 = f(t0 + dt / 3, y + dt * k1 / 3)
        k3 = f(t0 + dt * 2 / 3, y + dt * (2 / 3) * k2)
        return dt * (0.25 * k1 + 0.75 * k3)

    step = {"euler": euler, "midpoint": midpoint, "rk4": rk4, "heun2": heun2, "heun3": heun3}[method]
    sol = [y0]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        y = y + step(func, t0, t1 - t0, t1, y)
        sol.append(y)
    return torch.stack(sol, dim=0)


def install_shims():
    if "refnodes" in sys.modules:
        return
    # transformers must probe the real environment before librosa is stubbed
    from transformers import Wav2Vec2Model, Wav2Vec2FeatureExtractor  # noqa: F401

    @contextlib.contextmanager
    def model_to_target(logger, model):
        yield

    _mod("seconohe")
    _mod("seconohe.logger", initialize_logger=lambda name, *a, **k: logging.getLogger(name))
    _mod("seconohe.torch", model_to_target=model_to_target,
         get_torch_device_options=lambda *a, **k: (["cpu"], "cpu"),
         get_canonical_device=lambda d: torch.device(d),
         get_offload_device=lambda *a, **k: torch.device("cpu"))
    _mod("seconohe.downloader", download_file=lambda *a, **k: None)
    _mod("seconohe.register_nodes", register_nodes=lambda *a, **k: ({}, {}))

    class _PBar:
        def __init__(self, total):
            self.total, self.n = total, 0

        def update(self, k):
            self.n += k

    comfy = _mod("comfy")
    comfy.utils = _mod("comfy.utils", ProgressBar=_PBar, load_torch_file=None)
    comfy.model_management = _mod("comfy.model_management",
                                  unet_offload_device=lambda: torch.device("cpu"),
                                  get_torch_device=lambda: torch.device("cpu"),
                                  soft_empty_cache=lambda: None)
    _mod("folder_paths", models_dir="/tmp/models", get_folder_paths=lambda *a: [],
         get_filename_list=lambda *a: [], get_full_path=lambda *a: None,
         add_model_folder_path=lambda *a, **k: None)
    _mod("face_alignment")
    _mod("librosa")
    timm = _mod("timm")
    timm.layers = _mod("timm.layers", use_fused_attn=lambda *a, **k: True)
    timm.models = _mod("timm.models")
    timm.models.vision_transformer = _mod("timm.models.vision_transformer", Mlp=_Mlp)
    _mod("torchdiffeq", odeint=_fixed_grid_odeint)


def load_reference():
    """Returns the reference ``src/nodes`` package, imported in place as ``refnodes``."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    install_shims()
    if "refnodes" in sys.modules:
        return sys.modules["refnodes"]
    pkg_dir = os.path.join(REF_ROOT, "src", "nodes")
    spec = importlib.util.spec_from_file_location(
        "refnodes", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refnodes"] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_fmt(state_dict, **opt_overrides):
    """Reference FlowMatchingTransformer(BaseOptions(**overrides)) loaded with ``state_dict``."""
    ref = load_reference()
    import importlib
    FMT = importlib.import_module("refnodes.models.float.FMT")
    BaseOptions = importlib.import_module("refnodes.options.base_options").BaseOptions
    opt = BaseOptions()
    for k, v in opt_overrides.items():
        setattr(opt, k, v)
    opt.rank = torch.device("cpu")
    logging.getLogger().setLevel(logging.WARNING)
    model = FMT.FlowMatchingTransformer(opt)
    sd = {k: v for k, v in state_dict.items() if k not in ("pos_embed", "alignment_mask")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert set(missing) <= {"pos_embed", "alignment_mask"}, missing
    assert not unexpected, unexpected
    model.eval()
    model.target_device = torch.device("cpu")
    model.final_construction_options = {k: getattr(opt, k) for k in vars(opt) if k != "rank"}
    return ref, model, opt
