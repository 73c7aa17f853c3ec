"""Shared helpers: load golden fixtures and rebuild their seeded inputs (recipes in manifest.json)."""
import functools
import json
import os

import numpy as np
import torch

from golden.make_golden import CASES, case_inputs, cfv_extra_inputs, dims_of  # noqa: F401  (same recipe code that made the fixtures)
from oracle.synth import FmtDims, SMALL_DIMS, synth_state_dict, synth_projection, synth_wav2vec_features

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def manifest():
    return json.load(open(os.path.join(GOLDEN_DIR, "manifest.json")))


def golden(name) -> torch.Tensor:
    return torch.from_numpy(np.load(os.path.join(GOLDEN_DIR, name + ".npz"))["out"])


@functools.lru_cache(maxsize=None)
def weights(dims_name: str):
    d = FmtDims() if dims_name == "full" else SMALL_DIMS
    return synth_state_dict(d, seed=0)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())


def projection_weights(rec) -> dict:
    return synth_projection(rec["in_dim"], FmtDims().dim_a, seed=rec["seed"])


def projection_input(rec) -> torch.Tensor:
    return synth_wav2vec_features(rec["B"], rec["T"], rec["in_dim"], seed=rec["seed"])
