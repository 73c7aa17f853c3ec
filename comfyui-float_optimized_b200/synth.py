"""Seeded synthetic FMT weights and sampler inputs for benchmarks, smoke() and tests (holds no algorithm of the path).

The reference zero-initialises every adaLN layer and ``decoder.linear``
(``/root/reference/src/nodes/models/float/FMT.py:260-269``), which makes the velocity field
identically zero - useless for parity.  SURVEY.md §8d therefore prescribes non-degenerate
random weights.  They are generated here from a ``torch.Generator`` in a fixed key order so
that the build container (where the real reference consumes them through ``load_state_dict``)
and the GPU box (where the oracle and the CUDA path consume them) see bit-identical tensors.

Key names / shapes follow the reference state dict (SURVEY.md §8a "Weights").
"""
import math
from dataclasses import dataclass, asdict

import torch


@dataclass(frozen=True)
class FmtDims:
    """Mirror of the BaseOptions fields the FMT reads (``options/base_options.py:10-60``)."""
    dim_w: int = 512
    dim_a: int = 512
    dim_e: int = 7
    dim_h: int = 1024
    fmt_depth: int = 8
    num_heads: int = 8
    mlp_ratio: float = 4.0
    num_prev_frames: int = 10
    wav2vec_sec: float = 2.0
    fps: float = 25.0
    attention_window: int = 2

    @property
    def frames_per_clip(self) -> int:          # FMT.py:209
        return int(self.wav2vec_sec * self.fps)

    @property
    def total_frames(self) -> int:             # FMT.py:211
        return self.num_prev_frames + self.frames_per_clip

    @property
    def mlp_hidden(self) -> int:               # FMT.py:160
        return int(self.dim_h * self.mlp_ratio)

    def as_dict(self):
        return asdict(self)


SMALL_DIMS = FmtDims(dim_w=64, dim_a=64, dim_e=7, dim_h=128, fmt_depth=2, num_heads=2, mlp_ratio=2.0,
                     num_prev_frames=4, wav2vec_sec=0.48, fps=25.0, attention_window=1)  # 12-frame window


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """FMT.py:22-40: angle = pos / 10000^(2*(j//2)/d); even j -> sin, odd j -> cos (fp32 table)."""
    pos = torch.arange(n_position, dtype=torch.float64)[:, None]
    j = torch.arange(d_hid, dtype=torch.float64)[None, :]
    # the reference computes the angles in Python floats (double) and stores into torch.Tensor
    angle = (pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / d_hid)).float()
    out = angle.clone()
    out[:, 0::2] = torch.sin(angle[:, 0::2])
    out[:, 1::2] = torch.cos(angle[:, 1::2])
    return out


def weight_shapes(d: FmtDims):
    """Ordered (key, shape, kind) list; kind selects the distribution."""
    H, W, A, E, M = d.dim_h, d.dim_w, d.dim_a, d.dim_e, d.mlp_hidden
    out = [
        ("x_embedder.proj.weight", (H, W), "xavier"), ("x_embedder.proj.bias", (H,), "bias"),
        ("t_embedder.mlp.0.weight", (H, 256), "t"), ("t_embedder.mlp.0.bias", (H,), "bias"),
        ("t_embedder.mlp.2.weight", (H, H), "t"), ("t_embedder.mlp.2.bias", (H,), "bias"),
        ("c_embedder.weight", (H, W + A + E), "xavier"), ("c_embedder.bias", (H,), "bias"),
    ]
    for i in range(d.fmt_depth):
        p = f"blocks.{i}."
        out += [
            (p + "attn.qkv.weight", (3 * H, H), "xavier"), (p + "attn.qkv.bias", (3 * H,), "bias"),
            (p + "attn.proj.weight", (H, H), "xavier"), (p + "attn.proj.bias", (H,), "bias"),
            (p + "mlp.fc1.weight", (M, H), "xavier"), (p + "mlp.fc1.bias", (M,), "bias"),
            (p + "mlp.fc2.weight", (H, M), "xavier"), (p + "mlp.fc2.bias", (H,), "bias"),
            (p + "adaLN_modulation.1.weight", (6 * H, H), "ada"), (p + "adaLN_modulation.1.bias", (6 * H,), "ada"),
        ]
    out += [
        ("decoder.adaLN_modulation.1.weight", (2 * H, H), "ada"), ("decoder.adaLN_modulation.1.bias", (2 * H,), "ada"),
        ("decoder.linear.weight", (W, H), "ada"), ("decoder.linear.bias", (W,), "ada"),
    ]
    return out


def synth_state_dict(d: FmtDims = FmtDims(), seed: int = 0, ada_std: float = 0.01) -> dict:
    """fp32 state dict with the reference's key set (incl. pos_embed); deterministic in ``seed``."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in weight_shapes(d):
        if kind == "xavier":
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "t":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.02
        else:
            t = torch.randn(shape, generator=g) * ada_std
        sd[key] = t.contiguous()
    sd["pos_embed"] = sinusoid_table(d.total_frames, d.dim_h).unsqueeze(0).contiguous()
    return sd


def synth_inputs(d: FmtDims, batch: int, num_frames: int, seed: int = 7, dynamic_we: bool = False,
                 onehot_we: bool = False):
    """(r_s (B,W), wa (B,T,A), we (B,1|T,E)) with the statistics of SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    r_s = 0.5 * torch.randn(batch, d.dim_w, generator=g)
    wa = torch.randn(batch, num_frames, d.dim_a, generator=g)
    wa = torch.nn.functional.silu(torch.nn.functional.layer_norm(wa, (d.dim_a,)))   # FLOAT.py:338-342 statistics
    if dynamic_we:
        chunk = max(1, d.frames_per_clip)         # one emotion vector per window-length chunk ...
        n_chunks = math.ceil(num_frames / chunk)
        per_chunk = torch.softmax(torch.randn(batch, n_chunks, d.dim_e, generator=g), dim=-1)
        idx = (torch.arange(num_frames) * n_chunks // num_frames).clamp(max=n_chunks - 1)  # ... nearest-upsampled (nodes_vadv.py:832-840)
        we = per_chunk[:, idx]
    elif onehot_we:
        k = torch.randint(0, d.dim_e, (batch,), generator=g)
        we = torch.nn.functional.one_hot(k, d.dim_e).unsqueeze(1)   # int64, as FLOAT.py:200 produces
    else:
        we = torch.softmax(torch.randn(batch, 1, d.dim_e, generator=g), dim=-1)
    return r_s.contiguous(), wa.contiguous(), we.contiguous()


def synth_projection(in_dim: int, dim_a: int = 512, seed: int = 0) -> dict:
    """Seeded weights of the audio projection Sequential(Linear(in_dim, dim_a), LayerNorm(dim_a), SiLU)
    (FLOAT.py:338-342, nodes_vadv_loader.py:233-240) under its state-dict keys; the LayerNorm affine is perturbed so that
    it is exercised (the reference initialises it to weight 1, bias 0)."""
    g = torch.Generator().manual_seed(3000 + seed)
    bound = 1.0 / math.sqrt(in_dim)
    return {
        "0.weight": (torch.rand(dim_a, in_dim, generator=g) * 2 - 1) * bound,
        "0.bias": (torch.rand(dim_a, generator=g) * 2 - 1) * bound,
        "1.weight": 1.0 + 0.1 * torch.randn(dim_a, generator=g),
        "1.bias": 0.05 * torch.randn(dim_a, generator=g),
    }


def synth_wav2vec_features(batch: int, num_frames: int, in_dim: int, seed: int = 7) -> torch.Tensor:
    """Stand-in for the interpolated wav2vec2 hidden states (B, T, in_dim): unit-variance features with a per-layer offset."""
    g = torch.Generator().manual_seed(4000 + seed)
    return torch.randn(batch, num_frames, in_dim, generator=g) + 0.2 * torch.randn(1, 1, in_dim, generator=g)
