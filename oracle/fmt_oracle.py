"""CPU/torch-fp32 restatement of the reference FMT sampling path.  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this module - and only as the checker or the timed CPU baseline, never as a
fallback of the product path (the product fails loudly if its CUDA library is missing).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference (imported in place
from /root/reference through ``tests/golden/refshim.py``) on the seeded inputs of ``oracle/synth.py``
and commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against every fixture.  The reference itself ships no tests or golden vectors
(SURVEY.md §4).

Every function cites the reference lines it restates (paths relative to /root/reference).
It is written as plain functions over a weight dict - no nn.Module, no SDPA - so that it is an
independent statement of the algorithm rather than a copy of the reference's module tree.

Third-party arithmetic restated here (absent from /root/reference):
  * torchdiffeq (unpinned, requirements.txt:3): fixed-grid solvers, ``odeint_fixed`` below.
  * timm>=1.0.9 ``Mlp`` (requirements.txt:4): fc1 -> GELU(tanh) -> fc2.
"""
import math
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .synth import FmtDims


class Quant:
    """Optional numerics emulation of the CUDA path (design exploration + tolerance budgeting).

    ``act``   : applied to every GEMM input activation
    ``weight``: applied to every GEMM weight
    ``table`` : applied to the AdaLN shift/scale/gate vectors
    Identity by default (= exact fp32 restatement).
    """

    def __init__(self, act=None, weight=None, table=None, keep_fp32=()):
        ident = lambda t: t  # noqa: E731
        self.act, self.weight, self.table = act or ident, weight or ident, table or ident
        self._identity_w = weight is None
        self._wcache = {}
        self.keep_fp32 = tuple(keep_fp32)      # substrings of layer names whose GEMM stays exact (precision ablations)

    def _kept(self, name):
        return any(k in name for k in self.keep_fp32)

    def a(self, x, name):
        return x if self._kept(name) else self.act(x)

    def w(self, W, key):
        if self._identity_w or self._kept(key):
            return W[key]
        ck = (id(W), key)
        if ck not in self._wcache:
            self._wcache[ck] = self.weight(W[key])
        return self._wcache[ck]


def bf16_quant(table_dtype=torch.bfloat16) -> Quant:
    rt = lambda t: t.to(torch.bfloat16).to(torch.float32)  # noqa: E731
    return Quant(act=rt, weight=rt, table=lambda t: t.to(table_dtype).to(torch.float32))


_EXACT = Quant()


def _linear(x, W, name, q: Quant):
    return F.linear(q.a(x, name), q.w(W, name + ".weight"), W[name + ".bias"])


# --------------------------------------------------------------------------------------
# FMT pieces
# --------------------------------------------------------------------------------------
def band_mask(n: int, expansion: int, device=None) -> torch.Tensor:
    """FMT.py:15-19 (frame_width=1): True = blocked; row i attends j in [max(0,i-e), i+e]."""
    i = torch.arange(n, device=device)[:, None]
    j = torch.arange(n, device=device)[None, :]
    return (j < i - expansion) | (j > i + expansion)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: int = 10000) -> torch.Tensor:
    """FMT.py:107-126: freqs = exp(-ln(max_period) * k / half); [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def t_embedder(W, t: torch.Tensor) -> torch.Tensor:
    """FMT.py:100-104,128-131: Linear(256,H) -> SiLU -> Linear(H,H).  (fp32 in every mode)"""
    h = F.linear(timestep_embedding(t), W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])
    return F.linear(F.silu(h), W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])


def attention(W, p: str, x: torch.Tensor, blocked: torch.Tensor, num_heads: int, q: Quant) -> torch.Tensor:
    """FMT.py:69-91 (fused/masked path): qkv columns are [q|k|v], each head-major; scale hd^-1/2."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = q.act(_linear(x, W, p + "attn.qkv", q))   # CUDA path keeps qkv in the GEMM-operand dtype
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    qq, kk, vv = qkv[0], qkv[1], qkv[2]
    s = (qq @ kk.transpose(-2, -1)) * (hd ** -0.5)
    s = s.masked_fill(blocked, float("-inf"))
    o = torch.softmax(s, dim=-1) @ vv
    o = o.transpose(1, 2).reshape(B, N, C)
    return _linear(o, W, p + "attn.proj", q)


def fmt_block(W, i: int, x, c_silu, blocked, num_heads, q: Quant):
    """FMT.py:171-176: adaLN chunk order shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp."""
    p = f"blocks.{i}."
    H = x.shape[-1]
    mod = q.table(F.linear(q.act(c_silu), q.w(W, p + "adaLN_modulation.1.weight"), W[p + "adaLN_modulation.1.bias"]))
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
    h = F.layer_norm(x, (H,), eps=1e-6) * (1 + sc1) + sh1
    x = x + g1 * attention(W, p, h, blocked, num_heads, q)
    h = F.layer_norm(x, (H,), eps=1e-6) * (1 + sc2) + sh2
    h = F.gelu(_linear(h, W, p + "mlp.fc1", q), approximate="tanh")
    x = x + g2 * _linear(h, W, p + "mlp.fc2", q)
    return x


def fmt_forward(W, d: FmtDims, t, x, wa, wr, we, prev_x=None, prev_wa=None, prev_we=None, q: Quant = _EXACT):
    """FMT.py:277-340 with train=False (sequence_embedder is then the identity, :271-275)."""
    t_emb = t_embedder(W, t).unsqueeze(1)
    wr = wr.unsqueeze(1)
    if prev_x is not None:
        if prev_wa is None:
            raise ValueError("prev_x was provided, but prev_wa was not.")
        if we.shape[1] > 1 and prev_we is None:
            raise ValueError("`we` is dynamic (T>1), but prev_we was not provided with prev_x/prev_wa.")
        x = torch.cat([prev_x, x], dim=1)
        wa = torch.cat([prev_wa, wa], dim=1)
        if we.shape[1] > 1:
            we = torch.cat([prev_we, we], dim=1)
    x = _linear(x, W, "x_embedder.proj", q) + W["pos_embed"]
    n = wa.shape[1]
    wr = wr.repeat(1, n, 1)
    if we.shape[1] == 1:
        we = we.repeat(1, n, 1)          # static emotion also covers the context frames (FMT.py:325-326)
    elif we.shape[1] != n:
        raise ValueError(f"Dynamic emotion latent `we` time dimension ({we.shape[1]}) does not match "
                         f"audio latent `wa` time dimension ({n}).")
    c = torch.cat([wr, wa, we], dim=-1)    # int64 one-hot `we` is promoted to float here (FMT.py:333)
    c = _linear(c, W, "c_embedder", q) + t_emb
    c_silu = F.silu(c)
    blocked = band_mask(d.total_frames, d.attention_window, device=x.device)
    for i in range(d.fmt_depth):
        x = fmt_block(W, i, x, c_silu, blocked, d.num_heads, q)
    # Decoder, FMT.py:195-198
    mod = q.table(F.linear(q.act(c_silu), q.w(W, "decoder.adaLN_modulation.1.weight"), W["decoder.adaLN_modulation.1.bias"]))
    shift, scale = mod.chunk(2, dim=-1)
    h = F.layer_norm(x, (x.shape[-1],), eps=1e-6) * (1 + scale) + shift
    return _linear(h, W, "decoder.linear", q)


def forward_with_cfv(W, d: FmtDims, t, x, wa, wr, we, prev_x, prev_wa, prev_we=None,
                     a_cfg_scale=1.0, r_cfg_scale=1.0, e_cfg_scale=1.0, include_r_cfg=False, q: Quant = _EXACT):
    """FMT.py:342-401.  Branch order [uncond | all | audio-only] (3) or [truly-uncond | uncond | all | audio-only] (4);
    x / prev_x / prev_wa are replicated un-nulled, prev_we is nulled like we."""
    if a_cfg_scale != 1.0 or r_cfg_scale != 1.0 or e_cfg_scale != 1.0:
        z_wa, z_we, z_wr = torch.zeros_like(wa), torch.zeros_like(we), torch.zeros_like(wr)
        z_pwe = torch.zeros_like(prev_we) if prev_we is not None else None
        if not include_r_cfg:
            a_cat, r_cat, e_cat = [z_wa, wa, wa], [wr, wr, wr], [z_we, we, z_we]
            pe_cat = [z_pwe, prev_we, z_pwe]
        else:
            a_cat, r_cat, e_cat = [z_wa, z_wa, wa, wa], [z_wr, wr, wr, wr], [z_we, z_we, we, z_we]
            pe_cat = [z_pwe, z_pwe, prev_we, z_pwe]
        nb = len(a_cat)
        out = fmt_forward(W, d, t, torch.cat([x] * nb), torch.cat(a_cat), torch.cat(r_cat), torch.cat(e_cat),
                          torch.cat([prev_x] * nb), torch.cat([prev_wa] * nb),
                          torch.cat(pe_cat) if prev_we is not None else None, q=q)
        if not include_r_cfg:
            u, c_all, a_only = out.chunk(3, dim=0)
            return u + a_cfg_scale * (a_only - u) + e_cfg_scale * (c_all - a_only)
        tu, u, c_all, a_only = out.chunk(4, dim=0)
        return tu + r_cfg_scale * (u - tu) + a_cfg_scale * (a_only - u) + e_cfg_scale * (c_all - a_only)
    return fmt_forward(W, d, t, x, wa, wr, we, prev_x, prev_wa, prev_we, q=q)


# --------------------------------------------------------------------------------------
# torchdiffeq fixed-grid solvers (call sites nodes_adv.py:658, FLOAT.py:247)
# --------------------------------------------------------------------------------------
def _step_euler(f, t0, dt, t1, y):
    return dt * f(t0, y)


def _step_midpoint(f, t0, dt, t1, y):
    half = 0.5 * dt
    return dt * f(t0 + half, y + f(t0, y) * half)


def _step_rk4(f, t0, dt, t1, y):      # torchdiffeq's fixed-grid "rk4" = 3/8 rule (rk4_alt_step_func)
    k1 = f(t0, y)
    k2 = f(t0 + dt / 3, y + dt * k1 / 3)
    k3 = f(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
    k4 = f(t1, y + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def _step_heun2(f, t0, dt, t1, y):
    k1 = f(t0, y)
    k2 = f(t0 + dt, y + dt * k1)
    return dt * (0.5 * k1 + 0.5 * k2)


def _step_heun3(f, t0, dt, t1, y):
    k1 = f(t0, y)
    k2 = f(t0 + dt / 3, y + dt * k1 / 3)
    k3 = f(t0 + dt * 2 / 3, y + dt * (2 / 3) * k2)
    return dt * (0.25 * k1 + 0.75 * k3)


SOLVERS = {"euler": _step_euler, "midpoint": _step_midpoint, "rk4": _step_rk4, "heun2": _step_heun2, "heun3": _step_heun3}


def odeint_fixed(f: Callable, y0: torch.Tensor, t: torch.Tensor, method: str = "euler") -> torch.Tensor:
    """Returns y(t[-1]).  nfe grid points => nfe-1 steps; nfe == 1 returns y0 unchanged."""
    step = SOLVERS[method]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        y = y + step(f, t0, t1 - t0, t1, y)
    return y


# --------------------------------------------------------------------------------------
# Window loops
# --------------------------------------------------------------------------------------
def _pad_replicate(x: torch.Tensor, n: int) -> torch.Tensor:
    """F.pad(..., mode='replicate') along dim 1 up to n frames (nodes_adv.py:614-616)."""
    if x.shape[1] >= n:
        return x
    return torch.cat([x, x[:, -1:].expand(-1, n - x.shape[1], -1)], dim=1)


def sample_loop(W, d: FmtDims, r_s, wa, we, audio_num_frames: int, nfe: int = 10, method: str = "euler",
                a_cfg_scale=2.0, r_cfg_scale=1.0, e_cfg_scale=1.0, include_r_cfg=False,
                generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None,
                q: Quant = _EXACT, progress: Optional[Callable[[int], None]] = None) -> torch.Tensor:
    """nodes_adv.py:545-694 ``_perform_ode_sampling_loop``.

    ``noise`` (n_windows, B, L, dim_w), if given, replaces the per-window ``torch.randn`` draws
    (injected-noise parity); otherwise one ``randn(B, L, dim_w)`` per window is drawn from
    ``generator`` on the device of ``wa`` in window order (nodes_adv.py:606).
    """
    dev = wa.device
    B = wa.shape[0]
    dynamic = we.shape[1] > 1
    E = we.shape[2]
    L, P = d.frames_per_clip, d.num_prev_frames
    time = torch.linspace(0, 1, nfe, device=dev)
    prev_x = torch.zeros(B, P, d.dim_w, device=dev)
    prev_wa = torch.zeros(B, P, d.dim_w, device=dev)
    prev_we = torch.zeros(B, P, E, device=dev)
    n_win = math.ceil(audio_num_frames / L)
    outs = []
    for w in range(n_win):
        if noise is not None:
            x0 = noise[w].to(dev)
        else:
            x0 = torch.randn(B, L, d.dim_w, device=dev, generator=generator)
        wa_c = _pad_replicate(wa[:, w * L:(w + 1) * L], L)
        we_c = _pad_replicate(we[:, w * L:(w + 1) * L], L) if dynamic else we

        def f(t_scalar, xb, wa_c=wa_c, we_c=we_c, prev_x=prev_x, prev_wa=prev_wa, prev_we=prev_we):
            out = forward_with_cfv(W, d, t_scalar.reshape(1), xb, wa_c, r_s, we_c, prev_x, prev_wa, prev_we,
                                   a_cfg_scale, r_cfg_scale, e_cfg_scale, include_r_cfg, q=q)
            return out[:, P:]

        sample = odeint_fixed(f, x0, time, method)
        outs.append(sample)
        prev_x = _pad_replicate(sample, P)[:, -P:]
        prev_wa = _pad_replicate(wa_c, P)[:, -P:]
        prev_we = we_c[:, -P:] if dynamic else we_c.repeat(1, P, 1)
        if progress is not None:
            progress(w)
    return torch.cat(outs, dim=1)[:, :audio_num_frames]


def float_sample_legacy(W, d: FmtDims, r_s, wa, we, opt_nfe: int = 10, a_cfg_scale=1.0, r_cfg_scale=1.0,
                        e_cfg_scale=1.0, generator=None, noise=None, q: Quant = _EXACT) -> torch.Tensor:
    """FLOAT.py:172-253 ``FLOAT.sample`` from the point where ``wa``/``we`` exist: uses ``opt.nfe`` (its
    ``nfe`` argument is ignored, :188), static emotion only (possibly int64 one-hot, :200), never passes
    ``prev_we`` and never enables ``include_r_cfg``; Euler only in practice (``odeint_kwargs``, :78)."""
    dev = wa.device
    B, T = wa.shape[0], wa.shape[1]
    L, P = d.frames_per_clip, d.num_prev_frames
    time = torch.linspace(0, 1, opt_nfe, device=dev)
    outs, sample, wa_t = [], None, None
    for w in range(int(math.ceil(T / L))):
        x0 = noise[w].to(dev) if noise is not None else torch.randn(B, L, d.dim_w, device=dev, generator=generator)
        if w == 0:
            prev_x = torch.zeros(B, P, d.dim_w, device=dev)
            prev_wa = torch.zeros(B, P, d.dim_w, device=dev)
        else:
            prev_x, prev_wa = sample[:, -P:], wa_t[:, -P:]
        wa_t = _pad_replicate(wa[:, w * L:(w + 1) * L], L)

        def f(tt, zt, wa_t=wa_t, prev_x=prev_x, prev_wa=prev_wa):
            return forward_with_cfv(W, d, tt.reshape(1), zt, wa_t, r_s, we, prev_x, prev_wa, None,
                                    a_cfg_scale, r_cfg_scale, e_cfg_scale, False, q=q)[:, P:]

        sample = odeint_fixed(f, x0, time, "euler")
        outs.append(sample)
    return torch.cat(outs, dim=1)[:, :T]


def audio_projection(P, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``AudioEncoder.audio_projection`` (FLOAT.py:338-342) as applied by FloatApplyAudioProjection (nodes_vadv.py:188-192):
    SiLU(LayerNorm(Linear(x))) over the last dimension; ``P`` holds the Sequential's state-dict keys."""
    y = F.linear(x, P["0.weight"], P["0.bias"])
    y = F.layer_norm(y, (y.shape[-1],), P["1.weight"], P["1.bias"], eps)
    return F.silu(y)
