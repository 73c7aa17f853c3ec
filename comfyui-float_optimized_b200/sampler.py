"""Host side of the FMT sampler: the reference's window loop and solver schedule, expressed as one C-ABI call.

Mirrors (paths relative to the reference repo):
  * ``_perform_ode_sampling_loop``  src/nodes/nodes_adv.py:545-694   -> :func:`perform_ode_sampling_loop`
  * ``FLOAT.sample`` (from ``wa``/``we`` on)  src/nodes/models/float/FLOAT.py:172-253 -> :func:`float_sample`
  * ``FlowMatchingTransformer.forward_with_cfv``  FMT.py:342-401     -> :meth:`FmtBackend.velocity`
All arithmetic happens in libfmt_b200.so (hand-written CUDA, sm_100a).  PyTorch is used for device memory, the
RNG (one ``torch.randn`` per window, exactly like the reference) and streams.  There is no fallback path.
"""
import ctypes as C
import functools
import math
import os
import weakref
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from . import _cabi
from ._cabi import FmtError

# Butcher tableaux of torchdiffeq's fixed-grid solvers (src/nodes/__init__.py:15-23 lists what the node offers).
# y_i = y0 + dt * sum_j a[i][j] k_j ; t_i = t0 + c[i] * dt ; y1 = y0 + dt * sum_j b[j] k_j
SOLVERS = {
    "euler": dict(c=[0.0], a=[[0.0]], b=[1.0]),
    "midpoint": dict(c=[0.0, 0.5], a=[[0, 0], [0.5, 0]], b=[0.0, 1.0]),
    "heun2": dict(c=[0.0, 1.0], a=[[0, 0], [1.0, 0]], b=[0.5, 0.5]),
    "heun3": dict(c=[0.0, 1 / 3, 2 / 3], a=[[0, 0, 0], [1 / 3, 0, 0], [0, 2 / 3, 0]], b=[0.25, 0.0, 0.75]),
    # torchdiffeq's fixed-grid "rk4" is the 3/8-rule variant (rk4_alt_step_func)
    "rk4": dict(c=[0.0, 1 / 3, 2 / 3, 1.0], a=[[0, 0, 0, 0], [1 / 3, 0, 0, 0], [-1 / 3, 1, 0, 0], [1, -1, 1, 0]],
                b=[0.125, 0.375, 0.375, 0.125]),
}


@dataclass(frozen=True)
class Dims:
    dim_w: int = 512
    dim_a: int = 512
    dim_e: int = 7
    dim_h: int = 1024
    fmt_depth: int = 8
    num_heads: int = 8
    mlp_hidden: int = 4096
    num_prev_frames: int = 10
    frames_per_clip: int = 50
    attention_window: int = 2

    @staticmethod
    def from_options(opt, state_dict=None) -> "Dims":
        """Dims from a BaseOptions-like object (or dict); weight shapes win where both are known
        (the loader infers dim_h/depth/mlp_ratio/dim_a from the checkpoint, nodes_vadv_loader.py:655-866)."""
        get = (lambda k, dflt: opt.get(k, dflt)) if isinstance(opt, dict) else (lambda k, dflt: getattr(opt, k, dflt))
        dim_h = int(get("dim_h", 1024))
        d = dict(dim_w=int(get("dim_w", 512)), dim_a=int(get("dim_a", 512)), dim_e=int(get("dim_e", 7)), dim_h=dim_h,
                 fmt_depth=int(get("fmt_depth", 8)), num_heads=int(get("num_heads", 8)),
                 mlp_hidden=int(dim_h * float(get("mlp_ratio", 4.0))),
                 num_prev_frames=int(get("num_prev_frames", 10)),
                 frames_per_clip=int(float(get("wav2vec_sec", 2.0)) * float(get("fps", 25.0))),
                 attention_window=int(get("attention_window", 2)))
        if state_dict is not None:
            xw = state_dict["x_embedder.proj.weight"]
            d["dim_h"], d["dim_w"] = int(xw.shape[0]), int(xw.shape[1])
            d["mlp_hidden"] = int(state_dict["blocks.0.mlp.fc1.weight"].shape[0])
            d["fmt_depth"] = 1 + max(int(k.split(".")[1]) for k in state_dict if k.startswith("blocks."))
            d["dim_a"] = int(state_dict["c_embedder.weight"].shape[1]) - d["dim_w"] - d["dim_e"]
            if "pos_embed" in state_dict:
                d["frames_per_clip"] = int(state_dict["pos_embed"].shape[-2]) - d["num_prev_frames"]
        return Dims(**d)


def _f32_array(values):
    arr = (C.c_float * max(1, len(values)))()
    for i, v in enumerate(values):
        arr[i] = float(v)
    return arr


def build_schedule(nfe: int, method: str):
    """Time grid of ``torch.linspace(0, 1, nfe)`` (nodes_adv.py:587) walked by a fixed-grid solver:
    nfe points => nfe-1 steps; stage times t0 + c_i*dt evaluated in fp32 like torchdiffeq does.  (Cached: ~0.1 ms of torch scalar
    arithmetic that every sampler call would otherwise repeat.)"""
    return dict(_build_schedule(int(nfe), method))


@functools.lru_cache(maxsize=64)
def _build_schedule(nfe: int, method: str):
    if method not in SOLVERS:
        raise ValueError(f"Unknown fixed-step solver '{method}' (supported: {sorted(SOLVERS)})")
    tab = SOLVERS[method]
    t = torch.linspace(0, 1, int(nfe), dtype=torch.float32)
    dt = (t[1:] - t[:-1])
    stages = len(tab["c"])
    t_eval = []
    for i in range(max(0, int(nfe) - 1)):
        for c in tab["c"]:
            if c == 0.0:
                t_eval.append(float(t[i]))
            elif c == 1.0:
                t_eval.append(float(t[i + 1]) if method == "rk4" else float(t[i] + dt[i]))
            else:
                t_eval.append(float(t[i] + dt[i] * torch.tensor(c, dtype=torch.float32)))
    a_flat = [x for row in tab["a"] for x in row]
    return dict(n_steps=max(0, int(nfe) - 1), n_stages=stages, t_eval=t_eval, dt=[float(x) for x in dt], a=a_flat, b=tab["b"])


PRECISION_MODES = ("bf16", "fp32")


def resolve_mode(mode: Optional[str] = None) -> str:
    """Operand precision of the GEMMs: "bf16" (tcgen05, fp32 accumulate / residual / softmax / ODE state; latents within 2e-2
    max-abs of the reference) or "fp32" (validation mode, fp32 FFMA GEMMs; within 1e-4 relative, ~10x slower).  An explicit
    argument wins, then the environment variable FMT_MODE, then "bf16".  The reference itself computes in fp32."""
    m = mode if mode not in (None, "", "default") else os.environ.get("FMT_MODE", "bf16")
    m = str(m).lower()
    if m not in PRECISION_MODES:
        raise ValueError(f"precision mode '{m}' is not one of {PRECISION_MODES}")
    return m


def n_branches_for(a_cfg_scale, r_cfg_scale, e_cfg_scale, include_r_cfg) -> int:
    """FMT.py:346,359,380,400: one conditional forward when every scale == 1, else 3 (or 4) batched branches."""
    if a_cfg_scale != 1.0 or r_cfg_scale != 1.0 or e_cfg_scale != 1.0:
        return 4 if include_r_cfg else 3
    return 1


class FmtBackend:
    """Packed weights + workspace + captured window graph for one FMT on one CUDA device (owns an FmtHandle*)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dims: Dims, device):
        self.lib = _cabi.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise FmtError(f"FmtBackend needs a CUDA device (sm_100a); got '{self.device}'. There is no CPU fallback.")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dims = dims
        self._plan_key = None
        self._n_eval = 0
        self._handle = C.c_void_p()
        keys = list(_cabi.GLOBAL_KEYS) + [f"blocks.{i}.{k}" for i in range(dims.fmt_depth) for k in _cabi.BLOCK_KEYS]
        missing = [k for k in keys if k not in state_dict]
        if missing:
            raise KeyError(f"FMT state dict is missing {missing[:4]}{'...' if len(missing) > 4 else ''}")
        tensors = [state_dict[k].detach() for k in keys]
        on_host = all(t.device.type == "cpu" for t in tensors)
        if on_host:
            tensors = [t.to(torch.float32).contiguous() for t in tensors]
        else:
            tensors = [t.to(device=self.device, dtype=torch.float32).contiguous() for t in tensors]
        self._check_shapes(dict(zip(keys, tensors)))
        ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        cd = _cabi.FmtDims(dims.dim_w, dims.dim_a, dims.dim_e, dims.dim_h, dims.fmt_depth, dims.num_heads, dims.mlp_hidden,
                           dims.num_prev_frames, dims.frames_per_clip, dims.attention_window)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.fmt_create(C.byref(cd), ptrs, len(tensors),
                                            _cabi.FMT_LOC_HOST if on_host else _cabi.FMT_LOC_DEVICE,
                                            self.device.index, C.byref(self._handle)), "fmt_create")
        self._finalizer = weakref.finalize(self, self.lib.fmt_destroy, self._handle)

    def _check_shapes(self, sd):
        d = self.dims
        H, N = d.dim_h, d.num_prev_frames + d.frames_per_clip
        expect = {"x_embedder.proj.weight": (H, d.dim_w), "c_embedder.weight": (H, d.dim_w + d.dim_a + d.dim_e),
                  "t_embedder.mlp.0.weight": (H, 256), "decoder.linear.weight": (d.dim_w, H),
                  "blocks.0.attn.qkv.weight": (3 * H, H), "blocks.0.mlp.fc1.weight": (d.mlp_hidden, H),
                  "blocks.0.adaLN_modulation.1.weight": (6 * H, H)}
        for k, shp in expect.items():
            if tuple(sd[k].shape) != shp:
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != expected {shp} for {d}")
        if sd["pos_embed"].numel() != N * H:
            raise ValueError(f"pos_embed has {sd['pos_embed'].numel()} elements, expected {N}x{H}")

    # ------------------------------------------------------------------------------------------------
    def close(self):
        self._finalizer()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def configure(self, batch: int, n_branches: int, we_dynamic: bool, nfe: int, method: str = "euler", mode: str = "bf16"):
        mode_id = {"bf16": _cabi.FMT_MODE_BF16, "fp32": _cabi.FMT_MODE_FP32_VALIDATE}[resolve_mode(mode)]
        key = (batch, n_branches, bool(we_dynamic), int(nfe), method, mode_id)
        if key == self._plan_key:
            return
        sched = build_schedule(nfe, method)
        t_eval, dt, a, b = (_f32_array(sched[k]) for k in ("t_eval", "dt", "a", "b"))
        plan = _cabi.FmtPlan(batch, n_branches, int(bool(we_dynamic)), mode_id, sched["n_steps"], sched["n_stages"],
                             C.cast(t_eval, C.POINTER(C.c_float)), C.cast(dt, C.POINTER(C.c_float)),
                             C.cast(a, C.POINTER(C.c_float)), C.cast(b, C.POINTER(C.c_float)))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.fmt_configure(self._handle, C.byref(plan), self._stream()), "fmt_configure")
        self._plan_key = key
        self._n_eval = sched["n_steps"] * sched["n_stages"]

    def workspace_bytes(self) -> int:
        return int(self.lib.fmt_workspace_bytes(self._handle))

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.fmt_launch_count(self._handle, int(reset)))

    def window_kernel_status(self) -> int:
        """-1: the current plan runs one kernel per op; 0: the persistent window kernel is in use; >0: it trapped."""
        return int(self.lib.fmt_window_kernel_status(self._handle))

    def graph_kernel_nodes(self) -> int:
        return int(self.lib.fmt_graph_kernel_nodes(self._handle))

    # ------------------------------------------------------------------------------------------------
    def _check_clip(self, r_s, wa, we, audio_num_frames, noise, host: bool):
        """Every tensor the C call indexes by the plan's dimensions is checked here (the reference would raise a torch shape
        error inside F.linear / torch.cat; the kernels would read out of bounds)."""
        d = self.dims
        if self._plan_key is None:
            raise FmtError("sample_clip: configure() has not been called")
        B = self._plan_key[0]
        want_dev = "cpu" if host else "cuda"
        for name, t in (("r_s", r_s), ("wa", wa), ("we", we), ("noise", noise)):
            if not torch.is_tensor(t):
                raise TypeError(f"{name} must be a torch.Tensor")
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous float32 (got {t.dtype}, contiguous={t.is_contiguous()})")
            if t.device.type != want_dev or (not host and t.device != self.device):
                raise ValueError(f"{name} is on {t.device}, expected {'the CPU' if host else self.device}")
        if wa.dim() != 3 or wa.shape[0] != B or wa.shape[2] != d.dim_a or wa.shape[1] < 1:
            raise ValueError(f"wa shape {tuple(wa.shape)} != (B={B}, T>=1, dim_a={d.dim_a})")
        if r_s.numel() != B * d.dim_w or r_s.shape[0] != B:
            raise ValueError(f"r_s shape {tuple(r_s.shape)} != (B={B}, dim_w={d.dim_w})")
        if we.dim() != 3 or we.shape[0] != B or we.shape[2] != d.dim_e or we.shape[1] < 1:
            raise ValueError(f"we shape {tuple(we.shape)} != (B={B}, 1 or T, dim_e={d.dim_e})")
        dynamic = self._plan_key[2]
        if dynamic != (we.shape[1] > 1):
            raise ValueError(f"we has {we.shape[1]} frames but the plan was configured with we_dynamic={dynamic}")
        if int(audio_num_frames) < 1:
            raise ValueError(f"audio_num_frames must be >= 1 (got {audio_num_frames})")
        n_win = -(-int(audio_num_frames) // d.frames_per_clip)
        if tuple(noise.shape) != (n_win, B, d.frames_per_clip, d.dim_w):
            raise ValueError(f"noise shape {tuple(noise.shape)} != {(n_win, B, d.frames_per_clip, d.dim_w)}")

    def sample_clip(self, r_s, wa, we, audio_num_frames: int, noise, a_cfg_scale, r_cfg_scale, e_cfg_scale,
                    progress: Optional[Callable[[int, int], None]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All windows of one batch of clips.  Tensors on ``self.device`` (fp32, contiguous); asynchronous."""
        d = self.dims
        self._check_clip(r_s, wa, we, audio_num_frames, noise, host=False)
        B = wa.shape[0]
        if out is None:
            out = torch.empty(B, audio_num_frames, d.dim_w, device=self.device, dtype=torch.float32)
        elif tuple(out.shape) != (B, int(audio_num_frames), d.dim_w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise ValueError(f"out must be a contiguous float32 {(B, int(audio_num_frames), d.dim_w)} tensor on {self.device}")
        cb = _cabi.PROGRESS_FN(lambda w, n, _u: progress(w, n)) if progress is not None else _cabi.PROGRESS_FN(0)
        clip = _cabi.FmtClip(_cabi.FMT_LOC_DEVICE, r_s.data_ptr(), wa.data_ptr(), we.data_ptr(), noise.data_ptr(), out.data_ptr(),
                             wa.shape[1], we.shape[1], int(audio_num_frames), float(a_cfg_scale), float(r_cfg_scale),
                             float(e_cfg_scale), cb, None)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.fmt_sample_clip(self._handle, C.byref(clip), self._stream()), "fmt_sample_clip")
        return out

    def sample_clip_host(self, r_s, wa, we, audio_num_frames: int, noise, a_cfg_scale, r_cfg_scale, e_cfg_scale) -> torch.Tensor:
        """Same with HOST (CPU, fp32, contiguous) tensors in and out; copies happen inside the C call (synchronous)."""
        d = self.dims
        self._check_clip(r_s, wa, we, audio_num_frames, noise, host=True)
        out = torch.empty(wa.shape[0], audio_num_frames, d.dim_w, dtype=torch.float32)
        clip = _cabi.FmtClip(_cabi.FMT_LOC_HOST, r_s.data_ptr(), wa.data_ptr(), we.data_ptr(), noise.data_ptr(), out.data_ptr(),
                             wa.shape[1], we.shape[1], int(audio_num_frames), float(a_cfg_scale), float(r_cfg_scale),
                             float(e_cfg_scale), _cabi.PROGRESS_FN(0), None)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.fmt_sample_clip(self._handle, C.byref(clip), self._stream()), "fmt_sample_clip")
        return out

    def velocity(self, eval_index: int, x, wa, r_s, we, prev_x, prev_wa, prev_we, a_cfg_scale, r_cfg_scale, e_cfg_scale) -> torch.Tensor:
        """One ``forward_with_cfv`` (FMT.py:342-401) at plan time ``t_eval[eval_index]``; returns (B, P+L, dim_w)."""
        d = self.dims
        if self._plan_key is None:
            raise FmtError("velocity: configure() has not been called")
        B, L, P = self._plan_key[0], d.frames_per_clip, d.num_prev_frames
        dyn = self._plan_key[2]
        expect = {"x": (B, L, d.dim_w), "wa": (B, L, d.dim_a), "prev_x": (B, P, d.dim_w), "prev_wa": (B, P, d.dim_a),
                  "we": (B, L if dyn else 1, d.dim_e)}
        if dyn:
            if prev_we is None:
                raise ValueError("`we` is dynamic (T>1), but prev_we was not provided with prev_x/prev_wa.")     # FMT.py:304-307
            expect["prev_we"] = (B, P, d.dim_e)
        got = {"x": x, "wa": wa, "prev_x": prev_x, "prev_wa": prev_wa, "we": we, "prev_we": prev_we}
        for name, shp in expect.items():
            t = got[name]
            if not torch.is_tensor(t) or tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"{name}: expected a contiguous float32 {shp} tensor on {self.device}, got "
                                 f"{tuple(t.shape) if torch.is_tensor(t) else type(t)}")
        if not torch.is_tensor(r_s) or r_s.numel() != B * d.dim_w or r_s.dtype != torch.float32 or not r_s.is_contiguous() or r_s.device != self.device:
            raise ValueError(f"r_s: expected a contiguous float32 (B={B}, dim_w={d.dim_w}) tensor on {self.device}")
        if not 0 <= int(eval_index) < max(1, self._n_eval):
            raise ValueError(f"eval_index {eval_index} outside the plan's {self._n_eval} evaluations")
        v = torch.empty(x.shape[0], d.num_prev_frames + d.frames_per_clip, d.dim_w, device=self.device, dtype=torch.float32)
        ev = _cabi.FmtEval(x.data_ptr(), prev_x.data_ptr(), wa.data_ptr(), prev_wa.data_ptr(), we.data_ptr(),
                           prev_we.data_ptr() if prev_we is not None else None, r_s.data_ptr(), v.data_ptr(), int(eval_index),
                           float(a_cfg_scale), float(r_cfg_scale), float(e_cfg_scale))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.fmt_velocity(self._handle, C.byref(ev), self._stream()), "fmt_velocity")
        return v


# ----------------------------------------------------------------------------------------------------
# backend cache: one packed copy per (FMT module, device); rebuilt when the module's weights change
# ----------------------------------------------------------------------------------------------------
_BACKENDS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _weights_fingerprint(sd) -> tuple:
    return tuple((k, v.data_ptr(), getattr(v, "_version", 0), tuple(v.shape)) for k, v in sd.items() if torch.is_tensor(v))


def backend_for(fmt_model, device, dims: Optional[Dims] = None) -> FmtBackend:
    """``fmt_model``: anything with ``state_dict()`` (the reference ``FlowMatchingTransformer`` or :class:`FmtModel`)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise FmtError(f"float_fmt_model.target_device is '{device}': the B200 FMT sampler runs on CUDA (sm_100a) only; "
                       "there is no CPU fallback.")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    sd = fmt_model.state_dict()
    fp = _weights_fingerprint(sd)
    per_model = _BACKENDS.setdefault(fmt_model, {})
    hit = per_model.get(device)
    if hit is not None and hit[0] == fp:
        return hit[1]
    if dims is None:
        src = getattr(fmt_model, "final_construction_options", None) or getattr(fmt_model, "opt", None) or {}
        dims = Dims.from_options(src, sd)
    if hit is not None:
        hit[1].close()
    be = FmtBackend(sd, dims, device)
    per_model[device] = (fp, be)
    return be


def _as_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def draw_window_noise(batch: int, dims: Dims, n_windows: int, device, generator: Optional[torch.Generator]) -> torch.Tensor:
    """One ``torch.randn(B, L, dim_w)`` per window, in window order, on ``device`` (nodes_adv.py:606)."""
    gen_dev = generator.device if generator is not None else torch.device(device)
    chunks = [torch.randn(batch, dims.frames_per_clip, dims.dim_w, device=gen_dev, generator=generator) for _ in range(n_windows)]
    if not chunks:
        return torch.empty(0, batch, dims.frames_per_clip, dims.dim_w, device=device)
    return torch.stack(chunks, dim=0).to(device)


def perform_ode_sampling_loop(fmt_model, r_s_latent_dev, wa_latent_dev, we_latent_dev, audio_num_frames,
                              model_num_prev_frames, model_num_frames_for_clip, model_dim_w,
                              ode_nfe, ode_method, ode_atol, ode_rtol, target_device,
                              a_cfg_scale, r_cfg_scale, e_cfg_scale, include_r_cfg, noise_seed_generator,
                              progress_bar=None, mode: Optional[str] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for ``_perform_ode_sampling_loop`` (nodes_adv.py:545-694): same arguments, returns r_d on ``target_device``.

    ``ode_atol`` / ``ode_rtol`` are accepted and ignored, as torchdiffeq's fixed-grid solvers ignore them.
    ``noise`` (n_windows, B, L, dim_w) optionally injects the per-window x0 instead of drawing it.
    """
    device = torch.device(target_device)
    be = backend_for(fmt_model, device)
    d = be.dims
    if (d.num_prev_frames, d.frames_per_clip, d.dim_w) != (int(model_num_prev_frames), int(model_num_frames_for_clip), int(model_dim_w)):
        raise ValueError(f"window geometry (prev={model_num_prev_frames}, clip={model_num_frames_for_clip}, dim_w={model_dim_w}) "
                         f"does not match the FMT weights ({d.num_prev_frames}, {d.frames_per_clip}, {d.dim_w})")
    mode = resolve_mode(mode)
    # the shape errors the reference raises from inside torch (F.linear on cat[wr, wa, we], FMT.py:327-335) - raised here, before
    # any pointer reaches the kernels
    if wa_latent_dev.dim() != 3 or wa_latent_dev.shape[2] != d.dim_a:
        raise ValueError(f"wa_latent shape {tuple(wa_latent_dev.shape)}: expected (B, T, {d.dim_a}) - e.g. project the wav2vec features first")
    B = wa_latent_dev.shape[0]
    if we_latent_dev.dim() != 3 or we_latent_dev.shape[0] != B or we_latent_dev.shape[2] != d.dim_e:
        raise ValueError(f"we_latent shape {tuple(we_latent_dev.shape)}: expected ({B}, 1 or T, {d.dim_e})")
    if r_s_latent_dev.shape[0] != B or r_s_latent_dev.numel() != B * d.dim_w:
        raise ValueError(f"r_s_latent shape {tuple(r_s_latent_dev.shape)}: expected ({B}, {d.dim_w})")
    if int(audio_num_frames) < 1:
        raise ValueError(f"audio_num_frames must be >= 1 (got {audio_num_frames})")
    dynamic = we_latent_dev.shape[1] > 1
    nb = n_branches_for(a_cfg_scale, r_cfg_scale, e_cfg_scale, include_r_cfg)
    n_win = math.ceil(audio_num_frames / d.frames_per_clip)
    be.configure(B, nb, dynamic, ode_nfe, ode_method, mode)
    r_s, wa, we = _as_f32(r_s_latent_dev, be.device), _as_f32(wa_latent_dev, be.device), _as_f32(we_latent_dev, be.device)
    if noise is None:
        noise = draw_window_noise(B, d, n_win, be.device, noise_seed_generator)
    noise = _as_f32(noise, be.device)
    if noise.shape != (n_win, B, d.frames_per_clip, d.dim_w):
        raise ValueError(f"noise shape {tuple(noise.shape)} != {(n_win, B, d.frames_per_clip, d.dim_w)}")
    progress = (lambda w, n: progress_bar.update(1)) if progress_bar is not None else None
    return be.sample_clip(r_s, wa, we, int(audio_num_frames), noise, a_cfg_scale, r_cfg_scale, e_cfg_scale, progress)


def float_sample(fmt_model, opt, r_s, wa, we, a_cfg_scale=1.0, r_cfg_scale=1.0, e_cfg_scale=1.0, seed=None,
                 mode: Optional[str] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for the sampling part of ``FLOAT.sample`` (FLOAT.py:172-253) once ``wa`` (B,T,dim_a) and ``we`` (B,1,dim_e;
    float or the int64 one-hot of :200) exist.  Uses ``opt.nfe`` (the method's own ``nfe`` argument is ignored by the
    reference, :188), static emotion, 3 CFG branches at most, Euler (``opt.torchdiffeq_ode_method``)."""
    device = torch.device(opt.rank)
    g = None
    if getattr(opt, "fix_noise_seed", True):
        g = torch.Generator(device)
        g.manual_seed(opt.seed if seed is None else seed)
    be = backend_for(fmt_model, device)
    d = be.dims
    return perform_ode_sampling_loop(fmt_model, r_s, wa, we.to(torch.float32), wa.shape[1], d.num_prev_frames, d.frames_per_clip,
                                     d.dim_w, opt.nfe, getattr(opt, "torchdiffeq_ode_method", "euler"), getattr(opt, "ode_atol", 1e-5),
                                     getattr(opt, "ode_rtol", 1e-5), device, a_cfg_scale, r_cfg_scale, e_cfg_scale, False, g,
                                     mode=mode, noise=noise)


def float_sample_from_audio(agent_G, data, a_cfg_scale=1.0, r_cfg_scale=1.0, e_cfg_scale=1.0, emo=None, nfe=10, seed=None,
                            mode: Optional[str] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for the WHOLE of ``FLOAT.sample`` (FLOAT.py:172-253), bound in place of the method by
    :func:`use_b200_sampler`.  The conditioning front end keeps running the reference's own encoders (north_star): the audio
    encoder on the whole waveform (:191-194) and either the SER prediction or the forced int64 one-hot emotion (:196-200);
    the window loop + FMT + ODE solver behind it run in libfmt_b200.so.  ``nfe`` is accepted and ignored exactly as the
    reference ignores it (:188 uses ``self.opt.nfe``)."""
    import torch.nn.functional as F
    opt = agent_G.opt
    r_s = data["r_s"]
    a = data["a"].to(opt.rank)
    T = math.ceil(a.shape[-1] * opt.fps / opt.sampling_rate)
    wa = agent_G.audio_encoder.inference(a, seq_len=T)
    emo_idx = agent_G.emotion_encoder.label2id.get(str(emo).lower(), None)
    if emo_idx is None:
        we = agent_G.emotion_encoder.predict_emotion(a).unsqueeze(1)
    else:
        we = F.one_hot(torch.tensor(emo_idx, device=a.device), num_classes=opt.dim_e).unsqueeze(0).unsqueeze(0)
    return float_sample(agent_G.fmt, opt, r_s, wa, we, a_cfg_scale, r_cfg_scale, e_cfg_scale, seed=seed, mode=mode, noise=noise)


class use_b200_sampler:
    """``with use_b200_sampler(float_pipe.G):`` - inside the block ``G.sample`` (called by ``G.inference``, FLOAT.py:255-300) is
    this backend; the instance attribute is removed again on exit, so the reference module is left as it was."""

    def __init__(self, agent_G, mode: Optional[str] = None):
        self.G, self.mode = agent_G, mode

    def __enter__(self):
        G, mode = self.G, self.mode
        self._had = "sample" in vars(G)
        self._old = vars(G).get("sample")

        def sample(data, a_cfg_scale=1.0, r_cfg_scale=1.0, e_cfg_scale=1.0, emo=None, nfe=10, seed=None):
            return float_sample_from_audio(G, data, a_cfg_scale, r_cfg_scale, e_cfg_scale, emo, nfe, seed, mode=mode)
        try:
            object.__setattr__(G, "sample", sample)
        except Exception:
            G.sample = sample
        return G

    def __exit__(self, *exc):
        if self._had:
            object.__setattr__(self.G, "sample", self._old)
        else:
            try:
                object.__delattr__(self.G, "sample")
            except AttributeError:
                pass
        return False
