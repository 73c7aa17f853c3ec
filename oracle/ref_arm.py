"""CPU arm of ``bench.py``: the REAL reference sampler node timed on the host cores (TEST / BENCH INFRASTRUCTURE).

``make_cpu_sampler`` returns the reference's own ``FloatSampleMotionSequenceRD_VA.sample_rd_sequence_va``
(``src/nodes/nodes_vadv.py:618-735``) bound to a reference ``FlowMatchingTransformer`` that carries the seeded synthetic
weights, imported unmodified through ``oracle/refshim.py`` from ``/root/reference`` (build container) or from the git-ignored
copy ``oracle/_ref/reference`` that ``oracle/make_ref.py`` ships to the GPU box.  ``kind`` is "reference" then.  Only when no
copy of the reference exists does it fall back to the oracle port (``oracle/fmt_oracle.py``), ``kind`` "port".
"""
import importlib
import os

import torch

from . import fmt_oracle as O
from . import refshim


def make_cpu_sampler(W, dims, nfe, a_cfg, r_cfg, e_cfg, force_port=False, threads=None):
    """-> (kind, fn) with fn(r_s, wa, we, T, seed) -> r_d (B, T, dim_w) on the CPU, fp32, all host threads (or `threads` of them
    when several processes share the host, one per rank)."""
    torch.set_num_threads(threads or os.cpu_count() or 1)
    if refshim.reference_available() and not force_port:
        _, model, _ = refshim.build_reference_fmt(W)
        node = importlib.import_module("refnodes.nodes_vadv").FloatSampleMotionSequenceRD_VA()

        @torch.no_grad()
        def run_ref(r_s, wa, we, T, seed):
            out, _ = node.sample_rd_sequence_va(
                r_s_latent=r_s, wa_latent=wa, we_latent=we, audio_num_frames=T, float_fmt_model=model, a_cfg_scale=a_cfg,
                r_cfg_scale=r_cfg, e_cfg_scale=e_cfg, include_r_cfg=False, nfe=nfe, torchdiffeq_ode_method="euler", ode_atol=1e-5,
                ode_rtol=1e-5, audio_dropout_prob=0.1, ref_dropout_prob=0.1, emotion_dropout_prob=0.1, fix_noise_seed=True, seed=seed)
            return out
        return "reference", run_ref

    @torch.no_grad()
    def run_port(r_s, wa, we, T, seed):
        g = torch.Generator().manual_seed(seed)
        L = dims.frames_per_clip
        n_win = -(-T // L)
        noise = torch.stack([torch.randn(r_s.shape[0], L, dims.dim_w, generator=g) for _ in range(n_win)])
        return O.sample_loop(W, dims, r_s, wa, we, T, nfe=nfe, a_cfg_scale=a_cfg, r_cfg_scale=r_cfg, e_cfg_scale=e_cfg, noise=noise)
    return "port", run_port


def describe(kind):
    return ("unmodified reference node FloatSampleMotionSequenceRD_VA (nodes_vadv.py:618-735; torch fp32 eager, CPU)" if kind == "reference"
            else "oracle port of the reference (oracle/fmt_oracle.py, torch fp32, CPU)")
