"""GPU side of the PSNR >= 40 dB acceptance check (BASELINE.json north_star): samples config 1 (1 clip, 100 frames, nfe 10,
a_cfg 2, e_cfg 1) with the CUDA path in bf16 and fp32-validation mode and with the oracle on the same device and noise, and
writes the three latent sequences to gpurun_out/psnr_latents.npz.  tools/psnr_check.py (build container, where the reference
decoder exists) turns them into frames and PSNR."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import fmt_oracle as O  # noqa: E402
from oracle.synth import FmtDims, synth_inputs, synth_state_dict  # noqa: E402

pkg = load_package()
d = FmtDims()
dev = "cuda:0"
W = synth_state_dict(d, seed=0)
T = 100
r_s, wa, we = synth_inputs(d, 1, T, seed=7)
g = torch.Generator().manual_seed(15)
noise = torch.stack([torch.randn(1, d.frames_per_clip, d.dim_w, generator=g) for _ in range(2)])
model = pkg.FmtModel(W, target_device=dev)
node = pkg.FloatSampleMotionSequenceRD_VA()
args = (2.0, 1.0, 1.0, False, 10, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 15)
out = {}
for mode in ("bf16", "fp32"):
    out[mode] = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _mode=mode, _noise=noise)[0].numpy()
torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    ref = O.sample_loop({k: v.to(dev) for k, v in W.items()}, d, r_s.to(dev), wa.to(dev), we.to(dev), T, nfe=10, a_cfg_scale=2.0,
                        e_cfg_scale=1.0, noise=noise.to(dev)).cpu().numpy()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "psnr_latents.npz"), bf16=out["bf16"], fp32=out["fp32"], ref=ref)
for mode in ("bf16", "fp32"):
    print(f"{mode}: max|r_d - oracle| = {np.abs(out[mode] - ref).max():.3e}, rel = {np.linalg.norm(out[mode] - ref) / np.linalg.norm(ref):.3e}")
