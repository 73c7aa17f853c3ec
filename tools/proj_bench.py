"""Audio projection timing on the B200 box: our kernel chain (fp32 -> bf16 convert, tcgen05 GEMM K = 9216,
LayerNorm + SiLU) against eager PyTorch (fp32 without TF32, as the reference runs it) on the same device, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle import fmt_oracle as O  # noqa: E402
from oracle.synth import synth_projection, synth_wav2vec_features  # noqa: E402

pkg = load_package()
dev = torch.device("cuda:0")
layer = pkg.AudioProjectionLayer(9216, 512, target_device=dev)
layer.load_state_dict(synth_projection(9216, 512, seed=0))
be = pkg.projection_backend_for(layer, dev)
P = {k: v.to(dev) for k, v in synth_projection(9216, 512, seed=0).items()}
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B, T in ((1, 100), (32, 200), (64, 200), (256, 200)):
    x = synth_wav2vec_features(B, T, 9216, seed=1).to(dev)
    rows = B * T
    t_ours = timed(lambda: be.apply(x, "bf16"))
    with torch.no_grad():
        t_ref = timed(lambda: O.audio_projection(P, x), n=5)
        err = float((be.apply(x, "bf16") - O.audio_projection(P, x)).abs().max())
    flops = 2.0 * rows * 9216 * 512
    bytes_alg = rows * 9216 * 4 + 512 * 9216 * 2 + rows * 512 * 4          # fp32 features in, bf16 weights, fp32 wa out
    print(f"rows {rows:6d} ({B} clips x {T} frames): ours {t_ours:9.1f} us  {flops / t_ours / 1e6:7.1f} TFLOP/s  {bytes_alg / t_ours / 1e3:7.1f} GB/s algorithmic"
          f" | eager fp32 {t_ref:9.1f} us  x{t_ref / t_ours:5.1f} | max|diff| {err:.2e}")
