"""Dataflow window kernel (FMT_WINDOW=3) against the one-kernel-per-op schedule (FMT_WINDOW=0) + per-step timing.  Run on the B200 box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
synth = pkg.synth
d = synth.FmtDims()
dev = torch.device("cuda:0")
W = synth.synth_state_dict(d, seed=0)
cfgs = [(1, 3, "euler", 10, (2.0, 1.0, 1.0, False)), (1, 1, "euler", 4, (1.0, 1.0, 1.0, False)), (1, 4, "midpoint", 3, (2.0, 0.5, 1.5, True)),
        (2, 1, "heun3", 3, (1.0, 1.0, 1.0, False))]
if len(sys.argv) > 1:
    cfgs = cfgs[:int(sys.argv[1])]
for B, nb, method, nfe, (a, r, e, inc) in cfgs:
    T = 120
    r_s, wa, we = [t.to(dev) for t in synth.synth_inputs(d, B, T, seed=21)]
    g = torch.Generator().manual_seed(5)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(3)]).to(dev)
    outs = {}
    for flag in ("0", "3", "1"):
        os.environ["FMT_WINDOW"] = flag
        be = pkg.FmtBackend(W, pkg.Dims(), dev)
        be.configure(B, pkg.n_branches_for(a, r, e, inc), False, nfe, method, "bf16")
        st0 = be.window_kernel_status()
        outs[flag] = be.sample_clip(r_s, wa, we, T, noise, a, r, e).cpu()
        torch.cuda.synchronize()
        st1 = be.window_kernel_status()
        # timing
        for _ in range(3):
            be.sample_clip(r_s, wa, we, T, noise, a, r, e)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_it = 10
        for _ in range(n_it):
            be.sample_clip(r_s, wa, we, T, noise, a, r, e)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_it
        n_eval = 3 * (nfe - 1) * {"euler": 1, "midpoint": 2, "heun3": 3}[method]
        again = be.sample_clip(r_s, wa, we, T, noise, a, r, e).cpu()
        print(f"B={B} nb={nb} {method} nfe={nfe} FMT_WINDOW={flag}: status {st0}->{st1}  {dt * 1e6 / n_eval:8.1f} us/eval  finite={bool(torch.isfinite(outs[flag]).all())}"
              f"  max|x-per_op|={float((outs[flag] - outs['0']).abs().max()):.3e}  repeat_equal={bool(torch.equal(again, outs[flag]))}", flush=True)
        be.close()
