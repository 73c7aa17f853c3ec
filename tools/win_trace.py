"""Per-stage timing of the persistent window kernel from its barrier stamps (FMT_WIN_TRACE=1).  Run on the B200 box."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["FMT_WIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle.synth import FmtDims, synth_inputs, synth_state_dict  # noqa: E402

pkg = load_package()
d = FmtDims()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nfe = 10
dev = torch.device("cuda:0")
be = pkg.FmtBackend(synth_state_dict(d, seed=0), pkg.Dims(), dev)
be.configure(B, 3, False, nfe, "euler", "bf16")
r_s, wa, we = [t.to(dev) for t in synth_inputs(d, B, 50, seed=7)]
noise = torch.randn(1, B, 50, 512, device=dev)
for _ in range(3):
    be.sample_clip(r_s, wa, we, 50, noise, 2.0, 1.0, 1.0)
torch.cuda.synchronize()
n = be.lib.fmt_debug_window_trace(be._handle, None, 0)
buf = np.zeros(n, dtype=np.int64)
be.lib.fmt_debug_window_trace(be._handle, buf.ctypes.data_as(C.c_void_p), n)
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
full = buf.reshape(n_cta, -1, 6).astype(np.float64)
full = full[full[:, :, 0].max(axis=1) > 0]      # CTAs without work (grouped schedule: 148 - G * Cg) leave no stamps
tr = full[:, :, :2]
marks = full[:, :, 2:]
nbar = tr.shape[1]
per_eval = nbar // (nfe - 1)
if os.environ.get("FMT_WINDOW") == "2" and os.environ.get("FMT_WIN_FUSE_GELU", "1") != "0":
    per_eval = 4 + 7 * 8          # grouped schedule with the GELU fused into fc1's epilogue (the trace buffer keeps the 68-stage stride)
blk_names = ["qkv G", "attn", "proj G", "row", "fc1 G", "gelu", "fc2 G", "row2"] if per_eval == 4 + 8 * 8 else ["qkv G", "attn", "proj G", "row", "fc1 G", "fc2 G", "row2"]
names = ["x_emb G", "row0"] + sum([[f"b{b} {n}" for n in blk_names] for b in range(8)], []) + ["dec G", "comb"]
clk = 1.9  # GHz, approximate (SM clocks)
work = tr[:, 1:, 0] - tr[:, :-1, 1]          # stage s work = arrive[s] - pass[s-1]   (per CTA)
wait = tr[:, :, 1] - tr[:, :, 0]             # barrier wait
span = tr[:, 1:, 1] - tr[:, :-1, 1]          # pass-to-pass = full stage duration
e = 4                                         # a steady-state evaluation
print(f"{'stage':12s} {'span us':>8s} {'work max':>9s} {'work med':>9s} {'wait min':>9s}")
tot = 0.0
for s in range(per_eval):
    k = e * per_eval + s - 1                  # index into the diff arrays
    sp = np.median(span[:, k]) / clk / 1e3
    tot += sp
    print(f"{names[s]:12s} {sp:8.2f} {work[:, k].max() / clk / 1e3:9.2f} {np.median(work[:, k]) / clk / 1e3:9.2f} {wait[:, k + 1].min() / clk / 1e3:9.2f}")
print("sum of spans per evaluation: %.1f us" % tot)
# intra-stage marks (thread 32) of the stage FOLLOWING barrier k, relative to that barrier's pass stamp
for s_ in range(per_eval):
    k = e * per_eval + s_ - 1
    m = marks[:, k, :] - tr[:, k, 1:2]
    if (marks[:, k, :] > 0).any():
        mm = np.where(marks[:, k, :] > 0, m, np.nan)
        med = np.nanmedian(mm, axis=0) / clk / 1e3
        print(f"  marks {names[s_]:12s} " + " ".join(f"{x:7.2f}" for x in med) + f"   arrive {np.median(tr[:, k + 1, 0] - tr[:, k, 1]) / clk / 1e3:6.2f}")
kinds = {}
for s in range(per_eval):
    k = e * per_eval + s - 1
    key = names[s].split(" ", 1)[-1] if names[s][0] == "b" else names[s]
    kinds.setdefault(key, []).append(np.median(span[:, k]) / clk / 1e3)
for k_, v in kinds.items():
    print(f"  {k_:8s} mean span {np.mean(v):6.2f} us x {len(v)}")
