"""Host-side profile of one end-to-end node call (CPU tensors in, CPU tensors out) for configs[1].  Run on the B200 box."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
synth = pkg.synth
d = synth.FmtDims()
dev = torch.device("cuda:0")
model = pkg.FmtModel(synth.synth_state_dict(d, seed=0), target_device=dev)
r_s, wa, we = [t.pin_memory() for t in synth.synth_inputs(d, 1, 100, seed=7)]
node = pkg.FloatSampleMotionSequenceRD_VA()
args = (2.0, 1.0, 1.0, False, 10, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 15)


def call():
    return node.sample_rd_sequence_va(r_s, wa, we, 100, model, *args)[0]


for _ in range(5):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    call()
torch.cuda.synchronize()
print("e2e %.3f ms per call" % ((time.perf_counter() - t0) * 20))
be = pkg.backend_for(model, dev)
rd, wd, ed = r_s.to(dev), wa.to(dev), we.to(dev)
noise = torch.randn(2, 1, 50, 512, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    be.sample_clip(rd, wd, ed, 100, noise, 2.0, 1.0, 1.0)
torch.cuda.synchronize()
print("resident %.3f ms per call" % ((time.perf_counter() - t0) * 20))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
