"""Timeline of one evaluation of the dataflow window kernel (FMT_WIN_TRACE=1): per (stage, chunk) when the dependency of the GEMM
items was satisfied, when accumulators were ready, when the chunk was published, when the SIMT units saw it and released it.
SM clocks are mapped to global time with two (globaltimer, clock64) pairs per CTA.  Run on the B200 box."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["FMT_WIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
synth = pkg.synth
d = synth.FmtDims()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
verbose = len(sys.argv) > 2
nfe = 10
dev = torch.device("cuda:0")
be = pkg.FmtBackend(synth.synth_state_dict(d, seed=0), pkg.Dims(), dev)
be.configure(B, 3, False, nfe, "euler", "bf16")
r_s, wa, we = [t.to(dev) for t in synth.synth_inputs(d, B, 50, seed=7)]
noise = torch.randn(1, B, 50, 512, device=dev)
for _ in range(3):
    be.sample_clip(r_s, wa, we, 50, noise, 2.0, 1.0, 1.0)
torch.cuda.synchronize()
n = be.lib.fmt_debug_window_trace(be._handle, None, 0)
buf = np.zeros(n, dtype=np.int64)
be.lib.fmt_debug_window_trace(be._handle, buf.ctypes.data_as(C.c_void_p), n)
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
CH = int(os.environ.get("FMT_FLOW_CH", "64"))
n_gemms = 2 + 4 * d.fmt_depth
nsub = -(-(d.num_prev_frames + d.frames_per_clip) // CH)
n_chunks = 3 * B * nsub
n_main = n_cta * 2 * n_gemms * n_chunks * 8
tr = buf[:n_main].reshape(n_cta, 2, n_gemms, n_chunks, 8).astype(np.float64)
cal = buf[n_main:n_main + n_cta * 4].reshape(n_cta, 2, 2).astype(np.float64)
dclk = cal[:, 1, 1] - cal[:, 0, 1]                                            # SM clocks between the two rendezvous, per CTA (GPC clocks differ by up to ~0.5 %)
T_ns = float(np.median(cal[:, 1, 0] - cal[:, 0, 0]))
freq = dclk / T_ns                                                            # SM clocks per ns, per CTA
print("SM clock %.4f .. %.4f GHz; kernel %.1f us" % (freq.min(), freq.max(), T_ns / 1e3))
t = np.where(tr > 0, (tr - cal[:, 0, 1][:, None, None, None, None]) / freq[:, None, None, None, None], np.nan)
t[:, 0, :, :, 6] = np.nan
t0 = np.nanmin(t)
t = (t - t0) / 1e3                                                             # us since the first stamp of the traced evaluation
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "flow_trace_%s.npz" % os.environ.get("FLOW_TRACE_TAG", "t")),
                    t=t, raw=tr, n_gemms=n_gemms, n_chunks=n_chunks)   # [cta][engine][stage][chunk][slot] in us, for offline analysis
names = ["x_emb", "qkv", "proj", "fc1", "fc2"]
snames = ["row0"] + ["attn", "row", "gelu", "row2"] * d.fmt_depth + ["comb"]


def gname(g):
    return "x_emb" if g == 0 else "dec" if g == n_gemms - 1 else f"b{(g - 1) // 4} " + ["qkv", "proj", "fc1", "fc2"][(g - 1) % 4]


print(f"{'stage':10s} {'c':>2s} | {'G wait0':>8s} {'dep ok(first/last)':>19s} {'load':>7s} {'acc rdy':>8s} {'red iss':>8s} {'pub(med/last)':>15s} | {'S wait0':>8s} {'S ok(first/last)':>17s} {'work':>7s} {'fence':>7s} {'rel last':>8s}")
prev_end = 0.0
stage_rows = []
for g in range(n_gemms):
    for c in range(n_chunks):
        G = t[:, 0, g, c, :]
        S = t[:, 1, g, c, :]
        with np.errstate(all="ignore"):
            row = (np.nanmedian(G[:, 0]), np.nanmin(G[:, 1]), np.nanmax(G[:, 1]), np.nanmedian(G[:, 2]), np.nanmedian(G[:, 3]), np.nanmedian(G[:, 4]),
                   np.nanmedian(G[:, 5]), np.nanmax(G[:, 5]), np.nanmedian(S[:, 0]), np.nanmin(S[:, 1]), np.nanmax(S[:, 1]), np.nanmedian(S[:, 2]),
                   np.nanmedian(S[:, 3]), np.nanmax(S[:, 4]))
        stage_rows.append((g, c) + row)
        if verbose or g < 6 or g >= n_gemms - 2:
            print(f"{gname(g):10s} {c:2d} | {row[0]:8.2f} {row[1]:9.2f} {row[2]:9.2f} {row[3]:7.2f} {row[4]:8.2f} {row[5]:8.2f} {row[6]:7.2f} {row[7]:7.2f} | {row[8]:8.2f} {row[9]:8.2f} {row[10]:8.2f} {row[11]:7.2f} {row[12]:7.2f} {row[13]:8.2f}   {snames[g]}")
a = np.array(stage_rows)
print("\nper stage (all chunks): GEMM span = first dep ok -> last publish; SIMT span = first ok -> last release; stage end-to-end")
tot = {}
for g in range(n_gemms):
    r = a[a[:, 0] == g]
    g_first, g_last = np.nanmin(r[:, 3]), np.nanmax(r[:, 9])
    s_first, s_last = np.nanmin(r[:, 11]), np.nanmax(r[:, 15])
    print(f"{gname(g):10s} GEMM {g_first:8.2f} -> {g_last:8.2f} ({g_last - g_first:5.2f})   {snames[g]:5s} {s_first:8.2f} -> {s_last:8.2f} ({s_last - s_first:5.2f})   chunk-chain step (release c0 -> release c0 of the previous stage): {r[0, 15] - prev_end:5.2f}")
    prev_end = r[0, 15]
print("evaluation span: %.1f us" % np.nanmax(t))
# per-chunk chain decomposition, averaged over the block stages: dep ok (last CTA) - previous SIMT release; acc ready - dep ok; publish - acc; S ok - publish; S release - S ok
for c in range(n_chunks):
    r = a[a[:, 1] == c]
    dep = r[1:, 4] - r[:-1, 15]
    print(f"chunk {c}: release->dep ok(last) {np.nanmean(dep):5.2f} | dep ok(last)->acc(med) {np.nanmean(r[1:, 6] - r[1:, 4]):5.2f} | acc->pub(last) {np.nanmean(r[1:, 9] - r[1:, 6]):5.2f} | pub->S ok(last) {np.nanmean(r[1:, 12] - r[1:, 9]):5.2f} | S ok->release(last) {np.nanmean(r[1:, 15] - r[1:, 12]):5.2f}")
# ---- engine view: how long each engine of a CTA waits / works per (stage, chunk), averaged over block stages and CTAs
with np.errstate(all="ignore"):
    G = t[:, 0, 1:n_gemms - 1]
    S = t[:, 1, 1:n_gemms - 1]
    print("\nGEMM engine per (stage, chunk), mean over CTAs with an item: loader wait for the flag %.2f | for a ring slot %.2f | load+MMA (slot -> acc) %.2f | "
          "acc -> reduce issued %.2f | issued -> published %.2f" % (np.nanmean(G[..., 1] - G[..., 0]), np.nanmean(G[..., 2] - G[..., 1]), np.nanmean(G[..., 3] - G[..., 2]),
                                                                   np.nanmean(G[..., 4] - G[..., 3]), np.nanmean(G[..., 5] - G[..., 4])))
    print("  GEMM front end: copy issued -> tile landed (seen by the MMA warp) %.2f | landed -> accumulator ready (seen by the epilogue) %.2f" % (
        np.nanmean(G[..., 7] - G[..., 2]), np.nanmean(G[..., 3] - G[..., 7])))
    for j_, nm_ in enumerate(("qkv", "proj", "fc1", "fc2")):
        Gj = G[:, j_::4]
        print("    %-4s: issued -> landed %.2f | landed -> acc ready %.2f | acc -> reduce issued %.2f | issued -> published %.2f" % (
            nm_, np.nanmean(Gj[..., 7] - Gj[..., 2]), np.nanmean(Gj[..., 3] - Gj[..., 7]), np.nanmean(Gj[..., 4] - Gj[..., 3]), np.nanmean(Gj[..., 5] - Gj[..., 4])))
    print("SIMT engine per (stage, chunk) with units: wait %.2f | work %.2f | fence+bar %.2f | release %.2f" % (
        np.nanmean(S[..., 1] - S[..., 0]), np.nanmean(S[..., 2] - S[..., 1]), np.nanmean(S[..., 3] - S[..., 2]), np.nanmean(S[..., 4] - S[..., 3])))
    for kind, sl in (("attn", slice(0, None, 4)), ("row", slice(1, None, 4)), ("gelu", slice(2, None, 4)), ("row2", slice(3, None, 4))):
        K = S[:, sl]
        print(f"  {kind:5s}: wait {np.nanmean(K[..., 1] - K[..., 0]):5.2f} work {np.nanmean(K[..., 2] - K[..., 1]):5.2f} fence {np.nanmean(K[..., 3] - K[..., 2]):5.2f} "
              f"release {np.nanmean(K[..., 4] - K[..., 3]):5.2f}  (stage, chunk) pairs per CTA and stage: {np.mean(np.sum(~np.isnan(K[..., 1]), axis=2)):.2f}")
    A = S[:, 0::4]
    print("  attn detail (thread 0 of the SIMT engine): ok -> scores computed (loads landed) %.2f | shuffles %.2f | softmax %.2f | outputs + stores %.2f" % (
        np.nanmean(A[..., 5] - A[..., 1]), np.nanmean(A[..., 6] - A[..., 5]), np.nanmean(A[..., 7] - A[..., 6]), np.nanmean(A[..., 2] - A[..., 7])))
cta = os.environ.get("FLOW_TRACE_CTA")
if cta is not None:
    cta = int(cta)
    ev = []
    lab = [["G wait", "G dep ok", "G slot", "G acc", "G red", "G pub"], ["S wait", "S ok", "S work", "S fence", "S rel"]]
    for eng in range(2):
        for g in range(n_gemms):
            for c in range(n_chunks):
                for k, name in enumerate(lab[eng]):
                    v = t[cta, eng, g, c, k]
                    if not np.isnan(v):
                        ev.append((v, f"{name:9s} {gname(g) if eng == 0 else snames[g]:9s} st{g:2d} c{c}"))
    ev.sort()
    lo, hi = float(os.environ.get("FLOW_TRACE_T0", "150")), float(os.environ.get("FLOW_TRACE_T1", "215"))
    print(f"\nevents of CTA {cta} between {lo} and {hi} us")
    for v, s_ in ev:
        if lo <= v <= hi:
            print(f"{v:9.2f}  {s_}")
# ---- spread over CTAs: when the items of a (stage, chunk) published / when its SIMT units released, relative to the first one (quantiles, mean over block stages)
with np.errstate(all="ignore"):
    def spread(x):   # x: [cta, stage, chunk]
        lo = np.nanmin(x, axis=0, keepdims=True)
        q = np.nanpercentile(x - lo, [10, 50, 90, 100], axis=0)          # [4, stage, chunk]
        return np.nanmean(q, axis=(1, 2))
    print("\nspread over CTAs relative to the earliest (10 / 50 / 90 / 100 %%): GEMM dep ok %s | acc ready %s | published %s" % (
        np.round(spread(G[..., 1]), 2), np.round(spread(G[..., 3]), 2), np.round(spread(G[..., 5]), 2)))
    print("                                                              SIMT ok %s | work done %s | released %s" % (
        np.round(spread(S[..., 1]), 2), np.round(spread(S[..., 2]), 2), np.round(spread(S[..., 4]), 2)))
    polls = tr[:, 0, 1:n_gemms - 1, :, 6]
    waited = (G[..., 1] - G[..., 0])
    m = polls > 2
    print("loader polls per flag wait: mean %.1f; poll period (waits with > 2 polls) %.3f us" % (np.mean(polls[polls > 0]), np.nanmean(waited[m] / (polls[m] - 1))))
    # which CTAs publish last?  rank of each CTA's publish time per (stage, chunk), averaged
    rk = np.argsort(np.argsort(np.where(np.isnan(G[..., 5]), 1e9, G[..., 5]), axis=0), axis=0).astype(np.float64)
    rk[np.isnan(G[..., 5])] = np.nan
    mean_rank = np.nanmean(rk, axis=(1, 2))
    order = np.argsort(-np.nan_to_num(mean_rank))
    print("CTAs that publish latest on average (cta: mean rank of %d): %s" % (n_cta, ", ".join(f"{c_}:{mean_rank[c_]:.0f}" for c_ in order[:12])))
    print("CTAs that publish earliest: %s" % ", ".join(f"{c_}:{mean_rank[c_]:.0f}" for c_ in order[-8:]))
