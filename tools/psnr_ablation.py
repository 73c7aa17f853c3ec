"""PSNR >= 40 dB gate (BASELINE.json north_star), precision ablation on the CPU - build container or any box with oracle/_ref.

Question ("Next round" 2): which GEMM operands would have to stay above bf16 for frames decoded by the REFERENCE
decoder to reach 40 dB against frames decoded from the reference latents, and what would that cost?  Everything here is the
reference algorithm (oracle/fmt_oracle.py, pinned to the reference's fixtures) with the ROUNDING of a candidate operand format
emulated (``Quant``); the decoder is the reference ``Generator`` (random-init seed 0, random 512x512 portrait).
Also measured: the decoder's own sensitivity (PSNR of frames decoded from reference latents + Gaussian noise of a given size).

    python tools/psnr_ablation.py [--out profiles/r02_psnr.json] [--frames 0,13,49,50,77,99]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fmt_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402
from oracle.synth import FmtDims, synth_inputs, synth_state_dict  # noqa: E402


def round_to(dtype):
    return lambda t: t.to(dtype).to(torch.float32)


def round_mantissa(bits):
    """keep `bits` explicit mantissa bits (round to nearest even): 10 = tf32 / fp16 mantissa without fp16's range limits"""
    drop = 23 - bits

    def f(t):
        i = t.contiguous().view(torch.int32)
        bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
        return ((i + bias) >> drop << drop).view(torch.float32)
    return f


def split2(t):   # hi + lo bf16 pair: what a 3-MMA bf16x2 scheme feeds the tensor core (~16 mantissa bits)
    hi = t.to(torch.bfloat16).to(torch.float32)
    return hi + (t - hi).to(torch.bfloat16).to(torch.float32)


def build_decoder():
    refshim.load_reference()
    Generator = importlib.import_module("refnodes.models.float.generator").Generator
    torch.manual_seed(0)
    gen = Generator(512, 512, 20).eval()
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    with torch.no_grad():
        s_r, _, feats = gen.enc(img, None, None)
    return gen, s_r, feats


@torch.no_grad()
def decode(gen, s_r, feats, r_d, frames):
    """FLOAT.py:137-153: frame t = dec(s_r + r_d[:, t]) -> clamp(-1, 1) -> [0, 1]"""
    return [((gen.dec(s_r + r_d[:, t], None, feats)[0].clamp(-1, 1) + 1) / 2) for t in frames]


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else float(10 * np.log10(1.0 / mse))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_psnr.json"))
    ap.add_argument("--frames", default="0,13,49,50,77,99")
    ap.add_argument("--gpu-latents", default=os.path.join(ROOT, "gpurun_out", "psnr_latents.npz"),
                    help="optional: latents of the CUDA path (tools/psnr_dump.py) to put beside the emulations")
    a = ap.parse_args()
    frames = [int(x) for x in a.frames.split(",")]
    torch.set_num_threads(os.cpu_count() or 1)
    d = FmtDims()
    W = synth_state_dict(d, seed=0)
    T = 100
    r_s, wa, we = synth_inputs(d, 1, T, seed=7)
    g = torch.Generator().manual_seed(15)
    noise = torch.stack([torch.randn(1, d.frames_per_clip, d.dim_w, generator=g) for _ in range(2)])

    def sample(q):
        with torch.no_grad():
            return O.sample_loop(W, d, r_s, wa, we, T, nfe=10, a_cfg_scale=2.0, e_cfg_scale=1.0, noise=noise, q=q)

    bf, f16, m10 = round_to(torch.bfloat16), round_to(torch.float16), round_mantissa(10)
    edge = ("x_embedder", "decoder.linear")
    variants = {
        "bf16 weights + activations + table (production mode)": (O.Quant(act=bf, weight=bf, table=bf), 1.0),
        "bf16 weights only": (O.Quant(weight=bf), 1.0),
        "bf16 activations only": (O.Quant(act=bf), 1.0),
        "bf16 table only": (O.Quant(table=bf), 1.0),
        "bf16, x_embedder + decoder.linear exact": (O.Quant(act=bf, weight=bf, table=bf, keep_fp32=edge), 1.0),
        "bf16, fp32 table": (O.Quant(act=bf, weight=bf), 1.09),
        "bf16, all attention GEMMs (qkv, proj) exact": (O.Quant(act=bf, weight=bf, table=bf, keep_fp32=("attn.",)), 1.33),
        "bf16, all MLP GEMMs (fc1, fc2) exact": (O.Quant(act=bf, weight=bf, table=bf, keep_fp32=("mlp.",)), 1.66),
        "fp16 weights + activations + table (same bytes as bf16)": (O.Quant(act=f16, weight=f16, table=f16), 1.0),
        "10-bit mantissa (tf32 operands: 2x weight bytes)": (O.Quant(act=m10, weight=m10, table=m10), 2.0),
        "bf16 x2 split operands (3 MMAs, 2x weight bytes)": (O.Quant(act=split2, weight=split2, table=split2), 2.0),
    }
    gen, s_r, feats = build_decoder()
    ref = sample(O.Quant())
    ref_frames = decode(gen, s_r, feats, ref, frames)
    res = dict(gate="PSNR >= 40 dB on frames decoded by the reference decoder (random-init seed 0) vs frames decoded from the reference latents",
               workload="configs[1]: 1 clip, 100 frames, nfe 10, a_cfg 2, e_cfg 1; frames " + a.frames, variants={}, decoder_sensitivity={})
    for name, (q, bytes_x) in variants.items():
        lat = sample(q)
        fr = decode(gen, s_r, feats, lat, frames)
        p = [psnr(x, y) for x, y in zip(fr, ref_frames)]
        err = (lat - ref)
        res["variants"][name] = dict(min_psnr_db=min(p), mean_psnr_db=float(np.mean(p)), max_abs_latent_err=float(err.abs().max()),
                                     rms_latent_err=float(err.pow(2).mean().sqrt()), weight_bytes_vs_bf16=bytes_x, **{"pass": min(p) >= 40})
        print(f"{name:62s} max|err| {err.abs().max():.2e}  rms {err.pow(2).mean().sqrt():.2e}  PSNR min {min(p):6.2f} dB")
    gn = torch.Generator().manual_seed(11)
    for sigma in (1e-7, 1e-6, 3e-6, 1e-5, 1e-4, 1e-3, 4e-3):
        lat = ref + sigma * torch.randn(ref.shape, generator=gn)
        p = [psnr(x, y) for x, y in zip(decode(gen, s_r, feats, lat, frames), ref_frames)]
        res["decoder_sensitivity"][f"{sigma:.0e}"] = dict(latent_noise_rms=sigma, min_psnr_db=min(p), mean_psnr_db=float(np.mean(p)))
        print(f"decoder sensitivity: latent noise rms {sigma:.0e} -> PSNR min {min(p):6.2f} dB")
    if os.path.exists(a.gpu_latents):
        z = np.load(a.gpu_latents)
        zref = torch.from_numpy(z["ref"])
        zf = decode(gen, s_r, feats, zref, frames)
        res["cuda_path"] = {}
        for key in ("bf16", "fp32"):
            if key in z:
                lat = torch.from_numpy(z[key])
                p = [psnr(x, y) for x, y in zip(decode(gen, s_r, feats, lat, frames), zf)]
                res["cuda_path"][key] = dict(min_psnr_db=min(p), max_abs_latent_err=float((lat - zref).abs().max()), **{"pass": min(p) >= 40},
                                             source=os.path.relpath(a.gpu_latents, ROOT))
                print(f"CUDA path {key}: PSNR min {min(p):6.2f} dB")
    json.dump(res, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
