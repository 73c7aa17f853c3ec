"""PSNR >= 40 dB acceptance check (BASELINE.json north_star): frames decoded by the REFERENCE decoder from the CUDA path's
motion latents vs frames decoded from the reference-algorithm latents.  Build container only (needs /root/reference).

    python tools/psnr_check.py [gpurun_out/psnr_latents.npz] [--frames 0,13,49,50,77,99]

Decoder = the reference ``Generator`` (motion auto-encoder), random-init with seed 0, on a random 512x512 portrait
(SURVEY.md §8d); frame t = dec(s_r + r_d[t]) -> clamp(-1,1) -> [0,1]  (FLOAT.py:137-153)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import refshim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("latents", nargs="?", default=os.path.join(ROOT, "gpurun_out", "psnr_latents.npz"))
    ap.add_argument("--frames", default="0,13,49,50,77,99")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_psnr.json"))
    a = ap.parse_args()
    z = np.load(a.latents)
    refshim.load_reference()
    Generator = importlib.import_module("refnodes.models.float.generator").Generator
    torch.manual_seed(0)
    gen = Generator(512, 512, 20).eval()
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    frames = [int(x) for x in a.frames.split(",")]
    with torch.no_grad():
        s_r, _, feats = gen.enc(img, None, None)      # appearance latent + feature pyramid of the portrait
        res = {}
        for name in ("bf16", "fp32"):
            worst = float("inf")
            for t in frames:
                outs = []
                for key in ("ref", name):
                    r_d = torch.from_numpy(z[key])[:, t]
                    im, _ = gen.dec(s_r + r_d, None, feats)
                    outs.append(((im.clamp(-1, 1) + 1) / 2))
                mse = float(((outs[0] - outs[1]) ** 2).mean())
                psnr = 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)
                worst = min(worst, psnr)
                print(f"{name} frame {t:3d}: PSNR {psnr:6.2f} dB")
            res[name] = dict(min_psnr_db=worst, frames=frames, max_abs_latent_err=float(np.abs(z[name] - z["ref"]).max()))
    res["gate"] = "PSNR >= 40 dB"
    res["pass"] = all(v["min_psnr_db"] >= 40 for k, v in res.items() if isinstance(v, dict))
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
