"""The oracle (oracle/fmt_oracle.py) against every golden fixture produced by the real reference.

CPU only.  This is what makes the oracle "pinned" (see oracle/fmt_oracle.py header)."""
import pytest
import torch

import cases
from oracle import fmt_oracle as O

FAST = [n for n, r in cases.CASES.items() if r["dims"] == "small" or r["entry"] in ("cfv", "proj") or r.get("nfe", 99) <= 4]
ALL = list(cases.CASES)


def run_oracle(name, device="cpu", q=None, noise=None, W=None):
    rec = cases.CASES[name]
    if rec["entry"] == "proj":      # FloatApplyAudioProjection
        P = {k: v.to(device) for k, v in cases.projection_weights(rec).items()}
        with torch.no_grad():
            return O.audio_projection(P, cases.projection_input(rec).to(device))
    d = cases.dims_of(rec)
    W = W if W is not None else cases.weights(rec["dims"])
    r_s, wa, we = [t.to(device) for t in cases.case_inputs(rec)]
    kw = {} if q is None else {"q": q}
    entry = rec["entry"]
    with torch.no_grad():
        if entry in ("va", "adv"):
            gen = None if noise is not None else torch.Generator(device).manual_seed(rec["seed"])
            return O.sample_loop(W, d, r_s, wa, we, rec["T"], nfe=rec["nfe"], method=rec.get("method", "euler"),
                                 a_cfg_scale=rec["a"], r_cfg_scale=rec["r"], e_cfg_scale=rec["e"],
                                 include_r_cfg=rec.get("include_r_cfg", False), generator=gen, noise=noise, **kw)
        if entry == "legacy":
            gen = None if noise is not None else torch.Generator(device).manual_seed(rec["seed"])
            return O.float_sample_legacy(W, d, r_s, wa, we, opt_nfe=rec["nfe"], a_cfg_scale=rec["a"],
                                         r_cfg_scale=rec["r"], e_cfg_scale=rec["e"], generator=gen, noise=noise, **kw)
        x, prev_x, prev_wa, prev_we = [t.to(device) for t in cases.cfv_extra_inputs(rec)]
        L = d.frames_per_clip
        return O.forward_with_cfv(W, d, torch.tensor([rec["t"]], device=device), x, wa[:, :L], r_s,
                                  we[:, :L] if we.shape[1] > 1 else we, prev_x, prev_wa, prev_we,
                                  rec["a"], rec["r"], rec["e"], rec.get("include_r_cfg", False), **kw)


@pytest.mark.parametrize("name", ALL)
def test_oracle_matches_reference_fixture(name):
    ref = cases.golden(name)
    out = run_oracle(name)
    assert out.shape == ref.shape and out.dtype == torch.float32
    # fp32 vs fp32 on the same CPU; only op order differs (explicit softmax vs SDPA)
    assert cases.rel_err(out, ref) <= 2e-5, (name, cases.rel_err(out, ref))
    assert cases.max_abs(out, ref) <= 2e-4, (name, cases.max_abs(out, ref))


def test_manifest_covers_all_cases():
    m = cases.manifest()
    for name, rec in cases.CASES.items():
        assert name in m, f"fixture {name} missing: run tests/golden/make_golden.py"
        for k, v in rec.items():
            assert m[name][k] == v, (name, k)


def test_structural_known_answers():
    """doc/NETWORKS.md:16-22,72-80: 156.698 M parameters incl. pos_embed, 18.890 M per block; 94 keys."""
    W = cases.weights("full")
    assert len(W) == 93   # + the bool buffer alignment_mask = the reference's 94 state-dict keys
    total = sum(v.numel() for v in W.values())
    assert total == 156_698_112
    blk = sum(v.numel() for k, v in W.items() if k.startswith("blocks.0."))
    assert blk == 18_889_728 and round(blk / 1e6, 3) == 18.890


def test_nfe1_returns_noise():
    rec = cases.CASES["va_nfe1"]
    d = cases.dims_of(rec)
    g = torch.Generator().manual_seed(rec["seed"])
    n0 = torch.randn(1, d.frames_per_clip, d.dim_w, generator=g)
    n1 = torch.randn(1, d.frames_per_clip, d.dim_w, generator=g)
    ref = cases.golden("va_nfe1")
    assert torch.equal(ref, torch.cat([n0, n1], 1)[:, :rec["T"]])
