"""Parity of the CUDA path (through the node classes / C ABI) with the reference's golden fixtures and the oracle.

Gates (BASELINE.json north_star):
  fp32 validation mode : ||r_d - ref|| / ||ref|| <= 1e-4
  bf16 mode            : max |r_d - ref| <= 2e-2
Noise is injected (the fixtures were drawn from a CPU generator; the same draws are replayed here).
"""
import types

import pytest
import torch

import cases
from __graft_entry__ import load_package
from oracle import fmt_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return load_package()


_models = {}


def model_for(pkg, dims_name):
    if dims_name not in _models:
        rec_dims = cases.FmtDims() if dims_name == "full" else cases.SMALL_DIMS
        opt = pkg.BaseOptions()
        for k, v in rec_dims.as_dict().items():
            if hasattr(opt, k):
                setattr(opt, k, v)
        _models[dims_name] = pkg.FmtModel(cases.weights(dims_name), opt, target_device=DEV)
    return _models[dims_name]


def cpu_noise(rec):
    d = cases.dims_of(rec)
    g = torch.Generator().manual_seed(rec["seed"])
    n_win = -(-rec["T"] // d.frames_per_clip)
    return torch.stack([torch.randn(rec["B"], d.frames_per_clip, d.dim_w, generator=g) for _ in range(n_win)])


def run_node(pkg, name, mode):
    rec = cases.CASES[name]
    model = model_for(pkg, rec["dims"])
    r_s, wa, we = cases.case_inputs(rec)
    noise = cpu_noise(rec)
    if rec["entry"] == "va":
        out, passthrough = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(
            r_s, wa, we, rec["T"], model, rec["a"], rec["r"], rec["e"], rec.get("include_r_cfg", False), rec["nfe"],
            rec.get("method", "euler"), 1e-5, 1e-5, 0.1, 0.1, 0.1, True, rec["seed"], _mode=mode, _noise=noise)
        assert passthrough is model
    elif rec["entry"] == "adv":
        d = cases.dims_of(rec)
        opt = pkg.BaseOptions()
        opt.rank, opt.nfe = torch.device(DEV), rec["nfe"]
        pipe = types.SimpleNamespace(opt=opt, G=types.SimpleNamespace(fmt=model, num_prev_frames=d.num_prev_frames,
                                                                      num_frames_for_clip=d.frames_per_clip))
        out, _ = pkg.FloatSampleMotionSequenceRD().sample_rd_sequence(r_s, wa, rec["T"], we, pipe, rec["a"], rec["e"], rec["seed"],
                                                                      _mode=mode, _noise=noise)
    else:  # legacy FLOAT.sample
        opt = pkg.BaseOptions()
        opt.rank, opt.nfe = torch.device(DEV), rec["nfe"]
        assert we.dtype == torch.int64
        out = pkg.float_sample(model, opt, r_s.to(DEV), wa.to(DEV), we.to(DEV), rec["a"], rec["r"], rec["e"], seed=rec["seed"],
                               mode=mode, noise=noise).cpu()
    assert out.device.type == "cpu" and out.dtype == torch.float32
    return out


CLIP_CASES = [n for n, r in cases.CASES.items() if r["entry"] in ("va", "adv", "legacy")]
PROJ_CASES = [n for n, r in cases.CASES.items() if r["entry"] == "proj"]
CFV_CASES = [n for n, r in cases.CASES.items() if r["entry"] == "cfv"]


@pytest.mark.parametrize("name", CLIP_CASES)
def test_fp32_validation_mode_matches_reference(pkg, name):
    ref = cases.golden(name)
    out = run_node(pkg, name, "fp32")
    assert out.shape == ref.shape
    rel = cases.rel_err(out, ref)
    assert rel <= 1e-4, (name, rel)
    big = ref.abs() > 1e-2
    per_elem = ((out - ref).abs()[big] / ref.abs()[big]).max().item()
    assert per_elem <= 2e-2, (name, per_elem)    # per-element relative error where |ref| > 1e-2


@pytest.mark.parametrize("name", CLIP_CASES)
def test_bf16_mode_matches_reference(pkg, name):
    ref = cases.golden(name)
    out = run_node(pkg, name, "bf16")
    assert out.shape == ref.shape
    err = cases.max_abs(out, ref)
    assert err <= 2e-2, (name, err)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", None)])
@pytest.mark.parametrize("name", CFV_CASES)
def test_single_evaluation_matches_forward_with_cfv(pkg, name, mode, tol):
    rec = cases.CASES[name]
    d = cases.dims_of(rec)
    model = model_for(pkg, rec["dims"])
    be = pkg.backend_for(model, DEV)
    r_s, wa, we = [t.to(DEV) for t in cases.case_inputs(rec)]
    x, prev_x, prev_wa, prev_we = [t.to(DEV).contiguous() for t in cases.cfv_extra_inputs(rec)]
    L = d.frames_per_clip
    dynamic = we.shape[1] > 1
    nb = pkg.n_branches_for(rec["a"], rec["r"], rec["e"], rec.get("include_r_cfg", False))
    # a 2-point grid whose only evaluation time is t: linspace(0,1,2) evaluates at t=0; emulate arbitrary t via nfe grid search
    # -> use the plan with nfe chosen so that some grid point equals t (t = k/(nfe-1))
    t = rec["t"]
    nfe = 11
    k = round(t * (nfe - 1))
    assert abs(k / (nfe - 1) - t) < 1e-6
    be.configure(rec["B"], nb, dynamic, nfe, "euler", mode)
    v = be.velocity(k, x, wa[:, :L].contiguous(), r_s, (we[:, :L] if dynamic else we).contiguous(), prev_x, prev_wa,
                    prev_we if dynamic else None, rec["a"], rec["r"], rec["e"]).cpu()
    ref = cases.golden(name)
    assert v.shape == ref.shape
    if tol is not None:
        assert cases.rel_err(v, ref) <= tol, cases.rel_err(v, ref)
    else:
        assert cases.max_abs(v, ref) <= 1e-2, cases.max_abs(v, ref)


def test_same_seed_same_device_as_oracle(pkg):
    """The node with a CUDA generator vs the oracle on the same device with the same seed."""
    rec = cases.CASES["va_config1"]
    d = cases.dims_of(rec)
    model = model_for(pkg, "full")
    r_s, wa, we = cases.case_inputs(rec)
    out, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(
        r_s, wa, we, rec["T"], model, rec["a"], rec["r"], rec["e"], False, rec["nfe"], "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1,
        True, rec["seed"], _mode="fp32")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    Wd = {k: v.to(DEV) for k, v in cases.weights("full").items()}
    g = torch.Generator(DEV).manual_seed(rec["seed"])
    with torch.no_grad():
        ref = O.sample_loop(Wd, d, r_s.to(DEV), wa.to(DEV), we.to(DEV), rec["T"], nfe=rec["nfe"], a_cfg_scale=rec["a"],
                            r_cfg_scale=rec["r"], e_cfg_scale=rec["e"], generator=g).cpu()
    assert cases.rel_err(out, ref) <= 1e-4, cases.rel_err(out, ref)


def test_properties_full_size(pkg):
    """Size-independent properties at the full architecture: determinism, batch independence, nfe=1 identity."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    from oracle.synth import synth_inputs
    B, T = 4, 120
    r_s, wa, we = synth_inputs(d, B, T, seed=99)
    g = torch.Generator().manual_seed(3)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(3)])
    node = pkg.FloatSampleMotionSequenceRD_VA()
    args = (2.0, 1.0, 1.0, False, 4, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 3)
    a, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _noise=noise)
    b, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _noise=noise)
    assert torch.equal(a, b)                                    # B=4 runs the tiled GEMM path: bitwise deterministic
    for i in (0, 3):                                            # clips never mix: batch of 4 == 4 batches of 1
        s, _ = node.sample_rd_sequence_va(r_s[i:i + 1], wa[i:i + 1], we[i:i + 1], T, model, *args, _noise=noise[:, i:i + 1].contiguous())
        # B=1 runs the dataflow window kernel (split-K partial sums, attention on fp32 q/k/v): a different schedule of the same
        # arithmetic -> equal to the batched run within bf16 noise, not bitwise
        assert cases.max_abs(s, a[i:i + 1]) <= 5e-3, cases.max_abs(s, a[i:i + 1])
        s2, _ = node.sample_rd_sequence_va(r_s[i:i + 1], wa[i:i + 1], we[i:i + 1], T, model, *args, _noise=noise[:, i:i + 1].contiguous())
        # run-to-run: the K slices are rounded to multiples of 2^-14 before they meet in L2, their fp32 sums are exact and so
        # order-independent: the same seed gives the same bits, as the reference does under fix_noise_seed (nodes_vadv.py:673-689)
        assert torch.equal(s, s2)
    args1 = (2.0, 1.0, 1.0, False, 1, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 3)
    n1, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args1, _noise=noise)
    assert torch.equal(n1, torch.cat(list(noise), dim=1)[:, :T])  # nfe=1: the solver returns the noise unchanged
    assert torch.isfinite(a).all()


def test_errors_mirror_reference(pkg):
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    from oracle.synth import synth_inputs
    r_s, wa, we = synth_inputs(d, 2, 50, seed=1)
    node = pkg.FloatSampleMotionSequenceRD_VA()
    args = (2.0, 1.0, 1.0, False, 2, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 3)
    with pytest.raises(TypeError):
        node.sample_rd_sequence_va(r_s.numpy(), wa, we, 50, model, *args)
    with pytest.raises(ValueError):
        node.sample_rd_sequence_va(r_s[:1], wa, we, 50, model, *args)
    before = (model.opt.audio_dropout_prob, model.opt.ref_dropout_prob, model.opt.emotion_dropout_prob)
    with pytest.raises(ValueError):
        node.sample_rd_sequence_va(r_s, wa, we, 50, model, 2.0, 1.0, 1.0, False, 2, "dopri5", 1e-5, 1e-5, 0.3, 0.3, 0.3, True, 3)
    assert before == (model.opt.audio_dropout_prob, model.opt.ref_dropout_prob, model.opt.emotion_dropout_prob)
    cpu_model = pkg.FmtModel(cases.weights("full"), target_device="cpu")
    with pytest.raises(pkg.FmtError):
        node.sample_rd_sequence_va(r_s, wa, we, 50, cpu_model, *args)


@pytest.mark.parametrize("method,nfe,nb_scales", [("euler", 4, (2.0, 1.0, 1.0, False)), ("heun3", 3, (2.0, 1.0, 1.0, False)),
                                                   ("euler", 3, (1.0, 1.0, 1.0, False)), ("midpoint", 3, (2.0, 0.5, 1.5, True))])
def test_window_kernel_matches_per_op_path(pkg, monkeypatch, method, nfe, nb_scales):
    """The persistent window kernel (<= 256 token rows) and the one-kernel-per-op path are two schedules of the same
    arithmetic (bf16 operands, fp32 accumulation): same result up to bf16 rounding of intermediates."""
    d = cases.FmtDims()
    from oracle.synth import synth_inputs
    a, r, e, inc = nb_scales
    B, T = 1, 120
    r_s, wa, we = [t.to(DEV) for t in synth_inputs(d, B, T, seed=21)]
    g = torch.Generator().manual_seed(5)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(3)]).to(DEV)
    outs = {}
    # "3": dataflow window kernel (flow.cuh, the default), "3c": the same with 32-row chunks (attention crosses chunk boundaries);
    # "2": grouped window kernel (one slice of the SMs per sequence, GELU fused into fc1's epilogue); "2s": same with the
    # separate GELU stage and a 2-way K split of fc1; "1": barrier-stepped split-K window kernel; "0": one kernel per op
    for flag in ("3", "3c", "2", "2s", "1", "0"):
        monkeypatch.setenv("FMT_FLOW_CH", "32" if flag == "3c" else "64")
        monkeypatch.setenv("FMT_WINDOW", flag[0])
        monkeypatch.setenv("FMT_WIN_FUSE_GELU", "0" if flag == "2s" else "1")
        monkeypatch.setenv("FMT_WIN_PK", "0,0,2,0" if flag == "2s" else "0,0,0,0")
        be = pkg.FmtBackend(cases.weights("full"), pkg.Dims(), DEV)
        be.configure(B, pkg.n_branches_for(a, r, e, inc), False, nfe, method, "bf16")
        assert be.window_kernel_status() == (0 if flag != "0" else -1)
        outs[flag] = be.sample_clip(r_s, wa, we, T, noise, a, r, e).cpu()
        torch.cuda.synchronize()
        assert be.window_kernel_status() == (0 if flag != "0" else -1)
        be.close()
    for flag in ("3", "3c", "2", "2s", "1"):
        assert torch.isfinite(outs[flag]).all(), flag
        assert cases.max_abs(outs[flag], outs["0"]) <= 1e-2, (flag, cases.max_abs(outs[flag], outs["0"]))


def test_large_batch_matches_oracle(pkg):
    """32 clips per GPU (the tensor-pipe regime of BASELINE.json configs[3]): CTA-pair tcgen05 GEMMs + shared-memory band
    attention, checked against the oracle run on the same device with the same injected noise."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    from oracle.synth import synth_inputs
    B, T, nfe = 32, 70, 4
    r_s, wa, we = synth_inputs(d, B, T, seed=123)
    g = torch.Generator().manual_seed(9)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(2)])
    out, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(r_s, wa, we, T, model, 2.0, 1.0, 1.0, False, nfe, "euler", 1e-5, 1e-5,
                                                                       0.1, 0.1, 0.1, True, 9, _noise=noise)
    torch.backends.cuda.matmul.allow_tf32 = False
    Wd = {k: v.to(DEV) for k, v in cases.weights("full").items()}
    with torch.no_grad():
        ref = O.sample_loop(Wd, d, r_s.to(DEV), wa.to(DEV), we.to(DEV), T, nfe=nfe, a_cfg_scale=2.0, r_cfg_scale=1.0, e_cfg_scale=1.0,
                            noise=noise.to(DEV)).cpu()
    assert out.shape == ref.shape == (B, T, d.dim_w)
    assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)
    # clips never mix: clip 5 of the batch == the same clip sampled alone (different kernels -> bf16-level agreement)
    solo, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(r_s[5:6], wa[5:6], we[5:6], T, model, 2.0, 1.0, 1.0, False, nfe, "euler",
                                                                        1e-5, 1e-5, 0.1, 0.1, 0.1, True, 9, _noise=noise[:, 5:6].contiguous())
    assert cases.max_abs(solo, out[5:6]) <= 1e-2, cases.max_abs(solo, out[5:6])


def _oracle_on_gpu(d, r_s, wa, we, T, noise, **kw):
    torch.backends.cuda.matmul.allow_tf32 = False
    Wd = {k: v.to(DEV) for k, v in cases.weights("full").items()}
    with torch.no_grad():
        return O.sample_loop(Wd, d, r_s.to(DEV), wa.to(DEV), we.to(DEV), T, noise=noise.to(DEV), **kw).cpu()


@pytest.mark.parametrize("window_flag", ["3", "1", "2", "0"])
def test_non_default_window_geometry(pkg, monkeypatch, window_flag):
    """loader widgets other than the defaults (nodes_vadv_loader.py:684-704) - attention_window 4 (the
    general band-attention code, not the 5-key fast path), 6 context frames, 1.6 s windows (L = 40) - on the full-width
    architecture, through both schedules of the window kernel and the per-op path, against the oracle on the same device."""
    from oracle.synth import FmtDims, synth_inputs, synth_state_dict
    d = FmtDims(attention_window=4, num_prev_frames=6, wav2vec_sec=1.6, fmt_depth=3)
    W = synth_state_dict(d, seed=3)
    opt = pkg.BaseOptions()
    for k, v in d.as_dict().items():
        if hasattr(opt, k):
            setattr(opt, k, v)
    monkeypatch.setenv("FMT_WINDOW", window_flag)
    model = pkg.FmtModel(W, opt, target_device=DEV)
    B, T, nfe = 1, 95, 4                                            # 3 windows, ragged last one
    r_s, wa, we = synth_inputs(d, B, T, seed=77, dynamic_we=True)
    g = torch.Generator().manual_seed(11)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(3)])
    torch.backends.cuda.matmul.allow_tf32 = False
    Wd = {k: v.to(DEV) for k, v in W.items()}
    with torch.no_grad():
        ref = O.sample_loop(Wd, d, r_s.to(DEV), wa.to(DEV), we.to(DEV), T, nfe=nfe, a_cfg_scale=2.0, r_cfg_scale=1.0, e_cfg_scale=1.5,
                            noise=noise.to(DEV)).cpu()
    node = pkg.FloatSampleMotionSequenceRD_VA()
    args = (2.0, 1.0, 1.5, False, nfe, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 11)
    out, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _noise=noise)
    assert out.shape == ref.shape == (B, T, d.dim_w)
    assert pkg.backend_for(model, DEV).window_kernel_status() == (-1 if window_flag == "0" else 0)
    assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)
    if window_flag == "0":
        out32, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _mode="fp32", _noise=noise)
        assert cases.rel_err(out32, ref) <= 1e-4, cases.rel_err(out32, ref)


def test_long_clip_chained_windows(pkg):
    """BASELINE.json configs[2] in miniature: one clip, 12 sequential windows chained through prev_x / prev_wa (bf16 error
    accumulates across windows, SURVEY.md §7 'hard parts'), free-running against the oracle on the same noise."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    from oracle.synth import synth_inputs
    T, nfe = 590, 10                                              # 12 windows, ragged last one
    r_s, wa, we = synth_inputs(d, 1, T, seed=31)
    g = torch.Generator().manual_seed(4)
    noise = torch.stack([torch.randn(1, d.frames_per_clip, d.dim_w, generator=g) for _ in range(12)])
    node = pkg.FloatSampleMotionSequenceRD_VA()
    args = (2.0, 1.0, 1.0, False, nfe, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 4)
    out, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _noise=noise)
    ref = _oracle_on_gpu(d, r_s, wa, we, T, noise, nfe=nfe, a_cfg_scale=2.0, r_cfg_scale=1.0, e_cfg_scale=1.0)
    assert out.shape == ref.shape == (1, T, d.dim_w)
    assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)
    per_window = [(cases.max_abs(out[:, s:s + 50], ref[:, s:s + 50])) for s in range(0, T, 50)]
    assert max(per_window[6:]) <= 4 * max(max(per_window[:6]), 1e-3), per_window      # no blow-up along the chain
    out32, _ = node.sample_rd_sequence_va(r_s, wa, we, T, model, *args, _mode="fp32", _noise=noise)
    assert cases.rel_err(out32, ref) <= 1e-4, cases.rel_err(out32, ref)


@pytest.mark.parametrize("nfe,a,e", [(5, 1.0, 3.0), (20, 3.0, 1.0), (10, 4.0, 3.0), (10, 1.0, 1.0)])
def test_dynamic_emotion_sweeps(pkg, nfe, a, e):
    """BASELINE.json configs[4]: per-window (dynamic) emotion vectors, nfe sweep 5/10/20, a_cfg sweep 1-4 (a = e = 1 takes the
    single-branch path, FMT.py:400-401); 2 clips, 2.5 windows."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    from oracle.synth import synth_inputs
    B, T = 2, 125
    r_s, wa, we = synth_inputs(d, B, T, seed=77, dynamic_we=True)
    assert we.shape == (B, T, d.dim_e)
    g = torch.Generator().manual_seed(8)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(3)])
    out, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(r_s, wa, we, T, model, a, 1.0, e, False, nfe, "euler", 1e-5, 1e-5,
                                                                       0.1, 0.1, 0.1, True, 8, _noise=noise)
    ref = _oracle_on_gpu(d, r_s, wa, we, T, noise, nfe=nfe, a_cfg_scale=a, r_cfg_scale=1.0, e_cfg_scale=e)
    assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)


# ---------------------------------------------------------------------------------------------- SURVEY.md §8f rank 2
def _projection_layer(pkg, rec, device=DEV):
    layer = pkg.AudioProjectionLayer(rec["in_dim"], 512, target_device=device)
    layer.load_state_dict(cases.projection_weights(rec))
    return layer


@pytest.mark.parametrize("name", PROJ_CASES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_audio_projection_matches_reference(pkg, name, mode):
    """FloatApplyAudioProjection (nodes_vadv.py:147-198) against the fixture the reference node produced: CPU tensor in,
    CPU tensor out; fp32 validation mode 1e-4 relative, bf16 mode 2e-2 max-abs."""
    rec = cases.CASES[name]
    ref = cases.golden(name)
    (out,) = pkg.FloatApplyAudioProjection().apply_projection(cases.projection_input(rec), _projection_layer(pkg, rec), _mode=mode)
    assert out.device.type == "cpu" and out.dtype == torch.float32 and out.shape == ref.shape
    if mode == "fp32":
        assert cases.rel_err(out, ref) <= 1e-4, cases.rel_err(out, ref)
    else:
        assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)


def test_audio_projection_full_size(pkg):
    """configs[3] shape in front of the sampler: 64 clips x 200 frames of stacked wav2vec features (12 800 rows, K = 9216) against
    the oracle on the same device; rows are independent (a sub-batch gives the same rows), the backend is reused across calls."""
    from oracle import fmt_oracle as O
    from oracle.synth import synth_wav2vec_features
    rec = dict(in_dim=9216, seed=62)
    layer = _projection_layer(pkg, rec)
    be = pkg.projection_backend_for(layer, DEV)
    x = synth_wav2vec_features(64, 200, 9216, seed=5).to(DEV)
    P = {k: v.to(DEV) for k, v in cases.projection_weights(rec).items()}
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = O.audio_projection(P, x)
    out = be.apply(x, "bf16")
    assert out.shape == (64, 200, 512) and torch.isfinite(out).all()
    assert cases.max_abs(out.cpu(), ref.cpu()) <= 2e-2
    sub = be.apply(x[5:7], "bf16")
    assert torch.equal(sub, out[5:7])
    assert pkg.projection_backend_for(layer, DEV) is be and be.launch_count() >= 6
    out32 = be.apply(x[:4], "fp32")
    assert cases.rel_err(out32.cpu(), ref[:4].cpu()) <= 1e-4


def test_device_resident_handoff(pkg):
    """SURVEY.md 8f rank 2, second half: FloatApplyAudioProjection -> sampler node without the device -> host -> device round
    trip the reference makes between nodes (nodes_vadv.py:197,692-694,719).  keep_on_device hands CUDA tensors over; the result
    is bit-identical to the CPU hand-off (the default, which stays the reference's behaviour)."""
    from oracle.synth import synth_wav2vec_features
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    rec = dict(in_dim=9216, seed=62)
    layer = _projection_layer(pkg, rec)
    B, T = 2, 100
    feats = synth_wav2vec_features(B, T, 9216, seed=9)
    r_s, _, we = pkg.synth.synth_inputs(d, B, T, seed=7)
    g = torch.Generator().manual_seed(15)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g) for _ in range(2)])
    proj, samp = pkg.FloatApplyAudioProjection(), pkg.FloatSampleMotionSequenceRD_VA()
    args = (2.0, 1.0, 1.0, False, 4, "euler", 1e-5, 1e-5, 0.1, 0.1, 0.1, True, 15)
    (wa_cpu,) = proj.apply_projection(feats, layer)
    assert wa_cpu.device.type == "cpu"
    out_cpu, _ = samp.sample_rd_sequence_va(r_s, wa_cpu, we, T, model, *args, _noise=noise)
    assert out_cpu.device.type == "cpu"
    (wa_dev,) = proj.apply_projection(feats, layer, keep_on_device=True)
    assert wa_dev.is_cuda and torch.equal(wa_dev.cpu(), wa_cpu)
    out_dev, _ = samp.sample_rd_sequence_va(r_s.to(DEV), wa_dev, we.to(DEV), T, model, *args, keep_on_device=True, _noise=noise.to(DEV))
    assert out_dev.is_cuda and torch.equal(out_dev.cpu(), out_cpu)


def test_configs2_sixty_second_clip(pkg):
    """BASELINE.json configs[2] at full size: one 60 s clip = 1500 frames = 30 sequential windows chained through prev_x / prev_wa,
    free-running against the oracle on the same device and noise (270 dependent ODE steps)."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    T, nfe = 1500, 10
    r_s, wa, we = pkg.synth.synth_inputs(d, 1, T, seed=41)
    g = torch.Generator().manual_seed(6)
    noise = torch.stack([torch.randn(1, d.frames_per_clip, d.dim_w, generator=g) for _ in range(30)])
    out, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(r_s, wa, we, T, model, 2.0, 1.0, 1.0, False, nfe, "euler", 1e-5, 1e-5,
                                                                       0.1, 0.1, 0.1, True, 6, _noise=noise)
    ref = _oracle_on_gpu(d, r_s, wa, we, T, noise, nfe=nfe, a_cfg_scale=2.0, r_cfg_scale=1.0, e_cfg_scale=1.0)
    assert out.shape == ref.shape == (1, T, d.dim_w)
    per_window = [cases.max_abs(out[:, s:s + 50], ref[:, s:s + 50]) for s in range(0, T, 50)]
    assert max(per_window) <= 2e-2, per_window
    assert max(per_window[15:]) <= 4 * max(max(per_window[:15]), 1e-3), per_window      # no blow-up along the chain


def test_configs3_full_batch(pkg):
    """BASELINE.json configs[3] at its single-GPU batch: 256 independent clips in one call (46 080 token rows per evaluation, CTA-pair
    GEMMs, chunked AdaLN tables), one window, nfe 4 to keep the oracle short; against the oracle on the same device and noise."""
    d = cases.FmtDims()
    model = model_for(pkg, "full")
    B, T, nfe = 256, 50, 4
    r_s, wa, we = pkg.synth.synth_inputs(d, B, T, seed=321)
    g = torch.Generator().manual_seed(12)
    noise = torch.stack([torch.randn(B, d.frames_per_clip, d.dim_w, generator=g)])
    out, _ = pkg.FloatSampleMotionSequenceRD_VA().sample_rd_sequence_va(r_s, wa, we, T, model, 2.0, 1.0, 1.0, False, nfe, "euler", 1e-5, 1e-5,
                                                                       0.1, 0.1, 0.1, True, 12, _noise=noise)
    torch.backends.cuda.matmul.allow_tf32 = False
    Wd = {k: v.to(DEV) for k, v in cases.weights("full").items()}
    ref = torch.empty(B, T, d.dim_w)
    with torch.no_grad():
        for b0 in range(0, B, 64):                               # the eager oracle in slices of 64 clips (clips never mix)
            sl = slice(b0, b0 + 64)
            ref[sl] = O.sample_loop(Wd, d, r_s[sl].to(DEV), wa[sl].to(DEV), we[sl].to(DEV), T, nfe=nfe, a_cfg_scale=2.0, r_cfg_scale=1.0,
                                    e_cfg_scale=1.0, noise=noise[:, sl].to(DEV)).cpu()
    assert out.shape == (B, T, d.dim_w) and torch.isfinite(out).all()
    assert cases.max_abs(out, ref) <= 2e-2, cases.max_abs(out, ref)
