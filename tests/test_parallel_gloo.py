"""world_size-2 test (gloo, CPU) of the data-parallel host logic in parallel.py: sharding, per-clip noise, ragged gather.
The compute function injected here is the oracle (the checker standing in for the CUDA backend, which needs a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from __graft_entry__ import load_package
from oracle import fmt_oracle as O
from oracle.synth import synth_inputs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_fn(W, d):
    def fn(r_s, wa, we, T, noise):
        with torch.no_grad():
            return O.sample_loop(W, d, r_s, wa, we, T, nfe=3, a_cfg_scale=2.0, e_cfg_scale=1.0, noise=noise)
    return fn


def _worker(rank, world, port, B, T, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        pkg = load_package()
        import sys
        par = __import__(pkg.__name__ + ".parallel", fromlist=["x"])
        d = cases.SMALL_DIMS
        W = cases.weights("small")
        r_s, wa, we = synth_inputs(d, B, T, seed=5)
        seeds = [15 + i for i in range(B)]
        full = par.sample_clips_data_parallel(_oracle_fn(W, d), r_s, wa, we, T, seeds, d.frames_per_clip, d.dim_w, "cpu")
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    pkg = load_package()
    par = __import__(pkg.__name__ + ".parallel", fromlist=["x"])
    for n in (0, 1, 3, 8, 256):
        for world in (1, 2, 3, 8):
            b = [par.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    with pytest.raises(ValueError):
        par.shard_bounds(4, 2, 2)


@pytest.mark.parametrize("B", [3, 4, 1])
def test_two_ranks_equal_one_rank(tmp_path, B):
    """B clips over 2 gloo ranks (ragged when B is odd, one empty shard when B == 1) == the same clips in one process."""
    T = 30
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, T, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2)]
    assert torch.equal(got[0], got[1]) and got[0].shape == (B, T, cases.SMALL_DIMS.dim_w)
    pkg = load_package()
    par = __import__(pkg.__name__ + ".parallel", fromlist=["x"])
    d = cases.SMALL_DIMS
    r_s, wa, we = synth_inputs(d, B, T, seed=5)
    n_win = -(-T // d.frames_per_clip)
    noise = par.per_clip_noise([15 + i for i in range(B)], n_win, d.frames_per_clip, d.dim_w, "cpu")
    ref = _oracle_fn(cases.weights("small"), d)(r_s, wa, we, T, noise)
    assert cases.max_abs(got[0], ref) <= 1e-5       # batch-of-B vs per-shard batches: fp32 reassociation only
