"""Data-parallel sampling over independent clips: one process per GPU, no collective inside a step.

Windows of one clip are sequential (``prev_x`` chaining, nodes_adv.py:663-664), clips never mix (``forward_with_cfv`` only
concatenates along the batch, FMT.py:360-372), so the batch dimension is the only thing that shards.  Rank ``r`` of ``R``
samples clips ``shard_bounds(B, R, r)`` with the full (replicated) weights and its own per-clip noise; the only
collective is one final gather of the motion latents ``r_d`` (NCCL over NVLink on the GPUs; ``gloo`` in the CPU tests
of this host logic).  Results do not depend on ``R`` because the per-window noise is drawn per clip.
"""
from typing import Callable, List, Optional, Tuple

import torch


def shard_bounds(n_clips: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shards: the first ``n_clips % world`` ranks hold one clip more.  Empty shards are legal."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(int(n_clips), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def per_clip_noise(seeds: List[int], n_windows: int, frames_per_clip: int, dim_w: int, device) -> torch.Tensor:
    """(n_windows, len(seeds), L, dim_w): clip ``i`` draws its windows from its own ``torch.Generator(device)`` seeded
    ``seeds[i]`` (one ``randn(1, L, dim_w)`` per window, in window order - the B=1 draw order of nodes_adv.py:606), so a
    clip's noise is the same whichever rank samples it."""
    out = torch.empty(n_windows, len(seeds), frames_per_clip, dim_w, device=device, dtype=torch.float32)
    for i, s in enumerate(seeds):
        g = torch.Generator(device).manual_seed(int(s))
        for w in range(n_windows):
            out[w, i] = torch.randn(1, frames_per_clip, dim_w, device=device, generator=g)[0]
    return out


def gather_clips(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """All ranks end up with the (n_clips, T, dim_w) latents in clip order.  Ragged shards are padded to the largest one
    for ``all_gather_into_tensor`` and trimmed afterwards."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_clips, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} clips, expected {hi - lo}")
    padded = local
    if hi - lo < biggest:
        padded = torch.zeros(biggest, *local.shape[1:], device=local.device, dtype=local.dtype)
        padded[: hi - lo] = local
    buf = torch.empty(world * biggest, *local.shape[1:], device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, padded.contiguous(), group=group)
    parts = [buf[r * biggest: r * biggest + (h - l)] for r, (l, h) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def sample_clips_data_parallel(sample_fn: Callable[..., torch.Tensor], r_s: torch.Tensor, wa: torch.Tensor, we: torch.Tensor,
                               audio_num_frames: int, seeds: List[int], frames_per_clip: int, dim_w: int, device,
                               group=None, gather: bool = True) -> Optional[torch.Tensor]:
    """Shards ``B`` clips over the ranks of ``group`` and samples the local shard with
    ``sample_fn(r_s, wa, we, audio_num_frames, noise) -> (b_local, T, dim_w)`` (on the GPUs: ``FmtBackend.sample_clip``
    behind :func:`float_fmt_b200.perform_ode_sampling_loop`), then gathers.  ``r_s/wa/we`` are the FULL batch on every
    rank (they are tiny next to the weights); only the shard is moved to ``device``."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = wa.shape[0]
    if len(seeds) != B or r_s.shape[0] != B or we.shape[0] != B:
        raise ValueError("Batch size mismatch among r_s, wa, we latents and seeds.")
    lo, hi = shard_bounds(B, world, rank)
    n_win = -(-int(audio_num_frames) // frames_per_clip)
    if hi > lo:
        noise = per_clip_noise(seeds[lo:hi], n_win, frames_per_clip, dim_w, device)
        local = sample_fn(r_s[lo:hi].to(device), wa[lo:hi].to(device), we[lo:hi].to(device), int(audio_num_frames), noise)
    else:
        local = torch.empty(0, int(audio_num_frames), dim_w, device=device, dtype=torch.float32)
    if not gather or world == 1:
        return local
    return gather_clips(local, B, group)
