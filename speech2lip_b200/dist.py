"""Multi-GPU plumbing: frames are independent, so the path shards with no data-path collective
(SURVEY §8(e)).  One process per GPU; the only communication is ONE broadcast of the hot-path
parameters from rank 0 at start-up (replaces DistributedDataParallel's constructor broadcast,
src/face_simple/training.py:40), packed into a single flat buffer so it is one NCCL call."""
import torch
import torch.distributed as dist

from ._cabi import PARAM_NAMES


def shard_frames(n_frames, rank, world_size):
    """Contiguous block partition: rank r renders frames [lo, hi) — keeps the output video order trivial."""
    per = (n_frames + world_size - 1) // world_size
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def broadcast_params(params, src=0, group=None):
    """In-place broadcast of the hot-path tensors (reference state_dict names) from `src` as ONE flat
    fp32 buffer (<= 2.8 MB).  Works on NCCL (cuda tensors) and gloo (cpu tensors)."""
    names = [n for n in PARAM_NAMES]
    tensors = [params[n] for n in names]
    flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    return params


def gather_frames(local_rgb, n_frames, group=None):
    """Optional: collect every rank's [F_r,H,W,3] block on all ranks in frame order (all_gather of padded blocks)."""
    world = dist.get_world_size(group)
    per = (n_frames + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype, device=local_rgb.device)
    pad[:local_rgb.shape[0]] = local_rgb
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat(outs, 0)[:n_frames]
