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
    """In-place broadcast of the 42 HOT-PATH tensors (reference state_dict names) from `src` as ONE flat fp32 buffer
    (<= 2.8 MB): all an inference renderer needs.  It is NOT a replacement for DDP's constructor broadcast of a training
    module — use broadcast_module for that.  Works on NCCL (cuda tensors) and gloo (cpu tensors)."""
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


def broadcast_module(module, src=0, group=None):
    """What DistributedDataParallel's constructor does for the WHOLE module (training.py:40): every parameter AND buffer
    (the post-fusion UNet, its BatchNorm running statistics, canonical_depth_head, coord_linears included) from `src`,
    one flat broadcast per dtype.  broadcast_params above covers only the 42 hot-path tensors a renderer needs."""
    by_dtype = {}
    for t in list(module.parameters()) + list(module.buffers()):
        by_dtype.setdefault(t.dtype, []).append(t)
    with torch.no_grad():
        for dtype, ts in by_dtype.items():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=src, group=group)
            off = 0
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
    return module


def gather_frames(local_rgb, n_frames, group=None):
    """Optional: collect every rank's [F_r,H,W,3] block on all ranks in frame order (all_gather of padded blocks)."""
    world = dist.get_world_size(group)
    per = (n_frames + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype, device=local_rgb.device)
    pad[:local_rgb.shape[0]] = local_rgb
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat(outs, 0)[:n_frames]
