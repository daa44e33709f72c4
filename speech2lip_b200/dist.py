"""Multi-GPU plumbing: frames are independent, so the path shards with no data-path collective
(SURVEY §8(e)).  One process per GPU; the only communication is ONE broadcast of the hot-path
parameters from rank 0 at start-up (replaces DistributedDataParallel's constructor broadcast,
src/face_simple/training.py:40), packed into a single flat buffer so it is one NCCL call.
Training is data-parallel over frames; its one exchange step — the gradient average — is GradExchange below."""
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


# ------------------------------------------------------------------------------------------------------------------------
# Data-parallel TRAINING (BASELINE configs[4]): DP over frames as the reference does it (DistributedSampler,
# train.py:102; DDP, training.py:40) — every rank renders its own frames, the gradients of one flat bucket are averaged.


class _DeviceSpan:
    """A [n] float32 window of device memory the library allocated, wrapped zero-copy by torch.as_tensor."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class GradExchange:
    """Averages (or sums) the gradients of `params` over the ranks of `group` as ONE flat bucket per step.

    method="peer" (CUDA): the library's one-kernel all-reduce over NVLink peer memory (s2l_allreduce_peer): every rank's
        bucket lives in a CUDA-IPC buffer its peers have mapped; the kernel signals, waits and sums all payloads in rank order
        — bit-identical results on every rank, one launch per step, no NCCL on the path.  The 64-byte IPC handles travel
        through `group` once, at construction (any backend).  Raises if the handles cannot be exchanged or opened.
    method="collective": one torch.distributed.all_reduce of the flat bucket (NCCL for CUDA tensors, gloo for CPU tensors) —
        the library baseline the peer kernel is measured against, and the host-logic path the CPU tests cover."""

    def __init__(self, params, group=None, method="peer", average=True):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("GradExchange: no parameters")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.average = average
        self.method = method
        self.n = sum(p.numel() for p in self.params)
        self.device = self.params[0].device
        self.epoch = 0
        if method == "collective":
            self.flat = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self._views_out = self._views(self.flat)
            return
        if method != "peer":
            raise ValueError("GradExchange: method must be 'peer' or 'collective'")
        if self.device.type != "cuda":
            raise RuntimeError("GradExchange(method='peer') needs CUDA parameters; use method='collective' for CPU tensors")
        import ctypes as C
        from . import _cabi
        self._C, self._cabi, self._lib = C, _cabi, _cabi.lib()
        with torch.cuda.device(self.device):
            own, handle = C.c_void_p(), (C.c_uint8 * 64)()
            _cabi.check(self._lib.s2l_peer_alloc(self.n, C.byref(own), handle), "s2l_peer_alloc")
            self._own = own.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._ptrs = (C.c_void_p * self.world)()
            self._opened = []
            for r in range(self.world):
                if r == self.rank:
                    self._ptrs[r] = self._own
                    continue
                p, h = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(handles[r])
                _cabi.check(self._lib.s2l_peer_open(h, C.byref(p)), "s2l_peer_open (rank %d)" % r)
                self._ptrs[r] = p.value
                self._opened.append(p.value)
            self._payload = [torch.as_tensor(_DeviceSpan(self._own + self._lib.s2l_peer_payload_offset(self.n, e), self.n),
                                             device=self.device) for e in (0, 1)]
            self._views_in = [self._views(t) for t in self._payload]
            self.flat = torch.empty(self.n, dtype=torch.float32, device=self.device)
            self._views_out = self._views(self.flat)
        dist.barrier(group=group)          # every rank has mapped every buffer before the first signal is written

    def _views(self, flat):
        out, off = [], 0
        for p in self.params:
            out.append(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return out

    def _fill(self, views):
        grads = [p.grad if p.grad is not None else None for p in self.params]
        have = [(v, g) for v, g in zip(views, grads) if g is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g.detach() for _, g in have])
        for v, g in zip(views, grads):
            if g is None:
                v.zero_()

    def allreduce(self):
        """grads of `params` <- mean (or sum) over ranks; afterwards every p.grad is a view into self.flat."""
        self.epoch += 1
        scale = 1.0 / self.world if self.average else 1.0
        with torch.no_grad():
            if self.method == "collective":
                self._fill(self._views_out)
                dist.all_reduce(self.flat, group=self.group)
                if scale != 1.0:
                    self.flat.mul_(scale)
            else:
                self._fill(self._views_in[self.epoch & 1])
                with torch.cuda.device(self.device):
                    self._cabi.check(self._lib.s2l_allreduce_peer(self._ptrs, self.rank, self.world, self.n, scale, self.epoch,
                                                                  self.flat.data_ptr(), torch.cuda.current_stream().cuda_stream),
                                     "s2l_allreduce_peer")
            for p, v in zip(self.params, self._views_out):
                p.grad = v
        return self.flat

    def close(self):
        if self.method == "peer" and getattr(self, "_own", None):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody is still reading a buffer that is about to be unmapped
            for p in self._opened:
                self._lib.s2l_peer_close(p)
            self._payload = self._views_in = None
            self._lib.s2l_peer_free(self._own)
            self._own, self._opened = None, []
