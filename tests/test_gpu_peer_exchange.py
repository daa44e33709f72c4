"""The data-parallel training exchange (training): s2l_allreduce_peer through speech2lip_b200.dist.GradExchange.
Two processes share cuda:0 (CUDA IPC works between processes on one device; the handles travel over gloo), so the test runs
on a one-GPU box; bench.py --config 5 under torchrun exercises the same code with one GPU per rank over NVLink."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        from speech2lip_b200.dist import GradExchange
        g = torch.Generator().manual_seed(1234)
        shapes = [(256, 256), (256,), (256, 42), (3, 256), (5,), (1, 3)]        # the bucket length is not a multiple of 4
        ps = [torch.nn.Parameter(torch.zeros(s, device="cuda")) for s in shapes]
        per_rank = [[torch.randn(s, generator=g) for s in shapes] for _ in range(world * 3)]
        peer = GradExchange(ps, method="peer", average=True)
        ok, worst = True, 0.0
        sums = []
        for step in range(3):
            for i, p in enumerate(ps):
                p.grad = None if (rank == world - 1 and i == 4 and step == 1) else per_rank[step * world + rank][i].cuda()
            peer.allreduce()
            for i, p in enumerate(ps):
                terms = [per_rank[step * world + r][i] for r in range(world) if not (r == world - 1 and i == 4 and step == 1)]
                want = torch.zeros(shapes[i])
                for t in terms:                       # rank order, like the kernel
                    want = want + t
                want = want * (1.0 / world)
                err = (p.grad.cpu() - want).abs().max().item()
                worst = max(worst, err)
                ok = ok and err <= 1e-6 and p.grad.shape == p.shape
            sums.append(peer.flat.clone())
        # bit-identical on every rank: compare the raw words of the reduced buckets across ranks
        mine = torch.stack(sums).view(torch.int32).sum(dtype=torch.int64).cpu().reshape(1)
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        ok = ok and all(torch.equal(b, both[0]) for b in both)
        peer.close()
        q.put((rank, ok, worst))
        dist.destroy_process_group()
    except Exception as e:                            # surface the failure instead of a queue timeout
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("world", [1, 2])
def test_peer_allreduce_matches_sum_in_rank_order(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    print("peer all-reduce world %d:" % world, res)
    assert all(r[1] for r in res), res


def test_peer_argument_errors():
    import ctypes as C
    from speech2lip_b200 import _cabi
    lib = _cabi.lib()
    ptrs = (C.c_void_p * 1)()
    out = torch.zeros(8, device="cuda")
    assert lib.s2l_allreduce_peer(ptrs, 0, 1, 8, 1.0, 1, out.data_ptr(), None) == 1          # null peer buffer
    assert lib.s2l_allreduce_peer(ptrs, 0, 17, 8, 1.0, 1, out.data_ptr(), None) == 2         # world beyond the limit
    assert lib.s2l_allreduce_peer(ptrs, 0, 1, 8, 1.0, 0, out.data_ptr(), None) == 2          # epoch 0 is reserved
    assert lib.s2l_peer_buffer_bytes(1000) == 4096 + 2 * 4096
