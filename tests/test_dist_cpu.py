"""CPU, gloo, world_size 2: the multi-GPU host logic (frame sharding + one parameter broadcast)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_partitions_exactly():
    from speech2lip_b200.dist import shard_frames
    for n in (0, 1, 7, 8, 125, 1000):
        for world in (1, 2, 4, 8):
            spans = [shard_frames(n, r, world) for r in range(world)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= (n + world - 1) // world


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import synth
    from speech2lip_b200.dist import broadcast_params, broadcast_module, gather_frames, shard_frames
    # rank 0 holds the real weights, rank 1 garbage: after ONE broadcast both must agree
    sd = {k: torch.from_numpy(v).clone() for k, v in synth.make_state_dict(seed=rank, kind="default").items()}
    broadcast_params(sd, src=0)
    ref = synth.make_state_dict(seed=0, kind="default")
    same = all(torch.equal(sd[k], torch.from_numpy(ref[k])) for k in ref)
    lo, hi = shard_frames(5, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1, 1).expand(-1, 2, 2, 3).contiguous()
    full = gather_frames(local, 5)
    ok_gather = torch.equal(full[:, 0, 0, 0], torch.arange(5, dtype=torch.float32))
    # the whole training module (UNet weights, BatchNorm running statistics / int64 counters, depth head): DDP's constructor broadcast
    import json
    import speech2lip_b200 as s2l
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "may_cfg.json")))
    torch.manual_seed(100 + rank)
    m = s2l.TalkingFace(device=torch.device("cpu"), cfg=cfg)
    m.post_fusion_unet.inc.double_conv[1].running_mean.fill_(float(rank) + 0.5)
    m.post_fusion_unet.inc.double_conv[1].num_batches_tracked.fill_(7 + rank)
    broadcast_module(m, src=0)
    digest = torch.stack([t.double().sum() for t in list(m.parameters()) + list(m.buffers())]).sum().reshape(1)
    both = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    same_module = bool(torch.equal(both[0], both[1])) and float(m.post_fusion_unet.inc.double_conv[1].running_mean[0]) == 0.5 \
        and int(m.post_fusion_unet.inc.double_conv[1].num_batches_tracked) == 7
    # data-parallel training exchange, host logic (the flat-bucket path; the one-kernel peer path is a GPU test):
    # rank-specific gradients, one parameter without a gradient on rank 1, two consecutive steps
    from speech2lip_b200.dist import GradExchange
    ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    ex = GradExchange(ps, method="collective", average=True)
    ok_ex = True
    for step in (1, 2):
        for i, p in enumerate(ps):
            p.grad = None if (rank == 1 and i == 1) else torch.full_like(p, float((rank + 1) * (i + 1) * step))
        ex.allreduce()
        for i, p in enumerate(ps):
            want = ((1 + (0 if i == 1 else 2)) * (i + 1) * step) / 2.0
            ok_ex = ok_ex and bool((p.grad == want).all()) and p.grad.shape == p.shape
    try:
        GradExchange(ps, method="peer")
        ok_ex = False                                 # CPU tensors must be refused loudly, not silently routed elsewhere
    except RuntimeError:
        pass
    q.put((rank, same and same_module and ok_ex, ok_gather, (lo, hi)))
    dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], "parameters differ after broadcast"
    assert res[0][2] and res[1][2], "gather_frames lost frame order"
    assert res[0][3] == (0, 3) and res[1][3] == (3, 5)
