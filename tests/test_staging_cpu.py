"""Host logic of the file-backed input staging (speech2lip_b200.staging.NpyPrefetcher) on CPU tensors: order, values,
dtype conversion, more files than slots, a failing load surfaces at the consumer."""
import os

import numpy as np
import pytest
import torch

from speech2lip_b200.staging import NpyPrefetcher


def _files(tmp_path, n, shape=(12, 10, 2)):
    rng = np.random.default_rng(3)
    paths, arrs = [], []
    for i in range(n):
        a = rng.standard_normal(shape).astype(np.float64 if i % 2 else np.float32)
        p = os.path.join(tmp_path, "%05d.npy" % i)
        np.save(p, a)
        paths.append(p)
        arrs.append(a)
    return paths, arrs


def test_prefetcher_yields_files_in_order(tmp_path):
    paths, arrs = _files(str(tmp_path), 11)
    pf = NpyPrefetcher(paths, "cpu", depth=3)
    assert len(pf) == 11
    got = list(pf)
    assert len(got) == 11
    for g, a in zip(got, arrs):
        assert g.dtype == torch.float32 and tuple(g.shape) == a.shape
        assert torch.equal(g, torch.from_numpy(a).float())


def test_prefetcher_surfaces_a_failed_load(tmp_path):
    paths, _ = _files(str(tmp_path), 3)
    paths.insert(2, os.path.join(str(tmp_path), "missing.npy"))
    it = iter(NpyPrefetcher(paths, "cpu", depth=2))
    next(it)
    next(it)
    with pytest.raises(RuntimeError):
        next(it)


@pytest.mark.gpu
def test_prefetcher_on_the_gpu_matches_the_files(tmp_path):
    paths, arrs = _files(str(tmp_path), 9, shape=(50, 40, 2))
    total = torch.zeros(50, 40, 2, device="cuda")
    for g in NpyPrefetcher(paths, "cuda", depth=4):
        assert g.is_cuda
        total += g                                    # consumed on the current stream, ordered after the copy by the event
    want = sum(torch.from_numpy(a).float() for a in arrs)
    assert torch.allclose(total.cpu(), want, atol=1e-5)
