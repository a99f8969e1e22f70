"""world_size-2 gloo tests of the sharding / all-gather plumbing the N>1 bench path uses (no GPU, 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest

import dist_util as du


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 512, 4096, 1000):
        for w in (1, 2, 3, 8):
            b = [du.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_embed(lo, hi, d):
    i = np.arange(lo, hi, dtype=np.float32)[:, None]
    return np.sin(i * 0.37 + np.arange(d, dtype=np.float32)[None, :] * 0.11).astype(np.float32)


def _worker(rank, world, port, n_img, n_txt, d, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = du.shard_bounds(n_img, rank, world)
        local = torch.from_numpy(_fake_embed(lo, hi, d))
        full = du.all_gather_rows(dist, local, n_img, world)
        tlo, thi = du.shard_bounds(n_txt, rank, world)
        tloc = torch.from_numpy(_fake_embed(1000 + tlo, 1000 + thi, d))
        s, i = du.zero_shot_sharded(dist, rank, world, local, tloc, n_txt, 5)
        q.put((rank, full.numpy(), s.numpy(), i.numpy(), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img,n_txt", [(10, 7), (16, 8)])
def test_all_gather_and_sharded_zero_shot_world2(n_img, n_txt):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, d, world = _free_port(), 16, 2
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_img, n_txt, d, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    img, txt = _fake_embed(0, n_img, d), _fake_embed(1000, 1000 + n_txt, d)
    e = np.exp((img @ txt.T).astype(np.float64)).astype(np.float32) + np.float32(1e-9)
    p_ref = (e / e.astype(np.float64).sum(1, keepdims=True)).astype(np.float32)
    for rank, full, s, i, lo, hi in res:
        assert np.array_equal(full, img)                          # every rank ends with the full matrix, in order
        order = np.argsort(-p_ref[lo:hi], axis=1, kind="stable")[:, :5]
        assert np.array_equal(i, order)
        assert np.allclose(s, np.take_along_axis(p_ref[lo:hi], order, 1), rtol=1e-5)
