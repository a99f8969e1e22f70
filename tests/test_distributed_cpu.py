"""world_size-2 gloo tests of the sharding / all-gather plumbing the N>1 bench path uses (no GPU, 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest

import gloo_mirror as du


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


# ---- the library's own multi-GPU plumbing, the parts that run without a GPU ------------------------------------------
def test_library_shard_rule_matches_python_helper(prod):
    import ctypes as C
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    for n in (0, 1, 7, 512, 4096, 1000):
        for w in (1, 2, 3, 8):
            for r in range(w):
                prod.lib.clip_b200_debug_shard_bounds(n, r, w, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == du.shard_bounds(n, r, w)


def _rdzv_worker(rank, world, path, q):
    import ctypes as C
    import binding as bd
    lib = bd.ClipLib(bd.PRODUCT_LIB)
    buf = (C.c_ubyte * 128)()
    rc = lib.lib.clip_b200_debug_rendezvous(rank, world, path.encode() if path else None, buf)
    q.put((rank, rc, bytes(buf), lib.last_error()))


@pytest.mark.parametrize("explicit_path", [True, False])
def test_ranks_mode_rendezvous_world2(prod, explicit_path, tmp_path):
    """world_size-2 exchange of the NCCL unique id through the rendezvous file (csrc/dist.cpp), as the rank processes under
    torchrun do it: rank 0 creates the id with ncclGetUniqueId and publishes it, rank 1 must read the same 128 bytes."""
    if prod.lib.clip_b200_nccl_version() == 0:
        pytest.skip("libnccl.so.2 not loadable here")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    os.environ["MASTER_PORT"] = str(_free_port())           # default path = /tmp/clip_b200_rdzv_<port>_<parent pid>.<seq>
    path = str(tmp_path / "rdzv") if explicit_path else ""
    ps = [ctx.Process(target=_rdzv_worker, args=(r, 2, path, q)) for r in (1, 0)]      # the reader starts first
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[1][1] == 0, (res[0][3], res[1][3])
    assert res[0][2] == res[1][2] and any(res[0][2])
