"""N1: clip_image_preprocess / clip_image_batch_preprocess (clip.cpp:797-1008) on the device.
The device result must be bit-identical to the library's host path (which tests/test_host_side.py pins against reference outputs),
and the fused u8 -> embedding call must equal preprocess-on-host + clip_image_batch_encode exactly."""
import numpy as np
import pytest

from _util import model_file

pytestmark = pytest.mark.gpu

SIZES = [(224, 224), (300, 200), (200, 300), (640, 480), (100, 150), (64, 64), (500, 1000), (225, 224), (37, 91), (1024, 768)]


def _images(sizes, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i, (nx, ny) in enumerate(sizes):
        if i % 3 == 0:      # smooth gradient + noise, exercises clamping less; pure noise exercises it more
            yy, xx = np.mgrid[0:ny, 0:nx]
            img = np.stack([(xx * 255 // max(nx - 1, 1)), (yy * 255 // max(ny - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
        else:
            img = rng.integers(0, 256, (ny, nx, 3), dtype=np.uint8)
        out.append(img)
    return out


@pytest.fixture(scope="module")
def ctx(prod):
    c = prod.load(model_file("tiny", "f16", prod), 0)
    prod.lib.clip_b200_set_micro_batch(c, 64, 0)        # 700 images below = 11 micro-batches through both staging arenas
    yield c
    prod.free(c)


def test_device_preprocess_is_bit_identical_to_host(prod, ctx):
    imgs = _images(SIZES, 1)
    dev = prod.preprocess_device(ctx, imgs)
    for i, im in enumerate(imgs):
        host = prod.preprocess(ctx, im)
        assert host.shape == dev[i].shape
        assert np.array_equal(host, dev[i]), (SIZES[i], np.abs(host - dev[i]).max())


def test_reference_batch_call_on_host_threads_and_on_device(prod, ctx, monkeypatch):
    """clip_image_batch_preprocess (clip.h:95-96) as a reference caller uses it: host threads by default, the same call on the GPU
    with CLIP_B200_PREPROCESS=device -- identical buffers either way."""
    imgs = _images(SIZES, 5)
    want = np.stack([prod.preprocess(ctx, im) for im in imgs])
    monkeypatch.delenv("CLIP_B200_PREPROCESS", raising=False)
    assert np.array_equal(prod.batch_preprocess(ctx, imgs, n_threads=3), want)
    monkeypatch.setenv("CLIP_B200_PREPROCESS", "device")
    assert np.array_equal(prod.batch_preprocess(ctx, imgs, n_threads=3), want)


def test_fused_u8_encode_equals_host_preprocess_plus_encode(prod, ctx):
    imgs = _images(SIZES * 3, 2)
    fused = prod.image_batch_encode_u8(ctx, imgs)
    host_px = np.stack([prod.preprocess(ctx, im) for im in imgs])
    ref = prod.image_batch_encode(ctx, host_px)
    assert np.array_equal(fused, ref)


def test_many_images_several_micro_batches(prod, ctx):
    rng = np.random.default_rng(3)
    sizes = [(int(rng.integers(40, 400)), int(rng.integers(40, 400))) for _ in range(700)]
    imgs = _images(sizes, 4)
    fused = prod.image_batch_encode_u8(ctx, imgs)
    assert np.isfinite(fused).all()
    idx = [0, 1, 350, 699]
    host_px = np.stack([prod.preprocess(ctx, imgs[i]) for i in idx])
    assert np.array_equal(fused[idx], prod.image_batch_encode(ctx, host_px))
    # a second call re-uses the staging arenas
    assert np.array_equal(prod.image_batch_encode_u8(ctx, imgs[:50]), fused[:50])


def test_bad_inputs(prod, ctx):
    assert prod.image_batch_encode_u8(ctx, []).shape[0] == 0
    with pytest.raises(RuntimeError):
        prod.image_batch_encode_u8(ctx, [np.zeros((0, 5, 3), np.uint8)])


def test_device_preprocess_matches_reference_golden_directly(prod, ctx):
    """the DEVICE output against the sha256 of what the reference's clip_image_preprocess produced (tests/golden/host_ops.json,
    written by make_host_golden.py from oracle/_ref) -- no detour through the library's own host path"""
    import hashlib
    import json
    import os
    from _util import GOLDEN
    host = json.load(open(os.path.join(GOLDEN, "host_ops.json")))

    def synth_u8(nx, ny, seed):        # same generator as tests/golden/make_host_golden.py
        rng = np.random.Generator(np.random.PCG64(seed))
        base = rng.integers(0, 256, size=(ny // 8 + 2, nx // 8 + 2, 3)).astype(np.float32)
        img = np.kron(base, np.ones((8, 8, 1), np.float32))[:ny, :nx]
        img += rng.normal(0, 12, size=img.shape).astype(np.float32)
        return np.clip(img, 0, 255).astype(np.uint8)

    imgs = [synth_u8(e["nx"], e["ny"], e["seed"]) for e in host["preprocess"]]
    dev = prod.preprocess_device(ctx, imgs)
    for e, out in zip(host["preprocess"], dev):
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == e["sha256"], e
