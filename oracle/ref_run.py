"""oracle/ref_run.py -- TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference (oracle/_ref/libclip_ref.so, built by
oracle/Makefile from /root/reference) on given inputs, ONE MODEL PER PROCESS.

Why a subprocess: the reference's batched-output node reads an uninitialised arena tensor
(clip.cpp:1446-1454: `output` is created with ggml_new_tensor_2d and only ever ggml_acc'ed into), so results are
only reproducible from a fresh process with single-image calls (fresh mmap'ed arena == zeros).  Mixing batch sizes
or creating several contexts in one process yields garbage (verified, see DESIGN.md "oracle caveats").
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_LIB = os.path.join(HERE, "_ref", "libclip_ref.so")      # the unmodified reference (oracle/Makefile)


REF_LIB_V4 = os.path.join(HERE, "_ref", "libclip_ref_v4.so")   # same sources, -march=x86-64-v4 (AVX-512 hosts, timing only)


def available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libclip_ref.so"))


def host_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= fl
    except OSError:
        pass
    return False


def timing_lib():
    """(path, isa description) of the build the CPU-baseline legs should time: the AVX-512 build on AVX-512 hosts (what the
    reference's own CLIP_NATIVE=ON build selects there), else the AVX2 one.  Parity work always uses REF_LIB."""
    if os.environ.get("CLIP_REF_ISA", "") != "avx2" and os.path.exists(REF_LIB_V4) and host_has_avx512():
        return REF_LIB_V4, "reference built -O3 -march=x86-64-v4 (AVX-512 host)"
    return REF_LIB, "reference built -O3 -mavx2 -mfma -mf16c (ggml's AVX2 kernels)"


def run_reference(model: str, images=None, token_seqs=None, n_threads: int = 0, normalize: bool = True, timing: bool = False,
                  u8_image=None, texts=None, lib_path: str = None):
    """images: [n,S,S,3] f32 or None; token_seqs: list of int32 arrays or None.  Returns dict(img=, txt=[, img_s=, txt_s=]).
    u8_image [ny,nx,3] + texts (list of str): additionally runs clip_compare_text_and_image per text (res["cmp"]) and
    clip_zero_shot_label_image over all texts (res["zsl_scores"], res["zsl_idx"]) -- clip.cpp:1534-1571, 1624-1659."""
    with tempfile.TemporaryDirectory() as td:
        inp, out = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        d = {"normalize": np.array(int(normalize)), "n_threads": np.array(n_threads), "lib": np.array(lib_path or REF_LIB)}
        if images is not None:
            d["images"] = np.ascontiguousarray(images, np.float32)
        if token_seqs is not None:
            d["tok_flat"] = np.concatenate([np.asarray(t, np.int32) for t in token_seqs]) if len(token_seqs) else np.zeros(0, np.int32)
            d["tok_lens"] = np.array([len(t) for t in token_seqs], np.int64)
        if u8_image is not None and texts:
            d["u8"] = np.ascontiguousarray(u8_image, np.uint8)
            d["texts"] = np.array(list(texts))
        np.savez(inp, **d)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), model, inp, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError("reference run failed:\n" + r.stderr.decode()[-2000:])
        z = np.load(out)
        return {k: z[k] for k in z.files}


def _main():
    import time
    sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200"))
    import binding as bd
    model, inp, out = sys.argv[1:4]
    z = np.load(inp)
    nt = int(z["n_threads"]) or (os.cpu_count() or 4)
    normalize = bool(int(z["normalize"]))
    ref = bd.ClipLib(str(z["lib"]) if "lib" in z.files else REF_LIB)
    ctx = ref.load(model, 0)
    res = {}
    if "images" in z.files:
        imgs = z["images"]
        t0 = time.perf_counter()
        res["img"] = np.stack([ref.image_encode(ctx, np.ascontiguousarray(imgs[i]), normalize, nt) for i in range(len(imgs))]) \
            if len(imgs) else np.zeros((0, 1), np.float32)
        res["img_s"] = np.array(time.perf_counter() - t0)
    if "tok_lens" in z.files:
        lens, flat = z["tok_lens"], z["tok_flat"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        t0 = time.perf_counter()
        res["txt"] = np.stack([ref.text_encode(ctx, flat[offs[i]:offs[i + 1]], normalize, nt) for i in range(len(lens))]) \
            if len(lens) else np.zeros((0, 1), np.float32)
        res["txt_s"] = np.array(time.perf_counter() - t0)
    if "u8" in z.files:
        # scoring entry points LAST and zero-shot before compare: they run text + image graphs back to back in one context
        u8, texts = z["u8"], [str(t) for t in z["texts"]]
        sc, ix = ref.zero_shot_label_image(ctx, u8, texts, nt)
        res["zsl_scores"], res["zsl_idx"] = sc, ix
        res["cmp"] = np.array([ref.compare_text_and_image(ctx, t, u8, nt) for t in texts], np.float32)
    res["threads"] = np.array(nt)
    np.savez(out, **res)


if __name__ == "__main__":
    _main()
