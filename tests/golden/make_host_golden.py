"""Golden vectors for the host-side interface members (tokenizer, preprocess), produced by the REFERENCE
(oracle/_ref).  Run in the build container:  python tests/golden/make_host_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import binding as bd          # noqa: E402
import ref_run                # noqa: E402
import synth_gguf as sg       # noqa: E402

TEXTS = ["a photo of a cat", "a photo of the dog's red apple", "two  cats,   3 dogs!", "it's 12 o'clock... we've won",
         " leading space", "trailing space ", "tabs\tand\nnewlines", "", "apple", "UPPER lower 123abc", "a" * 40,
         "caf\xc3\xa9 na\xc3\xafve".encode("latin1").decode("latin1"), "x'll y'd z'm 're"]
IMAGES = [(64, 64, 1), (300, 200, 2), (97, 211, 3), (640, 480, 4), (50, 80, 5)]   # (nx, ny, seed)


def synth_u8(nx, ny, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = rng.integers(0, 256, size=(ny // 8 + 2, nx // 8 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:ny, :nx]
    img += rng.normal(0, 12, size=img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    model = sg.model_path("tiny", 1234, "f16")
    if not os.path.exists(model):
        sg.write_model(model, sg.GEOMETRIES["tiny"], 1234, 1)
    ref = bd.ClipLib(ref_run.REF_LIB)
    ctx = ref.load(model, 0)
    out = {"model_sha": sg.sha256_file(model), "tokens": {}, "preprocess": []}
    for t in TEXTS:
        out["tokens"][t] = [int(v) for v in ref.tokenize(ctx, t)]
    for nx, ny, seed in IMAGES:
        u8 = synth_u8(nx, ny, seed)
        f = ref.preprocess(ctx, u8)
        out["preprocess"].append({"nx": nx, "ny": ny, "seed": seed, "sha256": hashlib.sha256(f.tobytes()).hexdigest(),
                                  "sum": float(f.astype(np.float64).sum()), "first": [float(v) for v in f.ravel()[:6]]})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "host_ops.json"), "w"), indent=1)
    print("wrote host_ops.json:", len(out["tokens"]), "texts,", len(out["preprocess"]), "images")


if __name__ == "__main__":
    main()
