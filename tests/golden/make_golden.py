"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref, built from /root/reference by
oracle/Makefile) on seeded synthetic models and inputs.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py            # small models, all file types
    python tests/golden/make_golden.py --full     # + true ViT-B/32 and ViT-L/14 geometries (minutes)

Each fixture stores the sha256 of every model file it was produced from; tests regenerate the files from the
same seeds (clip.cpp_b200/synth_gguf.py + clip_model_quantize, which is byte-identical to the reference's
quantizer) and refuse to compare if a hash drifts.  The reference is called with batch = 1 from a fresh process
per model (see oracle/ref_run.py for why).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import binding as bd          # noqa: E402
import ref_run                # noqa: E402
import synth_gguf as sg       # noqa: E402

FTYPES = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}
SEED = 1234
IMG_SEED, TOK_SEED = 77, 99
TOK_LENS = [77, 5, 2, 33, 16, 77, 9, 64]

PLAN = {
    "tiny": dict(ftypes=list(FTYPES), n_img=4, n_txt=8),
    "tiny-gelu": dict(ftypes=["f16", "q4_0"], n_img=2, n_txt=2),
    "small-p14": dict(ftypes=["f16", "q4_0", "q8_0"], n_img=2, n_txt=2),
    "small-p14-336": dict(ftypes=["f16", "q4_0"], n_img=2, n_txt=2),
}
PLAN_FULL = {
    "vit-b32": dict(ftypes=["f16", "q4_0", "q4_1", "q5_0", "q5_1", "q8_0"], n_img=3, n_txt=3),
    "vit-l14": dict(ftypes=["f16", "q4_0", "q4_1", "q5_0", "q5_1", "q8_0"], n_img=3, n_txt=3),
}


def ensure_model(geom: str, ftype: str, ref_quantizer) -> str:
    """Create (or reuse) the model file for (geom, SEED, ftype) in the cache; returns its path."""
    path = sg.model_path(geom, SEED, ftype)
    if os.path.exists(path):
        return path
    if ftype in ("f32", "f16"):
        sg.write_model(path, sg.GEOMETRIES[geom], SEED, FTYPES[ftype])
    else:
        src = ensure_model(geom, "f16", ref_quantizer)
        assert ref_quantizer(src, path, FTYPES[ftype])
    return path


def token_seqs(n):
    seqs = []
    for i in range(n):
        L = TOK_LENS[i % len(TOK_LENS)]
        seqs.append(sg.synth_tokens(1, L, TOK_SEED + i)[0])
    return seqs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    plan = dict(PLAN)
    if a.full:
        plan.update(PLAN_FULL)
    if a.only:
        plan = {k: v for k, v in {**PLAN, **PLAN_FULL}.items() if k in a.only.split(",")}
    ref = bd.ClipLib(ref_run.REF_LIB)
    # the reference prints one line per tensor while quantizing: silence fd 1 around it
    def quiet_quant(src, dst, it):
        sys.stdout.flush()
        old = os.dup(1)
        dn = os.open(os.devnull, os.O_WRONLY)
        os.dup2(dn, 1)
        try:
            return ref.quantize(src, dst, it)
        finally:
            os.dup2(old, 1)
            os.close(dn)
            os.close(old)
    for geom, p in plan.items():
        g = sg.GEOMETRIES[geom]
        imgs = sg.synth_images(p["n_img"], g.image_size, IMG_SEED)
        seqs = token_seqs(p["n_txt"])
        out = {"seed": SEED, "img_seed": IMG_SEED, "tok_seed": TOK_SEED, "n_img": p["n_img"], "n_txt": p["n_txt"],
               "tok_lens": np.array([len(s) for s in seqs])}
        for ft in p["ftypes"]:
            path = ensure_model(geom, ft, quiet_quant)
            r = ref_run.run_reference(path, imgs, seqs)
            out["sha_" + ft] = sg.sha256_file(path)
            out["img_" + ft] = r["img"].astype(np.float32)
            out["txt_" + ft] = r["txt"].astype(np.float32)
            print("%-10s %-5s img %.2fs txt %.2fs  |img0|=%.4f" % (geom, ft, float(r["img_s"]), float(r["txt_s"]),
                                                               float(np.linalg.norm(r["img"][0]))), flush=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "%s-s%d.npz" % (geom, SEED)), **out)


if __name__ == "__main__":
    main()
