"""N4: the reference's own zero-shot accuracy benchmark, /root/reference/tests/benchmark.cpp, compiled UNCHANGED
twice by oracle/Makefile (target `harness`): against the reference library and against this repository's headers + libclip_b200.so.
Both binaries walk the same class-per-directory tree (synthetic PNGs, ViT-B/32 geometry -- the one shape for which the reference's
batched conv is correct) and must print the same acc@1 / acc@5 table; the timing lines are reported."""
import os
import re
import subprocess

import numpy as np
import pytest

from _util import ROOT, model_file

pytestmark = pytest.mark.gpu
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "benchmark_ref")
B200_BIN = os.path.join(ROOT, "oracle", "_ref", "benchmark_b200")
CLASSES = ["apple", "dog", "cat", "car", "tree", "house", "bird", "fish"]


def make_tree(root, per_class=4, seed=3):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(seed)
    for c in CLASSES:
        os.makedirs(os.path.join(root, c), exist_ok=True)
        for k in range(per_class):
            nx, ny = int(rng.integers(100, 400)), int(rng.integers(100, 400))
            base = rng.integers(0, 256, (ny // 16 + 2, nx // 16 + 2, 3)).astype(np.float32)
            img = np.kron(base, np.ones((16, 16, 1), np.float32))[:ny, :nx] + rng.normal(0, 10, (ny, nx, 3))
            Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(root, c, "img%d.png" % k))


def run(binary, model, tree, out):
    r = subprocess.run([binary, model, tree, "0", out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (binary, r.stdout[-500:], r.stderr[-500:])
    txt = open(out).read()
    table = {m.group(1).strip(): (float(m.group(2)), float(m.group(3))) for m in re.finditer(r"\| ([a-z ]+?)\s+\| ([0-9.]+) \| ([0-9.]+) \|", txt)}
    timing = re.findall(r"- (\d+) (texts|images) encoded in\s+([0-9.]+) ms", txt)
    return table, timing


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(B200_BIN)), reason="oracle/_ref/benchmark_* not built (make -C oracle harness)")
def test_reference_benchmark_source_gives_identical_accuracy_table(prod, tmp_path):
    model = model_file("vit-b32", "f16", prod)
    tree = str(tmp_path / "tree")
    make_tree(tree)
    ref_table, ref_t = run(REF_BIN, model, tree, str(tmp_path / "ref.txt"))
    got_table, got_t = run(B200_BIN, model, tree, str(tmp_path / "b200.txt"))
    print("reference:", ref_table, ref_t)
    print("b200     :", got_table, got_t)
    assert set(ref_table) == set(CLASSES) | {"total"}
    assert got_table == ref_table
