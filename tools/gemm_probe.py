"""Experiment driver (not part of the product): times one fused-dequant GEMM shape through the C-ABI test hook.
CLIP_B200_GEMM_DBG switches off pipeline stages to attribute time:  1 unpack math, 2 X loads, 4 Q loads, 8 epilogue."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import binding as bd, oracle as orc
M, N, K = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (9252, 1024, 4096))]
epi = int(sys.argv[4]) if len(sys.argv) > 4 else 0
qt = int(sys.argv[5]) if len(sys.argv) > 5 else 2
lib = bd.ClipLib()
rng = np.random.default_rng(0)
w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
raw = orc.quantize_rows(qt, w)
x = rng.standard_normal((M, K)).astype(np.float32)
y = np.empty((M, N), np.float32); resid = np.zeros((M, N), np.float32)
fp = C.POINTER(C.c_float); ms = C.c_float(0)
buf = np.frombuffer(raw, np.uint8)
best = 1e9
for i in range(4):
    rc = lib.lib.clip_b200_debug_gemm(qt, 1, M, N, K, epi, 0, x.ctypes.data_as(fp), buf.ctypes.data, None, resid.ctypes.data_as(fp), y.ctypes.data_as(fp), C.byref(ms))
    assert rc == 0, lib.last_error()
    best = min(best, ms.value)
pair = (N % 256 == 0)
tiles = (N // (256 if pair else 128)) * ((M + 191) // 192); kb = K // 64
per_cta = -(-tiles // (74 if pair else 148)) * kb
print("dbg=%s M=%d N=%d K=%d epi=%d qt=%d: %.1f us  %.0f TFLOP/s  ~%.0f cycles/k-block (%d k-blocks on the busiest CTA)" % (
    os.environ.get("CLIP_B200_GEMM_DBG", "0"), M, N, K, epi, qt, best * 1e3, 2.0 * M * N * K / best / 1e9, best * 1e-3 * 1.965e9 / per_cta, per_cta))
