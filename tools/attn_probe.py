import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "/root/repo/clip.cpp_b200")
import binding as bd
lib = bd.ClipLib()
nseq, T, H = 82, 257, 16
rng = np.random.default_rng(0)
qkv = rng.standard_normal((nseq * T, 3 * H * 64)).astype(np.float32); qkv[:, :H*64] *= 0.3
out = np.empty((nseq * T, H * 64), np.float32)
fp = C.POINTER(C.c_float); ms = C.c_float(0)
best = 1e9
for i in range(3):
    rc = lib.lib.clip_b200_debug_attention(1, nseq, T, H, 0, 0, qkv.ctypes.data_as(fp), out.ctypes.data_as(fp), C.byref(ms))
    assert rc == 0, lib.last_error()
    best = min(best, ms.value)
print("skew=%s: %.1f us" % (os.environ.get("CLIP_B200_ATTN_SKEW", "0"), best * 1e3))
