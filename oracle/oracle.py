"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of clip.cpp's encode path: numpy drives the graph, `liboracle.so` (clip_oracle.c)
does the arithmetic with the reference's rounding points.  Paths below are relative to /root/reference.

  vision forward  : clip.cpp:1247-1523        text forward : clip.cpp:1016-1233
  GGUF container  : ggml/src/ggml.c:19751-20063 (reader), tensor/KV names clip.cpp:41-79

Pinning: tests/test_oracle_pin.py checks this module against the reference itself (oracle/_ref,
built from the reference's own sources by oracle/Makefile) and against the committed golden vectors
in tests/golden/ that were produced by the reference (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libclip_ref.so")

F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0 = 0, 1, 2, 3, 6, 7, 8
TYPE_NAMES = {0: "f32", 1: "f16", 2: "q4_0", 3: "q4_1", 6: "q5_0", 7: "q5_1", 8: "q8_0"}
FTYPE_BY_NAME = {v: k for k, v in TYPE_NAMES.items()}
BLOCK_BYTES = {Q4_0: 18, Q4_1: 20, Q5_0: 22, Q5_1: 24, Q8_0: 34}


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when the reference sources are present)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "clip_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/clip.cpp") and not os.path.exists(REF_LIB):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        fp, vp, i64 = C.POINTER(C.c_float), C.c_void_p, C.c_int64
        L.orc_row_size.restype = C.c_size_t
        L.orc_row_size.argtypes = [C.c_int, i64]
        L.orc_quantize_row.argtypes = [C.c_int, fp, vp, i64]
        L.orc_dequantize_row.argtypes = [C.c_int, vp, fp, i64]
        L.orc_mul_mat.argtypes = [C.c_int, vp, i64, i64, fp, i64, fp, C.c_int]
        L.orc_layer_norm.argtypes = [fp, fp, i64, i64, C.c_float]
        L.orc_layer_norm.restype = None
        L.orc_gelu.argtypes = [fp, fp, i64, C.c_int]
        L.orc_gelu.restype = None
        L.orc_softmax_rows.argtypes = [fp, i64, i64]
        L.orc_softmax_rows.restype = None
        L.orc_attention.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_attention.restype = None
        L.orc_f32_to_f16_to_f32.argtypes = [fp, fp, i64]
        L.orc_f32_to_f16_to_f32.restype = None
        L.orc_sum_sq_sqrt.argtypes = [fp, i64]
        L.orc_sum_sq_sqrt.restype = C.c_float
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ------------------------------------------------------------------------------------------------
# GGUF reader (ggml.c:19751-20063): header, KV section, tensor infos, aligned data section
# ------------------------------------------------------------------------------------------------
class Tensor:
    __slots__ = ("name", "shape", "type", "data")

    def __init__(self, name, shape, typ, data):
        self.name, self.shape, self.type, self.data = name, shape, typ, data   # shape: numpy order [out, in]

    def f32(self) -> np.ndarray:
        """Dequantise to float32 exactly as ggml's dequantize_row_* (ggml.c:1496-1606)."""
        n = int(np.prod(self.shape))
        out = np.empty(n, np.float32)
        buf = np.frombuffer(self.data, np.uint8)
        rc = lib().orc_dequantize_row(self.type, buf.ctypes.data, _fp(out), n)
        assert rc == 0
        return out.reshape(self.shape)


class GGUF:
    def __init__(self, path: str):
        self.path = path
        self.f = open(path, "rb")
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        mm = self.mm
        magic, version, n_t, n_kv = struct.unpack_from("<IIQQ", mm, 0)
        assert magic == 0x46554747, "not a GGUF file"
        assert version >= 2
        self.version = version
        self.off = 24
        self.kv = {}
        self.kv_order = []
        for _ in range(n_kv):
            key = self._str()
            typ = self._u32()
            self.kv[key] = (typ, self._val(typ))
            self.kv_order.append(key)
        infos = []
        for _ in range(n_t):
            name = self._str()
            nd = self._u32()
            ne = struct.unpack_from("<%dQ" % nd, mm, self.off)
            self.off += 8 * nd
            typ = self._u32()
            (toff,) = struct.unpack_from("<Q", mm, self.off)
            self.off += 8
            infos.append((name, tuple(reversed(ne)), typ, toff))
        align = self.kv.get("general.alignment", (4, 32))[1]
        self.data_start = (self.off + align - 1) // align * align
        self.tensors = {}
        self.order = []
        for name, shape, typ, toff in infos:
            n = int(np.prod(shape))
            nbytes = n * 4 if typ == F32 else n * 2 if typ == F16 else n // 32 * BLOCK_BYTES[typ]
            s = self.data_start + toff
            self.tensors[name] = Tensor(name, shape, typ, memoryview(mm)[s:s + nbytes])
            self.order.append(name)

    def _u32(self):
        (v,) = struct.unpack_from("<I", self.mm, self.off)
        self.off += 4
        return v

    def _str(self):
        (n,) = struct.unpack_from("<Q", self.mm, self.off)
        s = bytes(self.mm[self.off + 8:self.off + 8 + n]).decode("utf-8", "replace")
        self.off += 8 + n
        return s

    _FMT = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<B", 10: "<Q", 11: "<q", 12: "<d"}

    def _val(self, typ):
        if typ == 8:
            return self._str()
        if typ == 9:
            et = self._u32()
            (n,) = struct.unpack_from("<Q", self.mm, self.off)
            self.off += 8
            return [self._val(et) for _ in range(n)]
        fmt = self._FMT[typ]
        (v,) = struct.unpack_from(fmt, self.mm, self.off)
        self.off += struct.calcsize(fmt)
        return bool(v) if typ == 7 else v

    def get(self, key, default=None):
        return self.kv[key][1] if key in self.kv else default


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
class OracleModel:
    def __init__(self, path: str, n_threads: int = 0):
        self.g = g = GGUF(path)
        self.nt = n_threads
        self.use_gelu = bool(g.get("clip.use_gelu"))
        self.has_text = bool(g.get("clip.has_text_encoder"))
        self.has_vision = bool(g.get("clip.has_vision_encoder"))
        if self.has_vision:
            self.image_size = g.get("clip.vision.image_size")
            self.patch = g.get("clip.vision.patch_size")
            self.v_hidden = g.get("clip.vision.embedding_length")
            self.v_heads = g.get("clip.vision.attention.head_count")
            self.v_layers = g.get("clip.vision.block_count")
            self.v_eps = np.float32(g.get("clip.vision.attention.layer_norm_epsilon"))
            self.v_proj = g.get("clip.vision.projection_dim")
        if self.has_text:
            self.t_hidden = g.get("clip.text.embedding_length")
            self.t_heads = g.get("clip.text.attention.head_count")
            self.t_layers = g.get("clip.text.block_count")
            self.t_eps = np.float32(g.get("clip.text.attention.layer_norm_epsilon"))
            self.t_proj = g.get("clip.text.projection_dim")
        self._f32 = {}

    def t(self, name) -> Tensor:
        return self.g.tensors[name]

    def f32(self, name) -> np.ndarray:
        if name not in self._f32:
            self._f32[name] = self.t(name).f32()
        return self._f32[name]

    # y = x . W^T through ggml's mul_mat semantics for W's storage type
    def mul_mat(self, wname: str, x: np.ndarray) -> np.ndarray:
        w = self.t(wname)
        n, k = w.shape
        x = np.ascontiguousarray(x, np.float32)
        m = x.shape[0]
        y = np.empty((m, n), np.float32)
        buf = np.frombuffer(w.data, np.uint8)
        rc = lib().orc_mul_mat(w.type, buf.ctypes.data, n, k, _fp(x), m, _fp(y), self.nt)
        assert rc == 0
        return y

    def ln(self, x, wname, bname, eps):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        lib().orc_layer_norm(_fp(x), _fp(y), x.shape[0], x.shape[1], eps)
        # ggml_mul(repeat(w), y) then ggml_add(.., repeat(b)): two fp32 roundings (clip.cpp:1073-1074)
        return (self.f32(wname) * y) + self.f32(bname)

    def gelu(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        lib().orc_gelu(_fp(x), _fp(y), x.size, 0 if self.use_gelu else 1)
        return y

    def _blocks(self, p, x, n_layers, heads, eps, T, causal):
        hid = x.shape[1]
        dh = hid // heads
        scale = np.float32(1.0) / np.float32(np.sqrt(np.float32(dh)))
        nseq = x.shape[0] // T
        for il in range(n_layers):
            b = "%s.blk.%d." % (p, il)
            cur = self.ln(x, b + "ln1.weight", b + "ln1.bias", eps)
            q = (self.f32(b + "attn_q.bias") + self.mul_mat(b + "attn_q.weight", cur)) * scale
            k = self.f32(b + "attn_k.bias") + self.mul_mat(b + "attn_k.weight", cur)
            v = self.f32(b + "attn_v.bias") + self.mul_mat(b + "attn_v.weight", cur)
            att = np.empty_like(q)
            for s in range(nseq):
                sl = slice(s * T, (s + 1) * T)
                qs, ks, vs = (np.ascontiguousarray(a[sl]) for a in (q, k, v))
                o = np.empty_like(qs)
                lib().orc_attention(_fp(qs), _fp(ks), _fp(vs), _fp(o), T, heads, dh, 1 if causal else 0, self.nt)
                att[sl] = o
            cur = self.f32(b + "attn_out.bias") + self.mul_mat(b + "attn_out.weight", att)
            x = cur + x
            cur = self.ln(x, b + "ln2.weight", b + "ln2.bias", eps)
            cur = self.f32(b + "ffn_down.bias") + self.mul_mat(b + "ffn_down.weight", cur)
            cur = self.gelu(cur)
            cur = self.f32(b + "ffn_up.bias") + self.mul_mat(b + "ffn_up.weight", cur)
            x = x + cur
        return x

    def _normalize(self, e):
        ln = lib().orc_sum_sq_sqrt(_fp(e), e.size)
        return e * (np.float32(1.0) / np.float32(ln))

    def encode_image(self, img: np.ndarray, normalize: bool = True) -> np.ndarray:
        """img: [S, S, 3] f32 NHWC.  One image per call == the reference's batch-of-1 semantics."""
        S, P, hid = self.image_size, self.patch, self.v_hidden
        assert img.shape == (S, S, 3)
        n = S // P
        # conv_2d stride P, pixels rounded to fp16, (c, ky, kx) order (ggml.c:13570-13666)
        planar = np.ascontiguousarray(img.transpose(2, 0, 1), np.float32)                       # [3, S, S]
        patches = planar.reshape(3, n, P, n, P).transpose(1, 3, 0, 2, 4).reshape(n * n, 3 * P * P)
        patches = np.ascontiguousarray(patches)
        pw = self.t("v.patch_embd.weight")
        assert pw.type == F16
        y = np.empty((n * n, hid), np.float32)
        buf = np.frombuffer(pw.data, np.uint8)
        rc = lib().orc_mul_mat(F16, buf.ctypes.data, hid, 3 * P * P, _fp(patches), n * n, _fp(y), self.nt)
        assert rc == 0
        x = np.zeros((n * n + 1, hid), np.float32)
        x[0] = self.f32("v.class_embd")
        x[1:] = y
        x = x + self.f32("v.position_embd.weight")                                             # clip.cpp:1325-1331
        x = self.ln(x, "v.pre_ln.weight", "v.pre_ln.bias", self.v_eps)
        x = self._blocks("v", x, self.v_layers, self.v_heads, self.v_eps, n * n + 1, False)
        cls = self.ln(x[0:1], "v.post_ln.weight", "v.post_ln.bias", self.v_eps)
        e = self.mul_mat("visual_projection.weight", cls)[0]
        return self._normalize(e) if normalize else e

    def encode_text(self, ids, normalize: bool = True) -> np.ndarray:
        ids = np.asarray(ids, np.int64)
        N = ids.size
        tok = self.f32("t.token_embd.weight")[ids]                                             # get_rows: dequantised rows
        x = self.f32("t.position_embd.weight")[:N] + tok                                       # clip.cpp:1059-1061
        x = self._blocks("t", x, self.t_layers, self.t_heads, self.t_eps, N, True)
        x = self.ln(x, "t.post_ln.weight", "t.post_ln.bias", self.t_eps)                       # all tokens, then row N-1
        e = self.mul_mat("text_projection.weight", x[N - 1:N])[0]
        return self._normalize(e) if normalize else e


# ------------------------------------------------------------------------------------------------
# helpers for tests
# ------------------------------------------------------------------------------------------------
def quantize_rows(typ: int, w: np.ndarray) -> bytes:
    """Quantise a [rows, k] float32 matrix with the reference quantizer rows (ggml.c:914-1116)."""
    w = np.ascontiguousarray(w, np.float32)
    rows, k = w.shape
    rs = lib().orc_row_size(typ, k)
    out = np.empty(rows * rs, np.uint8)
    for r in range(rows):
        rc = lib().orc_quantize_row(typ, _fp(w[r]), out.ctypes.data + r * rs, k)
        assert rc == 0
    return out.tobytes()


def cos(a, b) -> float:
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
