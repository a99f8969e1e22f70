"""Generates tests/golden/jpeg/*.jpg and manifest.json: small JPEG files of every layout the decoder in clip.cpp_b200/csrc/jpeg.cpp
claims, each with the sha256 of the RGB pixels the REFERENCE ITSELF (oracle/_ref/libclip_ref.so: clip_image_load_from_file ->
stb_image, clip.cpp:709-726) decodes it to.  Run in the build container (needs the reference library and PIL):

    python tests/golden/make_jpeg_golden.py

Two producers:
  * PIL / libjpeg for the layouts it can write: 4:4:4 / 4:2:2 / 4:2:0, baseline and progressive (spectral selection + successive
    approximation), optimised Huffman tables, restart intervals, greyscale, Adobe CMYK;
  * `encode()` below, a plain baseline encoder (numpy DCT, the Annex-K Huffman tables lifted from a libjpeg file), for what libjpeg
    does not emit: 4:4:0, 4:1:1, 4x2 and mixed ratios, luma sampled below chroma, non-interleaved sequential scans, 16-bit
    quantisation tables, 'R','G','B' component ids, Adobe transform 0 / 2 with 3 / 4 components, fill bytes and comments.
The test (tests/test_host_side.py) only reads the committed files and the manifest.
"""
import ctypes as C
import hashlib
import io
import json
import os
import sys

import numpy as np
from PIL import Image
from scipy.fft import dctn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "jpeg")
sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import binding as bd          # noqa: E402
import ref_run                # noqa: E402

ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56,
          57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def picture(w, h, kind, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    if kind == "noise":
        a = rng.integers(0, 256, (h, w, 3))
    elif kind == "smooth":
        a = np.stack([x * 255 // max(w - 1, 1), y * 255 // max(h - 1, 1), (x + y) * 127 // max(w + h - 2, 1)], -1)
    else:
        a = np.stack([np.sin(x / 3.0) * 127 + 128, np.cos(y / 5.0) * 127 + 128, ((x // 4 + y // 4) % 2) * 255], -1)
    return a.astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------------
# libjpeg's default tables, read back from a file it wrote
# ---------------------------------------------------------------------------------------------------------------------------
def segments(data):
    p = 2
    while p + 4 <= len(data) and data[p] == 0xFF:
        m, ln = data[p + 1], (data[p + 2] << 8) | data[p + 3]
        yield m, data[p + 4:p + 2 + ln]
        if m == 0xDA:
            return
        p += 2 + ln


def libjpeg_tables():
    buf = io.BytesIO()
    Image.fromarray(picture(16, 16, "noise")).save(buf, "JPEG", quality=75)
    huff, quant = {}, {}
    for m, body in segments(buf.getvalue()):
        while m == 0xC4 and body:
            n = sum(body[1:17])
            huff[(body[0] >> 4, body[0] & 15)] = (list(body[1:17]), list(body[17:17 + n]))
            body = body[17 + n:]
        while m == 0xDB and body:
            quant[body[0] & 15] = list(body[1:65])          # zig-zag order, 8-bit
            body = body[65:]
    assert len(huff) == 4 and len(quant) == 2
    return huff, quant


def code_map(bits, vals):
    out, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            out[vals[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    return out


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value, length):
        self.acc = (self.acc << length) | (value & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def seg(marker, body):
    return bytes([0xFF, marker]) + (len(body) + 2).to_bytes(2, "big") + bytes(body)


def encode(planes, samp, *, ids=None, tq=None, qtabs=None, interleaved=True, dri=0, jfif=True, adobe=None, wide_q=False, junk=False):
    """Baseline sequential JPEG.  planes: full-resolution float arrays, one per stored component (already in the stored colour space);
    samp: [(h, v)] per component."""
    huff, quant = libjpeg_tables()
    n = len(planes)
    H, W = planes[0].shape
    ids = ids or list(range(1, n + 1))
    tq = tq or [0] + [1] * (n - 1)
    qtabs = qtabs or quant
    hmax, vmax = max(h for h, _ in samp), max(v for _, v in samp)
    mx, my = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    blocks = []                                             # per component: [by][bx] -> 64 quantised coefficients (zig-zag)
    for c in range(n):
        h, v = samp[c]
        fx, fy = hmax // h, vmax // v
        px, py = -(-W * h // hmax), -(-H * v // vmax)
        src = np.pad(planes[c].astype(np.float64), ((0, py * fy - H), (0, px * fx - W)), mode="edge")
        low = src.reshape(py, fy, px, fx).mean(axis=(1, 3))
        low = np.pad(low, ((0, my * v * 8 - py), (0, mx * h * 8 - px)), mode="edge")
        q = np.array(qtabs[tq[c]], dtype=np.float64)
        rows = []
        for by in range(my * v):
            row = []
            for bx in range(mx * h):
                coef = dctn(low[8 * by:8 * by + 8, 8 * bx:8 * bx + 8] - 128.0, norm="ortho").reshape(64)
                row.append(np.rint(coef[ZIGZAG] / q).astype(int))
            rows.append(row)
        blocks.append((rows, px, py))
    dc_code = [code_map(*huff[(0, 0)])] + [code_map(*huff[(0, 1)])] * (n - 1)
    ac_code = [code_map(*huff[(1, 0)])] + [code_map(*huff[(1, 1)])] * (n - 1)

    out = bytearray(b"\xff\xd8")
    if jfif:
        out += seg(0xE0, b"JFIF\0\x01\x01\0\0\x01\0\x01\0\0")
    if adobe is not None:
        out += seg(0xEE, b"Adobe\0\x64\0\0\0\0" + bytes([adobe]))
    if junk:
        out += seg(0xFE, b"fixture written by tests/golden/make_jpeg_golden.py") + b"\xff\xff\xff"     # comment, then fill bytes before the next marker
    for t in sorted(set(tq)):
        if wide_q:
            out += seg(0xDB, bytes([0x10 | t]) + b"".join(int(x).to_bytes(2, "big") for x in qtabs[t]))
        else:
            out += seg(0xDB, bytes([t]) + bytes(qtabs[t]))
    out += seg(0xC0, bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([n]) +
               b"".join(bytes([ids[c], (samp[c][0] << 4) | samp[c][1], tq[c]]) for c in range(n)))
    for (cls, t), (bits, vals) in sorted(huff.items()):
        out += seg(0xC4, bytes([(cls << 4) | t]) + bytes(bits) + bytes(vals))
    if dri:
        out += seg(0xDD, dri.to_bytes(2, "big"))

    def put_block(bw, c, zz, pred):
        diff = int(zz[0]) - pred[c]
        pred[c] = int(zz[0])
        size = abs(diff).bit_length()
        bw.put(*dc_code[c][size])
        if size:
            bw.put(diff if diff > 0 else diff + (1 << size) - 1, size)
        run = 0
        last = max([k for k in range(1, 64) if zz[k]], default=0)
        for k in range(1, last + 1):
            if zz[k] == 0:
                run += 1
                continue
            while run > 15:
                bw.put(*ac_code[c][0xF0])
                run -= 16
            val = int(zz[k])
            size = abs(val).bit_length()
            bw.put(*ac_code[c][(run << 4) | size])
            bw.put(val if val > 0 else val + (1 << size) - 1, size)
            run = 0
        if last < 63:
            bw.put(*ac_code[c][0])

    def scan(comps, units):
        """units: list of MCUs, each a list of (component, by, bx)."""
        nonlocal out
        out += seg(0xDA, bytes([len(comps)]) + b"".join(bytes([ids[c], 0x00 if c == 0 else 0x11]) for c in comps) + bytes([0, 63, 0]))
        bw, pred, rst = BitWriter(), [0] * n, 0
        for i, unit in enumerate(units):
            if dri and i and i % dri == 0:
                bw.flush()
                bw.out += bytes([0xFF, 0xD0 + rst])
                rst = (rst + 1) & 7
                pred = [0] * n
            for c, by, bx in unit:
                put_block(bw, c, blocks[c][0][by][bx], pred)
        bw.flush()
        out += bw.out

    if interleaved and n > 1:
        units = [[(c, j * samp[c][1] + y, i * samp[c][0] + x) for c in range(n) for y in range(samp[c][1]) for x in range(samp[c][0])]
                 for j in range(my) for i in range(mx)]
        scan(list(range(n)), units)
    else:
        for c in range(n):
            _, px, py = blocks[c]
            scan([c], [[(c, by, bx)] for by in range(-(-py // 8)) for bx in range(-(-px // 8))])
    out += b"\xff\xd9"
    return bytes(out)


def ycc(rgb):
    r, g, b = [rgb[..., i].astype(np.float64) for i in range(3)]
    return [0.299 * r + 0.587 * g + 0.114 * b, 128 - 0.168736 * r - 0.331264 * g + 0.5 * b, 128 + 0.5 * r - 0.418688 * g - 0.081312 * b]


def pil_bytes(im, **kw):
    buf = io.BytesIO()
    im.save(buf, "JPEG", **kw)
    return buf.getvalue()


def cases():
    P = lambda w, h, k, s=0: Image.fromarray(picture(w, h, k, s))                      # noqa: E731
    planes = lambda w, h, k, s=0: ycc(picture(w, h, k, s))                             # noqa: E731
    rgbp = lambda w, h, k, s=0: [picture(w, h, k, s)[..., i].astype(np.float64) for i in range(3)]   # noqa: E731
    yield "pil_444_base", pil_bytes(P(37, 29, "pattern"), quality=85, subsampling="4:4:4")
    yield "pil_422_base_opt", pil_bytes(P(41, 30, "noise", 1), quality=60, subsampling="4:2:2", optimize=True)
    yield "pil_420_base", pil_bytes(P(50, 35, "smooth"), quality=90, subsampling="4:2:0")
    yield "pil_420_1x1", pil_bytes(P(1, 1, "noise", 2), quality=75, subsampling="4:2:0")
    yield "pil_420_17x33", pil_bytes(P(17, 33, "pattern"), quality=75, subsampling="4:2:0")
    yield "pil_444_prog", pil_bytes(P(37, 29, "noise", 3), quality=70, subsampling="4:4:4", progressive=True)
    yield "pil_422_prog", pil_bytes(P(33, 18, "pattern"), quality=80, subsampling="4:2:2", progressive=True)
    yield "pil_420_prog", pil_bytes(P(65, 49, "noise", 4), quality=50, subsampling="4:2:0", progressive=True)
    yield "pil_420_prog_q95", pil_bytes(P(40, 40, "pattern"), quality=95, subsampling="4:2:0", progressive=True, optimize=True)
    yield "pil_gray", pil_bytes(P(45, 31, "pattern").convert("L"), quality=80)
    yield "pil_gray_prog", pil_bytes(P(45, 31, "noise", 5).convert("L"), quality=60, progressive=True)
    yield "pil_cmyk", pil_bytes(P(30, 21, "pattern").convert("CMYK"), quality=85)
    yield "pil_cmyk_prog_420", pil_bytes(P(33, 40, "smooth").convert("CMYK"), quality=75, progressive=True, subsampling="4:2:0")
    yield "pil_rst_base", pil_bytes(P(72, 40, "pattern"), quality=75, subsampling="4:2:0", restart_marker_blocks=2)
    yield "pil_rst_prog", pil_bytes(P(56, 40, "noise", 6), quality=40, subsampling="4:2:0", progressive=True, restart_marker_rows=1)
    yield "pil_q1", pil_bytes(P(32, 32, "noise", 7), quality=1)
    yield "pil_q100_420", pil_bytes(P(32, 24, "noise", 8), quality=100, subsampling="4:2:0")
    yield "enc_440", encode(planes(29, 37, "pattern"), [(1, 2), (1, 1), (1, 1)])
    yield "enc_411", encode(planes(45, 20, "noise", 9), [(4, 1), (1, 1), (1, 1)])
    yield "enc_4x2", encode(planes(50, 30, "pattern"), [(4, 2), (1, 1), (1, 1)])                 # T.81 B.2.3: at most 10 blocks per MCU
    yield "enc_mixed_v2_h2", encode(planes(39, 30, "pattern"), [(2, 2), (2, 1), (1, 2)])         # Cb 1x2 (vertical only), Cr 2x1 (horizontal only)
    yield "enc_2x4", encode(planes(21, 43, "smooth"), [(2, 4), (1, 1), (1, 1)])
    yield "enc_luma_low", encode(planes(35, 27, "noise", 10), [(1, 1), (2, 2), (2, 2)])
    yield "enc_420_noninterleaved_dri", encode(planes(41, 37, "pattern"), [(2, 2), (1, 1), (1, 1)], interleaved=False, dri=5)
    yield "enc_422_w9", encode(planes(9, 9, "noise", 11), [(2, 1), (1, 1), (1, 1)], dri=1)
    yield "enc_wide_q_junk", encode(planes(24, 24, "pattern"), [(2, 2), (1, 1), (1, 1)], wide_q=True, junk=True,
                                    qtabs={0: [300] + [3] * 63, 1: [17] * 64})
    yield "enc_rgb_ids", encode(rgbp(26, 19, "pattern"), [(1, 1)] * 3, ids=[ord("R"), ord("G"), ord("B")], tq=[0, 0, 0])
    yield "enc_adobe0_rgb", encode(rgbp(26, 19, "noise", 12), [(1, 1)] * 3, jfif=False, adobe=0, tq=[0, 0, 0])
    yield "enc_adobe0_jfif_ycc", encode(planes(26, 19, "pattern"), [(2, 1), (1, 1), (1, 1)], jfif=True, adobe=0)
    k = np.full((19, 26), 200.0)
    yield "enc_ycck", encode(planes(26, 19, "pattern") + [k], [(2, 2), (1, 1), (1, 1), (2, 2)], jfif=False, adobe=2, tq=[0, 1, 1, 0])
    yield "enc_4comp_plain", encode(planes(26, 19, "smooth") + [k], [(1, 1)] * 4, tq=[0, 1, 1, 0])


def load_rgb(lib, path):
    lib.lib.clip_image_load_from_file.restype = C.c_bool
    lib.lib.clip_image_load_from_file.argtypes = [C.c_char_p, C.POINTER(bd.clip_image_u8)]
    im = bd.clip_image_u8()
    if not lib.lib.clip_image_load_from_file(path.encode(), C.byref(im)):
        return None
    return np.ctypeslib.as_array(im.data, shape=(im.ny, im.nx, 3)).copy()


def main():
    assert ref_run.available(), "build oracle/_ref first (make -C oracle ref)"
    ref = bd.ClipLib(ref_run.REF_LIB)
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, data in cases():
        path = os.path.join(OUT, name + ".jpg")
        with open(path, "wb") as f:
            f.write(data)
        rgb = load_rgb(ref, path)
        assert rgb is not None, f"the reference refuses {name}"
        try:
            pil = np.asarray(Image.open(path).convert("RGB")).astype(int)
        except OSError:
            pil = None                                     # libjpeg is stricter than stb_image about a few legal-but-unusual layouts
        manifest[name] = {"bytes": len(data), "nx": int(rgb.shape[1]), "ny": int(rgb.shape[0]), "sha256": hashlib.sha256(rgb.tobytes()).hexdigest(),
                          "mean": round(float(rgb.mean()), 4)}
        print(f"{name:28s} {len(data):6d} B  {rgb.shape[1]}x{rgb.shape[0]}  max |ref - libjpeg| = {'libjpeg refuses' if pil is None else np.abs(pil - rgb).max()}")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump({"source": "oracle/_ref/libclip_ref.so clip_image_load_from_file (stb_image, 3 channels)", "files": manifest}, f, indent=1, sort_keys=True)
    print("total", sum(v["bytes"] for v in manifest.values()), "bytes in", len(manifest), "files")


if __name__ == "__main__":
    main()
