"""ctypes view of the clip.cpp C ABI (clip.h:8-113 in the reference).

The same `ClipLib` class drives either shared library, because the product keeps the reference's
22 exported symbols and struct layouts byte for byte:

  * `libclip_b200.so`        -- the B200-native product (this repository)
  * the unmodified reference built under oracle/ (path: oracle/ref_run.py REF_LIB) -- tests / bench use it as the checker

It mirrors the shape of the reference's own Python binding
(/root/reference/examples/python_bindings/clip_cpp/clip.py:36-208) but carries the `size` members of
clip_image_u8 / clip_image_f32 that the upstream binding forgot (clip.h:50-64).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class clip_text_hparams(C.Structure):
    _fields_ = [("n_vocab", C.c_int32), ("num_positions", C.c_int32), ("hidden_size", C.c_int32),
                ("n_intermediate", C.c_int32), ("projection_dim", C.c_int32), ("n_head", C.c_int32),
                ("n_layer", C.c_int32), ("eps", C.c_float)]


class clip_vision_hparams(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("patch_size", C.c_int32), ("hidden_size", C.c_int32),
                ("n_intermediate", C.c_int32), ("projection_dim", C.c_int32), ("n_head", C.c_int32),
                ("n_layer", C.c_int32), ("eps", C.c_float)]


class clip_tokens(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_int32)), ("size", C.c_size_t)]


class clip_image_u8(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("data", C.POINTER(C.c_uint8)), ("size", C.c_size_t)]


class clip_image_f32(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("data", C.POINTER(C.c_float)), ("size", C.c_size_t)]


class clip_image_u8_batch(C.Structure):
    _fields_ = [("data", C.POINTER(clip_image_u8)), ("size", C.c_size_t)]


class clip_image_f32_batch(C.Structure):
    _fields_ = [("data", C.POINTER(clip_image_f32)), ("size", C.c_size_t)]


REFERENCE_SYMBOLS = [
    "clip_model_load", "clip_free", "clip_get_text_hparams", "clip_get_vision_hparams", "clip_tokenize",
    "clip_image_u8_make", "clip_image_f32_make", "clip_image_u8_clean", "clip_image_f32_clean",
    "clip_image_u8_free", "clip_image_f32_free", "clip_image_load_from_file", "clip_image_preprocess",
    "clip_text_encode", "clip_image_encode", "clip_image_batch_preprocess", "clip_image_batch_encode",
    "clip_compare_text_and_image", "clip_similarity_score", "softmax_with_sorting",
    "clip_zero_shot_label_image", "clip_model_quantize",
]

# additive B200 entry points (include/clip_b200.h); absent from the reference library
EXTENSION_SYMBOLS = [
    "clip_text_batch_encode", "clip_b200_image_encode_device", "clip_b200_text_encode_device",
    "clip_b200_zero_shot_batch", "clip_b200_device_malloc", "clip_b200_device_free",
    "clip_b200_host_malloc", "clip_b200_host_free", "clip_b200_memcpy_h2d", "clip_b200_memcpy_d2h",
    "clip_b200_synchronize", "clip_b200_last_error", "clip_b200_kernel_launches",
    "clip_b200_last_device_ms", "clip_b200_version", "clip_b200_set_micro_batch",
    "clip_b200_debug_gemm", "clip_b200_debug_attention", "clip_b200_image_batch_encode_u8",
    "clip_b200_image_batch_preprocess_device", "clip_b200_get_stream", "clip_b200_kernel_ms",
    "clip_b200_debug_repack_roundtrip", "clip_b200_debug_tokenize", "clip_b200_debug_preprocess",
    "clip_b200_mark", "clip_b200_mark_elapsed_ms",
    # scoring + multi-GPU (round 2)
    "clip_b200_topk_search", "clip_b200_zero_shot_images", "clip_b200_dist_unique_id", "clip_b200_dist_init_with_id",
    "clip_b200_dist_init", "clip_b200_dist_rank", "clip_b200_dist_world", "clip_b200_device_count", "clip_b200_nccl_version", "clip_b200_cuda_device_count",
    "clip_b200_dist_barrier", "clip_b200_dist_max_f64", "clip_b200_dist_all_gather", "clip_b200_image_encode_device_all",
    "clip_b200_text_encode_device_all", "clip_b200_image_batch_encode_all", "clip_b200_text_batch_encode_all",
    "clip_b200_debug_rendezvous", "clip_b200_debug_shard_bounds",
]

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get("CLIP_B200_LIB") or os.path.join(HERE, "libclip_b200.so")     # CLIP_B200_LIB: experiment builds only


class ClipLib:
    """Thin, explicit ctypes wrapper. `path` picks the implementation."""

    def __init__(self, path: str = PRODUCT_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not built -- run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        self.path = path
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.is_product = hasattr(L, "clip_b200_version")
        vp, ip, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
        L.clip_model_load.restype = vp
        L.clip_model_load.argtypes = [C.c_char_p, ip]
        L.clip_free.argtypes = [vp]
        L.clip_free.restype = None
        L.clip_get_text_hparams.restype = C.POINTER(clip_text_hparams)
        L.clip_get_text_hparams.argtypes = [vp]
        L.clip_get_vision_hparams.restype = C.POINTER(clip_vision_hparams)
        L.clip_get_vision_hparams.argtypes = [vp]
        L.clip_tokenize.restype = C.c_bool
        L.clip_tokenize.argtypes = [vp, C.c_char_p, C.POINTER(clip_tokens)]
        L.clip_image_preprocess.restype = C.c_bool
        L.clip_image_preprocess.argtypes = [vp, C.POINTER(clip_image_u8), C.POINTER(clip_image_f32)]
        L.clip_image_batch_preprocess.restype = None
        L.clip_image_batch_preprocess.argtypes = [vp, ip, C.POINTER(clip_image_u8_batch),
                                                  C.POINTER(clip_image_f32_batch)]
        L.clip_image_f32_clean.argtypes = [C.POINTER(clip_image_f32)]
        L.clip_image_f32_clean.restype = None
        L.clip_text_encode.restype = C.c_bool
        L.clip_text_encode.argtypes = [vp, ip, C.POINTER(clip_tokens), fp, C.c_bool]
        L.clip_image_encode.restype = C.c_bool
        L.clip_image_encode.argtypes = [vp, ip, C.POINTER(clip_image_f32), fp, C.c_bool]
        L.clip_image_batch_encode.restype = C.c_bool
        L.clip_image_batch_encode.argtypes = [vp, ip, C.POINTER(clip_image_f32_batch), fp, C.c_bool]
        L.clip_compare_text_and_image.restype = C.c_bool
        L.clip_compare_text_and_image.argtypes = [vp, ip, C.c_char_p, C.POINTER(clip_image_u8), fp]
        L.clip_similarity_score.restype = C.c_float
        L.clip_similarity_score.argtypes = [fp, fp, ip]
        L.softmax_with_sorting.restype = C.c_bool
        L.softmax_with_sorting.argtypes = [fp, ip, fp, C.POINTER(C.c_int)]
        L.clip_zero_shot_label_image.restype = C.c_bool
        L.clip_zero_shot_label_image.argtypes = [vp, ip, C.POINTER(clip_image_u8), C.POINTER(C.c_char_p),
                                                 C.c_size_t, fp, C.POINTER(C.c_int)]
        L.clip_model_quantize.restype = C.c_bool
        L.clip_model_quantize.argtypes = [C.c_char_p, C.c_char_p, ip]
        if self.is_product:
            L.clip_b200_version.restype = C.c_char_p
            L.clip_b200_last_error.restype = C.c_char_p
            L.clip_text_batch_encode.restype = C.c_bool
            L.clip_text_batch_encode.argtypes = [vp, ip, C.POINTER(clip_tokens), C.c_size_t, fp, C.c_bool]
            L.clip_b200_image_encode_device.restype = C.c_bool
            L.clip_b200_image_encode_device.argtypes = [vp, vp, C.c_size_t, vp, C.c_bool]
            L.clip_b200_text_encode_device.restype = C.c_bool
            L.clip_b200_text_encode_device.argtypes = [vp, vp, vp, C.c_size_t, ip, vp, C.c_bool]
            L.clip_b200_zero_shot_batch.restype = C.c_bool
            L.clip_b200_zero_shot_batch.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, fp, C.POINTER(C.c_int), ip]
            L.clip_b200_device_malloc.restype = vp
            L.clip_b200_device_malloc.argtypes = [vp, C.c_size_t]
            L.clip_b200_device_free.argtypes = [vp, vp]
            L.clip_b200_device_free.restype = None
            L.clip_b200_host_malloc.restype = vp
            L.clip_b200_host_malloc.argtypes = [C.c_size_t]
            L.clip_b200_host_free.argtypes = [vp]
            L.clip_b200_host_free.restype = None
            L.clip_b200_memcpy_h2d.restype = C.c_bool
            L.clip_b200_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
            L.clip_b200_memcpy_d2h.restype = C.c_bool
            L.clip_b200_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
            L.clip_b200_synchronize.restype = C.c_bool
            L.clip_b200_synchronize.argtypes = [vp]
            L.clip_b200_kernel_launches.restype = C.c_uint64
            L.clip_b200_kernel_launches.argtypes = [vp]
            L.clip_b200_last_device_ms.restype = C.c_float
            L.clip_b200_last_device_ms.argtypes = [vp]
            L.clip_b200_set_micro_batch.restype = None
            L.clip_b200_set_micro_batch.argtypes = [vp, ip, ip]
            L.clip_b200_get_stream.restype = vp
            L.clip_b200_get_stream.argtypes = [vp]
            L.clip_b200_kernel_ms.restype = C.c_float
            L.clip_b200_kernel_ms.argtypes = [vp, ip, C.POINTER(C.c_uint64)]
            L.clip_b200_mark.restype = C.c_bool
            L.clip_b200_mark.argtypes = [vp, ip]
            L.clip_b200_mark_elapsed_ms.restype = C.c_float
            L.clip_b200_mark_elapsed_ms.argtypes = [vp, ip, ip]
            L.clip_b200_debug_repack_roundtrip.restype = C.c_int
            L.clip_b200_debug_repack_roundtrip.argtypes = [ip, vp, ip, ip]
            L.clip_b200_debug_tokenize.restype = C.c_int
            L.clip_b200_debug_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), ip]
            L.clip_b200_debug_preprocess.restype = C.c_int
            L.clip_b200_debug_preprocess.argtypes = [C.POINTER(C.c_uint8), ip, ip, ip, fp, fp, fp]
            L.clip_b200_image_batch_encode_u8.restype = C.c_bool
            L.clip_b200_image_batch_encode_u8.argtypes = [vp, C.POINTER(clip_image_u8_batch), fp, C.c_bool]
            L.clip_b200_image_batch_preprocess_device.restype = C.c_bool
            L.clip_b200_image_batch_preprocess_device.argtypes = [vp, C.POINTER(clip_image_u8_batch), C.POINTER(clip_image_f32_batch)]
            L.clip_b200_debug_attention.restype = C.c_int
            L.clip_b200_debug_attention.argtypes = [ip, ip, ip, ip, ip, ip, fp, fp, C.POINTER(C.c_float)]
            L.clip_b200_debug_gemm.restype = C.c_int
            L.clip_b200_debug_gemm.argtypes = [ip, ip, ip, ip, ip, ip, ip, fp, vp, fp, fp, fp, C.POINTER(C.c_float)]
            ipp = C.POINTER(C.c_int)
            L.clip_b200_topk_search.restype = C.c_bool
            L.clip_b200_topk_search.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, ip, fp, ipp]
            L.clip_b200_zero_shot_images.restype = C.c_bool
            L.clip_b200_zero_shot_images.argtypes = [vp, ip, C.POINTER(clip_image_f32_batch), C.POINTER(clip_tokens), C.c_size_t,
                                                     C.c_bool, ip, fp, ipp]
            L.clip_b200_dist_unique_id.restype = C.c_bool
            L.clip_b200_dist_unique_id.argtypes = [vp]
            L.clip_b200_dist_init_with_id.restype = C.c_bool
            L.clip_b200_dist_init_with_id.argtypes = [vp, ip, ip, vp]
            L.clip_b200_dist_init.restype = C.c_bool
            L.clip_b200_dist_init.argtypes = [vp, ip, ip, C.c_char_p]
            for f in ("clip_b200_dist_rank", "clip_b200_dist_world", "clip_b200_device_count"):
                getattr(L, f).restype = C.c_int
                getattr(L, f).argtypes = [vp]
            L.clip_b200_nccl_version.restype = C.c_int
            L.clip_b200_nccl_version.argtypes = []
            L.clip_b200_cuda_device_count.restype = C.c_int
            L.clip_b200_cuda_device_count.argtypes = []
            L.clip_b200_debug_rendezvous.restype = C.c_int
            L.clip_b200_debug_rendezvous.argtypes = [ip, ip, C.c_char_p, vp]
            L.clip_b200_debug_shard_bounds.restype = None
            L.clip_b200_debug_shard_bounds.argtypes = [C.c_size_t, ip, ip, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
            L.clip_b200_dist_barrier.restype = C.c_bool
            L.clip_b200_dist_barrier.argtypes = [vp]
            L.clip_b200_dist_max_f64.restype = C.c_bool
            L.clip_b200_dist_max_f64.argtypes = [vp, C.POINTER(C.c_double), ip]
            L.clip_b200_dist_all_gather.restype = C.c_bool
            L.clip_b200_dist_all_gather.argtypes = [vp, vp, vp, C.c_size_t]
            L.clip_b200_image_encode_device_all.restype = C.c_bool
            L.clip_b200_image_encode_device_all.argtypes = [vp, vp, C.c_size_t, vp, C.c_bool]
            L.clip_b200_text_encode_device_all.restype = C.c_bool
            L.clip_b200_text_encode_device_all.argtypes = [vp, vp, vp, C.c_size_t, ip, vp, C.c_bool]
            L.clip_b200_image_batch_encode_all.restype = C.c_bool
            L.clip_b200_image_batch_encode_all.argtypes = [vp, ip, C.POINTER(clip_image_f32_batch), fp, C.c_bool]
            L.clip_b200_text_batch_encode_all.restype = C.c_bool
            L.clip_b200_text_batch_encode_all.argtypes = [vp, ip, C.POINTER(clip_tokens), C.c_size_t, fp, C.c_bool]

    # ---- model -----------------------------------------------------------------------------------
    def load(self, path: str, verbosity: int = 0):
        ctx = self.lib.clip_model_load(path.encode(), verbosity)
        if not ctx:
            raise RuntimeError("clip_model_load(%s) failed: %s" % (path, self.last_error()))
        return ctx

    def free(self, ctx):
        self.lib.clip_free(ctx)

    def last_error(self) -> str:
        if self.is_product:
            return self.lib.clip_b200_last_error().decode()
        return ""

    def vision_hparams(self, ctx) -> clip_vision_hparams:
        return self.lib.clip_get_vision_hparams(ctx).contents

    def text_hparams(self, ctx) -> clip_text_hparams:
        return self.lib.clip_get_text_hparams(ctx).contents

    def quantize(self, src: str, dst: str, itype: int) -> bool:
        return bool(self.lib.clip_model_quantize(src.encode(), dst.encode(), itype))

    # ---- encoders (host buffers; the reference-facing calls) ---------------------------------------
    @staticmethod
    def make_image_batch(images: np.ndarray):
        """images: [n, S, S, 3] float32 C-contiguous -> (batch struct, keep-alive tuple)."""
        assert images.dtype == np.float32 and images.flags["C_CONTIGUOUS"] and images.ndim == 4
        n, s = images.shape[0], images.shape[1]
        arr = (clip_image_f32 * n)()
        per = s * s * 3
        base = images.ctypes.data
        for i in range(n):
            arr[i].nx = s
            arr[i].ny = s
            arr[i].size = per
            arr[i].data = C.cast(base + i * per * 4, C.POINTER(C.c_float))
        batch = clip_image_f32_batch(arr, n)
        return batch, (arr, images)

    def image_batch_encode(self, ctx, images: np.ndarray, normalize=True, n_threads=4) -> np.ndarray:
        d = self.vision_hparams(ctx).projection_dim
        batch, keep = self.make_image_batch(images)
        out = np.empty((images.shape[0], d), np.float32)
        ok = self.lib.clip_image_batch_encode(ctx, n_threads, C.byref(batch),
                                              out.ctypes.data_as(C.POINTER(C.c_float)), normalize)
        if not ok:
            raise RuntimeError("clip_image_batch_encode failed: " + self.last_error())
        return out

    def image_encode(self, ctx, image: np.ndarray, normalize=True, n_threads=4) -> np.ndarray:
        assert image.dtype == np.float32 and image.ndim == 3 and image.flags["C_CONTIGUOUS"]
        d = self.vision_hparams(ctx).projection_dim
        s = image.shape[0]
        img = clip_image_f32(s, s, image.ctypes.data_as(C.POINTER(C.c_float)), s * s * 3)
        out = np.empty((d,), np.float32)
        ok = self.lib.clip_image_encode(ctx, n_threads, C.byref(img), out.ctypes.data_as(C.POINTER(C.c_float)),
                                        normalize)
        if not ok:
            raise RuntimeError("clip_image_encode failed: " + self.last_error())
        return out

    def text_encode(self, ctx, ids: np.ndarray, normalize=True, n_threads=4) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int32)
        d = self.text_hparams(ctx).projection_dim
        tk = clip_tokens(ids.ctypes.data_as(C.POINTER(C.c_int32)), ids.size)
        out = np.empty((d,), np.float32)
        ok = self.lib.clip_text_encode(ctx, n_threads, C.byref(tk), out.ctypes.data_as(C.POINTER(C.c_float)),
                                       normalize)
        if not ok:
            raise RuntimeError("clip_text_encode failed: " + self.last_error())
        return out

    def text_batch_encode(self, ctx, seqs, normalize=True, n_threads=4) -> np.ndarray:
        """seqs: list of int32 arrays (ragged).  Product-only extension (include/clip_b200.h)."""
        seqs = [np.ascontiguousarray(s, np.int32) for s in seqs]
        d = self.text_hparams(ctx).projection_dim
        arr = (clip_tokens * len(seqs))()
        for i, s in enumerate(seqs):
            arr[i].data = s.ctypes.data_as(C.POINTER(C.c_int32))
            arr[i].size = s.size
        out = np.empty((len(seqs), d), np.float32)
        ok = self.lib.clip_text_batch_encode(ctx, n_threads, arr, len(seqs),
                                             out.ctypes.data_as(C.POINTER(C.c_float)), normalize)
        if not ok:
            raise RuntimeError("clip_text_batch_encode failed: " + self.last_error())
        return out

    @staticmethod
    def make_token_array(seqs):
        seqs = [np.ascontiguousarray(s, np.int32) for s in seqs]
        arr = (clip_tokens * max(len(seqs), 1))()
        for i, s in enumerate(seqs):
            arr[i].data = s.ctypes.data_as(C.POINTER(C.c_int32))
            arr[i].size = s.size
        return arr, seqs

    def zero_shot_images(self, ctx, images: np.ndarray, label_seqs, top_k: int, normalize=False, n_threads=4):
        """extension: batch form of clip_zero_shot_label_image -> (scores [n, k], indices [n, k]); labels are token arrays.
        In ranks mode `label_seqs` is this rank's shard and the result ranks ALL ranks' labels (index = rank * len + j)."""
        batch, keep = self.make_image_batch(images)
        arr, keep2 = self.make_token_array(label_seqs)
        n_all = len(label_seqs) * max(1, self.lib.clip_b200_dist_world(ctx))
        k = n_all if top_k <= 0 or top_k > n_all else top_k
        scores = np.empty((images.shape[0], k), np.float32)
        idx = np.empty((images.shape[0], k), np.int32)
        ok = self.lib.clip_b200_zero_shot_images(ctx, n_threads, C.byref(batch), arr, len(label_seqs), normalize, k,
                                                 scores.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int)))
        if not ok:
            raise RuntimeError("clip_b200_zero_shot_images failed: " + self.last_error())
        return scores, idx

    def compare_text_and_image(self, ctx, text: str, img_u8: np.ndarray, n_threads=4) -> float:
        img_u8 = np.ascontiguousarray(img_u8, np.uint8)
        src = clip_image_u8(img_u8.shape[1], img_u8.shape[0], img_u8.ctypes.data_as(C.POINTER(C.c_uint8)), img_u8.size)
        score = C.c_float(0)
        if not self.lib.clip_compare_text_and_image(ctx, n_threads, text.encode(), C.byref(src), C.byref(score)):
            raise RuntimeError("clip_compare_text_and_image failed: " + self.last_error())
        return float(score.value)

    def zero_shot_label_image(self, ctx, img_u8: np.ndarray, labels, n_threads=4):
        img_u8 = np.ascontiguousarray(img_u8, np.uint8)
        src = clip_image_u8(img_u8.shape[1], img_u8.shape[0], img_u8.ctypes.data_as(C.POINTER(C.c_uint8)), img_u8.size)
        arr = (C.c_char_p * len(labels))(*[l.encode() for l in labels])
        scores = np.empty(len(labels), np.float32)
        idx = np.empty(len(labels), np.int32)
        if not self.lib.clip_zero_shot_label_image(ctx, n_threads, C.byref(src), arr, len(labels),
                                                   scores.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int))):
            raise RuntimeError("clip_zero_shot_label_image failed: " + self.last_error())
        return scores, idx

    def tokenize(self, ctx, text: str) -> np.ndarray:
        tk = clip_tokens()
        if not self.lib.clip_tokenize(ctx, text.encode(), C.byref(tk)):
            raise RuntimeError("clip_tokenize failed")
        return np.ctypeslib.as_array(tk.data, shape=(tk.size,)).copy()

    def _u8_batch(self, images):
        """list of [ny, nx, 3] uint8 arrays -> (clip_image_u8_batch, keep-alive objects)"""
        arrs = [np.ascontiguousarray(im, np.uint8) for im in images]
        items = (clip_image_u8 * len(arrs))()
        for i, a in enumerate(arrs):
            items[i] = clip_image_u8(a.shape[1], a.shape[0], a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size)
        return clip_image_u8_batch(items, len(arrs)), (arrs, items)

    def image_batch_encode_u8(self, ctx, images, normalize=True) -> np.ndarray:
        """extension (N1): raw u8 images of any size -> embeddings; resize / crop / normalise run on the GPU"""
        batch, keep = self._u8_batch(images)
        d = self.vision_hparams(ctx).projection_dim
        out = np.empty((len(images), d), np.float32)
        if not self.lib.clip_b200_image_batch_encode_u8(ctx, C.byref(batch), out.ctypes.data_as(C.POINTER(C.c_float)), normalize):
            raise RuntimeError("clip_b200_image_batch_encode_u8 failed: " + self.last_error())
        return out

    def preprocess_device(self, ctx, images) -> np.ndarray:
        """extension (N1): GPU preprocess alone -> [n, S, S, 3] float32"""
        batch, keep = self._u8_batch(images)
        res = (clip_image_f32 * len(images))()
        outb = clip_image_f32_batch(res, len(images))
        if not self.lib.clip_b200_image_batch_preprocess_device(ctx, C.byref(batch), C.byref(outb)):
            raise RuntimeError("clip_b200_image_batch_preprocess_device failed: " + self.last_error())
        s = res[0].nx
        out = np.stack([np.ctypeslib.as_array(res[i].data, shape=(s, s, 3)).copy() for i in range(len(images))])
        for i in range(len(images)):
            self.lib.clip_image_f32_clean(C.byref(res[i]))
        return out

    def batch_preprocess(self, ctx, images, n_threads: int = 4) -> np.ndarray:
        """clip_image_batch_preprocess (clip.h:95-96): the reference's batch call -> [n, S, S, 3] float32.  Host threads by default,
        the GPU when CLIP_B200_PREPROCESS=device is set in the environment."""
        batch, keep = self._u8_batch(images)
        res = (clip_image_f32 * len(images))()
        outb = clip_image_f32_batch(res, len(images))
        self.lib.clip_image_batch_preprocess(ctx, n_threads, C.byref(batch), C.byref(outb))
        if any(not res[i].data for i in range(len(images))):
            raise RuntimeError("clip_image_batch_preprocess failed: " + self.last_error())
        s = res[0].nx
        out = np.stack([np.ctypeslib.as_array(res[i].data, shape=(s, s, 3)).copy() for i in range(len(images))])
        for i in range(len(images)):
            self.lib.clip_image_f32_clean(C.byref(res[i]))
        return out

    def preprocess(self, ctx, img_u8: np.ndarray) -> np.ndarray:
        """img_u8: [ny, nx, 3] uint8 -> [S, S, 3] float32 (clip.cpp:797-927 semantics)."""
        img_u8 = np.ascontiguousarray(img_u8, np.uint8)
        ny, nx = img_u8.shape[:2]
        src = clip_image_u8(nx, ny, img_u8.ctypes.data_as(C.POINTER(C.c_uint8)), img_u8.size)
        dst = clip_image_f32()
        if not self.lib.clip_image_preprocess(ctx, C.byref(src), C.byref(dst)):
            raise RuntimeError("clip_image_preprocess failed")
        s = dst.nx
        out = np.ctypeslib.as_array(dst.data, shape=(s, s, 3)).copy()
        self.lib.clip_image_f32_clean(C.byref(dst))
        return out
