#!/usr/bin/env python
"""bench.py -- image-embeddings/sec, ViT-L/14 q4_0, b=512 per GPU (BASELINE.json metric), on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3                  # product (libclip_b200.so)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                           # the reference's CPU path on the host cores

A step = one pass of the hot path over one batch of 512 synthetic 224x224x3 images per GPU (weak scaling: the
images shard embarrassingly; the only exchange is ONE NCCL all-gather of the L2-normalised embeddings per step).
  value : whole-job img/s with the pixels already resident in HBM (clip_b200_image_encode_device)
  e2e   : the same metric through the reference-facing call clip_image_batch_encode with pinned HOST buffers,
          H2D of the pixels and D2H of the embeddings inside the timed region
Timing: CUDA events on the library's launch stream (clip_b200_mark), barrier + device sync on both sides, max over
ranks.  Inputs are 308 MB per step (> 126 MB L2), so every step re-reads them from HBM; no separate L2 flush.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "clip.cpp_b200"))

GEOM, FTYPE, SEED, BATCH = "vit-l14", "q4_0", 1234, 512
# algorithmic FLOPs per ViT-L/14 image: total, and the part executed by the fused-dequant GEMM kernel
F_IMG = 162_025_537_536
F_ATTN = 24 * 270_536_704
F_GEMM = F_IMG - F_ATTN
WORKLOAD = "ViT-L/14 q4_0 image encode, b=512 per GPU, 224x224x3 synthetic (configs[1]-style single-GPU run of the metric's config)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_model(lib, rank):
    import synth_gguf as sg
    path = sg.model_path(GEOM, SEED, FTYPE)
    if os.path.exists(path):
        return path
    if rank != 0:
        t0 = time.time()
        while not os.path.exists(path):
            time.sleep(1.0)
            if time.time() - t0 > 1200:
                raise RuntimeError("timed out waiting for rank 0 to write " + path)
        return path
    f16 = sg.model_path(GEOM, SEED, "f16")
    if not os.path.exists(f16):
        t0 = time.time()
        sg.write_model(f16 + ".tmp", sg.GEOMETRIES[GEOM], SEED, 1)
        os.replace(f16 + ".tmp", f16)
        log("bench: wrote %s in %.1fs" % (f16, time.time() - t0))
    t0 = time.time()
    # clip_model_quantize of the product library (byte-identical to the reference's: tests/test_host_side.py)
    assert lib.quantize(f16, path + ".tmp", 2), "quantize failed"
    os.replace(path + ".tmp", path)
    log("bench: quantized to %s in %.1fs" % (path, time.time() - t0))
    return path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0))), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


def cpu_reference_sample(model, n_images, threads):
    """The reference's own CPU implementation (oracle/_ref) on a bounded sample; falls back to the oracle port."""
    import synth_gguf as sg
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_run
    imgs = sg.synth_images(n_images, 224, 4321)
    if ref_run.available():
        r = ref_run.run_reference(model, images=imgs, n_threads=threads)
        return {"value": n_images / float(r["img_s"]), "unit": "img/s", "cores": int(r["threads"]), "kind": "reference",
                "sample": "%d single-image clip_image_encode calls (the reference cannot batch ViT-L/14), oracle/_ref built from the reference sources" % n_images}, r["img"]
    import oracle as orc
    om = orc.OracleModel(model, n_threads=threads)
    t0 = time.perf_counter()
    out = np.stack([om.encode_image(imgs[i]) for i in range(n_images)])
    dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": "img/s", "cores": threads or os.cpu_count(), "kind": "port",
            "sample": "%d images through the CPU oracle restatement (oracle/_ref absent)" % n_images}, out


def pick_threads(model):
    """ggml's spin-wait pool stops scaling long before 128 threads: take the fastest of a few candidates on one image."""
    import synth_gguf as sg
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_run
    n = os.cpu_count() or 8
    if not ref_run.available():
        return n
    img = sg.synth_images(1, 224, 1)
    best, best_t = None, 1e30
    for t in sorted({min(n, c) for c in (8, 16, 32, 64)}):
        r = ref_run.run_reference(model, images=np.concatenate([img, img]), n_threads=t)
        dt = float(r["img_s"])
        log("bench(reference): %d threads -> %.2f s / 2 images" % (t, dt))
        if dt < best_t:
            best, best_t = t, dt
    return best


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import binding as bd
    lib = bd.ClipLib(bd.PRODUCT_LIB) if os.path.exists(bd.PRODUCT_LIB) else None
    assert lib is not None, "libclip_b200.so is needed for clip_model_quantize when creating the synthetic model"
    model = ensure_model(lib, 0)
    threads = pick_threads(model)
    per_step = 4
    for _ in range(max(args.warmup - 1, 0)):       # one calibration pass above already warmed the page cache
        cpu_reference_sample(model, 1, threads)
    vals, base = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        base, _ = cpu_reference_sample(model, per_step, threads)
        vals.append(base["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    base["value"] = v
    out = {"impl": "reference", "metric": "image-embeddings/sec", "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * per_step / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "q4_0 x q8_0 int8 dot, fp32 accumulate (ggml CPU)", "data": "synthetic",
           "config": {"workload": WORKLOAD, "global_batch": BATCH * args.gpus, "step_sample": "%d images per step" % per_step},
           "cpu_baseline": base, "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": wall}
    print(json.dumps(out), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum per launch of gemm_dq_kernel, from the committed `ncu --set full` capture of the four
# layer GEMM shapes (FC2, QKV, out-proj, FC1) at the 256-image micro-batch this workload runs with (65792 token rows):
# (664.1 + 488.2 + 233.7 + 622.8) MB / 4.  It is BELOW the algorithmic bytes (X once + packed W once + Y once = 540.8 MB mean):
# activations written by the previous kernel are partly read from the 126 MB L2 and part of each output is still in L2 when
# the kernel retires.  profiles/r01_gemm_pair_ncu.md has the per-launch table.
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 502_200_000
ALGO_GEMM_BYTES_PER_LAUNCH = 540_800_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-text", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    os.environ["CLIP_B200_DEVICE"] = str(local)
    os.environ["CLIP_B200_PROFILE"] = "1"
    import binding as bd
    import synth_gguf as sg
    lib = bd.ClipLib(bd.PRODUCT_LIB)          # raises if the CUDA library is not built: there is no fallback path
    L = lib.lib

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = ensure_model(lib, rank)
    if dist:
        dist.barrier()
    ctx = lib.load(model, 0)
    d = lib.vision_hparams(ctx).projection_dim
    B = args.batch
    per = 224 * 224 * 3

    # ---- inputs: pinned host batch (e2e) and a device-resident copy (value) -------------------------------------
    h_pix = L.clip_b200_host_malloc(B * per * 4)
    h_out = L.clip_b200_host_malloc(B * d * 4)
    assert h_pix and h_out
    pix = np.ctypeslib.as_array(C.cast(h_pix, C.POINTER(C.c_float)), shape=(B, 224, 224, 3))
    pix[:] = sg.synth_images(B, 224, 1000 + rank)
    out_host = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(B, d))
    batch, keep = lib.make_image_batch(pix)
    d_pix = L.clip_b200_device_malloc(ctx, B * per * 4)
    assert d_pix and L.clip_b200_memcpy_h2d(ctx, d_pix, h_pix, B * per * 4)
    if dist:
        import torch
        t_local = torch.empty((B, d), dtype=torch.float32, device="cuda")
        t_all = torch.empty((B * world, d), dtype=torch.float32, device="cuda")
        d_out = t_local.data_ptr()
    else:
        d_out = L.clip_b200_device_malloc(ctx, B * d * 4)

    def sync_all():
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
        assert L.clip_b200_synchronize(ctx)

    def step_device():
        assert L.clip_b200_image_encode_device(ctx, d_pix, B, d_out, True), lib.last_error()
        if dist:
            dist.all_gather_into_tensor(t_all, t_local)       # the ONE collective of the path (NCCL over NVLink)

    def step_e2e():
        assert L.clip_image_batch_encode(ctx, 1, C.byref(batch), C.cast(h_out, C.POINTER(C.c_float)), True), lib.last_error()
        if dist:
            t_local.copy_(torch.from_numpy(out_host), non_blocking=False)
            dist.all_gather_into_tensor(t_all, t_local)

    def timed(fn, steps):
        sync_all()
        L.clip_b200_mark(ctx, 0)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        if dist:
            torch.cuda.synchronize()
        L.clip_b200_mark(ctx, 1)
        ms_dev = L.clip_b200_mark_elapsed_ms(ctx, 0, 1)
        sync_all()
        ms_wall = (time.perf_counter() - t0) * 1e3
        # events bracket the library's stream; the NCCL all-gather (torch stream) is covered by the wall clock between the
        # two device-wide synchronisations.  Take the larger of the two, then the max over ranks.
        ms = max(ms_dev, ms_wall) if dist else ms_dev
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, ms_wall / steps

    for _ in range(args.warmup):
        step_device()
    for k in range(4):
        L.clip_b200_kernel_ms(ctx, k, None)                     # drop warm-up profile records
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.clip_b200_kernel_launches(ctx)
    ms_step, ms_wall = timed(step_device, args.steps)
    launches = (L.clip_b200_kernel_launches(ctx) - l0) + (args.steps if dist else 0)
    clocks = sampler.stop() if rank == 0 else None
    kcount = C.c_uint64(0)
    kinds = {}
    for k, nm in enumerate(("gemm", "attention", "layernorm", "other")):
        ms = L.clip_b200_kernel_ms(ctx, k, C.byref(kcount))
        kinds[nm] = {"ms_per_step": ms / args.steps, "launches_per_step": kcount.value / args.steps}
    # ---- e2e: host buffers through the reference-facing call ----------------------------------------------------
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, max(2, args.steps // 2))
    for k in range(4):
        L.clip_b200_kernel_ms(ctx, k, None)

    value = B * world / (ms_step / 1e3)
    e2e = B * world / (ms_e2e / 1e3)

    # ---- N1 leg: raw u8 images in, resize + crop + normalise on the GPU, then the same encode ---------------------
    e2e_u8 = None
    if not dist and not args.no_text:
        SRC = 256
        rng8 = np.random.default_rng(7)
        pool = [rng8.integers(0, 256, (SRC, SRC, 3), dtype=np.uint8) for _ in range(16)]
        items = (bd.clip_image_u8 * B)()
        for i in range(B):
            a = pool[i % len(pool)]
            items[i] = bd.clip_image_u8(SRC, SRC, a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size)
        batch8 = bd.clip_image_u8_batch(items, B)

        def step_u8():
            assert L.clip_b200_image_batch_encode_u8(ctx, C.byref(batch8), C.cast(h_out, C.POINTER(C.c_float)), True), lib.last_error()

        step_u8()
        _, ms_u8 = timed(step_u8, 2)
        e2e_u8 = {"value": B / (ms_u8 / 1e3), "unit": "img/s", "ms_per_step": ms_u8, "h2d_bytes_per_step": B * SRC * SRC * 3,
                  "source": "%dx%d u8 RGB per image; resize/crop/normalise on the GPU, bit-identical to clip_image_preprocess; wall clock "
                            "around clip_b200_image_batch_encode_u8 (includes the host copy into pinned staging)" % (SRC, SRC)}
        for k in range(4):
            L.clip_b200_kernel_ms(ctx, k, None)

    # ---- secondary line: text-embeddings/sec, 2048 x 77-token sequences per GPU (BASELINE.json configs[3] shape) -------------
    text = None
    if lib.lib.clip_get_text_hparams(ctx).contents.n_layer > 0 and not args.no_text:
        TB, TL = 2048, 77
        ids = sg.synth_tokens(TB, TL, 2000 + rank)
        d_ids = L.clip_b200_device_malloc(ctx, ids.nbytes)
        d_tout = L.clip_b200_device_malloc(ctx, TB * d * 4)
        assert L.clip_b200_memcpy_h2d(ctx, d_ids, ids.ctypes.data, ids.nbytes)

        def step_text():
            assert L.clip_b200_text_encode_device(ctx, d_ids, None, TB, TL, d_tout, True), lib.last_error()

        for _ in range(3):
            step_text()
        ms_text, _ = timed(step_text, max(2, args.steps))
        for k in range(4):
            L.clip_b200_kernel_ms(ctx, k, None)
        text = {"metric": "text-embeddings/sec", "value": TB * world / (ms_text / 1e3), "unit": "seq/s", "ms_per_step": ms_text,
                "batch_per_gpu": TB, "tokens": TL, "tower": "ViT-L/14 text tower (h=768, 12 layers) q4_0"}
    if rank == 0:
        peak, peak_src = peaks()
        gemm_ms = kinds["gemm"]["ms_per_step"]
        achieved = F_GEMM * B / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        # parity of this very run against the reference-produced golden vectors (first images of the fixture)
        parity = None
        gpath = os.path.join(ROOT, "tests", "golden", "%s-s%d.npz" % (GEOM, SEED))
        if os.path.exists(gpath):
            g = np.load(gpath)
            if "img_" + FTYPE in g.files and str(g["sha_" + FTYPE]) == sg.sha256_file(model):
                gi = sg.synth_images(int(g["n_img"]), 224, int(g["img_seed"]))
                got = lib.image_batch_encode(ctx, gi)
                ref = g["img_" + FTYPE]
                c = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
                parity = {"one_minus_cos_max": float((1 - c).max()), "n": int(len(ref)), "tolerance": 1e-2,
                          "against": "reference ggml CPU embeddings (tests/golden, oracle/_ref)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu, _ = cpu_reference_sample(model, 8, pick_threads(model))
            except Exception as e:           # the baseline leg must never take the GPU number down with it
                cpu = {"value": None, "unit": "img/s", "cores": None, "kind": "unavailable", "sample": str(e)[:200]}
        out = {
            "metric": "image-embeddings/sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (q4_0 blocks unpacked to bf16 in-kernel, fp32 accumulate in TMEM; fp32 residual stream)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "l2": "inputs 308 MB/step > 126 MB L2, no flush needed", "collective": "1 NCCL all-gather of [B,768] f32 per step" if dist else "none (1 GPU)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * per * 4, "d2h_bytes_per_step": B * d * 4, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes/launch (dram read+write, mean of the 4 layer GEMM shapes, "
                         "256-image micro-batch; profiles/r01_gemm_pair_ncu.md) vs %d algorithmic" % ALGO_GEMM_BYTES_PER_LAUNCH,
                         "kernel": "gemm_dq_kernel (all fused-dequant GEMMs of the step)",
                         "flops_per_step": F_GEMM * B, "kernel_ms_per_step": gemm_ms, "peak_source": peak_src},
            "kernel_time_ms_per_step": kinds, "wall_ms_per_step": ms_wall, "parity": parity, "cpu_baseline": cpu, "text": text, "e2e_u8": e2e_u8,
        }
        print(json.dumps(out), flush=True)
    lib.free(ctx)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
