"""Worker of tests/test_gpu_multi.py: one process per GPU in RANKS mode (the way the driver launches bench.py under torchrun, but
without torch): RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT in the environment, NCCL communicator created by the library itself.
Writes <out>.rank<r>.npz with everything the parent compares."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "clip.cpp_b200"), HERE):
    sys.path.insert(0, p)
import binding as bd          # noqa: E402
import synth_gguf as sg       # noqa: E402


def main():
    model, out, n_img, n_lab = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["CLIP_B200_DEVICE"] = str(local)
    lib = bd.ClipLib(bd.PRODUCT_LIB)
    assert "torch" not in sys.modules
    ctx = lib.load(model, 0)
    assert lib.lib.clip_b200_dist_init(ctx, rank, world, None), lib.last_error()
    assert lib.lib.clip_b200_dist_world(ctx) == world and lib.lib.clip_b200_dist_rank(ctx) == rank
    import ctypes as C
    imgs = sg.synth_images(n_img, 64, 100 + rank)                       # this rank's shard
    seqs = [sg.synth_tokens(1, 5 + (i % 60), 1000 * rank + i)[0] for i in range(n_lab)]
    d = lib.vision_hparams(ctx).projection_dim
    batch, keep = lib.make_image_batch(imgs)
    img_all = np.empty((world * n_img, d), np.float32)
    assert lib.lib.clip_b200_image_batch_encode_all(ctx, 4, C.byref(batch), img_all.ctypes.data_as(C.POINTER(C.c_float)), True), lib.last_error()
    arr, keep2 = lib.make_token_array(seqs)
    txt_all = np.empty((world * n_lab, d), np.float32)
    assert lib.lib.clip_b200_text_batch_encode_all(ctx, 4, arr, n_lab, txt_all.ctypes.data_as(C.POINTER(C.c_float)), True), lib.last_error()
    local_img = lib.image_batch_encode(ctx, imgs)
    local_txt = lib.text_batch_encode(ctx, seqs)
    # device-resident variant: pixels in HBM, K5 writes into this rank's slot, in-place all-gather
    d_pix = lib.lib.clip_b200_device_malloc(ctx, imgs.nbytes)
    d_all = lib.lib.clip_b200_device_malloc(ctx, img_all.nbytes)
    assert lib.lib.clip_b200_memcpy_h2d(ctx, d_pix, imgs.ctypes.data, imgs.nbytes)
    assert lib.lib.clip_b200_image_encode_device_all(ctx, d_pix, n_img, d_all, True), lib.last_error()
    dev_all = np.empty_like(img_all)
    assert lib.lib.clip_b200_memcpy_d2h(ctx, dev_all.ctypes.data, d_all, dev_all.nbytes)
    # zero-shot: local images x ALL ranks' labels
    sc, ix = lib.zero_shot_images(ctx, imgs, seqs, 5, normalize=True)
    vals = (C.c_double * 2)(float(rank), -float(rank))
    assert lib.lib.clip_b200_dist_max_f64(ctx, vals, 2)
    assert lib.lib.clip_b200_dist_barrier(ctx)
    np.savez(out + ".rank%d.npz" % rank, img_all=img_all, txt_all=txt_all, local_img=local_img, local_txt=local_txt, dev_all=dev_all,
             zs_scores=sc, zs_idx=ix, maxes=np.array(list(vals)), nccl=np.array(lib.lib.clip_b200_nccl_version()))
    lib.free(ctx)


if __name__ == "__main__":
    main()
