// preprocess.cu -- N1: clip_image_preprocess / clip_image_batch_preprocess (clip.cpp:797-1008) on the device.
//
// Raw u8 RGB images of arbitrary size go up (3 B/pixel instead of 12 B/pixel of finished fp32 crops), two kernels per micro-batch
// produce the [S, S, 3] fp32 crops directly in the tower's pixel staging buffer:
//   K6a resize_h : horizontal bicubic pass for the columns of the centre crop and the source rows the vertical pass will read
//   K6b resize_v : vertical pass for the crop rows, clamp, (v/255 - mean)/std
// The result is BIT-IDENTICAL to this library's host path (host_ops.cpp: preprocess_image, itself pinned against reference outputs
// in tests/test_host_side.py): the taps are computed on the host by the very routine the host path uses (resize_taps, double
// precision, PIL antialias support), both passes accumulate in double in ascending tap order with separate IEEE multiply and add
// (__dmul_rn / __dadd_rn, no FMA contraction), each pass is clamped to [0,255] and stored as float exactly like clip.cpp:855-900,
// and the final normalise uses IEEE float division.  Only the crop window is computed -- the reference resizes the whole image and
// then crops.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "host_ops.h"
#include "kernels.h"
#include "model.h"

namespace cb {

namespace {

struct PreImgDev {
    const uint8_t* src;
    const double *kh, *kv;
    const int *loh, *cnth, *lov, *cntv;
    float* tmp;        // [ycnt][S][3] horizontally resized rows y0 .. y0 + ycnt
    int nx, ny, ksh, ksv, y0, ycnt;
};

__global__ void __launch_bounds__(256) resize_h_kernel(const PreImgDev* __restrict__ imgs, int S) {
    const PreImgDev im = imgs[blockIdx.x];
    const int E = S * 3;
    for (int y = blockIdx.y; y < im.ycnt; y += gridDim.y) {
        const uint8_t* row = im.src + (size_t)(im.y0 + y) * im.nx * 3;
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const int xx = e / 3, c = e - 3 * xx;
            const double* k = im.kh + (size_t)xx * im.ksh;
            const int lo = im.loh[xx], cnt = im.cnth[xx];
            double acc = 0.0;
            for (int x = 0; x < cnt; x++) acc = __dadd_rn(acc, __dmul_rn((double)row[3 * (x + lo) + c], k[x]));
            im.tmp[(size_t)y * E + e] = fminf(fmaxf((float)acc, 0.0f), 255.0f);
        }
    }
}

__global__ void __launch_bounds__(256) resize_v_kernel(const PreImgDev* __restrict__ imgs, int S, float m0, float m1, float m2, float s0,
                                                       float s1, float s2, float* __restrict__ out) {
    const PreImgDev im = imgs[blockIdx.x];
    const int E = S * 3;
    float* dst = out + (size_t)blockIdx.x * S * E;
    for (int yy = blockIdx.y; yy < S; yy += gridDim.y) {
        const double* k = im.kv + (size_t)yy * im.ksv;
        const int lo = im.lov[yy] - im.y0, cnt = im.cntv[yy];
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const int c = e % 3;
            double acc = 0.0;
            for (int y = 0; y < cnt; y++) acc = __dadd_rn(acc, __dmul_rn((double)im.tmp[(size_t)(lo + y) * E + e], k[y]));
            const float v = fminf(fmaxf((float)acc, 0.0f), 255.0f);
            const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
            dst[(size_t)yy * E + e] = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), mean), sd);
        }
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct AxisTaps {   // taps of one (in, out) size pair restricted to the crop window, shared by all images of that size
    ResizeTaps t;
    size_t off_k = 0, off_lo = 0, off_cnt = 0;
};

}  // namespace

// Enqueue the preprocessing of `nb` images into d_pixels [nb, S, S, 3]: the blob upload goes on `copy_st` (it overlaps the previous
// micro-batch's kernels), `copied` orders it before the two resize kernels on `st`.  `buf` selects one of the two staging arenas of
// the context; the caller must have made sure the work that last used arena `buf` has finished (it is rewritten here on the host).  Host-side cost: tap tables for each distinct
// image size (a few thousand cubic evaluations) and one memcpy of the raw pixels into pinned memory.
bool preprocess_device(clip_ctx* c, const clip_image_u8* imgs, int nb, int buf, float* d_pixels, cudaStream_t copy_st, cudaEvent_t copied,
                       cudaStream_t st, std::string& err) {
    const int S = c->vis.image_size;
    PreArena& a = c->pre[buf];
    // ---- plan the blob: [descriptors][tap tables][raw pixels]
    std::map<std::pair<int, int>, AxisTaps> hx, vy;    // key: (in size, resized size)
    struct Plan { int nx3, ny3, xo, yo; const AxisTaps *h, *v; size_t off_src, off_tmp; int y0, ycnt; };
    std::vector<Plan> plan(nb);
    size_t off = align_up(sizeof(PreImgDev) * (size_t)nb, 16);
    auto place = [&](AxisTaps& t) {
        t.off_k = off; off = align_up(off + t.t.k.size() * 8, 16);
        t.off_lo = off; off = align_up(off + t.t.lo.size() * 4, 16);
        t.off_cnt = off; off = align_up(off + t.t.cnt.size() * 4, 16);
    };
    for (int j = 0; j < nb; j++) {
        Plan& p = plan[j];
        if (!imgs[j].data || !preprocess_geometry(imgs[j].nx, imgs[j].ny, S, &p.nx3, &p.ny3)) { err = "clip_b200 preprocess: bad image " + std::to_string(j); return false; }
        p.xo = (p.nx3 - S) / 2; p.yo = (p.ny3 - S) / 2;
        auto ih = hx.find({imgs[j].nx, p.nx3});
        if (ih == hx.end()) { ih = hx.emplace(std::make_pair(imgs[j].nx, p.nx3), AxisTaps()).first; ih->second.t = resize_taps(imgs[j].nx, p.nx3, p.xo, S); place(ih->second); }
        auto iv = vy.find({imgs[j].ny, p.ny3});
        if (iv == vy.end()) { iv = vy.emplace(std::make_pair(imgs[j].ny, p.ny3), AxisTaps()).first; iv->second.t = resize_taps(imgs[j].ny, p.ny3, p.yo, S); place(iv->second); }
        p.h = &ih->second; p.v = &iv->second;
        p.y0 = p.v->t.lo[0];
        int yend = 0;
        for (int i = 0; i < S; i++) yend = std::max(yend, p.v->t.lo[i] + p.v->t.cnt[i]);
        p.ycnt = yend - p.y0;
    }
    size_t tmp_floats = 0;
    for (int j = 0; j < nb; j++) {
        plan[j].off_src = off; off = align_up(off + (size_t)imgs[j].nx * imgs[j].ny * 3, 16);
        plan[j].off_tmp = tmp_floats; tmp_floats += (size_t)plan[j].ycnt * S * 3;
    }
    // ---- (re)size the arenas
    if (off > a.cap) {
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        a.h = nullptr; a.d = nullptr; a.cap = 0;
        const size_t cap = off + off / 4;
        if (cudaHostAlloc((void**)&a.h, cap, cudaHostAllocDefault) != cudaSuccess || cudaMalloc((void**)&a.d, cap) != cudaSuccess) { err = "clip_b200 preprocess: staging allocation failed"; return false; }
        a.cap = cap;
    }
    if (tmp_floats > a.tmp_cap) {
        if (a.d_tmp) cudaFree(a.d_tmp);
        a.d_tmp = nullptr; a.tmp_cap = 0;
        const size_t cap = tmp_floats + tmp_floats / 4;
        if (cudaMalloc((void**)&a.d_tmp, cap * 4) != cudaSuccess) { err = "clip_b200 preprocess: scratch allocation failed"; return false; }
        a.tmp_cap = cap;
    }
    // ---- fill the pinned blob
    PreImgDev* desc = reinterpret_cast<PreImgDev*>(a.h);
    auto fill = [&](const AxisTaps& t) {
        memcpy(a.h + t.off_k, t.t.k.data(), t.t.k.size() * 8);
        memcpy(a.h + t.off_lo, t.t.lo.data(), t.t.lo.size() * 4);
        memcpy(a.h + t.off_cnt, t.t.cnt.data(), t.t.cnt.size() * 4);
    };
    for (auto& kv : hx) fill(kv.second);
    for (auto& kv : vy) fill(kv.second);
    for (int j = 0; j < nb; j++) {
        const Plan& p = plan[j];
        PreImgDev d;
        d.src = a.d + p.off_src;
        d.kh = reinterpret_cast<const double*>(a.d + p.h->off_k); d.loh = reinterpret_cast<const int*>(a.d + p.h->off_lo); d.cnth = reinterpret_cast<const int*>(a.d + p.h->off_cnt);
        d.kv = reinterpret_cast<const double*>(a.d + p.v->off_k); d.lov = reinterpret_cast<const int*>(a.d + p.v->off_lo); d.cntv = reinterpret_cast<const int*>(a.d + p.v->off_cnt);
        d.tmp = a.d_tmp + p.off_tmp;
        d.nx = imgs[j].nx; d.ny = imgs[j].ny; d.ksh = p.h->t.ksize; d.ksv = p.v->t.ksize; d.y0 = p.y0; d.ycnt = p.ycnt;
        desc[j] = d;
        memcpy(a.h + p.off_src, imgs[j].data, (size_t)imgs[j].nx * imgs[j].ny * 3);
    }
    a.last_bytes = off;
    if (cudaMemcpyAsync(a.d, a.h, off, cudaMemcpyHostToDevice, copy_st) != cudaSuccess || cudaEventRecord(copied, copy_st) != cudaSuccess ||
        cudaStreamWaitEvent(st, copied, 0) != cudaSuccess) { err = "clip_b200 preprocess: upload failed"; return false; }
    const dim3 grid((unsigned)nb, 32);
    resize_h_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PreImgDev*>(a.d), S);
    resize_v_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PreImgDev*>(a.d), S, c->image_mean[0], c->image_mean[1], c->image_mean[2],
                                          c->image_std[0], c->image_std[1], c->image_std[2], d_pixels);
    if (cudaGetLastError() != cudaSuccess) { err = "clip_b200 preprocess: kernel launch failed"; return false; }
    return true;
}

void preprocess_release(clip_ctx* c) {
    for (int i = 0; i < 2; i++) {
        PreArena& a = c->pre[i];
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        if (a.d_tmp) cudaFree(a.d_tmp);
        a = PreArena();
    }
}

}  // namespace cb
