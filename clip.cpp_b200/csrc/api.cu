// api.cu -- the C ABI of libclip_b200.so (include/clip_b200.h): loader, static kernel schedule of both towers,
// host<->device plumbing.  No graph, no dispatch: each encode call issues a fixed sequence of kernel launches on
// one stream.  There is NO CPU path -- if CUDA is unavailable clip_model_load fails loudly.
#include <chrono>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <numeric>
#include <thread>

#include "common.cuh"
#include "gguf.hpp"
#include "kernels.h"
#include "model.h"
#include "wpack.h"

using namespace cb;

namespace {

thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; fprintf(stderr, "clip_b200: %s\n", s.c_str()); }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_err(std::string(#call) + " failed: " + cudaGetErrorString(e_));                    \
            return false;                                                                          \
        }                                                                                          \
    } while (0)

uint16_t f32_to_bf16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x40u);
    x += 0x7FFFu + ((x >> 16) & 1u);
    return (uint16_t)(x >> 16);
}

template <class T>
bool dev_alloc(clip_ctx* c, T** p, size_t count) {
    void* d = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 256);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) { set_err(std::string("cudaMalloc(") + std::to_string(bytes) + ") failed: " + cudaGetErrorString(e)); return false; }
    c->allocs.push_back(d);
    *p = (T*)d;
    return true;
}
template <class T>
bool upload(clip_ctx* c, T** p, const void* host, size_t bytes) {
    uint8_t* d = nullptr;
    if (!dev_alloc(c, &d, bytes)) return false;
    CK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    *p = (T*)d;
    return true;
}

// ---- profiling hooks (CUDA events on the launch stream) -------------------------------------------------
cudaEvent_t get_event(clip_ctx* c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct Scope {
    clip_ctx* c; cudaEvent_t e0 = nullptr; int kind;
    Scope(clip_ctx* c_, int kind_) : c(c_), kind(kind_) {
        c->launches++;
        if (c->profile) { e0 = get_event(c); cudaEventRecord(e0, c->stream); }
    }
    ~Scope() {
        if (!c->profile) return;
        cudaEvent_t e1 = get_event(c);
        cudaEventRecord(e1, c->stream);
        c->prof.push_back({kind, e0, e1});
        if (c->prof.size() > 65536) {        // nobody drains the records: recycle the older half instead of growing without bound
            for (size_t i = 0; i < 32768; i++) { c->ev_pool.push_back(c->prof[i].e0); c->ev_pool.push_back(c->prof[i].e1); }
            c->prof.erase(c->prof.begin(), c->prof.begin() + 32768);
        }
    }
};
enum { K_GEMM = 0, K_ATTN = 1, K_LN = 2, K_OTHER = 3 };

// ---- tensor lookup helpers ---------------------------------------------------------------------------------
const GgufTensor* need(const GgufFile& g, const std::string& name) {
    const GgufTensor* t = g.tensor(name);
    if (!t) set_err("unable to find tensor " + name);
    return t;
}
bool kv_u32(const GgufFile& g, const std::string& k, int32_t& out) {
    const GgufKV* kv = g.find(k);
    if (!kv) { set_err("key " + k + " not found in file"); return false; }
    out = (int32_t)kv->u;
    return true;
}
bool kv_f32(const GgufFile& g, const std::string& k, float& out) {
    const GgufKV* kv = g.find(k);
    if (!kv) { set_err("key " + k + " not found in file"); return false; }
    out = (float)kv->f;
    return true;
}
bool kv_bool(const GgufFile& g, const std::string& k, bool& out) {
    const GgufKV* kv = g.find(k);
    if (!kv) { set_err("key " + k + " not found in file"); return false; }
    out = kv->u != 0;
    return true;
}

// fp32 vector/table on the device, dequantised exactly as ggml_get_rows / the f32 ops would see it
bool upload_f32(clip_ctx* c, const GgufFile& g, const std::string& name, float** out, int64_t expect_elems) {
    const GgufTensor* t = need(g, name);
    if (!t) return false;
    if ((int64_t)t->nelements() != expect_elems) { set_err("tensor " + name + " has unexpected size"); return false; }
    std::vector<float> h((size_t)expect_elems);
    const int64_t k = (int64_t)t->ne[0], rows = expect_elems / k;
    const size_t rb = (size_t)(k / (int64_t)ggml_type_block_elems(t->type)) * ggml_type_block_bytes(t->type);
    for (int64_t r = 0; r < rows; r++)
        if (!dequant_row((int)t->type, t->data + (size_t)r * rb, h.data() + r * k, k)) { set_err("cannot dequantise " + name); return false; }
    return upload(c, out, h.data(), h.size() * 4);
}

// GEMM weight from one or more row-concatenated tensors (fused QKV).  kpad > 0 zero-pads K (patch embedding).
bool make_linear(clip_ctx* c, const GgufFile& g, const std::vector<std::string>& wn, const std::vector<std::string>& bn, Linear& L,
                 int kpad = 0) {
    std::vector<const GgufTensor*> ts;
    for (auto& n : wn) { const GgufTensor* t = need(g, n); if (!t) return false; ts.push_back(t); }
    const uint32_t type = ts[0]->type;
    int64_t K = 1;
    for (uint32_t d = 0; d + 1 < std::max(ts[0]->n_dims, 2u); d++) K *= (int64_t)ts[0]->ne[d];   // 4-D conv weight: K = P*P*3
    int64_t N = 0;
    for (auto* t : ts) {
        if (t->type != type || (int64_t)(t->nelements() / (t->ne[ts[0]->n_dims - 1])) != K) { set_err("inconsistent fused weight " + wn[0]); return false; }
        N += (int64_t)t->ne[t->n_dims - 1];
    }
    const int64_t Kp = kpad > 0 ? kpad : K;
    if (N % GEMM_BM || Kp % GEMM_BK) { set_err("weight " + wn[0] + ": shape [" + std::to_string(N) + "," + std::to_string(Kp) + "] is not tileable (N%128, K%64)"); return false; }
    L.N = (int)N; L.K = (int)Kp;
    if (type == 0 || type == 1) {
        L.qtype = QT_F16;
        std::vector<uint16_t> h((size_t)N * Kp, 0);
        int64_t r0 = 0;
        for (auto* t : ts) {
            const int64_t rows = (int64_t)t->ne[t->n_dims - 1];
            for (int64_t r = 0; r < rows; r++) {
                uint16_t* dst = h.data() + (size_t)(r0 + r) * Kp;
                if (type == 1) memcpy(dst, t->data + (size_t)r * K * 2, (size_t)K * 2);
                else { const float* s = (const float*)(t->data) + (size_t)r * K; for (int64_t k = 0; k < K; k++) dst[k] = f32_to_f16(s[k]); }
            }
            r0 += rows;
        }
        if (!upload(c, &L.d_w, h.data(), h.size() * 2)) return false;
        if (!make_tma_2d_16bit(&L.w_map, L.d_w, (uint64_t)N, (uint64_t)Kp, (uint64_t)Kp, GEMM_BM)) { set_err("cuTensorMapEncodeTiled failed for " + wn[0]); return false; }
        L.d_raw = L.d_w;
    } else {
        if (kpad > 0) { set_err("quantized weights cannot be K-padded: " + wn[0]); return false; }
        L.qtype = (int)type;
        const size_t chunk = wpack_chunk_bytes((int)type), bb = wpack_ggml_block_bytes((int)type);
        if (!chunk) { set_err("unsupported weight type in " + wn[0]); return false; }
        std::vector<uint8_t> packed(wpack_total_bytes((int)type, N, K));
        int64_t r0 = 0;
        for (auto* t : ts) {
            const int64_t rows = (int64_t)t->ne[1];
            if (rows % GEMM_BM) { set_err("fused part of " + wn[0] + " is not a multiple of 128 rows"); return false; }
            if (!wpack_repack((int)type, t->data, rows, K, packed.data() + (size_t)(r0 / GEMM_BM) * (K / GEMM_BK) * chunk)) { set_err("re-tiling failed for " + wn[0]); return false; }
            r0 += rows;
        }
        if (!upload(c, &L.d_w, packed.data(), packed.size())) return false;
        if (c->debug_naive) {
            std::vector<uint8_t> raw((size_t)N * (K / 32) * bb);
            size_t o = 0;
            for (auto* t : ts) { memcpy(raw.data() + o, t->data, t->nbytes); o += t->nbytes; }
            if (!upload(c, &L.d_raw, raw.data(), raw.size())) return false;
        }
    }
    if (!bn.empty()) {
        std::vector<float> hb;
        for (auto& n : bn) {
            const GgufTensor* t = need(g, n);
            if (!t) return false;
            if (t->type != 0) { set_err("bias " + n + " must be f32"); return false; }
            const float* p = (const float*)t->data;
            hb.insert(hb.end(), p, p + t->nelements());
        }
        if ((int64_t)hb.size() != N) { set_err("bias size mismatch for " + wn[0]); return false; }
        if (!upload(c, &L.d_bias, hb.data(), hb.size() * 4)) return false;
    }
    return true;
}

bool load_blocks(clip_ctx* c, const GgufFile& g, const char* p, Tower& tw) {
    tw.L.resize(tw.layers);
    for (int il = 0; il < tw.layers; il++) {
        const std::string b = std::string(p) + ".blk." + std::to_string(il) + ".";
        Layer& l = tw.L[il];
        if (!make_linear(c, g, {b + "attn_q.weight", b + "attn_k.weight", b + "attn_v.weight"},
                         {b + "attn_q.bias", b + "attn_k.bias", b + "attn_v.bias"}, l.qkv)) return false;
        if (!make_linear(c, g, {b + "attn_out.weight"}, {b + "attn_out.bias"}, l.out)) return false;
        if (!make_linear(c, g, {b + "ffn_down.weight"}, {b + "ffn_down.bias"}, l.fc1)) return false;   // HF fc1: h -> f
        if (!make_linear(c, g, {b + "ffn_up.weight"}, {b + "ffn_up.bias"}, l.fc2)) return false;       // HF fc2: f -> h
        if (l.qkv.N != 3 * tw.hidden || l.qkv.K != tw.hidden || l.out.N != tw.hidden || l.out.K != tw.hidden ||
            l.fc1.N != tw.ff || l.fc1.K != tw.hidden || l.fc2.N != tw.hidden || l.fc2.K != tw.ff) { set_err("layer shape mismatch in " + b); return false; }
        if (!upload_f32(c, g, b + "ln1.weight", &l.ln1_g, tw.hidden) || !upload_f32(c, g, b + "ln1.bias", &l.ln1_b, tw.hidden) ||
            !upload_f32(c, g, b + "ln2.weight", &l.ln2_g, tw.hidden) || !upload_f32(c, g, b + "ln2.bias", &l.ln2_b, tw.hidden)) return false;
    }
    return true;
}

bool check_geometry(const char* what, const Tower& t) {
    if (t.hidden <= 0 || t.ff <= 0 || t.heads <= 0 || t.layers < 0 || t.layers > 1024 || t.proj <= 0 || !(t.eps >= 0.f)) { set_err(std::string(what) + ": non-positive hyper-parameter in the file"); return false; }
    if (t.heads <= 0 || t.hidden % t.heads || t.hidden / t.heads != 64) { set_err(std::string(what) + ": head_dim must be 64 (hidden " + std::to_string(t.hidden) + ", heads " + std::to_string(t.heads) + ")"); return false; }
    if (t.hidden > 2048) { set_err(std::string(what) + ": hidden size > 2048 is not supported by the register-resident LayerNorm"); return false; }
    if (t.hidden % 128 || t.ff % 128 || t.proj % 128) { set_err(std::string(what) + ": hidden/ff/projection sizes must be multiples of 128"); return false; }
    return true;
}

bool any_f16_gemm(const Tower& t) {
    if (!t.present) return false;
    bool f = t.proj_w.qtype == QT_F16;
    for (auto& l : t.L) f = f || l.qkv.qtype == QT_F16 || l.out.qtype == QT_F16 || l.fc1.qtype == QT_F16 || l.fc2.qtype == QT_F16;
    return f;
}

// ---- workspaces ----------------------------------------------------------------------------------------------
bool ensure_ws(clip_ctx* c, Tower& tw, int items, int T, bool vision) {
    Workspace& w = tw.ws;
    if (w.cap_items >= items && w.T == T) return true;
    if (w.cap_items) { set_err("workspace re-sizing is not supported after the first encode"); return false; }
    const int h = tw.hidden;
    w.cap_items = items; w.T = T; w.cap_rows = items * T;
    const size_t rows = (size_t)w.cap_rows;
    if (!dev_alloc(c, &w.x, rows * h) || !dev_alloc(c, &w.a, rows * h) || !dev_alloc(c, &w.qkv, rows * 3 * h) ||
        !dev_alloc(c, &w.g, rows * tw.ff) || !dev_alloc(c, &w.sel16, (size_t)items * h) || !dev_alloc(c, &w.sel32, (size_t)items * h) ||
        !dev_alloc(c, &w.proj32, (size_t)items * tw.proj)) return false;
    bool ok = make_tma_2d_16bit(&w.map_a, w.a, rows, h, h, GEMM_BN) && make_tma_2d_16bit(&w.map_g, w.g, rows, tw.ff, tw.ff, GEMM_BN) &&
              make_tma_2d_16bit(&w.map_sel, w.sel16, items, h, h, GEMM_BN) &&
              make_tma_2d_16bit(&w.map_a_half, w.a, rows, h, h, GEMM_BN / 2) && make_tma_2d_16bit(&w.map_g_half, w.g, rows, tw.ff, tw.ff, GEMM_BN / 2) &&
              make_tma_2d_16bit(&w.map_q128, w.qkv, rows, 3 * h, 3 * h, 128) &&
              make_tma_2d_16bit(&w.map_kv16, w.qkv, rows, 3 * h, 3 * h, 16) &&
              make_tma_2d_16bit_plain(&w.map_out_qkv, w.qkv, rows, 3 * h, 3 * h, GEMM_OUT_BOX, GEMM_OUT_BOX) &&
              make_tma_2d_f32_plain(&w.map_out_x32, w.x, rows, h, h, GEMM_OUT_BOX, GEMM_OUT_BOX) &&
              make_tma_2d_16bit_plain(&w.map_out_g, w.g, rows, tw.ff, tw.ff, GEMM_OUT_BOX, GEMM_OUT_BOX);
    if (vision) {
        const size_t per = (size_t)tw.image_size * tw.image_size * 3;
        if (!dev_alloc(c, &w.pixels[0], items * per) || !dev_alloc(c, &w.pixels[1], items * per) ||
            !dev_alloc(c, &w.patches, (size_t)items * tw.n_patches * tw.kpad) || !dev_alloc(c, &w.patch32, (size_t)items * tw.n_patches * h)) return false;
        ok = ok && make_tma_2d_16bit(&w.map_patches, w.patches, (uint64_t)items * tw.n_patches, tw.kpad, tw.kpad, GEMM_BN);
    } else {
        if (!dev_alloc(c, &w.ids, rows) || !dev_alloc(c, &w.last, (size_t)items)) return false;
    }
    if (!ok) { set_err("cuTensorMapEncodeTiled failed for a workspace buffer"); return false; }
    return true;
}

bool run_linear(clip_ctx* c, const Linear& L, const TmaMap* xmap, const void* xptr, bool x_bf16, int M, void* out, int ldo, int epi,
                int out_bf16, int scale_cols = 0, float scale = 1.f, const TmaMap* xmap_half = nullptr, const TmaMap* out_map = nullptr) {
    Scope s(c, K_GEMM);
    if (c->debug_naive) {
        launch_naive_gemm(xptr, x_bf16, L.d_raw, L.qtype, L.d_bias, out, M, L.N, L.K, ldo, epi, out_bf16, scale_cols, scale, c->stream);
        CK(cudaGetLastError());
        return true;
    }
    GemmArgs a;
    a.x_map = xmap; a.x_half_map = xmap_half; a.w_map = &L.w_map; a.w_packed = L.d_w; a.qtype = L.qtype; a.operand_bf16 = x_bf16;
    a.bias = L.d_bias; a.out = out; a.M = M; a.N = L.N; a.K = L.K; a.ldo = ldo; a.epi = epi; a.out_bf16 = out_bf16;
    a.scale_cols = scale_cols; a.scale = scale; a.out_map = out_map;
    CK(gemm_launch(a, c->stream, c->num_sms, nullptr));
    return true;
}

// The per-layer schedule shared by both towers (clip.cpp:1064-1143 text, 1342-1423 vision).
bool run_blocks(clip_ctx* c, Tower& tw, int nseq, int T, bool causal) {
    // The residual stream x stays fp32 (as in the reference).  The branch GEMMs (out-proj, FC2) add their result INTO x themselves: the
    // epilogue stages fp32 blocks in shared memory and a TMA tensor reduce-add (cp.reduce.async.bulk.tensor .add) performs x += acc + bias
    // in the memory system -- the SMs never read x for it, and the traffic hides under the tensor-bound GEMM.  LayerNorm then only reads
    // x and writes its 16-bit output: 6 B per element instead of the 12 B of the round-1 "deferred residual" scheme.
    Workspace& w = tw.ws;
    const int M = nseq * T, h = tw.hidden;
    const int bf = c->operand_bf16 ? 1 : 0;
    const float qscale = 1.0f / sqrtf(64.0f);
    for (auto& l : tw.L) {
        { Scope s(c, K_LN); launch_layernorm(w.x, h, M, h, tw.eps, l.ln1_g, l.ln1_b, nullptr, w.a, bf, c->stream); }
        if (!run_linear(c, l.qkv, &w.map_a, w.a, bf, M, w.qkv, 3 * h, EPI_STORE16, bf, h, qscale, &w.map_a_half, &w.map_out_qkv)) return false;
        if (c->attn_tc && attention_tc_supported(T)) {
            const int q_done = attention_tc_tiles(T) * 128;
            { Scope s(c, K_ATTN); CK(launch_attention_tc(&w.map_q128, w.qkv, &w.map_kv16, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, c->num_sms, c->stream)); }
            if (q_done < T) { Scope s(c, K_ATTN); launch_attention(w.qkv, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, q_done, c->stream); }
        } else if (c->attn_tc && attention_tc_long_supported(T, causal ? 1 : 0)) {
            Scope s(c, K_ATTN); CK(launch_attention_tc_long(&w.map_q128, w.qkv, w.a, nseq, T, tw.heads, bf, c->num_sms, c->stream));
        } else {
            Scope s(c, K_ATTN); launch_attention(w.qkv, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, 0, c->stream);
        }
        if (!run_linear(c, l.out, &w.map_a, w.a, bf, M, w.x, h, EPI_REDADD32, 0, 0, 1.f, &w.map_a_half, &w.map_out_x32)) return false;
        { Scope s(c, K_LN); launch_layernorm(w.x, h, M, h, tw.eps, l.ln2_g, l.ln2_b, nullptr, w.a, bf, c->stream); }
        if (!run_linear(c, l.fc1, &w.map_a, w.a, bf, M, w.g, tw.ff, c->use_gelu ? EPI_GELU16 : EPI_QGELU16, bf, 0, 1.f, &w.map_a_half, &w.map_out_g)) return false;
        if (!run_linear(c, l.fc2, &w.map_g, w.g, bf, M, w.x, h, EPI_REDADD32, 0, 0, 1.f, &w.map_g_half, &w.map_out_x32)) return false;
    }
    CK(cudaGetLastError());
    return true;
}

// Vision tower on nb images already resident at d_pixels (NHWC f32); writes d_out [nb, proj].
bool vision_forward(clip_ctx* c, const float* d_pixels, int nb, float* d_out, bool normalize) {
    Tower& tw = c->vis;
    Workspace& w = tw.ws;
    const int h = tw.hidden, bf = c->operand_bf16 ? 1 : 0;
    { Scope s(c, K_OTHER); launch_im2col(d_pixels, nb, tw.image_size, tw.patch, tw.kpad, w.patches, c->stream); }
    // patch embedding = stride-P conv = GEMM over fp16 patches, fp16 weights, fp32 out, no bias (clip.cpp:1309)
    if (!run_linear(c, tw.patch_w, &w.map_patches, w.patches, false, nb * tw.n_patches, w.patch32, h, EPI_STORE32, 0)) return false;
    { Scope s(c, K_OTHER); launch_assemble_preln(w.patch32, tw.class_embd, tw.pos, nb, tw.T, h, tw.eps, tw.pre_g, tw.pre_b, w.x, c->stream); }
    if (!run_blocks(c, tw, nb, tw.T, false)) return false;
    // CLS rows -> post-LN -> projection -> (L2 norm)   (clip.cpp:1426-1455)
    { Scope s(c, K_LN); launch_layernorm(w.x, (size_t)tw.T * h, nb, h, tw.eps, tw.post_g, tw.post_b, nullptr, w.sel16, bf, c->stream); }
    if (!run_linear(c, tw.proj_w, &w.map_sel, w.sel16, bf, nb, w.proj32, tw.proj, EPI_STORE32, 0)) return false;
    { Scope s(c, K_OTHER); launch_l2norm(w.proj32, d_out, nb, tw.proj, normalize ? 1 : 0, c->stream); }
    CK(cudaGetLastError());
    return true;
}

// Text tower on nb sequences padded to T tokens; ws.ids / ws.last already filled.
bool text_forward(clip_ctx* c, int nb, int T, float* d_out, bool normalize) {
    Tower& tw = c->txt;
    Workspace& w = tw.ws;
    const int h = tw.hidden, bf = c->operand_bf16 ? 1 : 0;
    { Scope s(c, K_OTHER); launch_text_embed(w.ids, tw.tok, tw.pos, nb, T, h, tw.n_vocab, w.x, c->stream); }
    if (!run_blocks(c, tw, nb, T, true)) return false;
    // final LN is row-wise, so LN(select(EOT)) == select(LN(all)) (clip.cpp:1146-1155)
    { Scope s(c, K_OTHER); launch_gather_rows(w.x, w.sel32, nb, h, T, w.last, nullptr, bf, c->stream); }
    { Scope s(c, K_LN); launch_layernorm(w.sel32, h, nb, h, tw.eps, tw.post_g, tw.post_b, nullptr, w.sel16, bf, c->stream); }
    if (!run_linear(c, tw.proj_w, &w.map_sel, w.sel16, bf, nb, w.proj32, tw.proj, EPI_STORE32, 0)) return false;
    { Scope s(c, K_OTHER); launch_l2norm(w.proj32, d_out, nb, tw.proj, normalize ? 1 : 0, c->stream); }
    CK(cudaGetLastError());
    return true;
}

void dev_free(clip_ctx* c, void* p) {
    if (!p) return;
    auto it = std::find(c->allocs.begin(), c->allocs.end(), p);
    if (it != c->allocs.end()) c->allocs.erase(it);
    cudaFree(p);
}
// grow-only device buffer; the outgrown one is released once the stream has drained (nothing may still read it)
template <class T>
bool ensure_buf(clip_ctx* c, T** buf, size_t* cap, size_t count) {
    if (*cap >= count) return true;
    const size_t want = std::max(count, *cap + *cap / 2);
    T* p = nullptr;
    if (!dev_alloc(c, &p, want)) return false;
    if (*buf) { cudaStreamSynchronize(c->stream); dev_free(c, *buf); }
    *buf = p; *cap = want;
    return true;
}
bool ensure_out(clip_ctx* c, size_t floats) { return ensure_buf(c, &c->d_out, &c->d_out_cap, floats); }

int default_micro_batch(int T) {   // 37 token tiles = 148/4: every N/128 that is a multiple of 4 fills whole waves
    // measured on B200 (ViT-L/14 q4_0, b=512, final kernels): 3 -> 4914, 4 -> 5062, 8 -> 5219, 12 -> 5316 img/s, one chunk of 512 -> 5394
    // device-resident but 5298 end to end (the H2D copy of a single chunk cannot hide under compute).  Fewer, longer launches: the
    // ~3 us drain/fill gap per launch is paid 1200 times per 512 images at 3 waves.  12 waves = 331 ViT-L/14 images = 1.9 GB of workspace.
    int waves = 12;
    if (const char* e = getenv("CLIP_B200_TOKEN_TILES_X37")) waves = std::max(1, atoi(e));
    const int mb = 37 * waves * GEMM_BN / T;
    return mb < 1 ? 1 : mb;
}

void chunking(size_t n, int mb, size_t& n_chunks, size_t& chunk) {
    n_chunks = (n + mb - 1) / mb;
    chunk = (n + n_chunks - 1) / n_chunks;
}

// Chunk boundaries for inputs that still have to cross PCIe: the host->device copy of the FIRST micro-batch cannot hide under any
// kernel, so it is made small (a quarter of a micro-batch); the rest is split into balanced micro-batches as usual.
std::vector<size_t> host_chunk_bounds(size_t n, int mb) {
    std::vector<size_t> b{0};
    if (n <= (size_t)mb) { b.push_back(n); return b; }
    const size_t head = std::max<size_t>(1, (size_t)mb / 4);
    b.push_back(head);
    size_t n_chunks, chunk;
    chunking(n - head, mb, n_chunks, chunk);
    for (size_t i0 = head; i0 < n; i0 += chunk) b.push_back(std::min(n, i0 + chunk));
    return b;
}

bool sync_and_time(clip_ctx* c) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { set_err(std::string("kernel execution failed: ") + cudaGetErrorString(e)); return false; }
    cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1);
    return true;
}

bool image_encode_device_locked(clip_ctx* c, const float* d_pixels, size_t n, float* d_vec, bool normalize) {
    Tower& tw = c->vis;
    if (!ensure_ws(c, tw, tw.micro_batch, tw.T, true)) return false;
    size_t n_chunks, chunk;
    chunking(n, tw.micro_batch, n_chunks, chunk);
    const size_t per = (size_t)tw.image_size * tw.image_size * 3;
    CK(cudaEventRecord(c->ev_t0, c->stream));
    for (size_t i0 = 0; i0 < n; i0 += chunk) {
        const int nb = (int)std::min(chunk, n - i0);
        if (!vision_forward(c, d_pixels + i0 * per, nb, d_vec + i0 * tw.proj, normalize)) return false;
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    return true;
}

// ---- devices mode helper: f(replica, r, lo, hi) on one host thread per GPU, contiguous balanced shards of n items --------------
template <class F>
bool for_each_replica(clip_ctx* c, size_t n, F f) {
    const int w = (int)c->replicas.size();
    std::vector<std::string> errs((size_t)w);
    std::vector<char> ok((size_t)w, 1);
    std::vector<std::thread> th;
    for (int r = 0; r < w; r++) {
        size_t lo, hi;
        shard_bounds(n, r, w, lo, hi);
        th.emplace_back([&, r, lo, hi]() {
            g_err.clear();
            ok[r] = f(c->replicas[r], r, lo, hi) ? 1 : 0;
            if (!ok[r]) errs[r] = g_err;
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(c->device);
    for (int r = 0; r < w; r++)
        if (!ok[r]) { set_err("device " + std::to_string(c->replicas[r]->device) + ": " + errs[r]); return false; }
    return true;
}

bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char* clip_b200_last_error(void) { return g_err.c_str(); }
const char* clip_b200_version(void) { return "clip_b200 0.1 (sm_100a; tcgen05 fused-dequant GEMM)"; }

// one replica of the model on `device` (weights are <= 0.33 GB: replicated, never sharded -- SURVEY.md section 8e)
static clip_ctx* load_on_device(const GgufFile& g, int device, const int verbosity) {
    {
        std::unique_ptr<clip_ctx, void (*)(clip_ctx*)> guard(new clip_ctx, [](clip_ctx* p) { clip_free(p); });
        clip_ctx* c = guard.get();
        c->device = device;
        if (cudaSetDevice(c->device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) { set_err("cudaGetDeviceProperties failed"); return nullptr; }
        if (prop.major != 10) { set_err(std::string("device '") + prop.name + "' is not sm_100: this library only contains sm_100a code"); return nullptr; }
        c->num_sms = prop.multiProcessorCount;
        c->debug_naive = getenv("CLIP_B200_DEBUG_NAIVE") && atoi(getenv("CLIP_B200_DEBUG_NAIVE")) != 0;
        c->profile = getenv("CLIP_B200_PROFILE") && atoi(getenv("CLIP_B200_PROFILE")) != 0;
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); return nullptr; }
        for (int i = 0; i < 2; i++) {
            cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming);
        }
        cudaEventCreate(&c->ev_t0);
        cudaEventCreate(&c->ev_t1);
        if (gemm_init() != cudaSuccess || attention_tc_init() != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic smem) failed"); return nullptr; }
        if (const char* at = getenv("CLIP_B200_ATTN")) c->attn_tc = strcmp(at, "legacy") != 0;

        if (!kv_bool(g, "clip.has_text_encoder", c->has_text) || !kv_bool(g, "clip.has_vision_encoder", c->has_vision) ||
            !kv_bool(g, "clip.use_gelu", c->use_gelu)) return nullptr;
        if (verbosity >= 1) {
            const GgufKV* nm = g.find("general.name");
            const GgufKV* ds = g.find("general.description");
            const GgufKV* ft = g.find("general.file_type");
            if (nm) printf("%s: model name:   %s\n", "clip_model_load", nm->s.c_str());
            if (ds) printf("%s: description:  %s\n", "clip_model_load", ds->s.c_str());
            printf("%s: GGUF version: %u\n", "clip_model_load", g.version);
            printf("%s: alignment:    %zu\n", "clip_model_load", g.alignment);
            printf("%s: n_tensors:    %zu\n", "clip_model_load", g.tensors.size());
            printf("%s: n_kv:         %zu\n", "clip_model_load", g.kvs.size());
            if (ft) printf("%s: ftype:        %d\n", "clip_model_load", (int)ft->u);
            printf("%s: text_encoder:   %d\n%s: vision_encoder: %d\n", "clip_model_load", c->has_text, "clip_model_load", c->has_vision);
            printf("%s: device:       %d (%s, %d SMs)\n", "clip_model_load", c->device, prop.name, c->num_sms);
        }

        if (c->has_text) {
            Tower& t = c->txt;
            auto& hp = c->thp;
            if (!kv_u32(g, "clip.text.embedding_length", hp.hidden_size) || !kv_u32(g, "clip.text.attention.head_count", hp.n_head) ||
                !kv_u32(g, "clip.text.feed_forward_length", hp.n_intermediate) || !kv_u32(g, "clip.text.block_count", hp.n_layer) ||
                !kv_u32(g, "clip.text.context_length", hp.num_positions) || !kv_u32(g, "clip.text.projection_dim", hp.projection_dim) ||
                !kv_f32(g, "clip.text.attention.layer_norm_epsilon", hp.eps)) return nullptr;
            const GgufKV* tk = g.find("tokenizer.ggml.tokens");
            if (!tk || tk->type != GT_ARR || tk->arr_type != GT_STR) { set_err("key tokenizer.ggml.tokens not found in file"); return nullptr; }
            hp.n_vocab = (int32_t)tk->strs.size();
            for (int32_t i = 0; i < hp.n_vocab; i++) c->vocab.token_to_id[tk->strs[i]] = i;
            c->vocab.n = hp.n_vocab;
            t.present = true; t.hidden = hp.hidden_size; t.ff = hp.n_intermediate; t.heads = hp.n_head; t.layers = hp.n_layer;
            t.proj = hp.projection_dim; t.eps = hp.eps; t.n_vocab = hp.n_vocab; t.n_ctx = hp.num_positions;
            if (t.n_ctx <= 0 || t.n_ctx > 4096 || t.n_vocab <= 0) { set_err("text tower: bad context_length / vocabulary size"); return nullptr; }
            if (!check_geometry("text tower", t)) return nullptr;
            if (!upload_f32(c, g, "t.token_embd.weight", &t.tok, (int64_t)t.n_vocab * t.hidden) ||
                !upload_f32(c, g, "t.position_embd.weight", &t.pos, (int64_t)t.n_ctx * t.hidden) ||
                !upload_f32(c, g, "t.post_ln.weight", &t.post_g, t.hidden) || !upload_f32(c, g, "t.post_ln.bias", &t.post_b, t.hidden)) return nullptr;
            if (!make_linear(c, g, {"text_projection.weight"}, {}, t.proj_w)) return nullptr;
            if (!load_blocks(c, g, "t", t)) return nullptr;
            t.micro_batch = default_micro_batch(t.n_ctx);
            if (verbosity >= 2)
                printf("\n%s: text model hparams\nn_vocab            %d\nnum_positions      %d\nt_hidden_size      %d\nt_n_intermediate   %d\nt_projection_dim   %d\nt_n_head           %d\nt_n_layer          %d\n",
                       "clip_model_load", hp.n_vocab, hp.num_positions, hp.hidden_size, hp.n_intermediate, hp.projection_dim, hp.n_head, hp.n_layer);
        }
        if (c->has_vision) {
            Tower& t = c->vis;
            auto& hp = c->vhp;
            if (!kv_u32(g, "clip.vision.embedding_length", hp.hidden_size) || !kv_u32(g, "clip.vision.attention.head_count", hp.n_head) ||
                !kv_u32(g, "clip.vision.feed_forward_length", hp.n_intermediate) || !kv_u32(g, "clip.vision.block_count", hp.n_layer) ||
                !kv_u32(g, "clip.vision.image_size", hp.image_size) || !kv_u32(g, "clip.vision.patch_size", hp.patch_size) ||
                !kv_u32(g, "clip.vision.projection_dim", hp.projection_dim) || !kv_f32(g, "clip.vision.attention.layer_norm_epsilon", hp.eps)) return nullptr;
            const GgufKV* km = g.find("clip.vision.image_mean");
            const GgufKV* ks = g.find("clip.vision.image_std");
            if (!km || !ks || km->type != GT_ARR || ks->type != GT_ARR || km->arr_type != GT_F32 || ks->arr_type != GT_F32 || km->arr_n < 3 || ks->arr_n < 3) {
                set_err("key clip.vision.image_mean / image_std not found in file"); return nullptr;
            }
            memcpy(c->image_mean, km->raw + 12, 12);
            memcpy(c->image_std, ks->raw + 12, 12);
            t.present = true; t.hidden = hp.hidden_size; t.ff = hp.n_intermediate; t.heads = hp.n_head; t.layers = hp.n_layer;
            t.proj = hp.projection_dim; t.eps = hp.eps; t.image_size = hp.image_size; t.patch = hp.patch_size;
            if (t.patch <= 0 || t.image_size <= 0 || t.image_size > 4096 || t.image_size % t.patch) { set_err("image_size must be a positive multiple of patch_size"); return nullptr; }
            t.n_patches = (t.image_size / t.patch) * (t.image_size / t.patch);
            t.T = t.n_patches + 1;
            t.kpad = (3 * t.patch * t.patch + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
            if (!check_geometry("vision tower", t)) return nullptr;
            if (!make_linear(c, g, {"v.patch_embd.weight"}, {}, t.patch_w, t.kpad)) return nullptr;
            if (t.patch_w.N != t.hidden) { set_err("v.patch_embd.weight shape mismatch"); return nullptr; }
            if (!upload_f32(c, g, "v.class_embd", &t.class_embd, t.hidden) ||
                !upload_f32(c, g, "v.position_embd.weight", &t.pos, (int64_t)t.T * t.hidden) ||
                !upload_f32(c, g, "v.pre_ln.weight", &t.pre_g, t.hidden) || !upload_f32(c, g, "v.pre_ln.bias", &t.pre_b, t.hidden) ||
                !upload_f32(c, g, "v.post_ln.weight", &t.post_g, t.hidden) || !upload_f32(c, g, "v.post_ln.bias", &t.post_b, t.hidden)) return nullptr;
            if (!make_linear(c, g, {"visual_projection.weight"}, {}, t.proj_w)) return nullptr;
            if (!load_blocks(c, g, "v", t)) return nullptr;
            t.micro_batch = default_micro_batch(t.T);
            if (verbosity >= 2)
                printf("\n%s: vision model hparams\nimage_size         %d\npatch_size         %d\nv_hidden_size      %d\nv_n_intermediate   %d\nv_projection_dim   %d\nv_n_head           %d\nv_n_layer          %d\n",
                       "clip_model_load", hp.image_size, hp.patch_size, hp.hidden_size, hp.n_intermediate, hp.projection_dim, hp.n_head, hp.n_layer);
        }
        // 16-bit operand type of the towers: fp16 when any tower GEMM weight is stored unquantized (the reference rounds
        // those activations to fp16 too, ggml.c:11333-11349); bf16 for fully quantized towers unless overridden.
        const bool f16w = any_f16_gemm(c->vis) || any_f16_gemm(c->txt);
        c->operand_bf16 = !f16w;
        if (const char* op = getenv("CLIP_B200_OPERAND")) {
            if (!strcmp(op, "f16")) c->operand_bf16 = false;
            else if (!strcmp(op, "bf16") && !f16w) c->operand_bf16 = true;
        }
        if (verbosity >= 1) printf("%s: operand type: %s%s\n", "clip_model_load", c->operand_bf16 ? "bf16" : "fp16", c->debug_naive ? "  [DEBUG naive GEMM]" : "");
        if (cudaDeviceSynchronize() != cudaSuccess) { set_err("upload failed"); return nullptr; }
        return guard.release();
    }
}

// CLIP_B200_DEVICES = "all" | "0,1,2,..." -> devices mode (several GPUs behind ONE context); otherwise CLIP_B200_DEVICE (default 0),
// or LOCAL_RANK when CLIP_B200_DIST=env asks the context to join the launcher's ranks (torchrun / mpirun style environment).
static bool parse_devices(int ndev, std::vector<int>& out) {
    out.clear();
    const char* dl = getenv("CLIP_B200_DEVICES");
    if (dl && *dl) {
        if (!strcmp(dl, "all")) { for (int i = 0; i < ndev; i++) out.push_back(i); return true; }
        const char* p = dl;
        while (*p) {
            char* e = nullptr;
            const long v = strtol(p, &e, 10);
            if (e == p || v < 0 || v >= ndev) { set_err(std::string("CLIP_B200_DEVICES: bad device list '") + dl + "'"); return false; }
            if (std::find(out.begin(), out.end(), (int)v) != out.end()) { set_err("CLIP_B200_DEVICES: duplicate device"); return false; }
            out.push_back((int)v);
            p = (*e == ',') ? e + 1 : e;
            if (*e && *e != ',') { set_err(std::string("CLIP_B200_DEVICES: bad device list '") + dl + "'"); return false; }
        }
        if (out.empty()) { set_err("CLIP_B200_DEVICES: empty device list"); return false; }
        return true;
    }
    const char* dist_env = getenv("CLIP_B200_DIST");
    const char* dv = getenv("CLIP_B200_DEVICE");
    int d = dv ? atoi(dv) : 0;
    if (!dv && dist_env && !strcmp(dist_env, "env") && getenv("LOCAL_RANK")) d = atoi(getenv("LOCAL_RANK"));
    if (d < 0 || d >= ndev) { set_err("CLIP_B200_DEVICE out of range"); return false; }
    out.push_back(d);
    return true;
}

struct clip_ctx* clip_model_load(const char* fname, const int verbosity) {
    try {
        g_err.clear();
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err("no CUDA device: libclip_b200 has no CPU path"); return nullptr; }
        GgufFile g;
        std::string err;
        if (!g.parse(fname, err)) { set_err(err); return nullptr; }
        std::vector<int> devs;
        if (!parse_devices(ndev, devs)) return nullptr;
        std::unique_ptr<clip_ctx, void (*)(clip_ctx*)> leader(load_on_device(g, devs[0], verbosity), [](clip_ctx* p) { clip_free(p); });
        if (!leader) return nullptr;
        if (devs.size() > 1) {
            // devices mode: one replica + one NCCL communicator per GPU, all owned by the leader context
            leader->replicas.push_back(leader.get());
            for (size_t i = 1; i < devs.size(); i++) {
                clip_ctx* r = load_on_device(g, devs[i], 0);
                if (!r) return nullptr;
                r->is_replica = true;
                leader->replicas.push_back(r);
            }
            std::vector<DistComm> comms(devs.size());
            std::string derr;
            if (!dist_init_all(comms.data(), devs.data(), (int)devs.size(), derr)) { set_err("devices mode: " + derr); return nullptr; }
            for (size_t i = 0; i < devs.size(); i++) leader->replicas[i]->dist = comms[i];
            if (verbosity >= 1) printf("%s: devices mode: %zu GPUs, NCCL %d\n", __func__, devs.size(), dist_nccl_version());
            cudaSetDevice(leader->device);
        } else if (const char* de = getenv("CLIP_B200_DIST")) {
            if (!strcmp(de, "env") && getenv("RANK") && getenv("WORLD_SIZE") && atoi(getenv("WORLD_SIZE")) > 1) {
                if (!clip_b200_dist_init(leader.get(), atoi(getenv("RANK")), atoi(getenv("WORLD_SIZE")), nullptr)) return nullptr;
            }
        }
        return leader.release();
    } catch (const std::exception& e) {
        set_err(std::string("clip_model_load: ") + e.what());
        return nullptr;
    }
}

void clip_free(struct clip_ctx* c) {
    if (!c) return;
    for (size_t i = 1; i < c->replicas.size(); i++) { c->replicas[i]->is_replica = false; clip_free(c->replicas[i]); }
    c->replicas.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    dist_destroy(c->dist);
    preprocess_release(c);
    for (void* p : c->allocs) cudaFree(p);
    for (int i = 0; i < 2; i++) if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    for (auto& p : c->prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) { if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]); }
    for (int i = 0; i < 4; i++) if (c->marks[i]) cudaEventDestroy(c->marks[i]);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

struct clip_text_hparams* clip_get_text_hparams(struct clip_ctx* ctx) { return &ctx->thp; }
struct clip_vision_hparams* clip_get_vision_hparams(struct clip_ctx* ctx) { return &ctx->vhp; }

bool clip_tokenize(const struct clip_ctx* ctx, const char* text, struct clip_tokens* tokens) {
    if (!ctx || !ctx->has_text) { printf("This GGUF file seems to have no text encoder\n"); return false; }
    try {
        std::vector<int32_t> v = tokenize(ctx->vocab, text);
        tokens->size = v.size();
        tokens->data = new clip_vocab_id[v.size()];
        std::copy(v.begin(), v.end(), tokens->data);
        return true;
    } catch (...) { return false; }
}

struct clip_image_u8* clip_image_u8_make() { return new clip_image_u8(); }
struct clip_image_f32* clip_image_f32_make() { return new clip_image_f32(); }
void clip_image_u8_clean(struct clip_image_u8* img) { if (img && img->data) { delete[] img->data; img->data = nullptr; } }
void clip_image_f32_clean(struct clip_image_f32* res) { if (res && res->data) { delete[] res->data; res->data = nullptr; } }
void clip_image_u8_free(struct clip_image_u8* img) { clip_image_u8_clean(img); delete img; }
void clip_image_f32_free(struct clip_image_f32* res) { clip_image_f32_clean(res); delete res; }

bool clip_image_load_from_file(const char* fname, struct clip_image_u8* img) {
    std::vector<uint8_t> rgb;
    int nx = 0, ny = 0;
    if (!load_image_file(fname, rgb, nx, ny)) {
        set_err(std::string("clip_image_load_from_file: cannot decode '") + fname + "': supported formats are JPEG (Huffman baseline / progressive, 8-bit), PNG, "
                "BMP (uncompressed), GIF (first frame) and binary PGM / PPM; TGA / PSD / HDR / PIC and arithmetic-coded or 12-bit JPEG are not decoded by this library -- convert first");
        return false;
    }
    img->nx = nx; img->ny = ny; img->size = rgb.size();
    img->data = new uint8_t[rgb.size()];
    memcpy(img->data, rgb.data(), rgb.size());
    return true;
}

bool clip_image_preprocess(const struct clip_ctx* ctx, const struct clip_image_u8* img, struct clip_image_f32* res) {
    if (!ctx || !ctx->has_vision) { printf("This gguf file seems to have no vision encoder\n"); return false; }
    const int S = ctx->vhp.image_size;
    res->nx = S; res->ny = S; res->size = (size_t)3 * S * S;
    res->data = new float[res->size]();
    if (!preprocess_image(img->data, img->nx, img->ny, S, ctx->image_mean, ctx->image_std, res->data)) {
        delete[] res->data; res->data = nullptr;
        return false;
    }
    return true;
}

void clip_image_batch_preprocess(const struct clip_ctx* ctx, const int n_threads, const struct clip_image_u8_batch* in,
                                 struct clip_image_f32_batch* out) {
    // CLIP_B200_PREPROCESS=device: the same call runs K6a/K6b on the GPU (bit-identical results, new[]-owned buffers as below); a
    // failure is reported and leaves the outputs empty -- no silent detour through the host loop
    const char* where = getenv("CLIP_B200_PREPROCESS");
    if (where && strcmp(where, "device") == 0 && in->size > 0) {
        if (!clip_b200_image_batch_preprocess_device(ctx, in, out)) {
            fprintf(stderr, "clip_image_batch_preprocess (device): %s\n", clip_b200_last_error());
            for (size_t i = 0; i < in->size; i++) { out->data[i].data = nullptr; out->data[i].size = 0; }
        }
        return;
    }
    out->size = in->size;
    const size_t n = in->size;
    int nt = std::max(1, std::min<int>(n_threads, (int)n));
    if (nt == 1) { for (size_t i = 0; i < n; i++) clip_image_preprocess(ctx, &in->data[i], &out->data[i]); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=]() { for (size_t i = t; i < n; i += nt) clip_image_preprocess(ctx, &in->data[i], &out->data[i]); });
    for (auto& t : th) t.join();
}

// ---- N1: preprocess on the device -------------------------------------------------------------------------
// Shared driver: micro-batches of raw u8 images -> pixel staging buffer (preprocess.cu) -> optional vision forward.
static bool preprocess_chunks(clip_ctx* c, const clip_image_u8* imgs, size_t n, float* h_pixels_out, float* vec, bool normalize) {
    Tower& tw = c->vis;
    if (!ensure_ws(c, tw, tw.micro_batch, tw.T, true)) return false;
    if (vec && !ensure_out(c, n * tw.proj)) return false;
    size_t n_chunks, chunk;
    chunking(n, tw.micro_batch, n_chunks, chunk);
    const size_t per = (size_t)tw.image_size * tw.image_size * 3;
    CK(cudaEventRecord(c->ev_t0, c->stream));
    size_t ci = 0;
    for (size_t i0 = 0; i0 < n; i0 += chunk, ci++) {
        const int nb = (int)std::min(chunk, n - i0), b = (int)(ci & 1);
        if (ci >= 2) CK(cudaEventSynchronize(c->ev_consumed[b]));      // arena b is rewritten on the host below
        std::string err;
        if (!preprocess_device(c, imgs + i0, nb, b, tw.ws.pixels[b], c->copy_stream, c->ev_copied[b], c->stream, err)) { set_err(err); return false; }
        if (h_pixels_out) CK(cudaMemcpyAsync(h_pixels_out + i0 * per, tw.ws.pixels[b], (size_t)nb * per * 4, cudaMemcpyDeviceToHost, c->stream));
        if (vec && !vision_forward(c, tw.ws.pixels[b], nb, c->d_out + i0 * tw.proj, normalize)) return false;
        CK(cudaEventRecord(c->ev_consumed[b], c->stream));
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    if (vec) CK(cudaMemcpyAsync(vec, c->d_out, n * tw.proj * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_time(c);
}

bool clip_b200_image_batch_encode_u8(const struct clip_ctx* cctx, const struct clip_image_u8_batch* imgs, float* vec, const bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { set_err("no vision encoder"); return false; }
    if (!imgs || imgs->size == 0) return true;
    if (c->replicas.size() > 1 && imgs->size >= c->replicas.size()) {
        const int d = c->vis.proj;
        return for_each_replica(c, imgs->size, [&](clip_ctx* r, int, size_t lo, size_t hi) {
            if (hi == lo) return true;
            std::lock_guard<std::mutex> lk(r->mu);
            CK(cudaSetDevice(r->device));
            return preprocess_chunks(r, imgs->data + lo, hi - lo, nullptr, vec + lo * d, normalize);
        });
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return preprocess_chunks(c, imgs->data, imgs->size, nullptr, vec, normalize);
}

bool clip_b200_image_batch_preprocess_device(const struct clip_ctx* cctx, const struct clip_image_u8_batch* in, struct clip_image_f32_batch* out) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { set_err("no vision encoder"); return false; }
    out->size = in->size;
    const size_t n = in->size;
    if (n == 0) return true;
    const int S = c->vhp.image_size;
    const size_t per = (size_t)3 * S * S;
    std::vector<float> host(n * per);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CK(cudaSetDevice(c->device));
        if (!preprocess_chunks(c, in->data, n, host.data(), nullptr, false)) return false;
    }
    for (size_t i = 0; i < n; i++) {        // same ownership convention as clip_image_preprocess: data is new[]-allocated here
        out->data[i].nx = S; out->data[i].ny = S; out->data[i].size = per;
        out->data[i].data = new float[per];
        memcpy(out->data[i].data, host.data() + i * per, per * 4);
    }
    return true;
}

bool clip_b200_image_encode_device(const struct clip_ctx* cctx, const void* d_pixels, size_t n, void* d_vec, bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { set_err("no vision encoder"); return false; }
    if (n == 0) return true;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!image_encode_device_locked(c, (const float*)d_pixels, n, (float*)d_vec, normalize)) return false;
    return sync_and_time(c);
}

// One GPU: n host images -> embeddings at d_vec (device, may be null) and / or vec (host, may be null).  Host pixels travel on the
// copy stream, double-buffered against the kernels of the previous micro-batch.  Pinned / registered caller buffers are copied
// directly; PAGEABLE ones (what clip_image_preprocess returns: new float[]) are first gathered into a pinned arena by host threads,
// in sub-batches so the H2D of sub-batch k overlaps the memcpy of k+1 -- a pageable cudaMemcpyAsync would be staged synchronously
// by the driver, one 602 KB image at a time.
static bool image_batch_encode_one(clip_ctx* c, int n_threads, const clip_image_f32* imgs, size_t n, float* vec, float* d_vec, bool normalize,
                                   bool sync) {
    Tower& tw = c->vis;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!ensure_ws(c, tw, tw.micro_batch, tw.T, true)) return false;
    if (!d_vec) { if (!ensure_out(c, n * tw.proj)) return false; d_vec = c->d_out; }
    const std::vector<size_t> bounds = host_chunk_bounds(n, tw.micro_batch);
    const size_t per = (size_t)tw.image_size * tw.image_size * 3;
    bool pinned = true;
    for (size_t i = 0; i < n && pinned; i++) pinned = host_ptr_is_pinned(imgs[i].data);
    if (!pinned && c->h_stage_cap < (size_t)tw.micro_batch * per) {
        for (int b = 0; b < 2; b++) {
            if (c->h_stage[b]) { cudaFreeHost(c->h_stage[b]); c->h_stage[b] = nullptr; }
            CK(cudaMallocHost((void**)&c->h_stage[b], (size_t)tw.micro_batch * per * 4));
        }
        c->h_stage_cap = (size_t)tw.micro_batch * per;
    }
    const int nt = std::max(1, std::min({std::max(n_threads, 4), 16, (int)std::thread::hardware_concurrency()}));
    CK(cudaEventRecord(c->ev_t0, c->stream));
    for (size_t ci = 0; ci + 1 < bounds.size(); ci++) {
        const size_t i0 = bounds[ci];
        const int nb = (int)(bounds[ci + 1] - i0), b = (int)(ci & 1);
        if (ci >= 2) CK(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
        if (pinned) {
            for (int j = 0; j < nb; j++)
                CK(cudaMemcpyAsync(tw.ws.pixels[b] + (size_t)j * per, imgs[i0 + j].data, per * 4, cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            if (ci >= 2) CK(cudaEventSynchronize(c->ev_copied[b]));      // the arena's previous H2D has left the host buffer
            const int sub = 32;
            for (int j0 = 0; j0 < nb; j0 += sub) {
                const int ns = std::min(sub, nb - j0);
                float* dst = c->h_stage[b] + (size_t)j0 * per;
                const clip_image_f32* src = imgs + i0 + j0;
                const int t_use = std::min(nt, ns);
                if (t_use <= 1) { for (int j = 0; j < ns; j++) memcpy(dst + (size_t)j * per, src[j].data, per * 4); }
                else {
                    std::vector<std::thread> th;
                    for (int t = 0; t < t_use; t++)
                        th.emplace_back([=]() { for (int j = t; j < ns; j += t_use) memcpy(dst + (size_t)j * per, src[j].data, per * 4); });
                    for (auto& t : th) t.join();
                }
                CK(cudaMemcpyAsync(tw.ws.pixels[b] + (size_t)j0 * per, dst, (size_t)ns * per * 4, cudaMemcpyHostToDevice, c->copy_stream));
            }
        }
        CK(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        if (!vision_forward(c, tw.ws.pixels[b], nb, d_vec + i0 * tw.proj, normalize)) return false;
        CK(cudaEventRecord(c->ev_consumed[b], c->stream));
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    if (vec) CK(cudaMemcpyAsync(vec, d_vec, n * tw.proj * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync ? sync_and_time(c) : true;
}

static bool check_images(clip_ctx* c, const clip_image_f32* imgs, size_t n, const char* who) {
    for (size_t i = 0; i < n; i++)
        if (imgs[i].nx != c->vis.image_size || imgs[i].ny != c->vis.image_size || !imgs[i].data) { set_err(std::string(who) + ": every image must be image_size x image_size"); return false; }
    return true;
}

bool clip_image_batch_encode(const struct clip_ctx* cctx, const int n_threads, const struct clip_image_f32_batch* imgs, float* vec,
                             const bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { printf("This gguf file seems to have no vision encoder\n"); return false; }
    const size_t n = imgs->size;
    if (n == 0) return true;
    if (!check_images(c, imgs->data, n, "clip_image_batch_encode")) return false;
    if (c->replicas.size() > 1 && n >= c->replicas.size()) {
        // devices mode: contiguous shards, every GPU copies its slice of the result straight into the caller's vec (no exchange needed)
        const int d = c->vis.proj;
        return for_each_replica(c, n, [&](clip_ctx* r, int, size_t lo, size_t hi) {
            return hi == lo || image_batch_encode_one(r, n_threads, imgs->data + lo, hi - lo, vec + lo * d, nullptr, normalize, true);
        });
    }
    return image_batch_encode_one(c, n_threads, imgs->data, n, vec, nullptr, normalize, true);
}

bool clip_image_encode(const struct clip_ctx* ctx, const int n_threads, struct clip_image_f32* img, float* vec, const bool normalize) {
    if (!ctx || !ctx->has_vision) { printf("This gguf file seems to have no vision encoder\n"); return false; }
    clip_image_f32_batch b;
    b.data = img; b.size = 1;
    return clip_image_batch_encode(ctx, n_threads, &b, vec, normalize);
}

// n sequences padded to T tokens on the host (h_ids / h_last) or on the device (d_ids / d_lens) -> d_out_dev (device) and / or vec_host
static bool text_encode_impl(clip_ctx* c, const int32_t* h_ids, const int32_t* h_last, const int32_t* d_ids, const int32_t* d_lens,
                             size_t n, int T, float* vec_host, float* vec_dev, bool normalize, bool sync = true) {
    Tower& tw = c->txt;
    // the padded length is fixed per context so that one workspace serves every call
    if (T > tw.n_ctx) { set_err("sequence longer than context_length"); return false; }
    if (!ensure_ws(c, tw, tw.micro_batch, tw.n_ctx, false)) return false;
    float* d_out = vec_dev;
    if (!d_out) { if (!ensure_out(c, n * tw.proj)) return false; d_out = c->d_out; }
    size_t n_chunks, chunk;
    chunking(n, tw.micro_batch, n_chunks, chunk);
    CK(cudaEventRecord(c->ev_t0, c->stream));
    for (size_t i0 = 0; i0 < n; i0 += chunk) {
        const int nb = (int)std::min(chunk, n - i0);
        if (h_ids) {
            CK(cudaMemcpyAsync(tw.ws.ids, h_ids + i0 * T, (size_t)nb * T * 4, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(tw.ws.last, h_last + i0, (size_t)nb * 4, cudaMemcpyHostToDevice, c->stream));
        } else {
            CK(cudaMemcpyAsync(tw.ws.ids, d_ids + i0 * T, (size_t)nb * T * 4, cudaMemcpyDeviceToDevice, c->stream));
            launch_fill_last(d_lens ? d_lens + i0 : nullptr, tw.ws.last, nb, T, c->stream);
        }
        if (!text_forward(c, nb, T, d_out + i0 * tw.proj, normalize)) return false;
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    if (vec_host) CK(cudaMemcpyAsync(vec_host, d_out, n * tw.proj * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync ? sync_and_time(c) : true;
}

// host token sequences -> padded id matrix + EOT row indices.  Padding to a multiple of 8 tokens is causally invisible to the EOT row.
static bool pack_tokens(clip_ctx* c, const clip_tokens* seqs, size_t n, std::vector<int32_t>& ids, std::vector<int32_t>& last, int& T_out) {
    size_t T = 0;
    for (size_t i = 0; i < n; i++) {
        if (seqs[i].size == 0 || !seqs[i].data) { set_err("empty token sequence"); return false; }
        T = std::max(T, seqs[i].size);
    }
    if ((int)T > c->txt.n_ctx) { set_err("token sequence longer than context_length (" + std::to_string(c->txt.n_ctx) + ")"); return false; }
    T = std::min<size_t>((T + 7) / 8 * 8, (size_t)c->txt.n_ctx);
    ids.assign(n * T, 0);
    last.resize(n);
    for (size_t i = 0; i < n; i++) {
        memcpy(ids.data() + i * T, seqs[i].data, seqs[i].size * 4);
        last[i] = (int32_t)seqs[i].size - 1;     // the reference selects row N-1 (clip.cpp:1154-1155)
    }
    T_out = (int)T;
    return true;
}

static bool text_batch_encode_one(clip_ctx* c, const clip_tokens* seqs, size_t n, float* vec, float* d_vec, bool normalize, bool sync) {
    std::vector<int32_t> ids, last;
    int T = 0;
    if (!pack_tokens(c, seqs, n, ids, last, T)) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // ids / last are pageable vectors: the async copies inside return only after the driver has staged them, so they may go out of
    // scope when this function returns even with sync == false
    return text_encode_impl(c, ids.data(), last.data(), nullptr, nullptr, n, T, vec, d_vec, normalize, sync);
}

bool clip_text_batch_encode(const struct clip_ctx* cctx, const int n_threads, const struct clip_tokens* seqs, const size_t n, float* vec,
                            const bool normalize) {
    (void)n_threads;
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_text) { printf("This GGUF file seems to have no text encoder\n"); return false; }
    if (n == 0) return true;
    if (c->replicas.size() > 1 && n >= 8 * c->replicas.size()) {
        const int d = c->txt.proj;
        return for_each_replica(c, n, [&](clip_ctx* r, int, size_t lo, size_t hi) {
            return hi == lo || text_batch_encode_one(r, seqs + lo, hi - lo, vec + lo * d, nullptr, normalize, true);
        });
    }
    return text_batch_encode_one(c, seqs, n, vec, nullptr, normalize, true);
}

bool clip_text_encode(const struct clip_ctx* ctx, const int n_threads, const struct clip_tokens* tokens, float* vec, const bool normalize) {
    if (!ctx || !ctx->has_text) { printf("This GGUF file seems to have no text encoder\n"); return false; }
    return clip_text_batch_encode(ctx, n_threads, tokens, 1, vec, normalize);
}

bool clip_b200_text_encode_device(const struct clip_ctx* cctx, const void* d_ids, const void* d_lens, size_t n, int seq_len, void* d_vec,
                                  bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_text) { set_err("no text encoder"); return false; }
    if (n == 0) return true;
    if (seq_len < 1 || seq_len > c->txt.n_ctx) { set_err("clip_b200_text_encode_device: seq_len must be in [1, context_length]"); return false; }
    if (!d_ids || !d_vec) { set_err("clip_b200_text_encode_device: null device pointer"); return false; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // lengths stay on the device: a tiny kernel turns them into EOT-row indices per micro-batch (clamped to [1, seq_len])
    return text_encode_impl(c, nullptr, nullptr, (const int32_t*)d_ids, (const int32_t*)d_lens, n, seq_len, nullptr, (float*)d_vec, normalize);
}

float clip_similarity_score(const float* vec1, const float* vec2, const int vec_dim) {
    float dot = 0.0f;
    for (int i = 0; i < vec_dim; i++) dot += vec1[i] * vec2[i];
    return dot;
}

bool softmax_with_sorting(float* arr, const int length, float* sorted_scores, int* indices) {
    if (length <= 0) return false;
    double sum = 0.0;
    for (int i = 0; i < length; i++) { arr[i] = (float)(exp(arr[i]) + 1e-9); sum += arr[i]; }
    std::vector<int> idx(length);
    for (int i = 0; i < length; i++) { arr[i] = (float)(arr[i] / sum); idx[i] = i; }
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return arr[a] > arr[b]; });
    for (int i = 0; i < length; i++) { sorted_scores[i] = arr[idx[i]]; indices[i] = idx[i]; }
    return true;
}

bool clip_compare_text_and_image(const struct clip_ctx* ctx, const int n_threads, const char* text, const struct clip_image_u8* image,
                                 float* score) {
    if (!ctx || !(ctx->has_text && ctx->has_vision)) { printf("clip_compare_text_and_image function can only be used with two-tower models\n"); return false; }
    const int d = ctx->vhp.projection_dim;
    std::vector<float> iv(d), tv(d);
    clip_tokens tk{nullptr, 0};
    if (!clip_tokenize(ctx, text, &tk)) return false;
    const bool ok_t = clip_text_encode(ctx, n_threads, &tk, tv.data(), true);
    delete[] tk.data;
    if (!ok_t) return false;
    clip_image_f32 res{0, 0, nullptr, 0};
    if (!clip_image_preprocess(ctx, image, &res)) return false;
    const bool ok_i = clip_image_encode(ctx, n_threads, &res, iv.data(), true);
    clip_image_f32_clean(&res);
    if (!ok_i) return false;
    *score = clip_similarity_score(iv.data(), tv.data(), d);
    return true;
}

bool clip_zero_shot_label_image(struct clip_ctx* ctx, const int n_threads, const struct clip_image_u8* input_img, const char** labels,
                                const size_t n_labels, float* scores, int* indices) {
    if (!ctx || !(ctx->has_text && ctx->has_vision)) { printf("clip_zero_shot_label_image function can only be used with two-tower models\n"); return false; }
    const int d = ctx->vhp.projection_dim;
    clip_image_f32 res{0, 0, nullptr, 0};
    if (!clip_image_preprocess(ctx, input_img, &res)) return false;
    std::vector<float> iv(d);
    const bool ok_i = clip_image_encode(ctx, n_threads, &res, iv.data(), false);
    clip_image_f32_clean(&res);
    if (!ok_i) return false;
    std::vector<clip_tokens> tks(n_labels);
    for (size_t i = 0; i < n_labels; i++) if (!clip_tokenize(ctx, labels[i], &tks[i])) return false;
    std::vector<float> tv(n_labels * d), sims(n_labels);
    const bool ok_t = clip_text_batch_encode(ctx, n_threads, tks.data(), n_labels, tv.data(), false);
    for (auto& t : tks) delete[] t.data;
    if (!ok_t) return false;
    for (size_t i = 0; i < n_labels; i++) sims[i] = clip_similarity_score(iv.data(), tv.data() + i * d, d);
    return softmax_with_sorting(sims.data(), (int)n_labels, scores, indices);
}

// =====================================================================================================================
// Scoring on the device (search.cu): similarity matrix -> [softmax_with_sorting arithmetic] -> top-k, results to the host
// =====================================================================================================================
// d_q [nq, d] x d_db [ndb, d] -> host scores / indices [nq, k], best first (ties: lower index first).  softmax: p = (exp(s)+1e-9)/sum
// over the whole row first (clip.cpp:1591-1622); otherwise the raw dot products are ranked (nearest-neighbour search).
// The stream is synchronised on return.  Caller holds c->mu and has set the device.
static bool score_topk_locked(clip_ctx* c, const float* d_q, size_t nq, const float* d_db, size_t ndb, int d, int k, bool softmax,
                              float* scores, int* indices) {
    if (nq == 0 || ndb == 0) return true;
    if (ndb > 0x7fffffffull || nq > 0x7fffffffull) { set_err("scoring: more than 2^31 rows"); return false; }
    if (k <= 0 || (size_t)k > ndb) k = (int)ndb;
    const int SL = topk_slice();
    const size_t budget = (size_t)32 << 20;                               // floats of similarity scratch per pass (128 MB)
    size_t CH = std::min(ndb, std::max<size_t>(budget / std::min<size_t>(nq, 256), (size_t)SL));
    if (softmax) CH = ndb;                                                // the normalisation needs whole rows
    const size_t RB = std::max<size_t>(1, std::min<size_t>({nq, budget / CH + 1, (size_t)32768}));
    size_t nsl = 0;
    for (size_t c0 = 0; c0 < ndb; c0 += CH) nsl += (std::min(CH, ndb - c0) + SL - 1) / SL;
    const int kk = std::min(k, SL);
    const bool device_select = nsl == 1 || k <= 1024;                    // each stage must shrink the candidate set
    if (!ensure_buf(c, &c->d_logits, &c->d_logits_cap, RB * CH)) return false;
    const size_t cand_ld = nsl * (size_t)kk;
    if (device_select) {
        size_t cap0 = c->d_cand_cap, cap1 = c->d_cand_cap, cap2 = c->d_cand_cap, cap3 = c->d_cand_cap;
        if (!ensure_buf(c, &c->d_cand_v[0], &cap0, RB * cand_ld) || !ensure_buf(c, &c->d_cand_v[1], &cap1, RB * cand_ld) ||
            !ensure_buf(c, &c->d_cand_i[0], &cap2, RB * cand_ld) || !ensure_buf(c, &c->d_cand_i[1], &cap3, RB * cand_ld)) return false;
        c->d_cand_cap = std::min(std::min(cap0, cap1), std::min(cap2, cap3));
    }
    std::vector<float> hrow;
    std::vector<int> hidx;
    for (size_t r0 = 0; r0 < nq; r0 += RB) {
        const int rows = (int)std::min(RB, nq - r0);
        size_t sl0 = 0;
        for (size_t c0 = 0; c0 < ndb; c0 += CH) {
            const int cols = (int)std::min(CH, ndb - c0);
            { Scope s(c, K_OTHER); launch_similarity(d_q + r0 * d, d_db + c0 * d, c->d_logits, rows, cols, d, CH, c->stream); }
            if (softmax) { Scope s(c, K_OTHER); launch_softmax_ref(c->d_logits, rows, cols, CH, c->stream); }
            if (device_select) {
                Scope s(c, K_OTHER);
                launch_topk_stage(c->d_logits, nullptr, CH, rows, cols, (int)c0, kk, c->d_cand_v[0], c->d_cand_i[0], cand_ld, (int)(sl0 * kk), c->stream);
                sl0 += ((size_t)cols + SL - 1) / SL;
            } else {
                // full ranking of rows longer than one sort slice (top_k > 1024): ranked on the host, as the reference does
                hrow.resize((size_t)rows * cols);
                CK(cudaMemcpy2DAsync(hrow.data(), (size_t)cols * 4, c->d_logits, CH * 4, (size_t)cols * 4, rows, cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                hidx.resize(cols);
                for (int i = 0; i < rows; i++) {
                    const float* r = hrow.data() + (size_t)i * cols;
                    std::iota(hidx.begin(), hidx.end(), 0);
                    std::partial_sort(hidx.begin(), hidx.begin() + k, hidx.end(), [&](int a, int b) { return r[a] > r[b] || (r[a] == r[b] && a < b); });
                    for (int j = 0; j < k; j++) { scores[(r0 + i) * k + j] = r[hidx[j]]; indices[(r0 + i) * k + j] = hidx[j]; }
                }
            }
        }
        if (!device_select) continue;
        int cur = 0;
        size_t cur_n = cand_ld;
        while (cur_n > (size_t)SL || (cur_n > (size_t)kk && nsl > 1)) {      // merge stages: candidates of all slices compete again
            const size_t nslices = (cur_n + SL - 1) / SL;
            { Scope s(c, K_OTHER); launch_topk_stage(c->d_cand_v[cur], c->d_cand_i[cur], cand_ld, rows, (int)cur_n, 0, kk, c->d_cand_v[cur ^ 1], c->d_cand_i[cur ^ 1], cand_ld, 0, c->stream); }
            cur ^= 1;
            cur_n = nslices * (size_t)kk;
            if (nslices == 1) break;
        }
        CK(cudaMemcpy2DAsync(scores + r0 * k, (size_t)k * 4, c->d_cand_v[cur], cand_ld * 4, (size_t)k * 4, rows, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpy2DAsync(indices + r0 * k, (size_t)k * 4, c->d_cand_i[cur], cand_ld * 4, (size_t)k * 4, rows, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaGetLastError());
    return true;
}

bool clip_b200_zero_shot_batch(const struct clip_ctx* cctx, const void* d_img, size_t n_img, const void* d_txt, size_t n_txt, float* scores,
                               int* indices, int top_k) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int d = c->has_vision ? c->vhp.projection_dim : c->thp.projection_dim;
    return score_topk_locked(c, (const float*)d_img, n_img, (const float*)d_txt, n_txt, d, top_k, true, scores, indices);
}

bool clip_b200_topk_search(const struct clip_ctx* cctx, const void* d_queries, size_t n_queries, const void* d_db, size_t n_db, int top_k,
                           float* scores, int* indices) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int d = c->has_vision ? c->vhp.projection_dim : c->thp.projection_dim;
    return score_topk_locked(c, (const float*)d_queries, n_queries, (const float*)d_db, n_db, d, top_k, false, scores, indices);
}

// =====================================================================================================================
// Multi-GPU inside the library (dist.h).  ranks mode: one process per GPU joined by clip_b200_dist_init; devices mode: one
// process, CLIP_B200_DEVICES at clip_model_load.  The ONLY exchange of the path is one NCCL all-gather of final embeddings.
// =====================================================================================================================
bool clip_b200_dist_unique_id(void* out128) {
    std::string e;
    if (!dist_unique_id(out128, e)) { set_err(e); return false; }
    return true;
}

bool clip_b200_dist_init_with_id(struct clip_ctx* c, int rank, int world, const void* id128) {
    if (!c || !id128) { set_err("clip_b200_dist_init: null argument"); return false; }
    if (c->replicas.size() > 1) { set_err("clip_b200_dist_init: this context is already in devices mode (CLIP_B200_DEVICES)"); return false; }
    if (c->dist.comm) { set_err("clip_b200_dist_init: already initialised"); return false; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    std::string e;
    if (!dist_init_rank(c->dist, rank, world, id128, e)) { set_err(e); return false; }
    return true;
}

bool clip_b200_dist_init(struct clip_ctx* c, int rank, int world, const char* rendezvous) {
    if (!c) { set_err("clip_b200_dist_init: null context"); return false; }
    if (world == 1) return true;
    unsigned char id[128];
    std::string e;
    if (!dist_rendezvous_id(rank, world, rendezvous, id, e)) { set_err(e); return false; }
    const bool ok = clip_b200_dist_init_with_id(c, rank, world, id);
    dist_rendezvous_done(rank);
    return ok;
}

int clip_b200_dist_rank(const struct clip_ctx* c) { return c ? c->dist.rank : 0; }
int clip_b200_dist_world(const struct clip_ctx* c) { return c ? (c->replicas.size() > 1 ? 1 : c->dist.world) : 1; }
int clip_b200_device_count(const struct clip_ctx* c) { return c ? std::max<int>(1, (int)c->replicas.size()) : 0; }
int clip_b200_nccl_version(void) { return dist_nccl_version(); }
int clip_b200_cuda_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

static bool ranks_mode(const clip_ctx* c) { return c->replicas.size() <= 1 && c->dist.comm && c->dist.world > 1; }

static bool ensure_scalars(clip_ctx* c) {
    if (c->d_scalars) return true;
    return dev_alloc(c, &c->d_scalars, 16);
}

bool clip_b200_dist_all_gather(const struct clip_ctx* cctx, const void* d_send, void* d_recv, size_t bytes_per_rank) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!ranks_mode(c)) {
        if (d_send != d_recv) CK(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        std::string e;
        if (!dist_all_gather(c->dist, d_send, d_recv, bytes_per_rank, c->stream, e)) { set_err(e); return false; }
    }
    CK(cudaStreamSynchronize(c->stream));
    return true;
}

bool clip_b200_dist_max_f64(const struct clip_ctx* cctx, double* vals, int n) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || n < 0 || n > 16) { set_err("clip_b200_dist_max_f64: n must be in [0, 16]"); return false; }
    if (!ranks_mode(c) || n == 0) return true;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!ensure_scalars(c)) return false;
    CK(cudaMemcpyAsync(c->d_scalars, vals, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    std::string e;
    if (!dist_all_reduce_max_f64(c->dist, c->d_scalars, (size_t)n, c->stream, e)) { set_err(e); return false; }
    CK(cudaMemcpyAsync(vals, c->d_scalars, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return true;
}

bool clip_b200_dist_barrier(const struct clip_ctx* cctx) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c) return false;
    if (ranks_mode(c)) { double v = 0.0; if (!clip_b200_dist_max_f64(c, &v, 1)) return false; }
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());
    return true;
}

// ranks mode: every rank encodes ITS n_local items and ends with ALL world * n_local embeddings, rank-major.  The last kernel of
// the tower (K5 l2norm) writes straight into this rank's slot of the gather buffer; the all-gather runs in place on the launch stream.
static bool gather_in_place(clip_ctx* c, float* d_all, size_t n_local, int d) {
    if (!ranks_mode(c)) return true;
    std::string e;
    const size_t bytes = n_local * (size_t)d * 4;
    if (!dist_all_gather(c->dist, (const char*)d_all + (size_t)c->dist.rank * bytes, d_all, bytes, c->stream, e)) { set_err(e); return false; }
    return true;
}

bool clip_b200_image_encode_device_all(const struct clip_ctx* cctx, const void* d_pixels, size_t n_local, void* d_vec_all, bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { set_err("no vision encoder"); return false; }
    if (n_local == 0) return true;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int d = c->vis.proj;
    float* slot = (float*)d_vec_all + (ranks_mode(c) ? (size_t)c->dist.rank * n_local * d : 0);
    if (!image_encode_device_locked(c, (const float*)d_pixels, n_local, slot, normalize)) return false;
    if (!gather_in_place(c, (float*)d_vec_all, n_local, d)) return false;
    CK(cudaEventRecord(c->ev_t1, c->stream));
    return sync_and_time(c);
}

bool clip_b200_image_batch_encode_all(const struct clip_ctx* cctx, const int n_threads, const struct clip_image_f32_batch* imgs, float* vec_all,
                                      const bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { set_err("no vision encoder"); return false; }
    if (!ranks_mode(c)) return clip_image_batch_encode(c, n_threads, imgs, vec_all, normalize);
    const size_t n = imgs->size;
    if (n == 0) { set_err("clip_b200_image_batch_encode_all: every rank must pass the same, non-zero number of images"); return false; }
    if (!check_images(c, imgs->data, n, "clip_b200_image_batch_encode_all")) return false;
    const int d = c->vis.proj, w = c->dist.world;
    { std::lock_guard<std::mutex> lk(c->mu); CK(cudaSetDevice(c->device)); if (!ensure_out(c, (size_t)w * n * d)) return false; }
    if (!image_batch_encode_one(c, n_threads, imgs->data, n, nullptr, c->d_out + (size_t)c->dist.rank * n * d, normalize, false)) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!gather_in_place(c, c->d_out, n, d)) return false;
    CK(cudaEventRecord(c->ev_t1, c->stream));
    CK(cudaMemcpyAsync(vec_all, c->d_out, (size_t)w * n * d * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_time(c);
}

bool clip_b200_text_batch_encode_all(const struct clip_ctx* cctx, const int n_threads, const struct clip_tokens* seqs, const size_t n, float* vec_all,
                                     const bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_text) { set_err("no text encoder"); return false; }
    if (!ranks_mode(c)) return clip_text_batch_encode(c, n_threads, seqs, n, vec_all, normalize);
    if (n == 0) { set_err("clip_b200_text_batch_encode_all: every rank must pass the same, non-zero number of sequences"); return false; }
    const int d = c->txt.proj, w = c->dist.world;
    { std::lock_guard<std::mutex> lk(c->mu); CK(cudaSetDevice(c->device)); if (!ensure_out(c, (size_t)w * n * d)) return false; }
    if (!text_batch_encode_one(c, seqs, n, nullptr, c->d_out + (size_t)c->dist.rank * n * d, normalize, false)) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!gather_in_place(c, c->d_out, n, d)) return false;
    CK(cudaEventRecord(c->ev_t1, c->stream));
    CK(cudaMemcpyAsync(vec_all, c->d_out, (size_t)w * n * d * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_time(c);
}

bool clip_b200_text_encode_device_all(const struct clip_ctx* cctx, const void* d_ids, const void* d_lens, size_t n_local, int seq_len,
                                      void* d_vec_all, bool normalize) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_text) { set_err("no text encoder"); return false; }
    if (n_local == 0) return true;
    if (seq_len < 1 || seq_len > c->txt.n_ctx || !d_ids || !d_vec_all) { set_err("clip_b200_text_encode_device_all: bad arguments"); return false; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int d = c->txt.proj;
    float* slot = (float*)d_vec_all + (ranks_mode(c) ? (size_t)c->dist.rank * n_local * d : 0);
    if (!text_encode_impl(c, nullptr, nullptr, (const int32_t*)d_ids, (const int32_t*)d_lens, n_local, seq_len, nullptr, slot, normalize, false)) return false;
    if (!gather_in_place(c, (float*)d_vec_all, n_local, d)) return false;
    CK(cudaEventRecord(c->ev_t1, c->stream));
    return sync_and_time(c);
}

// Batched zero-shot labelling (BASELINE.json configs[4]; per image the arithmetic of clip_zero_shot_label_image, clip.cpp:1624-1659):
//   single GPU   : n images x n_labels labels
//   devices mode : images AND labels are sharded over the GPUs, the label embeddings are all-gathered (NCCL), every GPU scores its images
//   ranks mode   : the caller passes THIS rank's images and THIS rank's label shard (same count on every rank); label j of rank r has
//                  global index r * n_labels + j; scores / indices cover this rank's images against all world * n_labels labels
// scores / indices: [n images, top_k], best first.  The embeddings never leave the GPUs.
bool clip_b200_zero_shot_images(const struct clip_ctx* cctx, const int n_threads, const struct clip_image_f32_batch* imgs,
                                const struct clip_tokens* labels, const size_t n_labels, const bool normalize, int top_k, float* scores,
                                int* indices) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !(c->has_text && c->has_vision)) { set_err("clip_b200_zero_shot_images needs a two-tower model"); return false; }
    const size_t n = imgs ? imgs->size : 0;
    if (n == 0 || n_labels == 0) return n == 0;
    if (!check_images(c, imgs->data, n, "clip_b200_zero_shot_images")) return false;
    const int d = c->vis.proj;
    if (c->replicas.size() > 1) {
        const int w = (int)c->replicas.size();
        const size_t per = (n_labels + w - 1) / w, tot = (size_t)w * per;      // equal slots; only the last ones may be partly empty
        if (top_k <= 0 || (size_t)top_k > n_labels) top_k = (int)n_labels;
        // phase 1: every GPU encodes its image shard and its label shard (slot r of its own gather buffer)
        if (!for_each_replica(c, n, [&](clip_ctx* r, int ri, size_t lo, size_t hi) {
                {
                    std::lock_guard<std::mutex> lk(r->mu);
                    CK(cudaSetDevice(r->device));
                    if (!ensure_buf(r, &r->d_emb[0], &r->d_emb_cap[0], std::max<size_t>(hi - lo, 1) * d) || !ensure_buf(r, &r->d_emb[1], &r->d_emb_cap[1], tot * d)) return false;
                    CK(cudaMemsetAsync(r->d_emb[1], 0, tot * d * 4, r->stream));
                }
                if (hi > lo && !image_batch_encode_one(r, n_threads, imgs->data + lo, hi - lo, nullptr, r->d_emb[0], normalize, false)) return false;
                const size_t l0 = std::min(n_labels, (size_t)ri * per), l1 = std::min(n_labels, l0 + per);
                if (l1 > l0 && !text_batch_encode_one(r, labels + l0, l1 - l0, nullptr, r->d_emb[1] + (size_t)ri * per * d, normalize, false)) return false;
                return true;
            })) return false;
        // phase 2: ONE all-gather of the label embeddings, in place on every GPU (group call: one thread drives all communicators)
        std::string e;
        if (!dist_group_start(e)) { set_err(e); return false; }
        for (int ri = 0; ri < w; ri++) {
            clip_ctx* r = c->replicas[ri];
            cudaSetDevice(r->device);
            if (!dist_all_gather(r->dist, r->d_emb[1] + (size_t)ri * per * d, r->d_emb[1], per * d * 4, r->stream, e)) { dist_group_end(e); set_err(e); return false; }
        }
        if (!dist_group_end(e)) { set_err(e); return false; }
        cudaSetDevice(c->device);
        // phase 3: every GPU ranks the labels for its images and copies its slice of the result to the caller
        return for_each_replica(c, n, [&](clip_ctx* r, int, size_t lo, size_t hi) {
            if (hi == lo) return true;
            std::lock_guard<std::mutex> lk(r->mu);
            CK(cudaSetDevice(r->device));
            return score_topk_locked(r, r->d_emb[0], hi - lo, r->d_emb[1], n_labels, d, top_k, true, scores + lo * top_k, indices + lo * top_k);
        });
    }
    const int w = ranks_mode(c) ? c->dist.world : 1, rank = ranks_mode(c) ? c->dist.rank : 0;
    const size_t n_all = (size_t)w * n_labels;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CK(cudaSetDevice(c->device));
        if (!ensure_buf(c, &c->d_emb[0], &c->d_emb_cap[0], n * d) || !ensure_buf(c, &c->d_emb[1], &c->d_emb_cap[1], n_all * d)) return false;
    }
    if (!image_batch_encode_one(c, n_threads, imgs->data, n, nullptr, c->d_emb[0], normalize, false)) return false;
    if (!text_batch_encode_one(c, labels, n_labels, nullptr, c->d_emb[1] + (size_t)rank * n_labels * d, normalize, false)) return false;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!gather_in_place(c, c->d_emb[1], n_labels, d)) return false;
    return score_topk_locked(c, c->d_emb[0], n, c->d_emb[1], n_all, d, top_k, true, scores, indices);
}

bool clip_model_quantize(const char* fname_inp, const char* fname_out, const int itype) {
    try {
        std::string err;
        if (!quantize_file(fname_inp, fname_out, itype, err)) { set_err("clip_model_quantize: " + err); return false; }
        return true;
    } catch (const std::exception& e) { set_err(std::string("clip_model_quantize: ") + e.what()); return false; }
}

// ---- helpers ---------------------------------------------------------------------------------------------------
void* clip_b200_device_malloc(const struct clip_ctx* c, size_t bytes) {
    if (c) cudaSetDevice(c->device);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { set_err("cudaMalloc failed"); return nullptr; }
    return p;
}
void clip_b200_device_free(const struct clip_ctx* c, void* p) { if (c) cudaSetDevice(c->device); cudaFree(p); }
void* clip_b200_host_malloc(size_t bytes) { void* p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr; return p; }
void clip_b200_host_free(void* p) { cudaFreeHost(p); }
bool clip_b200_memcpy_h2d(const struct clip_ctx* c, void* d, const void* h, size_t bytes) {
    if (c) cudaSetDevice(c->device);
    CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    return true;
}
bool clip_b200_memcpy_d2h(const struct clip_ctx* c, void* h, const void* d, size_t bytes) {
    if (c) cudaSetDevice(c->device);
    CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
    return true;
}
bool clip_b200_synchronize(const struct clip_ctx* c) { if (c) cudaSetDevice(c->device); CK(cudaDeviceSynchronize()); return true; }
void* clip_b200_get_stream(const struct clip_ctx* c) { return c ? (void*)c->stream : nullptr; }
void clip_b200_set_micro_batch(const struct clip_ctx* cc, int images, int sequences) {
    clip_ctx* c = const_cast<clip_ctx*>(cc);
    if (!c) return;
    if (images > 0 && c->has_vision && c->vis.ws.cap_items == 0) c->vis.micro_batch = images;
    if (sequences > 0 && c->has_text && c->txt.ws.cap_items == 0) c->txt.micro_batch = sequences;
}
// CUDA-event stopwatch on the launch stream (slots 0..3) so that callers can time a region on the device; the events belong to
// the context (= to its device)
bool clip_b200_mark(const struct clip_ctx* cc, int slot) {
    clip_ctx* c = const_cast<clip_ctx*>(cc);
    if (!c || slot < 0 || slot > 3) return false;
    CK(cudaSetDevice(c->device));
    if (!c->marks[slot]) CK(cudaEventCreate(&c->marks[slot]));
    CK(cudaEventRecord(c->marks[slot], c->stream));
    return true;
}
float clip_b200_mark_elapsed_ms(const struct clip_ctx* c, int a, int b) {
    if (!c || a < 0 || a > 3 || b < 0 || b > 3 || !c->marks[a] || !c->marks[b]) return -1.f;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(c->marks[b]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    cudaEventElapsedTime(&ms, c->marks[a], c->marks[b]);
    return ms;
}
uint64_t clip_b200_kernel_launches(const struct clip_ctx* c) { return c ? c->launches : 0; }
float clip_b200_last_device_ms(const struct clip_ctx* c) { return c ? c->last_ms : 0.f; }
float clip_b200_kernel_ms(const struct clip_ctx* cc, int kind, uint64_t* count) {
    clip_ctx* c = const_cast<clip_ctx*>(cc);
    if (!c) return 0.f;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    float total = 0.f;
    uint64_t n = 0;
    std::vector<ProfEvent> keep;
    for (auto& p : c->prof) {
        if (p.kind == kind) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.e0, p.e1);
            total += ms; n++;
            c->ev_pool.push_back(p.e0); c->ev_pool.push_back(p.e1);
        } else keep.push_back(p);
    }
    c->prof.swap(keep);
    if (count) *count = n;
    return total;
}

int clip_b200_debug_gemm(int qtype, int operand_bf16, int M, int N, int K, int epi, int use_naive, const float* x, const void* w_rows,
                         const float* bias, const float* resid_in, float* y_out, float* ms) {
    g_err.clear();
    if (gemm_init() != cudaSuccess) { set_err("gemm_init failed"); return 1; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const bool out32 = (epi == EPI_STORE32 || epi == EPI_REDADD32);
    int eff_q = qtype;
    std::vector<uint16_t> x16((size_t)M * K);
    for (size_t i = 0; i < x16.size(); i++) x16[i] = operand_bf16 ? f32_to_bf16(x[i]) : f32_to_f16(x[i]);
    std::vector<uint8_t> wdev;
    std::vector<uint16_t> wf16;
    size_t raw_bytes;
    if (qtype == 0 || qtype == 1) {
        eff_q = QT_F16;
        wf16.resize((size_t)N * K);
        if (qtype == 1) memcpy(wf16.data(), w_rows, wf16.size() * 2);
        else for (size_t i = 0; i < wf16.size(); i++) wf16[i] = f32_to_f16(((const float*)w_rows)[i]);
        raw_bytes = wf16.size() * 2;
    } else {
        raw_bytes = (size_t)N * (K / 32) * wpack_ggml_block_bytes(qtype);
        wdev.resize(wpack_total_bytes(qtype, N, K));
        if (!wpack_repack(qtype, (const uint8_t*)w_rows, N, K, wdev.data())) { set_err("repack failed"); return 2; }
    }
    void *d_x = nullptr, *d_w = nullptr, *d_raw = nullptr, *d_out = nullptr;
    float* d_bias = nullptr;
    const size_t out_bytes = (size_t)M * N * (out32 ? 4 : 2);
    int rc = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t e0, e1;
    cudaStreamCreate(&st); cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto fail = [&](const char* m, cudaError_t e) { set_err(std::string(m) + ": " + cudaGetErrorString(e)); rc = 3; };
    cudaError_t e;
    do {
        if ((e = cudaMalloc(&d_x, x16.size() * 2)) != cudaSuccess) { fail("malloc x", e); break; }
        if ((e = cudaMalloc(&d_out, out_bytes)) != cudaSuccess) { fail("malloc out", e); break; }
        if ((e = cudaMalloc(&d_raw, raw_bytes)) != cudaSuccess) { fail("malloc raw", e); break; }
        cudaMemcpy(d_x, x16.data(), x16.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(d_raw, eff_q == QT_F16 ? (const void*)wf16.data() : w_rows, raw_bytes, cudaMemcpyHostToDevice);
        if (eff_q == QT_F16) d_w = d_raw;
        else { if ((e = cudaMalloc(&d_w, wdev.size())) != cudaSuccess) { fail("malloc w", e); break; } cudaMemcpy(d_w, wdev.data(), wdev.size(), cudaMemcpyHostToDevice); }
        if (bias) { cudaMalloc(&d_bias, (size_t)N * 4); cudaMemcpy(d_bias, bias, (size_t)N * 4, cudaMemcpyHostToDevice); }
        if (epi == EPI_REDADD32 && resid_in) cudaMemcpy(d_out, resid_in, out_bytes, cudaMemcpyHostToDevice);
        else cudaMemset(d_out, 0, out_bytes);
        TmaMap xm, xh, wm, om;
        memset(&wm, 0, sizeof(wm));
        const bool use_out_map = epi != EPI_STORE32 && !(getenv("CLIP_B200_DEBUG_DIRECT_STORE") && atoi(getenv("CLIP_B200_DEBUG_DIRECT_STORE")) != 0);
        if (use_out_map && !(out32 ? make_tma_2d_f32_plain(&om, d_out, M, N, N, GEMM_OUT_BOX, GEMM_OUT_BOX)
                                   : make_tma_2d_16bit_plain(&om, d_out, M, N, N, GEMM_OUT_BOX, GEMM_OUT_BOX))) { set_err("tensor map OUT failed"); rc = 4; break; }
        if (!make_tma_2d_16bit(&xm, d_x, M, K, K, GEMM_BN) || !make_tma_2d_16bit(&xh, d_x, M, K, K, GEMM_BN / 2)) { set_err("tensor map X failed"); rc = 4; break; }
        if (eff_q == QT_F16 && !make_tma_2d_16bit(&wm, d_w, N, K, K, GEMM_BM)) { set_err("tensor map W failed"); rc = 4; break; }
        cudaEventRecord(e0, st);
        if (use_naive) {
            launch_naive_gemm(d_x, operand_bf16, d_raw, eff_q, d_bias, d_out, M, N, K, N, epi, operand_bf16, N / 2, 0.125f, st);
            e = cudaGetLastError();
        } else {
            GemmArgs a;
            a.x_map = &xm; a.x_half_map = &xh; a.w_map = &wm; a.w_packed = (const uint8_t*)d_w; a.qtype = eff_q; a.operand_bf16 = operand_bf16 != 0;
            a.bias = d_bias; a.out = d_out; a.M = M; a.N = N; a.K = K; a.ldo = N; a.epi = epi; a.out_bf16 = operand_bf16;
            a.scale_cols = N / 2; a.scale = 0.125f; a.out_map = use_out_map ? &om : nullptr;
            e = gemm_launch(a, st, prop.multiProcessorCount, nullptr);
        }
        cudaEventRecord(e1, st);
        if (e != cudaSuccess) { fail("launch", e); break; }
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { fail("kernel", e); break; }
        if (ms) cudaEventElapsedTime(ms, e0, e1);
        if (out32) cudaMemcpy(y_out, d_out, out_bytes, cudaMemcpyDeviceToHost);
        else {
            std::vector<uint16_t> o16((size_t)M * N);
            cudaMemcpy(o16.data(), d_out, out_bytes, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < o16.size(); i++) {
                if (operand_bf16) { uint32_t b = (uint32_t)o16[i] << 16; memcpy(&y_out[i], &b, 4); }
                else y_out[i] = f16_to_f32(o16[i]);
            }
        }
    } while (0);
    cudaFree(d_x); cudaFree(d_out); cudaFree(d_raw); if (d_w && d_w != d_raw) cudaFree(d_w); cudaFree(d_bias);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
    return rc;
}


// TEST HOOK: one attention launch on device 0.  qkv: fp32 host [nseq*T, 3*H*64] (columns Q | K | V, the Q columns already scaled),
// rounded to the operand type on the way in; out: fp32 host [nseq*T, H*64].  use_legacy=1 runs the mma.sync flash kernel.
int clip_b200_debug_attention(int operand_bf16, int nseq, int T, int H, int causal, int use_legacy, const float* qkv, float* out, float* ms) {
    g_err.clear();
    if (attention_tc_init() != cudaSuccess) { set_err("attention_tc_init failed"); return 1; }
    if (!use_legacy && !attention_tc_supported(T) && !attention_tc_long_supported(T, causal)) { set_err("T not supported by the tcgen05 attention kernels"); return 2; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const size_t rows = (size_t)nseq * T, hid = (size_t)H * 64;
    std::vector<uint16_t> h16(rows * 3 * hid);
    for (size_t i = 0; i < h16.size(); i++) h16[i] = operand_bf16 ? f32_to_bf16(qkv[i]) : f32_to_f16(qkv[i]);
    uint16_t *d_qkv = nullptr, *d_out = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    do {
        cudaError_t e;
        // pad by one 272-row box: the K/V boxes of the last sequence read (and discard) rows past the end
        if ((e = cudaMalloc(&d_qkv, (rows + 272) * 3 * hid * 2)) != cudaSuccess || (e = cudaMalloc(&d_out, rows * hid * 2)) != cudaSuccess) { set_err(std::string("cudaMalloc: ") + cudaGetErrorString(e)); rc = 3; break; }
        // the padding is POISONED (0xFFFF = NaN in fp16 and bf16): rows past the last sequence are stale memory in the model path
        // (ensure_ws never initialises the workspace tail), and no output may depend on them
        cudaMemset(d_qkv, 0xff, (rows + 272) * 3 * hid * 2);
        cudaMemset(d_out, 0xff, rows * hid * 2);
        cudaMemcpy(d_qkv, h16.data(), h16.size() * 2, cudaMemcpyHostToDevice);
        cudaStreamCreate(&st); cudaEventCreate(&e0); cudaEventCreate(&e1);
        TmaMap mq, m16;
        if (!make_tma_2d_16bit(&mq, d_qkv, rows, 3 * hid, 3 * hid, 128) ||
            !make_tma_2d_16bit(&m16, d_qkv, rows, 3 * hid, 3 * hid, 16)) { set_err("tensor map failed"); rc = 4; break; }
        for (int rep = 0; rep < (ms ? 3 : 1); rep++) {
            cudaEventRecord(e0, st);
            if (use_legacy) { launch_attention(d_qkv, d_out, nseq, T, H, causal, operand_bf16, 0, st); e = cudaGetLastError(); }
            else if (attention_tc_supported(T)) e = launch_attention_tc(&mq, d_qkv, &m16, d_out, nseq, T, H, causal, operand_bf16, prop.multiProcessorCount, st);
            else e = launch_attention_tc_long(&mq, d_qkv, d_out, nseq, T, H, operand_bf16, prop.multiProcessorCount, st);
            cudaEventRecord(e1, st);
            if (e != cudaSuccess) break;
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        }
        if (e != cudaSuccess) { set_err(std::string("attention: ") + cudaGetErrorString(e)); rc = 5; break; }
        if (ms) cudaEventElapsedTime(ms, e0, e1);
        std::vector<uint16_t> o16(rows * hid);
        cudaMemcpy(o16.data(), d_out, o16.size() * 2, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < o16.size(); i++) {
            if (operand_bf16) { uint32_t b = (uint32_t)o16[i] << 16; memcpy(&out[i], &b, 4); }
            else out[i] = f16_to_f32(o16[i]);
        }
    } while (0);
    cudaFree(d_qkv); cudaFree(d_out);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return rc;
}


// ---- CPU-only test hooks (no context / GPU needed) ----------------------------------------------------------
// ranks-mode rendezvous of the NCCL unique id (dist.cpp) without a GPU: rank 0 publishes, the others read the same 128 bytes
int clip_b200_debug_rendezvous(int rank, int world, const char* path, void* out128) {
    std::string e;
    if (!dist_rendezvous_id(rank, world, path, out128, e)) { set_err(e); return 1; }
    return 0;
}
void clip_b200_debug_shard_bounds(size_t n, int r, int w, size_t* lo, size_t* hi) { shard_bounds(n, r, w, *lo, *hi); }
int clip_b200_debug_repack_roundtrip(int qtype, const void* rows, int N, int K) {
    const size_t raw = (size_t)N * (K / 32) * wpack_ggml_block_bytes(qtype);
    std::vector<uint8_t> packed(wpack_total_bytes(qtype, N, K)), back(raw);
    if (!wpack_repack(qtype, (const uint8_t*)rows, N, K, packed.data())) return 1;
    if (!wpack_unpack(qtype, packed.data(), N, K, back.data())) return 2;
    return memcmp(back.data(), rows, raw) == 0 ? 0 : 3;
}
int clip_b200_debug_tokenize(const char* gguf_path, const char* text, int32_t* out, int cap) {
    GgufFile g;
    std::string err;
    if (!g.parse(gguf_path, err)) { set_err(err); return -1; }
    const GgufKV* tk = g.find("tokenizer.ggml.tokens");
    if (!tk) { set_err("no vocabulary"); return -1; }
    Vocab v;
    for (size_t i = 0; i < tk->strs.size(); i++) v.token_to_id[tk->strs[i]] = (int32_t)i;
    std::vector<int32_t> ids = tokenize(v, text);
    for (size_t i = 0; i < ids.size() && (int)i < cap; i++) out[i] = ids[i];
    return (int)ids.size();
}
int clip_b200_debug_preprocess(const uint8_t* rgb, int nx, int ny, int out_size, const float* mean, const float* stdv, float* out) {
    return preprocess_image(rgb, nx, ny, out_size, mean, stdv, out) ? 0 : 1;
}

// ---- ggml/ggml.h shim symbols (include/ggml/ggml.h) ---------------------------------------------------------
static std::chrono::steady_clock::time_point g_t0 = std::chrono::steady_clock::now();
void ggml_time_init(void) { g_t0 = std::chrono::steady_clock::now(); }
int64_t ggml_time_us(void) { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - g_t0).count(); }
int64_t ggml_time_ms(void) { return ggml_time_us() / 1000; }

}  // extern "C"
