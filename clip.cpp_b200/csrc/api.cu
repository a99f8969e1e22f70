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
        if (c->profile) { cudaEvent_t e1 = get_event(c); cudaEventRecord(e1, c->stream); c->prof.push_back({kind, e0, e1}); }
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
    if (!dev_alloc(c, &w.x, rows * h) || !dev_alloc(c, &w.a, rows * h) || !dev_alloc(c, &w.d, rows * h) || !dev_alloc(c, &w.qkv, rows * 3 * h) ||
        !dev_alloc(c, &w.g, rows * tw.ff) || !dev_alloc(c, &w.sel16, (size_t)items * h) || !dev_alloc(c, &w.sel32, (size_t)items * h) ||
        !dev_alloc(c, &w.proj32, (size_t)items * tw.proj)) return false;
    bool ok = make_tma_2d_16bit(&w.map_a, w.a, rows, h, h, GEMM_BN) && make_tma_2d_16bit(&w.map_g, w.g, rows, tw.ff, tw.ff, GEMM_BN) &&
              make_tma_2d_16bit(&w.map_sel, w.sel16, items, h, h, GEMM_BN) &&
              make_tma_2d_16bit(&w.map_a_half, w.a, rows, h, h, GEMM_BN / 2) && make_tma_2d_16bit(&w.map_g_half, w.g, rows, tw.ff, tw.ff, GEMM_BN / 2) &&
              make_tma_2d_16bit(&w.map_q128, w.qkv, rows, 3 * h, 3 * h, 128) && make_tma_2d_16bit(&w.map_kv256, w.qkv, rows, 3 * h, 3 * h, 256) &&
              make_tma_2d_16bit(&w.map_kv16, w.qkv, rows, 3 * h, 3 * h, 16);
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
                int out_bf16, int scale_cols = 0, float scale = 1.f, const TmaMap* xmap_half = nullptr) {
    Scope s(c, K_GEMM);
    if (c->debug_naive) {
        launch_naive_gemm(xptr, x_bf16, L.d_raw, L.qtype, L.d_bias, out, M, L.N, L.K, ldo, epi, out_bf16, scale_cols, scale, c->stream);
        CK(cudaGetLastError());
        return true;
    }
    GemmArgs a;
    a.x_map = xmap; a.x_half_map = xmap_half; a.w_map = &L.w_map; a.w_packed = L.d_w; a.qtype = L.qtype; a.operand_bf16 = x_bf16;
    a.bias = L.d_bias; a.out = out; a.M = M; a.N = L.N; a.K = L.K; a.ldo = ldo; a.epi = epi; a.out_bf16 = out_bf16;
    a.scale_cols = scale_cols; a.scale = scale;
    CK(gemm_launch(a, c->stream, c->num_sms, nullptr));
    return true;
}

// The per-layer schedule shared by both towers (clip.cpp:1064-1143 text, 1342-1423 vision).
bool run_blocks(clip_ctx* c, Tower& tw, int nseq, int T, bool causal) {
    // Residual adds are deferred: each branch GEMM (out-proj, FC2) stores its output 16-bit into w.d and the NEXT LayerNorm
    // applies x += d while it reads x anyway.  On return one delta (the last FC2) is still pending in w.d.
    Workspace& w = tw.ws;
    const int M = nseq * T, h = tw.hidden;
    const int bf = c->operand_bf16 ? 1 : 0;
    const float qscale = 1.0f / sqrtf(64.0f);
    bool pending = false;
    for (auto& l : tw.L) {
        { Scope s(c, K_LN); launch_layernorm(w.x, h, M, h, tw.eps, l.ln1_g, l.ln1_b, pending ? w.d : nullptr, w.a, bf, c->stream); }
        if (!run_linear(c, l.qkv, &w.map_a, w.a, bf, M, w.qkv, 3 * h, EPI_STORE16, bf, h, qscale, &w.map_a_half)) return false;
        if (c->attn_tc && attention_tc_supported(T)) {
            const int q_done = attention_tc_tiles(T) * 128;
            { Scope s(c, K_ATTN); CK(launch_attention_tc(&w.map_q128, &w.map_kv256, &w.map_kv16, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, c->num_sms, c->stream)); }
            if (q_done < T) { Scope s(c, K_ATTN); launch_attention(w.qkv, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, q_done, c->stream); }
        } else {
            Scope s(c, K_ATTN); launch_attention(w.qkv, w.a, nseq, T, tw.heads, causal ? 1 : 0, bf, 0, c->stream);
        }
        if (!run_linear(c, l.out, &w.map_a, w.a, bf, M, w.d, h, EPI_STORE16, bf, 0, 1.f, &w.map_a_half)) return false;
        { Scope s(c, K_LN); launch_layernorm(w.x, h, M, h, tw.eps, l.ln2_g, l.ln2_b, w.d, w.a, bf, c->stream); }
        if (!run_linear(c, l.fc1, &w.map_a, w.a, bf, M, w.g, tw.ff, c->use_gelu ? EPI_GELU16 : EPI_QGELU16, bf, 0, 1.f, &w.map_a_half)) return false;
        if (!run_linear(c, l.fc2, &w.map_g, w.g, bf, M, w.d, h, EPI_STORE16, bf, 0, 1.f, &w.map_g_half)) return false;
        pending = true;
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
    { Scope s(c, K_LN); launch_layernorm(w.x, (size_t)tw.T * h, nb, h, tw.eps, tw.post_g, tw.post_b, tw.layers ? w.d : nullptr, w.sel16, bf, c->stream); }
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
    { Scope s(c, K_OTHER); launch_gather_rows(w.x, w.sel32, nb, h, T, w.last, tw.layers ? w.d : nullptr, bf, c->stream); }
    { Scope s(c, K_LN); launch_layernorm(w.sel32, h, nb, h, tw.eps, tw.post_g, tw.post_b, nullptr, w.sel16, bf, c->stream); }
    if (!run_linear(c, tw.proj_w, &w.map_sel, w.sel16, bf, nb, w.proj32, tw.proj, EPI_STORE32, 0)) return false;
    { Scope s(c, K_OTHER); launch_l2norm(w.proj32, d_out, nb, tw.proj, normalize ? 1 : 0, c->stream); }
    CK(cudaGetLastError());
    return true;
}

bool ensure_out(clip_ctx* c, size_t floats) {
    if (c->d_out_cap >= floats) return true;
    float* p = nullptr;
    if (!dev_alloc(c, &p, floats)) return false;   // old buffer stays in allocs until clip_free (grow-only, rare)
    c->d_out = p; c->d_out_cap = floats;
    return true;
}

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

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char* clip_b200_last_error(void) { return g_err.c_str(); }
const char* clip_b200_version(void) { return "clip_b200 0.1 (sm_100a; tcgen05 fused-dequant GEMM)"; }

struct clip_ctx* clip_model_load(const char* fname, const int verbosity) {
    try {
        g_err.clear();
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err("no CUDA device: libclip_b200 has no CPU path"); return nullptr; }
        GgufFile g;
        std::string err;
        if (!g.parse(fname, err)) { set_err(err); return nullptr; }
        std::unique_ptr<clip_ctx, void (*)(clip_ctx*)> guard(new clip_ctx, [](clip_ctx* p) { clip_free(p); });
        clip_ctx* c = guard.get();
        const char* dv = getenv("CLIP_B200_DEVICE");
        c->device = dv ? atoi(dv) : 0;
        if (c->device < 0 || c->device >= ndev) { set_err("CLIP_B200_DEVICE out of range"); return nullptr; }
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
            if (nm) printf("%s: model name:   %s\n", __func__, nm->s.c_str());
            if (ds) printf("%s: description:  %s\n", __func__, ds->s.c_str());
            printf("%s: GGUF version: %u\n", __func__, g.version);
            printf("%s: alignment:    %zu\n", __func__, g.alignment);
            printf("%s: n_tensors:    %zu\n", __func__, g.tensors.size());
            printf("%s: n_kv:         %zu\n", __func__, g.kvs.size());
            if (ft) printf("%s: ftype:        %d\n", __func__, (int)ft->u);
            printf("%s: text_encoder:   %d\n%s: vision_encoder: %d\n", __func__, c->has_text, __func__, c->has_vision);
            printf("%s: device:       %d (%s, %d SMs)\n", __func__, c->device, prop.name, c->num_sms);
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
            if (!check_geometry("text tower", t)) return nullptr;
            if (!upload_f32(c, g, "t.token_embd.weight", &t.tok, (int64_t)t.n_vocab * t.hidden) ||
                !upload_f32(c, g, "t.position_embd.weight", &t.pos, (int64_t)t.n_ctx * t.hidden) ||
                !upload_f32(c, g, "t.post_ln.weight", &t.post_g, t.hidden) || !upload_f32(c, g, "t.post_ln.bias", &t.post_b, t.hidden)) return nullptr;
            if (!make_linear(c, g, {"text_projection.weight"}, {}, t.proj_w)) return nullptr;
            if (!load_blocks(c, g, "t", t)) return nullptr;
            t.micro_batch = default_micro_batch(t.n_ctx);
            if (verbosity >= 2)
                printf("\n%s: text model hparams\nn_vocab            %d\nnum_positions      %d\nt_hidden_size      %d\nt_n_intermediate   %d\nt_projection_dim   %d\nt_n_head           %d\nt_n_layer          %d\n",
                       __func__, hp.n_vocab, hp.num_positions, hp.hidden_size, hp.n_intermediate, hp.projection_dim, hp.n_head, hp.n_layer);
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
            if (t.patch <= 0 || t.image_size % t.patch) { set_err("image_size must be a multiple of patch_size"); return nullptr; }
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
                       __func__, hp.image_size, hp.patch_size, hp.hidden_size, hp.n_intermediate, hp.projection_dim, hp.n_head, hp.n_layer);
        }
        // 16-bit operand type of the towers: fp16 when any tower GEMM weight is stored unquantized (the reference rounds
        // those activations to fp16 too, ggml.c:11333-11349); bf16 for fully quantized towers unless overridden.
        const bool f16w = any_f16_gemm(c->vis) || any_f16_gemm(c->txt);
        c->operand_bf16 = !f16w;
        if (const char* op = getenv("CLIP_B200_OPERAND")) {
            if (!strcmp(op, "f16")) c->operand_bf16 = false;
            else if (!strcmp(op, "bf16") && !f16w) c->operand_bf16 = true;
        }
        if (verbosity >= 1) printf("%s: operand type: %s%s\n", __func__, c->operand_bf16 ? "bf16" : "fp16", c->debug_naive ? "  [DEBUG naive GEMM]" : "");
        if (cudaDeviceSynchronize() != cudaSuccess) { set_err("upload failed"); return nullptr; }
        return guard.release();
    } catch (const std::exception& e) {
        set_err(std::string("clip_model_load: ") + e.what());
        return nullptr;
    }
}

void clip_free(struct clip_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    preprocess_release(c);
    for (void* p : c->allocs) cudaFree(p);
    for (auto& p : c->prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) { if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]); }
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
    if (!load_image_file(fname, rgb, nx, ny)) { fprintf(stderr, "%s: failed to load '%s' (PPM P6 / 24-bit BMP only)\n", __func__, fname); return false; }
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

bool clip_image_batch_encode(const struct clip_ctx* cctx, const int n_threads, const struct clip_image_f32_batch* imgs, float* vec,
                             const bool normalize) {
    (void)n_threads;
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_vision) { printf("This gguf file seems to have no vision encoder\n"); return false; }
    const size_t n = imgs->size;
    if (n == 0) return true;
    Tower& tw = c->vis;
    for (size_t i = 0; i < n; i++)
        if (imgs->data[i].nx != tw.image_size || imgs->data[i].ny != tw.image_size || !imgs->data[i].data) { set_err("clip_image_batch_encode: every image must be image_size x image_size"); return false; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (!ensure_ws(c, tw, tw.micro_batch, tw.T, true) || !ensure_out(c, n * tw.proj)) return false;
    size_t n_chunks, chunk;
    chunking(n, tw.micro_batch, n_chunks, chunk);
    const size_t per = (size_t)tw.image_size * tw.image_size * 3;
    CK(cudaEventRecord(c->ev_t0, c->stream));
    size_t ci = 0;
    for (size_t i0 = 0; i0 < n; i0 += chunk, ci++) {
        const int nb = (int)std::min(chunk, n - i0), b = (int)(ci & 1);
        // H2D of chunk ci on the copy stream overlaps the kernels of chunk ci-1 (double-buffered staging)
        if (ci >= 2) CK(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
        for (int j = 0; j < nb; j++)
            CK(cudaMemcpyAsync(tw.ws.pixels[b] + (size_t)j * per, imgs->data[i0 + j].data, per * 4, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        if (!vision_forward(c, tw.ws.pixels[b], nb, c->d_out + i0 * tw.proj, normalize)) return false;
        CK(cudaEventRecord(c->ev_consumed[b], c->stream));
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    CK(cudaMemcpyAsync(vec, c->d_out, n * tw.proj * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_time(c);
}

bool clip_image_encode(const struct clip_ctx* ctx, const int n_threads, struct clip_image_f32* img, float* vec, const bool normalize) {
    if (!ctx || !ctx->has_vision) { printf("This gguf file seems to have no vision encoder\n"); return false; }
    clip_image_f32_batch b;
    b.data = img; b.size = 1;
    return clip_image_batch_encode(ctx, n_threads, &b, vec, normalize);
}

static bool text_encode_impl(clip_ctx* c, const int32_t* h_ids, const int32_t* h_last, const int32_t* d_ids, const int32_t* d_lens,
                             size_t n, int T, float* vec_host, float* vec_dev, bool normalize) {
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
            CK(cudaMemcpyAsync(tw.ws.last, d_lens + i0, (size_t)nb * 4, cudaMemcpyDeviceToDevice, c->stream));   // already len-1
        }
        if (!text_forward(c, nb, T, d_out + i0 * tw.proj, normalize)) return false;
    }
    CK(cudaEventRecord(c->ev_t1, c->stream));
    if (vec_host) CK(cudaMemcpyAsync(vec_host, d_out, n * tw.proj * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_time(c);
}

bool clip_text_batch_encode(const struct clip_ctx* cctx, const int n_threads, const struct clip_tokens* seqs, const size_t n, float* vec,
                            const bool normalize) {
    (void)n_threads;
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c || !c->has_text) { printf("This GGUF file seems to have no text encoder\n"); return false; }
    if (n == 0) return true;
    size_t T = 0;
    for (size_t i = 0; i < n; i++) {
        if (seqs[i].size == 0 || !seqs[i].data) { set_err("empty token sequence"); return false; }
        T = std::max(T, seqs[i].size);
    }
    if ((int)T > c->txt.n_ctx) { set_err("token sequence longer than context_length (" + std::to_string(c->txt.n_ctx) + ")"); return false; }
    // pad to a multiple of 8 tokens (cheap) -- padded positions are causally invisible to the EOT row
    T = std::min<size_t>((T + 7) / 8 * 8, (size_t)c->txt.n_ctx);
    std::vector<int32_t> ids(n * T, 0), last(n);
    for (size_t i = 0; i < n; i++) {
        memcpy(ids.data() + i * T, seqs[i].data, seqs[i].size * 4);
        last[i] = (int32_t)seqs[i].size - 1;     // the reference selects row N-1 (clip.cpp:1154-1155)
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return text_encode_impl(c, ids.data(), last.data(), nullptr, nullptr, n, (int)T, vec, nullptr, normalize);
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
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // device lengths are converted to last-row indices on the host side of this call (tiny D2H)
    std::vector<int32_t> last(n, seq_len - 1);
    if (d_lens) {
        CK(cudaMemcpy(last.data(), d_lens, n * 4, cudaMemcpyDeviceToHost));
        for (auto& v : last) { if (v < 1 || v > seq_len) { set_err("bad sequence length"); return false; } v -= 1; }
    }
    int32_t* d_last = nullptr;
    CK(cudaMalloc(&d_last, n * 4));
    cudaError_t e = cudaMemcpy(d_last, last.data(), n * 4, cudaMemcpyHostToDevice);
    bool ok = e == cudaSuccess && text_encode_impl(c, nullptr, nullptr, (const int32_t*)d_ids, d_last, n, seq_len, nullptr, (float*)d_vec, normalize);
    cudaFree(d_last);
    return ok;
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

bool clip_b200_zero_shot_batch(const struct clip_ctx* cctx, const void* d_img, size_t n_img, const void* d_txt, size_t n_txt, float* scores,
                               int* indices, int top_k) {
    clip_ctx* c = const_cast<clip_ctx*>(cctx);
    if (!c) return false;
    if (top_k <= 0 || (size_t)top_k > n_txt) top_k = (int)n_txt;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int d = c->has_vision ? c->vhp.projection_dim : c->thp.projection_dim;
    float* d_logits = nullptr;
    CK(cudaMalloc(&d_logits, n_img * n_txt * 4));
    { Scope s(c, K_OTHER); launch_logits((const float*)d_img, (const float*)d_txt, d_logits, (int)n_img, (int)n_txt, d, c->stream); }
    { Scope s(c, K_OTHER); launch_softmax_plain(d_logits, (int)n_img, (int)n_txt, c->stream); }
    std::vector<float> p(n_img * n_txt);
    cudaError_t e = cudaMemcpyAsync(p.data(), d_logits, p.size() * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_logits);
    if (e != cudaSuccess) { set_err(std::string("zero_shot_batch failed: ") + cudaGetErrorString(e)); return false; }
    std::vector<int> idx(n_txt);
    for (size_t i = 0; i < n_img; i++) {
        const float* r = p.data() + i * n_txt;
        std::iota(idx.begin(), idx.end(), 0);
        std::partial_sort(idx.begin(), idx.begin() + top_k, idx.end(), [&](int a, int b) { return r[a] > r[b] || (r[a] == r[b] && a < b); });
        for (int k = 0; k < top_k; k++) { scores[i * top_k + k] = r[idx[k]]; indices[i * top_k + k] = idx[k]; }
    }
    return true;
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
// CUDA-event stopwatch on the launch stream (slots 0..3) so that callers can time a region on the device
static cudaEvent_t g_marks[4] = {nullptr, nullptr, nullptr, nullptr};
bool clip_b200_mark(const struct clip_ctx* c, int slot) {
    if (!c || slot < 0 || slot > 3) return false;
    cudaSetDevice(c->device);
    if (!g_marks[slot]) CK(cudaEventCreate(&g_marks[slot]));
    CK(cudaEventRecord(g_marks[slot], c->stream));
    return true;
}
float clip_b200_mark_elapsed_ms(const struct clip_ctx* c, int a, int b) {
    if (!c || a < 0 || a > 3 || b < 0 || b > 3 || !g_marks[a] || !g_marks[b]) return -1.f;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(g_marks[b]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    cudaEventElapsedTime(&ms, g_marks[a], g_marks[b]);
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
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const bool out32 = (epi == EPI_RESID32 || epi == EPI_STORE32);
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
        if (epi == EPI_RESID32 && resid_in) cudaMemcpy(d_out, resid_in, out_bytes, cudaMemcpyHostToDevice);
        else cudaMemset(d_out, 0, out_bytes);
        TmaMap xm, xh, wm;
        memset(&wm, 0, sizeof(wm));
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
            a.scale_cols = N / 2; a.scale = 0.125f;
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
    if (!use_legacy && !attention_tc_supported(T)) { set_err("T not supported by the tcgen05 attention kernel"); return 2; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
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
        cudaMemset(d_qkv, 0, (rows + 272) * 3 * hid * 2);
        cudaMemset(d_out, 0xff, rows * hid * 2);
        cudaMemcpy(d_qkv, h16.data(), h16.size() * 2, cudaMemcpyHostToDevice);
        cudaStreamCreate(&st); cudaEventCreate(&e0); cudaEventCreate(&e1);
        TmaMap mq, mkv, m16;
        if (!make_tma_2d_16bit(&mq, d_qkv, rows, 3 * hid, 3 * hid, 128) || !make_tma_2d_16bit(&mkv, d_qkv, rows, 3 * hid, 3 * hid, 256) ||
            !make_tma_2d_16bit(&m16, d_qkv, rows, 3 * hid, 3 * hid, 16)) { set_err("tensor map failed"); rc = 4; break; }
        for (int rep = 0; rep < (ms ? 3 : 1); rep++) {
            cudaEventRecord(e0, st);
            if (use_legacy) { launch_attention(d_qkv, d_out, nseq, T, H, causal, operand_bf16, 0, st); e = cudaGetLastError(); }
            else e = launch_attention_tc(&mq, &mkv, &m16, d_out, nseq, T, H, causal, operand_bf16, prop.multiProcessorCount, st);
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
