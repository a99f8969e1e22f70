// kernels.h -- host launchers of the non-GEMM kernels (kernels.cu, attention.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

// K2 LayerNorm: y16[r, :] = ((x[r*in_stride ...] - mean) * rstd) * gamma + beta   (reference: ggml.c:10796-10845,
// affine clip.cpp:1071-1074).  x fp32, row r starts at x + r*in_stride.  One warp per row.
// If delta16 != NULL the kernel first applies a pending residual branch: x[r,:] += delta16[r,:] (written back in fp32); the model
// schedule passes NULL (round 2: the GEMM epilogues add into x).
void launch_layernorm(float* x, size_t in_stride, int rows, int h, float eps, const float* gamma, const float* beta,
                      const void* delta16, void* y16, int bf16, cudaStream_t st);

// K0a im2col for the stride-P patch conv: pixels NHWC f32 [B,S,S,3] -> fp16 [B*Np, kpad], k = c*P*P + ky*P + kx,
// zero padded to kpad (reference conv_2d im2col, ggml.c:13595-13631; pixel -> fp16 rounding ggml.c:13623).
void launch_im2col(const float* pixels, int B, int S, int P, int kpad, void* patches16, cudaStream_t st);

// K4v assemble + pre-LN: x[b,t,:] = LN( (t==0 ? class_embd : patch[b*Np+t-1,:]) + pos[t,:] ) * g + b  -> fp32
// (clip.cpp:1315-1339)
void launch_assemble_preln(const float* patch, const float* class_embd, const float* pos, int B, int T, int h, float eps,
                           const float* gamma, const float* beta, float* x, cudaStream_t st);

// K4t text embed: x[s*T+t,:] = tok[ids[s*T+t],:] + pos[t,:]   (clip.cpp:1059-1061), tables fp32 (dequantised at load)
void launch_text_embed(const int32_t* ids, const float* tok, const float* pos, int nseq, int T, int h, int n_vocab, float* x,
                       cudaStream_t st);

// K3 attention: softmax(Q K^T [causal]) V per (sequence, head); qkv 16-bit [nseq*T, 3*H*64] (Q pre-scaled),
// out 16-bit [nseq*T, H*64].  head_dim is fixed at 64.  (clip.cpp:1100-1108, 1382-1388)
// q_start (multiple of 64): first query row this launch computes (rows before it come from the tcgen05 kernel below).
void launch_attention(const void* qkv16, void* out16, int nseq, int T, int H, int causal, int bf16, int q_start, cudaStream_t st);

// K3 on tcgen05 (attention_tc.cu): 128-query tiles, S and O accumulators in TMEM.  map_q / map_kv16 are TMA views of the qkv matrix
// [rows, 3*H*64] with box rows 128 / 16; the K / V boxes come from a per-launch 3-D view [sequence][token][column] of qkv16 (zero fill
// past each sequence's T tokens).  Handles query tiles [0, attention_tc_tiles(T)); T must satisfy _supported().
struct TmaMap;
bool attention_tc_supported(int T);
int attention_tc_tiles(int T);
cudaError_t attention_tc_init();
cudaError_t launch_attention_tc(const TmaMap* map_q, const void* qkv16, const TmaMap* map_kv16, void* out16, int nseq, int T,
                                int H, int causal, int bf16, int num_sms, cudaStream_t st);
// long sequences (257 < T <= 640, non-causal; e.g. the 577 tokens of ViT-L/14@336): same data path, online softmax over 192-key blocks
bool attention_tc_long_supported(int T, int causal);
cudaError_t launch_attention_tc_long(const TmaMap* map_q, const void* qkv16, void* out16, int nseq, int T, int H, int bf16, int num_sms,
                                     cudaStream_t st);

// K5 head tail: out[r,:] = normalize ? v / sqrt(sum v^2) : v    (clip.cpp:1163-1166, 1448-1455)
void launch_l2norm(const float* v, float* out, int rows, int d, int normalize, cudaStream_t st);

// gather rows: dst[r,:] = src[idx(r),:] where idx(r) = r*stride_rows + offs[r] (offs may be null -> 0)
void launch_gather_rows(const float* src, float* dst, int rows, int h, int stride_rows, const int32_t* offs, const void* delta16,
                        int bf16, cudaStream_t st);   // dst = src[idx] (+ delta16[idx])

// text: last[i] = clamp(lens ? lens[i] : seq_len, 1, seq_len) - 1  (index of the EOT row, clip.cpp:1154-1155), on the device
void launch_fill_last(const int32_t* lens, int32_t* last, int n, int seq_len, cudaStream_t st);

// ---- scoring (search.cu) ----------------------------------------------------------------------------------------------
// S[i, j] = <a_i, b_j> fp32, one sequential fmaf chain per element (clip_similarity_score, clip.cpp:1525-1532); row stride lds
void launch_similarity(const float* a, const float* b, float* s, int na, int nb, int d, size_t lds, cudaStream_t st);
// in place p = (exp(s) + 1e-9) / sum with softmax_with_sorting's rounding points (clip.cpp:1599-1607)
void launch_softmax_ref(float* logits, int rows, int cols, size_t ld, cudaStream_t st);
// one top-k stage: every slice of topk_slice() candidates of a row is sorted (value desc, index asc) and its best kk are written to
// out[row, out_off + slice * kk ...]; in_idx == nullptr: candidate j has index idx_base + j
int topk_slice();
void launch_topk_stage(const float* vals, const int* in_idx, size_t ld, int rows, int n, int idx_base, int kk, float* out_v, int* out_i,
                       size_t out_ld, int out_off, cudaStream_t st);

}  // namespace cb
struct clip_ctx;
struct clip_image_u8;
#include <string>
namespace cb {
// N1 device-side preprocess (preprocess.cu): raw u8 images -> [nb, S, S, 3] fp32 crops, bit-identical to host_ops.cpp: preprocess_image
bool preprocess_device(clip_ctx* c, const clip_image_u8* imgs, int nb, int buf, float* d_pixels, cudaStream_t copy_st, cudaEvent_t copied,
                       cudaStream_t st, std::string& err);
void preprocess_release(clip_ctx* c);

// DEBUG ONLY (tests / CLIP_B200_DEBUG_NAIVE=1): scalar GEMM straight from ggml-format rows on the device.
void launch_naive_gemm(const void* x16, int x_bf16, const void* w_ggml, int qtype, const float* bias, void* out, int M, int N,
                       int K, int ldo, int epi, int out_bf16, int scale_cols, float scale, cudaStream_t st);

}  // namespace cb
