/*
 * clip_b200.h -- C ABI of libclip_b200.so, the B200-native drop-in for monatis/clip.cpp's encode path.
 *
 * PART 1 re-declares, symbol for symbol and struct for struct, the reference's public interface
 * (reference: /root/reference/clip.h:8-113, 22 extern "C" functions + 8 POD structs) so that the reference's
 * own callers -- examples/main.cpp, zsl.cpp, extract.cpp, simple.c, tests/benchmark.cpp, the ctypes binding in
 * examples/python_bindings/clip_cpp/clip.py -- link against this library unchanged.  Each declaration cites the
 * reference definition it replaces.  Differences in behaviour are limited to error handling: where the reference
 * throws a C++ exception through the C ABI, calls exit(1) or aborts (clip.cpp:85-115, 289-292, 1293), this
 * library returns false / NULL and records a message retrievable with clip_b200_last_error().
 *
 * PART 2 declares the additive entry points the reference lacks (batched text, device-resident buffers,
 * batched zero-shot scoring, measurement hooks).  Nothing in part 2 changes the meaning of part 1.
 *
 * Plain pointers and sizes only; no CUDA, torch or C++ types appear in any signature.
 */
#ifndef CLIP_B200_H
#define CLIP_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* =====================================================================================================
 * PART 1 -- the reference interface (clip.h)
 * ===================================================================================================== */

struct clip_ctx; /* opaque; clip.h:8 (definition clip.cpp:240-253) */

/* clip.h:14-23 -- text tower hyper-parameters, filled from GGUF KV (clip.cpp:464-477) */
struct clip_text_hparams {
    int32_t n_vocab;
    int32_t num_positions;
    int32_t hidden_size;
    int32_t n_intermediate;
    int32_t projection_dim;
    int32_t n_head;
    int32_t n_layer;
    float eps;
};

/* clip.h:25-34 -- vision tower hyper-parameters (clip.cpp:524-535) */
struct clip_vision_hparams {
    int32_t image_size;
    int32_t patch_size;
    int32_t hidden_size;
    int32_t n_intermediate;
    int32_t projection_dim;
    int32_t n_head;
    int32_t n_layer;
    float eps;
};

/* clip.h:36-40 -- token id sequence; `data` is allocated by clip_tokenize with new[] (clip.cpp:675) */
typedef int32_t clip_vocab_id;
struct clip_tokens {
    clip_vocab_id * data;
    size_t size;
};

/* clip.h:50-56 -- RGB uint8 image, interleaved */
struct clip_image_u8 {
    int nx;
    int ny;
    uint8_t * data;
    size_t size;
};

/* clip.h:58-64 -- RGB float32 image, NHWC (RGBRGB...), already mean/std normalised */
struct clip_image_f32 {
    int nx;
    int ny;
    float * data;
    size_t size;
};

/* clip.h:66-74 */
struct clip_image_u8_batch {
    struct clip_image_u8 * data;
    size_t size;
};
struct clip_image_f32_batch {
    struct clip_image_f32 * data;
    size_t size;
};

/* clip.h:42, clip.cpp:334-596.  Parses the GGUF file, uploads the weights to HBM (re-tiled for TMA) and returns a
 * context bound to one GPU (env CLIP_B200_DEVICE, default 0).  NULL on any failure -- missing key/tensor, unsupported
 * geometry, no CUDA device -- with the reason in clip_b200_last_error(); never throws, exits or falls back to a CPU. */
struct clip_ctx * clip_model_load(const char * fname, const int verbosity);

/* clip.h:44, clip.cpp:1010-1014 */
void clip_free(struct clip_ctx * ctx);

/* clip.h:46-47, clip.cpp:1846-1847 -- pointers into the context */
struct clip_text_hparams * clip_get_text_hparams(struct clip_ctx * ctx);
struct clip_vision_hparams * clip_get_vision_hparams(struct clip_ctx * ctx);

/* clip.h:76, clip.cpp:598-679 -- host side; same word split + greedy longest-match as the reference */
bool clip_tokenize(const struct clip_ctx * ctx, const char * text, struct clip_tokens * tokens);

/* clip.h:78-85, clip.cpp:681-707 */
struct clip_image_u8 * clip_image_u8_make();
struct clip_image_f32 * clip_image_f32_make();
void clip_image_u8_clean(struct clip_image_u8 * img);
void clip_image_f32_clean(struct clip_image_f32 * res);
void clip_image_u8_free(struct clip_image_u8 * img);
void clip_image_f32_free(struct clip_image_f32 * res);

/* clip.h:87, clip.cpp:709-726 -- the reference decodes with stb_image; this library has its own decoders for JPEG (Huffman baseline and
 * progressive, 8-bit, grey / YCbCr / RGB / CMYK / YCCK, any integer sampling ratio, restart intervals), PNG (every colour type and bit
 * depth, plain or Adam7-interlaced), BMP (palettes, 16 / 24 / 32 bits, no RLE), GIF (first frame) and binary PGM / PPM.  Each returns the same 3-channel
 * pixels as stb_image, byte for byte (JPEG included: the inverse DCT, chroma up-sampling and colour conversion follow stb_image's
 * integer arithmetic; tests/golden/jpeg).  TGA / PSD / HDR / PIC and arithmetic-coded or 12-bit JPEG return false with an explicit
 * message in clip_b200_last_error. */
bool clip_image_load_from_file(const char * fname, struct clip_image_u8 * img);

/* clip.h:88, clip.cpp:797-927 -- host side PIL-style bicubic resize + centre crop + normalise */
bool clip_image_preprocess(const struct clip_ctx * ctx, const struct clip_image_u8 * img, struct clip_image_f32 * res);

/* clip.h:90-91, clip.cpp:1016-1233 -- one token sequence (<= context_length) -> vec[projection_dim] */
bool clip_text_encode(const struct clip_ctx * ctx, const int n_threads, const struct clip_tokens * tokens, float * vec,
                      const bool normalize);
/* clip.h:92-93, clip.cpp:1235-1245 */
bool clip_image_encode(const struct clip_ctx * ctx, const int n_threads, struct clip_image_f32 * img, float * vec,
                       const bool normalize);

/* clip.h:95-96, clip.cpp:963-1008 -- host threads = min(n_threads, n images) */
void clip_image_batch_preprocess(const struct clip_ctx * ctx, const int n_threads,
                                 const struct clip_image_u8_batch * img_inputs, struct clip_image_f32_batch * imgs_resized);

/* clip.h:97-98, clip.cpp:1247-1523 -- THE HOT PATH.  vec receives imgs->size * projection_dim floats, image-major.
 * Semantics: B independent single-image encodes (the reference's own batched conv path is wrong for ViT-L/14,
 * SURVEY.md section 8c).  n_threads is accepted and ignored by the GPU forward. */
bool clip_image_batch_encode(const struct clip_ctx * ctx, const int n_threads, const struct clip_image_f32_batch * imgs,
                             float * vec, const bool normalize);

/* clip.h:102-103, clip.cpp:1534-1571 */
bool clip_compare_text_and_image(const struct clip_ctx * ctx, const int n_threads, const char * text,
                                 const struct clip_image_u8 * image, float * score);
/* clip.h:104, clip.cpp:1525-1532 */
float clip_similarity_score(const float * vec1, const float * vec2, const int vec_dim);
/* clip.h:105, clip.cpp:1591-1622 -- p_i = (exp(s_i) + 1e-9) / sum, sorted descending */
bool softmax_with_sorting(float * arr, const int length, float * sorted_scores, int * indices);
/* clip.h:106-107, clip.cpp:1624-1659 */
bool clip_zero_shot_label_image(struct clip_ctx * ctx, const int n_threads, const struct clip_image_u8 * input_img,
                                const char ** labels, const size_t n_labels, float * scores, int * indices);

/* clip.h:109, clip.cpp:1661-1844 -- f32/f16 GGUF -> q4_0(2) q4_1(3) q5_0(6) q5_1(7) q8_0(8); output is byte-identical
 * to the reference's (tests/test_host_side.py) */
bool clip_model_quantize(const char * fname_inp, const char * fname_out, const int itype);

/* =====================================================================================================
 * PART 2 -- additive B200 entry points (not in the reference)
 * ===================================================================================================== */

/* n independent clip_text_encode calls in one launch sequence; seqs may be ragged (each <= context_length).
 * vec receives n * projection_dim floats.  (SURVEY.md section 8b "needed extension") */
bool clip_text_batch_encode(const struct clip_ctx * ctx, const int n_threads, const struct clip_tokens * seqs,
                            const size_t n, float * vec, const bool normalize);

/* Device-resident variants: inputs/outputs are DEVICE pointers on the context's GPU (obtained from
 * clip_b200_device_malloc or any CUDA allocator in the same process).  No host<->device copy is made.
 *   d_pixels : n * image_size*image_size*3 floats, NHWC per image;  d_vec : n * projection_dim floats
 *   d_ids    : n * seq_len int32 (padded), d_lens : n int32 true lengths (NULL = all seq_len) */
bool clip_b200_image_encode_device(const struct clip_ctx * ctx, const void * d_pixels, size_t n, void * d_vec,
                                   bool normalize);
bool clip_b200_text_encode_device(const struct clip_ctx * ctx, const void * d_ids, const void * d_lens, size_t n,
                                  int seq_len, void * d_vec, bool normalize);

/* Batched zero-shot scoring on the device (clip.cpp:1624-1659 semantics per image): d_img_vec [n_img, d], d_txt_vec [n_txt, d]
 * -> host scores/indices [n_img, top_k], best first; similarity matrix, softmax and top-k selection all run on the GPU. */
bool clip_b200_zero_shot_batch(const struct clip_ctx * ctx, const void * d_img_vec, size_t n_img, const void * d_txt_vec,
                               size_t n_txt, float * scores, int * indices, int top_k);

/* Nearest-neighbour search on the device (stands in for USearch in examples/image-search/search.cpp:114-158): raw dot products
 * (cosine similarity for normalised embeddings) of n_queries x n_db device-resident vectors, best top_k per query to the host. */
bool clip_b200_topk_search(const struct clip_ctx * ctx, const void * d_queries, size_t n_queries, const void * d_db, size_t n_db,
                           int top_k, float * scores, int * indices);

/* Batched zero-shot labelling, the batch form of clip_zero_shot_label_image (clip.h:106-107, clip.cpp:1624-1659): n images x
 * n_labels token sequences -> scores / indices [n, top_k], best first (p = (exp(s)+1e-9)/sum over ALL labels, clip.cpp:1591-1622).
 * normalize = false reproduces the reference (un-normalised embeddings).  Works in every multi-GPU mode (see below): in devices
 * mode images and labels are sharded over the GPUs and the label embeddings all-gathered; in ranks mode the caller passes this
 * rank's images and this rank's label shard (same count on every rank; global label index = rank * n_labels + j). */
bool clip_b200_zero_shot_images(const struct clip_ctx * ctx, const int n_threads, const struct clip_image_f32_batch * imgs,
                                const struct clip_tokens * labels, const size_t n_labels, const bool normalize, int top_k,
                                float * scores, int * indices);

/* ---- multi-GPU inside the library (csrc/dist.h; the reference's analogue is ggml-cuda.cu:404-407, 5934-5957) -----------------
 * devices mode: set CLIP_B200_DEVICES=0,1,...|all before clip_model_load: ONE context drives a replica per GPU (NCCL communicators
 *   from ncclCommInitAll); clip_image_batch_encode / clip_text_batch_encode / clip_b200_zero_shot_images shard their batch
 *   contiguously and each GPU copies its slice of the result into the caller's buffer.
 * ranks mode:   one process per GPU (torchrun, mpirun, ...): clip_b200_dist_init joins an NCCL communicator (ncclCommInitRank);
 *   rendezvous = NULL takes the 128-byte unique id from a file named after MASTER_PORT and the launcher's pid (single node), or
 *   exchange clip_b200_dist_unique_id's output yourself and call clip_b200_dist_init_with_id.  CLIP_B200_DIST=env makes
 *   clip_model_load do this from RANK / WORLD_SIZE / LOCAL_RANK.  The *_all entry points encode this rank's items and return the
 *   embeddings of ALL ranks, rank-major (one in-place ncclAllGather on the launch stream; every rank must pass the same count).
 * libnccl.so.2 is dlopen'ed on first use; no torch, no MPI. */
bool clip_b200_dist_unique_id(void * out128);
bool clip_b200_dist_init_with_id(struct clip_ctx * ctx, int rank, int world, const void * id128);
bool clip_b200_dist_init(struct clip_ctx * ctx, int rank, int world, const char * rendezvous);
int  clip_b200_dist_rank(const struct clip_ctx * ctx);
int  clip_b200_dist_world(const struct clip_ctx * ctx);
int  clip_b200_device_count(const struct clip_ctx * ctx);           /* GPUs behind this context (devices mode), else 1 */
int  clip_b200_nccl_version(void);                                  /* 0 when libnccl cannot be loaded */
int  clip_b200_cuda_device_count(void);                             /* visible CUDA devices (0 without a driver) */
bool clip_b200_dist_barrier(const struct clip_ctx * ctx);           /* all ranks + device synchronize */
bool clip_b200_dist_max_f64(const struct clip_ctx * ctx, double * vals, int n);   /* element-wise max over ranks, n <= 16 */
bool clip_b200_dist_all_gather(const struct clip_ctx * ctx, const void * d_send, void * d_recv, size_t bytes_per_rank);
bool clip_b200_image_encode_device_all(const struct clip_ctx * ctx, const void * d_pixels, size_t n_local, void * d_vec_all,
                                       bool normalize);
bool clip_b200_text_encode_device_all(const struct clip_ctx * ctx, const void * d_ids, const void * d_lens, size_t n_local,
                                      int seq_len, void * d_vec_all, bool normalize);
bool clip_b200_image_batch_encode_all(const struct clip_ctx * ctx, const int n_threads, const struct clip_image_f32_batch * imgs,
                                      float * vec_all, const bool normalize);
bool clip_b200_text_batch_encode_all(const struct clip_ctx * ctx, const int n_threads, const struct clip_tokens * seqs,
                                     const size_t n, float * vec_all, const bool normalize);

/* memory + stream helpers so that a plain C caller needs no CUDA headers */
void * clip_b200_device_malloc(const struct clip_ctx * ctx, size_t bytes);
void   clip_b200_device_free(const struct clip_ctx * ctx, void * p);
void * clip_b200_host_malloc(size_t bytes);            /* pinned */
void   clip_b200_host_free(void * p);
bool   clip_b200_memcpy_h2d(const struct clip_ctx * ctx, void * d_dst, const void * h_src, size_t bytes);
bool   clip_b200_memcpy_d2h(const struct clip_ctx * ctx, void * h_dst, const void * d_src, size_t bytes);
bool   clip_b200_synchronize(const struct clip_ctx * ctx);
void * clip_b200_get_stream(const struct clip_ctx * ctx);          /* the cudaStream_t every kernel is launched on */

/* images (vision) / sequences (text) processed per pass through the layer stack; 0 keeps the current value.  Workspaces are sized
 * once, at the first encode of a tower: a call after that is ignored for that tower. */
void   clip_b200_set_micro_batch(const struct clip_ctx * ctx, int images, int sequences);

/* measurement hooks */
uint64_t clip_b200_kernel_launches(const struct clip_ctx * ctx);   /* kernels launched by this context so far */
float    clip_b200_last_device_ms(const struct clip_ctx * ctx);    /* CUDA-event time of the last encode's kernels */
/* per-kernel-class CUDA-event time accumulated since the last call with the same kind (0 GEMM, 1 attention,
 * 2 layernorm, 3 other); *count receives the number of launches.  Enabled by env CLIP_B200_PROFILE=1. */
float    clip_b200_kernel_ms(const struct clip_ctx * ctx, int kind, uint64_t * count);
/* CUDA-event stopwatch on the launch stream: mark(slot 0..3), then elapsed(a, b) in ms (-1 on error) */
bool     clip_b200_mark(const struct clip_ctx * ctx, int slot);
float    clip_b200_mark_elapsed_ms(const struct clip_ctx * ctx, int a, int b);
const char * clip_b200_last_error(void);
const char * clip_b200_version(void);

/* TEST HOOK: run one fused-dequant GEMM  Y[M,N] = X[M,K] . W[N,K]^T (+bias) on device 0.
 * w_rows: ggml-format rows of type qtype; x: fp32 host [M,K] (rounded to the operand type on the way in);
 * y_out: fp32 host [M,N].  use_naive=1 runs the scalar debug kernel instead.  Returns 0 on success. */
int clip_b200_debug_gemm(int qtype, int operand_bf16, int M, int N, int K, int epi, int use_naive, const float * x,
                         const void * w_rows, const float * bias, const float * resid_in, float * y_out, float * ms);

/* N1 -- clip_image_batch_preprocess (clip.h:98, clip.cpp:963-1008) on the device, fused with the encode:
 * raw u8 RGB images of any size are uploaded (3 B/pixel), resized (PIL-style antialiased bicubic, a = -0.5), centre-cropped and
 * normalised on the GPU -- bit-identical to clip_image_preprocess -- and encoded like clip_image_batch_encode (clip.h:104). */
bool clip_b200_image_batch_encode_u8(const struct clip_ctx * ctx, const struct clip_image_u8_batch * imgs, float * vec, const bool normalize);
/* the device preprocess alone, results copied back into new[]-allocated clip_image_f32 buffers exactly like
 * clip_image_batch_preprocess fills them (clip.cpp:985-1007) */
bool clip_b200_image_batch_preprocess_device(const struct clip_ctx * ctx, const struct clip_image_u8_batch * in,
                                             struct clip_image_f32_batch * out);

/* TEST HOOK: one attention launch (clip.cpp:1082-1108 / 1363-1388 semantics: softmax(Q K^T [+ causal mask]) V per head, head_dim 64).
 * qkv: fp32 host [nseq*T, 3*H*64], columns Q | K | V with the 1/sqrt(64) scale already applied to Q; out: fp32 host [nseq*T, H*64].
 * use_legacy=1 runs the warp-level mma.sync kernel instead of the tcgen05 one.  ms (optional): device time of the last of 3 launches. */
int clip_b200_debug_attention(int operand_bf16, int nseq, int T, int H, int causal, int use_legacy, const float * qkv, float * out,
                              float * ms);

/* CPU-only TEST HOOKS (no context, no GPU): lossless re-tiling check, tokenizer and preprocess without a model context */
int clip_b200_debug_repack_roundtrip(int qtype, const void * rows, int N, int K);
int clip_b200_debug_rendezvous(int rank, int world, const char * path, void * out128);     /* ranks-mode id exchange (csrc/dist.cpp) */
void clip_b200_debug_shard_bounds(size_t n, int r, int w, size_t * lo, size_t * hi);        /* devices-mode sharding rule */
int clip_b200_debug_tokenize(const char * gguf_path, const char * text, int32_t * out, int cap);
int clip_b200_debug_preprocess(const uint8_t * rgb, int nx, int ny, int out_size, const float * mean, const float * stdv,
                               float * out);

#ifdef __cplusplus
}
#endif
#endif /* CLIP_B200_H */
