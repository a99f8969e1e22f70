// host_ops.h -- host-side pieces of the clip.h interface that stay on the CPU (tokenizer, preprocess, scoring,
// file quantizer) plus small numeric helpers shared with the loader.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace cb {

float f16_to_f32(uint16_t h);
uint16_t f32_to_f16(float f);   // round to nearest even

// dequantize_row_* (reference: ggml/src/ggml.c:1496-1606): bit-exact fp32 values of a stored row
bool dequant_row(int qtype, const uint8_t* src, float* dst, int64_t k);
// quantize_row_*_reference (ggml/src/ggml.c:914-1116): the rows clip_model_quantize writes
bool quant_row(int qtype, const float* src, uint8_t* dst, int64_t k);

struct Vocab {
    std::map<std::string, int32_t> token_to_id;
    int32_t n = 0;
};
// clip_tokenize (clip.cpp:598-679): regex word split, whole-word "</w>" lookup, greedy longest match, SOT/EOT
std::vector<int32_t> tokenize(const Vocab& v, const char* text);

// resize geometry and bicubic taps of clip_image_preprocess (clip.cpp:743-794, 797-853), shared with the device-side preprocess
// (preprocess.cu): taps for the output positions [o0, o0 + n) of one axis, k[i * ksize + t] in double exactly as the reference
// normalises them; lo[i] = first source index, cnt[i] = number of taps.
struct ResizeTaps {
    int ksize = 0;
    std::vector<double> k;
    std::vector<int> lo, cnt;
};
ResizeTaps resize_taps(int in_size, int out_size, int o0, int n);
bool preprocess_geometry(int nx, int ny, int S, int* nx3, int* ny3);   // false when the resized image would be smaller than S

// clip_image_preprocess (clip.cpp:797-927)
bool preprocess_image(const uint8_t* src, int nx, int ny, int out_size, const float mean[3], const float stdv[3], float* dst);

bool quantize_file(const char* inp, const char* out, int itype, std::string& err);

bool load_image_file(const char* fname, std::vector<uint8_t>& rgb, int& nx, int& ny);
// jpeg.cpp: baseline / progressive Huffman JPEG -> interleaved RGB, pixels identical to the reference's stb_image path
bool decode_jpeg(const uint8_t* data, size_t size, std::vector<uint8_t>& rgb, int& nx, int& ny);

}  // namespace cb
