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

// clip_image_preprocess (clip.cpp:797-927)
bool preprocess_image(const uint8_t* src, int nx, int ny, int out_size, const float mean[3], const float stdv[3], float* dst);

bool quantize_file(const char* inp, const char* out, int itype, std::string& err);

bool load_image_file(const char* fname, std::vector<uint8_t>& rgb, int& nx, int& ny);

}  // namespace cb
