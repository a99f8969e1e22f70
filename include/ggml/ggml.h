/*
 * Minimal stand-in for the two ggml symbols the reference's callers use next to clip.h
 * (examples/main.cpp:7-8, tests/benchmark.cpp:57, models/quantize.cpp:30): wall-clock helpers.
 * Nothing else of ggml exists in this library.
 */
#ifndef CLIP_B200_GGML_SHIM_H
#define CLIP_B200_GGML_SHIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void    ggml_time_init(void);
int64_t ggml_time_ms(void);
int64_t ggml_time_us(void);
#ifdef __cplusplus
}
#endif
#endif
