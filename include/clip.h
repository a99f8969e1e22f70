/* Shim so that sources written against the reference (`#include "clip.h"`) build against libclip_b200.so. */
#ifndef CLIP_H
#define CLIP_H
#include "ggml/ggml.h"
#include "clip_b200.h"
#endif
