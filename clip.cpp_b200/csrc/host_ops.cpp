// host_ops.cpp -- CPU-side members of the clip.h interface.  These are the callers / data formats on either side of
// the GPU hot path.  Tokenizer, GGUF reader / writer and file handling are own
// designs; the bit-exact arithmetic kernels -- the bicubic tap table (keys_cubic / resize_taps, PIL's Resample.c as restated in
// clip.cpp:728-794) and the block quantizers (quant_row, ggml.c:914-1116) -- necessarily follow the reference's operation order
// step by step, because byte-identical output leaves no other choice.  Each function cites the lines whose results it reproduces.
#include "host_ops.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <zlib.h>

#include <algorithm>
#include <new>
#include <fstream>

#include "gguf.hpp"

namespace cb {

// ---------------------------------------------------------------------------------------------------
// fp16 <-> fp32 (IEEE binary16, round-to-nearest-even; equals the F16C conversions the reference uses on x86,
// ggml/src/ggml.c:326-333)
// ---------------------------------------------------------------------------------------------------
float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {   // subnormal: normalise
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0u));   // inf / nan
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                        // overflow -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                     // underflow -> 0
    uint32_t exp = x >> 23, man = x & 0x7FFFFFu;
    if (exp < 113) {   // subnormal half
        man |= 0x800000u;
        const uint32_t shift = 126 - exp;           // 14..24
        const uint32_t half = man >> shift, rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        uint32_t r = half;
        if (rem > mid || (rem == mid && (half & 1))) r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((exp - 112) << 10) | (man >> 13);
    const uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) r++;   // may carry into the exponent: still correct
    return (uint16_t)(sign | r);
}

// ---------------------------------------------------------------------------------------------------
// block (de)quantisation, ggml/src/ggml.c:866-911 layouts
// ---------------------------------------------------------------------------------------------------
bool dequant_row(int qt, const uint8_t* src, float* dst, int64_t k) {
    if (qt == 0) { memcpy(dst, src, (size_t)k * 4); return true; }
    if (qt == 1) { for (int64_t i = 0; i < k; i++) { uint16_t h; memcpy(&h, src + 2 * i, 2); dst[i] = f16_to_f32(h); } return true; }
    const size_t bb = ggml_type_block_bytes(qt);
    if (!bb || qt < 2 || k % 32) return false;
    for (int64_t b = 0; b < k / 32; b++) {
        const uint8_t* p = src + b * bb;
        float* y = dst + b * 32;
        uint16_t dh; memcpy(&dh, p, 2);
        const float d = f16_to_f32(dh);
        if (qt == 8) { for (int j = 0; j < 32; j++) y[j] = (float)((const int8_t*)(p + 2))[j] * d; continue; }
        float m = 0.f; uint32_t qh = 0; const uint8_t* qs;
        if (qt == 2) qs = p + 2;
        else if (qt == 3) { uint16_t mh; memcpy(&mh, p + 2, 2); m = f16_to_f32(mh); qs = p + 4; }
        else if (qt == 6) { memcpy(&qh, p + 2, 4); qs = p + 6; }
        else { uint16_t mh; memcpy(&mh, p + 2, 2); m = f16_to_f32(mh); memcpy(&qh, p + 4, 4); qs = p + 8; }
        for (int j = 0; j < 16; j++) {
            int q0 = qs[j] & 0x0F, q1 = qs[j] >> 4;
            if (qt == 6 || qt == 7) { q0 |= ((qh >> j) & 1) << 4; q1 |= ((qh >> (j + 16)) & 1) << 4; }
            if (qt == 2) { q0 -= 8; q1 -= 8; }
            if (qt == 6) { q0 -= 16; q1 -= 16; }
            if (qt == 3 || qt == 7) { y[j] = (float)q0 * d + m; y[j + 16] = (float)q1 * d + m; }
            else { y[j] = (float)q0 * d; y[j + 16] = (float)q1 * d; }
        }
    }
    return true;
}

bool quant_row(int qt, const float* x, uint8_t* dst, int64_t k) {
    const size_t bb = ggml_type_block_bytes(qt);
    if (!bb || qt < 2 || k % 32) return false;
    for (int64_t b = 0; b < k / 32; b++) {
        const float* xb = x + b * 32;
        uint8_t* p = dst + b * bb;
        if (qt == 8) {                                        // ggml.c:1097-1114
            float amax = 0.f;
            for (int j = 0; j < 32; j++) amax = std::max(amax, fabsf(xb[j]));
            const float d = amax / 127.0f, id = d ? 1.0f / d : 0.0f;
            const uint16_t dh = f32_to_f16(d); memcpy(p, &dh, 2);
            for (int j = 0; j < 32; j++) ((int8_t*)(p + 2))[j] = (int8_t)roundf(xb[j] * id);
            continue;
        }
        const bool sym = (qt == 2 || qt == 6);
        const int bits5 = (qt == 6 || qt == 7);
        float d, mn = 0.f, off;
        if (sym) {                                            // ggml.c:922-947, 1004-1036: signed abs-max
            float amax = 0.f, mx = 0.f;
            for (int j = 0; j < 32; j++) { const float v = xb[j]; if (amax < fabsf(v)) { amax = fabsf(v); mx = v; } }
            d = mx / (bits5 ? -16.0f : -8.0f);
            off = bits5 ? 16.5f : 8.5f;
        } else {                                              // ggml.c:963-988, 1052-1084: min/max
            float mx = -FLT_MAX; mn = FLT_MAX;
            for (int j = 0; j < 32; j++) { const float v = xb[j]; if (v < mn) mn = v; if (v > mx) mx = v; }
            d = (mx - mn) / (bits5 ? 31.0f : 15.0f);
            off = 0.5f;
        }
        const float id = d ? 1.0f / d : 0.0f;
        const uint16_t dh = f32_to_f16(d); memcpy(p, &dh, 2);
        uint8_t* qs = p + 2;
        if (!sym) { const uint16_t mh = f32_to_f16(mn); memcpy(p + 2, &mh, 2); qs += 2; }
        uint8_t* qhp = qs;
        if (bits5) qs += 4;
        uint32_t qh = 0;
        for (int j = 0; j < 16; j++) {
            uint8_t q0, q1;
            if (sym) {
                // x*id + 8.5f is ONE fused multiply-add in the reference binary: ggml is built with FMA enabled
                // (CLIP_NATIVE / -mfma) and GCC contracts the expression (ggml.c:937-941, 1020-1024)
                const int lim = bits5 ? 31 : 15;
                q0 = (uint8_t)std::min(lim, (int)(int8_t)fmaf(xb[j], id, off));
                q1 = (uint8_t)std::min(lim, (int)(int8_t)fmaf(xb[j + 16], id, off));
            } else {
                const float x0 = (xb[j] - mn) * id, x1 = (xb[j + 16] - mn) * id;
                if (qt == 7) { q0 = (uint8_t)(x0 + off); q1 = (uint8_t)(x1 + off); }   // no clamp (ggml.c:1074-1075)
                else { q0 = (uint8_t)std::min(15, (int)(int8_t)(x0 + off)); q1 = (uint8_t)std::min(15, (int)(int8_t)(x1 + off)); }
            }
            qs[j] = (uint8_t)((q0 & 0x0F) | ((q1 & 0x0F) << 4));
            qh |= (uint32_t)((q0 >> 4) & 1) << j;
            qh |= (uint32_t)((q1 >> 4) & 1) << (j + 16);
        }
        if (bits5) memcpy(qhp, &qh, 4);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// tokenizer: clip.cpp:598-679.  The reference splits with the std::regex
//   's|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+
// (ECMAScript, ordered alternation, "C" locale classes).  The scanner below implements that grammar directly.
// ---------------------------------------------------------------------------------------------------
namespace {
inline bool is_alpha(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_other(unsigned char c) { return !is_alpha(c) && !is_digit(c) && !is_space(c); }

size_t match_len(const std::string& s, size_t p) {
    static const char* const contractions[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
    const size_t n = s.size();
    for (const char* c : contractions) {
        const size_t l = strlen(c);
        if (s.compare(p, l, c) == 0) return l;
    }
    bool (*const classes[3])(unsigned char) = {is_alpha, is_digit, is_other};
    for (auto cls : classes) {
        size_t q = p;
        if (s[q] == ' ') q++;
        if (q < n && cls((unsigned char)s[q])) {
            while (q < n && cls((unsigned char)s[q])) q++;
            return q - p;
        }
    }
    if (is_space((unsigned char)s[p])) {
        size_t e = p;
        while (e < n && is_space((unsigned char)s[e])) e++;
        if (e == n) return e - p;          // \s+(?!\S) at end of input
        if (e - p >= 2) return e - p - 1;  // \s+(?!\S) backs off one character before a non-space
        return e - p;                      // \s+
    }
    return 1;
}
}  // namespace

std::vector<int32_t> tokenize(const Vocab& v, const char* text) {
    const std::string str = text ? text : "";
    std::vector<int32_t> out;
    out.push_back(49406);   // <|startoftext|>, hard-coded in the reference (clip.cpp:637)
    for (size_t p = 0; p < str.size();) {
        const size_t l = match_len(str, p);
        const std::string word = str.substr(p, l);
        p += l;
        const std::string whole = (word[0] == ' ' ? word.substr(1) : word) + "</w>";
        auto it = v.token_to_id.find(whole);
        if (it != v.token_to_id.end()) { out.push_back(it->second); continue; }
        for (size_t i = 0; i < word.size();) {          // greedy longest match on the raw word (clip.cpp:655-668)
            bool hit = false;
            for (size_t j = word.size(); j > i; j--) {
                auto c = v.token_to_id.find(word.substr(i, j - i));
                if (c != v.token_to_id.end()) { out.push_back(c->second); i = j; hit = true; break; }
            }
            if (!hit) { fprintf(stderr, "clip_tokenize: unknown token '%c'\n", word[i]); i++; }
        }
    }
    out.push_back(49407);   // <|endoftext|> (clip.cpp:671)
    return out;
}

// ---------------------------------------------------------------------------------------------------
// preprocess: clip.cpp:728-927 -- PIL-style separable bicubic (Keys a = -0.5) with antialias support, clamp to
// [0,255] after each pass, centre crop, (v/255 - mean)/std.  Arithmetic order is kept (double accumulation in
// ascending tap order) so results are bit-identical to the reference (tests/test_host_side.py).
// ---------------------------------------------------------------------------------------------------
namespace {
inline double keys_cubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
}  // namespace

ResizeTaps resize_taps(int in_size, int out_size, int o0, int n) {
    ResizeTaps t;
    const float in0 = 0.0f, in1 = (float)in_size;
    double support = 2.0, fs = (double)(in1 - in0) / out_size;
    if (fs < 1.0) fs = 1.0;
    support *= fs;
    t.ksize = (int)ceil(support) * 2 + 1;
    t.k.assign((size_t)n * t.ksize, 0.0);
    t.lo.resize(n);
    t.cnt.resize(n);
    const double ss = 1.0 / fs;
    for (int i = 0; i < n; i++) {
        const int o = o0 + i;
        const double center = in0 + (o + 0.5) * (in1 - in0) / out_size;
        int lo = (int)(center - support + 0.5);
        if (lo < 0) lo = 0;
        int hi = (int)(center + support + 0.5);
        if (hi > in_size) hi = in_size;
        const int cnt = hi - lo;
        double* k = &t.k[(size_t)i * t.ksize];
        double ww = 0.0;
        for (int x = 0; x < cnt; x++) { const double w = keys_cubic((x + lo - center + 0.5) * ss); k[x] = w; ww += w; }
        if (ww != 0.0) for (int x = 0; x < cnt; x++) k[x] /= ww;
        t.lo[i] = lo;
        t.cnt[i] = cnt;
    }
    return t;
}

bool preprocess_geometry(int nx, int ny, int S, int* nx3, int* ny3) {
    if (nx <= 0 || ny <= 0 || S <= 0) return false;
    const float scale = std::min((float)nx, (float)ny) / (float)S;
    *nx3 = (int)(nx / scale + 0.5f);
    *ny3 = (int)(ny / scale + 0.5f);
    return *nx3 >= S && *ny3 >= S;
}

namespace {
typedef ResizeTaps Taps;
inline float clamp255(double v) { return std::min(std::max((float)v, 0.0f), 255.0f); }
}  // namespace

// Only what the centre crop keeps is computed: output columns [xo, xo+S) of the horizontal pass, for the input rows the vertical taps
// of output rows [yo, yo+S) touch.  Every kept sample is produced by the same sequence of double operations as in the reference
// (taps in ascending order into one accumulator, clamp, float), so the result does not change -- the loops are only arranged so that
// the three channels (horizontal) and a whole row of samples (vertical) are independent accumulators side by side.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx2", "default")))      // wider lanes for the row accumulators; no FMA contraction either way
#endif
bool preprocess_image(const uint8_t* src, int nx, int ny, int S, const float mean[3], const float stdv[3], float* dst) {
    int nx3 = 0, ny3 = 0;
    if (!src || !preprocess_geometry(nx, ny, S, &nx3, &ny3)) return false;
    const int xo = (nx3 - S) / 2, yo = (ny3 - S) / 2;
    const Taps th = resize_taps(nx, nx3, xo, S), tv = resize_taps(ny, ny3, yo, S);
    const int r_lo = tv.lo[0], r_hi = tv.lo[S - 1] + tv.cnt[S - 1];           // tap windows move monotonically with the output row
    const size_t row_len = (size_t)3 * S;
    std::vector<float> tmp((size_t)(r_hi - r_lo) * row_len);
    for (int y = r_lo; y < r_hi; y++) {
        const uint8_t* row = src + (size_t)3 * nx * y;
        float* out = &tmp[(size_t)(y - r_lo) * row_len];
        for (int xx = 0; xx < S; xx++) {
            const double* k = &th.k[(size_t)xx * th.ksize];
            const uint8_t* p = row + (size_t)3 * th.lo[xx];
            const int cnt = th.cnt[xx];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int x = 0; x < cnt; x++, p += 3) {
                const double w = k[x];
                a0 += (double)p[0] * w; a1 += (double)p[1] * w; a2 += (double)p[2] * w;
            }
            out[3 * xx] = clamp255(a0); out[3 * xx + 1] = clamp255(a1); out[3 * xx + 2] = clamp255(a2);
        }
    }
    std::vector<double> acc(row_len);
    for (int yy = 0; yy < S; yy++) {
        const double* k = &tv.k[(size_t)yy * tv.ksize];
        const int lo = tv.lo[yy], cnt = tv.cnt[yy];
        std::fill(acc.begin(), acc.end(), 0.0);
        double* __restrict a = acc.data();
        for (int y = 0; y < cnt; y++) {
            const float* __restrict row = &tmp[(size_t)(lo + y - r_lo) * row_len];
            const double w = k[y];
            for (size_t i = 0; i < row_len; i++) a[i] += (double)row[i] * w;
        }
        float* out = dst + (size_t)yy * row_len;
        for (int x = 0; x < S; x++)
            for (int c = 0; c < 3; c++) out[3 * x + c] = ((clamp255(acc[3 * x + c]) / 255.0f) - mean[c]) / stdv[c];
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// clip_model_quantize: clip.cpp:1661-1844.  Quantises every 2-D tensor whose name ends in "weight"
// (regex ".*weight", clip.cpp:1711-1739) from f32/f16; everything else is copied.  The output container is
// written exactly as the reference's gguf writer lays it out (ggml.c:20541-20640): version 2 header, source KVs
// in order with general.file_type replaced and general.quantization_version = 2 appended, tensor infos with
// recomputed offsets, 32-byte alignment padding.
// ---------------------------------------------------------------------------------------------------
namespace {
void put_str(std::string& b, const std::string& s) { const uint64_t n = s.size(); b.append((const char*)&n, 8); b.append(s); }
template <class T> void put(std::string& b, T v) { b.append((const char*)&v, sizeof(T)); }
bool ends_with_weight(const std::string& n) { return n.size() >= 6 && n.compare(n.size() - 6, 6, "weight") == 0; }
}  // namespace

bool quantize_file(const char* inp, const char* outp, int itype, std::string& err) {
    if (!(itype == 2 || itype == 3 || itype == 6 || itype == 7 || itype == 8)) { err = "invalid quantization type"; return false; }
    GgufFile g;
    if (!g.parse(inp, err)) return false;
    const size_t align = 32;
    struct OutT { const GgufTensor* t; uint32_t type; std::vector<uint8_t> data; const uint8_t* ptr; size_t size; uint64_t offset; };
    std::vector<OutT> outs(g.tensors.size());
    uint64_t off = 0;
    std::vector<float> f32buf;
    for (size_t i = 0; i < g.tensors.size(); i++) {
        const GgufTensor& t = g.tensors[i];
        OutT& o = outs[i];
        o.t = &t;
        const bool q = ends_with_weight(t.name) && t.n_dims == 2;
        if (q) {
            if (t.type != 0 && t.type != 1) { err = "input must be f32 or f16"; return false; }
            const int64_t k = (int64_t)t.ne[0], rows = (int64_t)t.ne[1];
            if (k % 32) { err = "row length of " + t.name + " is not a multiple of 32"; return false; }
            const size_t rb = (size_t)(k / 32) * ggml_type_block_bytes(itype);
            o.data.resize((size_t)rows * rb);
            f32buf.resize((size_t)k);
            for (int64_t r = 0; r < rows; r++) {
                if (!dequant_row((int)t.type, t.data + (size_t)r * (t.type == 0 ? 4 : 2) * k, f32buf.data(), k) ||
                    !quant_row(itype, f32buf.data(), o.data.data() + (size_t)r * rb, k)) { err = "cannot convert tensor " + t.name; return false; }
            }
            o.type = (uint32_t)itype; o.ptr = o.data.data(); o.size = o.data.size();
        } else {
            o.type = t.type; o.ptr = t.data; o.size = t.nbytes;
        }
        o.offset = off;
        off += (o.size + align - 1) / align * align;
    }
    std::string meta;
    put<uint32_t>(meta, 0x46554747u);
    put<uint32_t>(meta, 2u);                                   // GGUF_VERSION of the vendored ggml
    put<uint64_t>(meta, (uint64_t)outs.size());
    const bool has_qv = g.find("general.quantization_version") != nullptr;
    put<uint64_t>(meta, (uint64_t)g.kvs.size() + (has_qv ? 0 : 1));
    for (const GgufKV& kv : g.kvs) {
        put_str(meta, kv.key);
        if (kv.key == "general.file_type") { put<uint32_t>(meta, GT_U32); put<uint32_t>(meta, (uint32_t)itype); continue; }
        if (kv.key == "general.quantization_version") { put<uint32_t>(meta, GT_U32); put<uint32_t>(meta, 2u); continue; }
        put<uint32_t>(meta, kv.type);
        meta.append((const char*)kv.raw, kv.raw_len);
    }
    if (!has_qv) { put_str(meta, "general.quantization_version"); put<uint32_t>(meta, GT_U32); put<uint32_t>(meta, 2u); }
    for (const OutT& o : outs) {
        put_str(meta, o.t->name);
        put<uint32_t>(meta, o.t->n_dims);
        for (uint32_t d = 0; d < o.t->n_dims; d++) put<uint64_t>(meta, o.t->ne[d]);
        put<uint32_t>(meta, o.type);
        put<uint64_t>(meta, o.offset);
    }
    meta.append((align - meta.size() % align) % align, '\0');
    std::ofstream f(outp, std::ios::binary);
    if (!f) { err = std::string("cannot open output ") + outp; return false; }
    f.write(meta.data(), (std::streamsize)meta.size());
    static const char zeros[32] = {0};
    for (const OutT& o : outs) {
        f.write((const char*)o.ptr, (std::streamsize)o.size);
        f.write(zeros, (std::streamsize)((align - o.size % align) % align));
    }
    f.close();
    if (!f) { err = "write failed"; return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// image files (the reference uses stb_image, clip.cpp:709-726): JPEG (jpeg.cpp), PNG, binary PPM (P6, maxval 255) and 24-bit
// uncompressed BMP, each decoded to the same 3-channel pixels stb_image returns.
// ---------------------------------------------------------------------------------------------------
// PNG (the reference decodes through stb_image, clip.cpp:709-726, and asks for 3 channels: alpha is dropped, grey is replicated, grey of
// 1 / 2 / 4 bits is scaled to 0..255, 16-bit samples keep their high byte).  Colour types 0/2/3/4/6, every legal bit depth, plain and
// Adam7-interlaced.  zlib does the inflate; the result is bit-identical to what stb_image returns.
static bool decode_png(const std::vector<uint8_t>& buf, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 + 25 || memcmp(buf.data(), sig, 8) != 0) return false;
    auto be32 = [&](size_t o) { return ((uint32_t)buf[o] << 24) | ((uint32_t)buf[o + 1] << 16) | ((uint32_t)buf[o + 2] << 8) | buf[o + 3]; };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, pal;
    for (size_t p = 8; p + 12 <= buf.size();) {
        const uint32_t len = be32(p);
        if (len > buf.size() - p - 12) return false;
        const uint8_t* tag = &buf[p + 4];
        const uint8_t* d = &buf[p + 8];
        if (!memcmp(tag, "IHDR", 4) && len >= 13) { w = be32(p + 8); h = be32(p + 12); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (!memcmp(tag, "PLTE", 4)) pal.assign(d, d + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        p += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || w > 65535 || h > 65535 || (uint64_t)w * h > (1ull << 28) || interlace > 1) return false;
    int ch;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break; default: return false; }
    const bool small = depth == 1 || depth == 2 || depth == 4;                 // packed samples: grey and palette only
    if (!(depth == 8 || (depth == 16 && ctype != 3) || (small && (ctype == 0 || ctype == 3)))) return false;
    const size_t bits_pp = (size_t)depth * ch;
    const size_t bps = depth == 16 ? 2 : 1, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;   // bpp: the distance the filters look back
    // Adam7 (PNG spec 8.2): pass p holds the pixels (x0 + i*dx, y0 + j*dy); a plain image is one pass covering everything
    static const int ax0[7] = {0, 4, 0, 2, 0, 1, 0}, ay0[7] = {0, 0, 4, 0, 2, 0, 1}, adx[7] = {8, 8, 4, 4, 2, 2, 1}, ady[7] = {8, 8, 8, 4, 4, 2, 2};
    struct Pass { uint32_t x0, y0, dx, dy, pw, ph; size_t stride; };
    std::vector<Pass> passes;
    size_t total = 0;
    for (int p = 0; p < (interlace ? 7 : 1); p++) {
        Pass q;
        if (interlace) { q.x0 = ax0[p]; q.y0 = ay0[p]; q.dx = adx[p]; q.dy = ady[p]; }
        else { q.x0 = q.y0 = 0; q.dx = q.dy = 1; }
        if (w <= q.x0 || h <= q.y0) continue;                                   // empty pass: no bytes in the stream
        q.pw = (w - q.x0 + q.dx - 1) / q.dx;
        q.ph = (h - q.y0 + q.dy - 1) / q.dy;
        q.stride = (q.pw * bits_pp + 7) / 8;
        total += (q.stride + 1) * q.ph;
        passes.push_back(q);
    }
    std::vector<uint8_t> raw(total);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) return false;
    const int scale = depth == 1 ? 255 : (depth == 2 ? 85 : 17), mask = (1 << depth) - 1;
    rgb.resize((size_t)w * h * 3);
    size_t at = 0;
    std::vector<uint8_t> prev, cur;
    for (const Pass& q : passes) {
        prev.assign(q.stride, 0);
        cur.resize(q.stride);
        for (uint32_t y = 0; y < q.ph; y++) {
            const uint8_t ft = raw[at];
            const uint8_t* in = &raw[at + 1];
            at += q.stride + 1;
            for (size_t i = 0; i < q.stride; i++) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int pred;
                switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                default: return false;
                }
                cur[i] = (uint8_t)(in[i] + pred);
            }
            uint8_t* row = &rgb[(size_t)(q.y0 + y * q.dy) * w * 3];
            for (uint32_t x = 0; x < q.pw; x++) {
                uint8_t* out = row + (size_t)(q.x0 + x * q.dx) * 3;
                if (small) {
                    const size_t bit = (size_t)x * depth;
                    const int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & mask;
                    if (ctype == 0) out[0] = out[1] = out[2] = (uint8_t)(v * scale);
                    else { const size_t k = (size_t)v * 3; if (k + 3 > pal.size()) return false; out[0] = pal[k]; out[1] = pal[k + 1]; out[2] = pal[k + 2]; }
                    continue;
                }
                const uint8_t* px = &cur[x * (bits_pp / 8)];       // 16-bit samples are big-endian: the first byte is the high byte stb keeps
                switch (ctype) {
                case 0: case 4: out[0] = out[1] = out[2] = px[0]; break;
                case 2: case 6: out[0] = px[0]; out[1] = px[bps]; out[2] = px[2 * bps]; break;
                case 3: { const size_t k = (size_t)px[0] * 3; if (k + 3 > pal.size()) return false; out[0] = pal[k]; out[1] = pal[k + 1]; out[2] = pal[k + 2]; } break;
                }
            }
            prev.swap(cur);
        }
    }
    nx = (int)w; ny = (int)h;
    return true;
}

// BMP as stb_image reads it (3 channels requested, so an alpha channel is simply dropped): OS/2 and Windows V3/V4/V5 headers,
// 1 / 4 / 8-bit palettes, 24-bit BGR, 16 / 32-bit with the default or BI_BITFIELDS masks (a field of n < 8 bits widens by repeating its
// bits), bottom-up or top-down.  RLE-compressed files are refused, as in the reference.
static bool decode_bmp(const std::vector<uint8_t>& buf, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    auto u16 = [&](size_t o) -> uint32_t { return o + 2 <= buf.size() ? (uint32_t)buf[o] | ((uint32_t)buf[o + 1] << 8) : 0u; };
    auto u32 = [&](size_t o) -> uint32_t { return u16(o) | (u16(o + 2) << 16); };
    const uint32_t offs = u32(10), hsz = u32(14);
    if (!(hsz == 12 || hsz == 40 || hsz == 56 || hsz == 108 || hsz == 124)) return false;
    int64_t w, h;
    uint32_t planes, bpp, comp = 0;
    if (hsz == 12) { w = u16(18); h = u16(20); planes = u16(22); bpp = u16(24); }
    else { w = (int32_t)u32(18); h = (int32_t)u32(22); planes = u16(26); bpp = u16(28); comp = u32(30); }
    if (planes != 1 || comp == 1 || comp == 2 || comp > 3 || (comp == 3 && bpp != 16 && bpp != 32)) return false;
    const bool bottom_up = h > 0;
    if (h < 0) h = -h;
    if (w <= 0 || h == 0 || w > (1 << 24) || h > (1 << 24) || w * h > (1ll << 28)) return false;
    uint32_t mask[3] = {0, 0, 0};                                    // R, G, B
    if (bpp == 16 || bpp == 32) {
        if (comp == 3) {
            if (hsz == 12) return false;
            // the masks follow the 40-byte core; for the 56-byte header the reference looks 16 bytes further (behind the header's
            // own mask fields) -- kept, the aim being the reference's pixels
            const size_t mo = hsz == 56 ? 70 : 54;
            for (int c = 0; c < 3; c++) mask[c] = u32(mo + 4 * c);
            if (hsz <= 56 && mask[0] == mask[1] && mask[1] == mask[2]) return false;
        } else if (bpp == 16) { mask[0] = 31u << 10; mask[1] = 31u << 5; mask[2] = 31u; }
        else { mask[0] = 0xffu << 16; mask[1] = 0xffu << 8; mask[2] = 0xffu; }
        if (!mask[0] || !mask[1] || !mask[2]) return false;
    }
    const size_t head = 14 + (size_t)hsz + ((hsz <= 56 && comp == 3) ? 12 : 0);     // palette (if any) starts here
    if (offs < head || offs > buf.size()) return false;
    size_t row_bytes;
    uint8_t pal[256][3];
    if (bpp == 1 || bpp == 4 || bpp == 8) {
        // (OS/2 header: the reference sizes the palette four entries short and reads uninitialised memory for the rest; every entry
        // it does load is the one read here)
        const size_t entry = hsz == 12 ? 3 : 4, n = (offs - head) / entry;
        if (n == 0 || n > 256 || head + n * entry > buf.size()) return false;
        memset(pal, 0, sizeof pal);
        for (size_t i = 0; i < n; i++) { pal[i][2] = buf[head + i * entry]; pal[i][1] = buf[head + i * entry + 1]; pal[i][0] = buf[head + i * entry + 2]; }
        row_bytes = ((size_t)w * bpp + 7) / 8;
    } else if (bpp == 16 || bpp == 24 || bpp == 32) {
        if (offs - head > 1024) return false;
        row_bytes = (size_t)w * (bpp / 8);
    } else return false;
    const size_t stride = (row_bytes + 3) & ~(size_t)3;
    if (offs + stride * (size_t)(h - 1) + row_bytes > buf.size()) return false;
    int shift[3] = {0, 0, 0}, bits[3] = {0, 0, 0};
    for (int c = 0; c < 3 && (bpp == 16 || bpp == 32); c++) {
        int hi = 31;
        while (!(mask[c] >> hi)) hi--;
        shift[c] = hi - 7;                                           // brings the field's top bit to bit 7
        bits[c] = __builtin_popcount(mask[c]);
        if (bits[c] > 8) return false;
    }
    auto widen = [](uint32_t v, int n) -> uint8_t {                   // n-bit value -> 8 bits by bit replication
        if (n == 0) return 0;
        uint32_t r = 0;
        for (int have = 0; have < 8; have += n) r = (r << n) | v;
        const int extra = ((8 + n - 1) / n) * n - 8;
        return (uint8_t)(r >> extra);
    };
    rgb.resize((size_t)w * h * 3);
    for (int64_t y = 0; y < h; y++) {
        const uint8_t* row = &buf[offs + stride * (size_t)(bottom_up ? h - 1 - y : y)];
        uint8_t* out = &rgb[(size_t)y * w * 3];
        for (int64_t x = 0; x < w; x++, out += 3) {
            if (bpp <= 8) {
                const int idx = bpp == 8 ? row[x] : (bpp == 4 ? (row[x >> 1] >> ((~x & 1) * 4)) & 15 : (row[x >> 3] >> (7 - (x & 7))) & 1);
                out[0] = pal[idx][0]; out[1] = pal[idx][1]; out[2] = pal[idx][2];
            } else if (bpp == 24 || (bpp == 32 && mask[0] == 0xff0000u && mask[1] == 0xff00u && mask[2] == 0xffu)) {
                const uint8_t* px = row + x * (bpp / 8);
                out[0] = px[2]; out[1] = px[1]; out[2] = px[0];
            } else {
                const uint8_t* px = row + x * (bpp / 8);
                const uint32_t v = bpp == 16 ? (uint32_t)px[0] | ((uint32_t)px[1] << 8) : (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16) | ((uint32_t)px[3] << 24);
                for (int c = 0; c < 3; c++) {
                    uint32_t f = v & mask[c];
                    f = shift[c] < 0 ? f << -shift[c] : f >> shift[c];
                    out[c] = widen(f >> (8 - bits[c]), bits[c]);
                }
            }
        }
    }
    nx = (int)w; ny = (int)h;
    return true;
}

static bool load_image_file_impl(const char* fname, std::vector<uint8_t>& rgb, int& nx, int& ny);
bool load_image_file(const char* fname, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    try {
        return load_image_file_impl(fname, rgb, nx, ny);
    } catch (const std::bad_alloc&) {          // a header announcing more pixels than there is memory: refuse, do not unwind through the C ABI
        return false;
    }
}

static bool load_image_file_impl(const char* fname, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    FILE* f = fopen(fname, "rb");
    if (!f) return false;
    std::vector<uint8_t> buf;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long len = ftell(f);
        if (len > 0) buf.resize((size_t)len);
    }
    rewind(f);
    const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (got != buf.size() || buf.empty()) return false;
    if (buf.size() >= 8 && buf[0] == 0x89 && buf[1] == 'P' && buf[2] == 'N' && buf[3] == 'G') return decode_png(buf, rgb, nx, ny);
    if (buf.size() >= 4 && buf[0] == 0xFF && buf[1] == 0xD8) return decode_jpeg(buf.data(), buf.size(), rgb, nx, ny);
    if (buf.size() >= 2 && buf[0] == 'P' && (buf[1] == '6' || buf[1] == '5')) {       // binary PPM / PGM, samples of at most 8 bits, no scaling
        size_t p = 2;
        int vals[3], got = 0;
        while (got < 3 && p < buf.size()) {
            while (p < buf.size() && is_space(buf[p])) p++;
            if (p < buf.size() && buf[p] == '#') { while (p < buf.size() && buf[p] != '\n') p++; continue; }
            int v = 0; bool any = false;
            while (p < buf.size() && is_digit(buf[p])) { v = v * 10 + (buf[p] - '0'); p++; any = true; }
            if (!any) return false;
            vals[got++] = v;
        }
        if (got < 3 || vals[2] < 1 || vals[2] > 255) return false;      // (16-bit PNM: the reference returns the LOW byte of every sample; refused here)
        p++;   // single whitespace after maxval
        nx = vals[0]; ny = vals[1];
        if (nx <= 0 || ny <= 0 || (uint64_t)nx * (uint64_t)ny > (1ull << 28)) return false;
        const size_t ch = buf[1] == '6' ? 3 : 1, need = (size_t)nx * ny * ch;
        if (p + need > buf.size()) return false;
        if (ch == 3) rgb.assign(buf.begin() + p, buf.begin() + p + need);
        else {
            rgb.resize(need * 3);
            for (size_t i = 0; i < need; i++) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = buf[p + i];
        }
        return true;
    }
    if (buf.size() >= 26 && buf[0] == 'B' && buf[1] == 'M') return decode_bmp(buf, rgb, nx, ny);
    return false;
}

}  // namespace cb
