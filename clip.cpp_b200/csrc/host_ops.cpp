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

#include <algorithm>
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

}  // namespace cb
