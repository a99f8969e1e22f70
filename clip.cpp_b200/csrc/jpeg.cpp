// jpeg.cpp -- Huffman JPEG decoder for clip_image_load_from_file: baseline and progressive DCT, 8-bit samples, 1 / 3 / 4 components,
// restart intervals, every integer sampling ratio.  Host-side only (no CUDA).
//
// The reference hands image files to stb_image and asks for 3 channels (clip.cpp:709-726).  JPEG decoders agree on the coefficients
// but not on the last bit of the pixels: the inverse DCT, the chroma up-sampling filter and the YCbCr->RGB rounding are each
// decoder's own choice, and a +-1..3 LSB difference per pixel moves an embedding by ~1e-4 in cosine.  To keep "same file in -> same
// embedding out" this decoder makes stb_image's three choices (restated below, each at its function):
//   * inverse DCT: the LL&M "islow" integer transform with 12-bit constants, 2 guard bits kept after the column pass;
//   * up-sampling: JFIF-centred triangle filter for 2x (3/4, 1/4 taps, fixed rounding), pixel replication for other ratios;
//   * colour: 20-bit fixed point with 12-bit coefficients, the Cb term of green truncated to its upper 16 bits.
// tests/test_host_side.py checks the result byte for byte against the reference's loader on tests/golden/jpeg/*.jpg (fixtures made by
// tests/golden/make_jpeg_golden.py) and, when /root/reference is there, on the reference's own two sample JPEGs.
// Speed (profiles/r02_host_side.md; vs the reference's stb_image with its SSE2 kernels): 600x500 baseline 2.3 ms vs 2.7 ms, progressive
// 3.8 ms vs 5.0 ms; 4000x3000 baseline 126 ms vs 140 ms, progressive 170 ms vs 235 ms -- single lookup for short AC codes, AVX2 integer
// IDCT and colour rows (portable loops otherwise, same bytes), IDCT / colour rows on up to 8 threads from a megapixel up.
// Not supported (load fails, as it does in the reference): arithmetic coding, lossless / hierarchical modes, 12-bit samples.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "host_ops.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

namespace cb {
namespace {

// k-th coefficient of the zig-zag scan -> index in the row-major 8x8 block (ITU T.81 figure A.6)
const uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ---------------------------------------------------------------------------------------------------
// entropy-coded segment reader: MSB-first bits, 0xFF00 un-stuffing, stops at the first marker and feeds zero bits after it
// ---------------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    uint64_t acc = 0;      // valid bits are the top `n`
    int n = 0;
    int marker = 0;        // marker byte that ended the segment (0 = none seen yet)
    bool dry = false;      // ran past the marker / end of file: bits are padding

    void restart() { acc = 0; n = 0; marker = 0; dry = false; }

    void fill() {
        if (n <= 32 && !marker && end - p >= 4) {                      // four ordinary bytes at once (no FF among them)
            const uint32_t w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
            const uint32_t inv = ~w;                                   // a byte of w is FF <=> that byte of inv is 00
            if (((inv - 0x01010101u) & ~inv & 0x80808080u) == 0) {
                acc |= (uint64_t)w << (32 - n);
                n += 32;
                p += 4;
            }
        }
        while (n <= 56) {
            uint32_t b = 0;
            if (marker || p >= end) dry = true;
            else if (*p != 0xFF) b = *p++;
            else {
                const uint8_t* q = p + 1;
                while (q < end && *q == 0xFF) q++;                     // fill bytes
                if (q >= end) { marker = 0xD9; dry = true; }           // a lone FF at the end of the file reads as EOI
                else if (*q == 0) { b = 0xFF; p = q + 1; }             // stuffed byte
                else { marker = *q; dry = true; p = q - 1; }           // p stays on the marker's FF for the segment parser
            }
            acc |= (uint64_t)b << (56 - n);
            n += 8;
        }
    }
    uint32_t peek16() { if (n < 16) fill(); return (uint32_t)(acc >> 48); }
    void skip(int k) { acc <<= k; n -= k; }
    uint32_t bits(int k) {                                             // 1 <= k <= 16
        if (n < k) fill();
        const uint32_t v = (uint32_t)(acc >> (64 - k));
        skip(k);
        return v;
    }
    uint32_t bit() { return bits(1); }
    // T.81 F.2.2.1 RECEIVE + EXTEND: k magnitude bits -> signed value
    int receive_extend(int k) {
        const int v = (int)bits(k);
        return v < (1 << (k - 1)) ? v - (1 << k) + 1 : v;
    }
};

// canonical Huffman table (T.81 annex C): codes of each length are consecutive, symbols listed in code order
struct Huffman {
    uint8_t sym[256];
    int first_code[17], first_idx[17], count[17];
    uint16_t quick[512];   // 9-bit prefix -> (length << 8) | symbol, 0 when the code is longer than 9 bits
    int32_t ac_quick[512]; // AC tables: 9-bit prefix holding a whole (code, magnitude bits) pair -> value * 65536 | run << 8 | bits used
    bool valid = false;

    bool build(const uint8_t counts[16], const uint8_t* symbols) {
        memset(quick, 0, sizeof quick);
        int code = 0, idx = 0;
        for (int l = 1; l <= 16; l++) {
            first_code[l] = code; first_idx[l] = idx; count[l] = counts[l - 1];
            if (code + count[l] > (1 << l) || idx + count[l] > 256) return false;
            for (int i = 0; i < count[l]; i++, idx++, code++) {
                sym[idx] = symbols[idx];
                if (l <= 9) {
                    const int lo = code << (9 - l), span = 1 << (9 - l);
                    for (int j = 0; j < span; j++) quick[lo + j] = (uint16_t)((l << 8) | symbols[idx]);
                }
            }
            code <<= 1;
        }
        for (int i = 0; i < 512; i++) {
            ac_quick[i] = 0;
            const int l = quick[i] >> 8, run = (quick[i] >> 4) & 15, sz = quick[i] & 15;
            if (l == 0 || sz == 0 || l + sz > 9) continue;
            int v = (i >> (9 - l - sz)) & ((1 << sz) - 1);
            if (v < (1 << (sz - 1))) v += 1 - (1 << sz);
            ac_quick[i] = v * 65536 + (run << 8) + l + sz;
        }
        valid = true;
        return true;
    }
    int decode(BitReader& br) const {
        const uint32_t top = br.peek16();
        const uint16_t q = quick[top >> 7];
        if (q) { br.skip(q >> 8); return q & 255; }
        for (int l = 10; l <= 16; l++) {
            const int off = (int)(top >> (16 - l)) - first_code[l];
            if (off >= 0 && off < count[l]) { br.skip(l); return sym[first_idx[l] + off]; }
        }
        return -1;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;      // frame header
    int td = 0, ta = 0;                    // Huffman table selectors of the current scan
    int px = 0, py = 0;                    // samples that carry image content: ceil(X*h/hmax), ceil(Y*v/vmax)
    int bw = 0, bh = 0;                    // blocks allocated per row / column (whole MCUs)
    int dc_pred = 0;
    std::vector<int16_t> coef;             // bw*bh blocks of 64, row-major inside the block
    std::vector<uint8_t> plane;            // (bw*8) x (bh*8) samples after the inverse DCT
};

// ---------------------------------------------------------------------------------------------------
// inverse DCT.  One 8-point pass of the Loeffler-Ligtenberg-Moschytz transform in 12-bit fixed point; the products, the order of
// the additions and both rounding steps follow stb_image's integer IDCT so that planes are bit-identical to the reference's.
// ---------------------------------------------------------------------------------------------------
constexpr int fx12(float x) { return (int)(x * 4096 + 0.5); }
#if defined(__x86_64__) && defined(__GNUC__)
#define CB_SIMD_CLONES __attribute__((target_clones("avx2", "default")))     // the lane loops below compile to SIMD; AVX2 has the 32-bit multiply
#else
#define CB_SIMD_CLONES
#endif

// Eight independent 8-point passes at once, lane l taking in[k][l], k = 0..7.  out[k][l] = (even_k + bias + odd_k) >> shift and
// out[7-k][l] = (even_k + bias - odd_k) >> shift.  Unsigned 32-bit arithmetic: two's-complement wrap-around, i.e. what the reference's
// int arithmetic does on every real image, and no undefined overflow on corrupt coefficients.
typedef uint32_t u32;
inline void llm_pass8(const int32_t in[8][8], int32_t out[8][8], u32 bias, int shift) {
    for (int l = 0; l < 8; l++) {
        const u32 s0 = (u32)in[0][l], s1 = (u32)in[1][l], s2 = (u32)in[2][l], s3 = (u32)in[3][l];
        const u32 s4 = (u32)in[4][l], s5 = (u32)in[5][l], s6 = (u32)in[6][l], s7 = (u32)in[7][l];
        const u32 z = (s2 + s6) * (u32)fx12(0.5411961f);
        const u32 c2 = z + s6 * (u32)fx12(-1.847759065f);
        const u32 c3 = z + s2 * (u32)fx12(0.765366865f);
        const u32 a = (s0 + s4) * 4096u + bias, b = (s0 - s4) * 4096u + bias;
        const u32 e0 = a + c3, e3 = a - c3, e1 = b + c2, e2 = b - c2;
        u32 p3 = s7 + s3, p4 = s5 + s1, p1 = s7 + s1, p2 = s5 + s3;
        const u32 p5 = (p3 + p4) * (u32)fx12(1.175875602f);
        const u32 t0 = s7 * (u32)fx12(0.298631336f), t1 = s5 * (u32)fx12(2.053119869f);
        const u32 t2 = s3 * (u32)fx12(3.072711026f), t3 = s1 * (u32)fx12(1.501321110f);
        p1 = p5 + p1 * (u32)fx12(-0.899976223f);
        p2 = p5 + p2 * (u32)fx12(-2.562915447f);
        p3 *= (u32)fx12(-1.961570560f);
        p4 *= (u32)fx12(-0.390180644f);
        const u32 o0 = t3 + p1 + p4, o1 = t2 + p2 + p3, o2 = t1 + p2 + p4, o3 = t0 + p1 + p3;
        out[0][l] = (int32_t)(e0 + o0) >> shift; out[7][l] = (int32_t)(e0 - o0) >> shift;
        out[1][l] = (int32_t)(e1 + o1) >> shift; out[6][l] = (int32_t)(e1 - o1) >> shift;
        out[2][l] = (int32_t)(e2 + o2) >> shift; out[5][l] = (int32_t)(e2 - o2) >> shift;
        out[3][l] = (int32_t)(e3 + o3) >> shift; out[4][l] = (int32_t)(e3 - o3) >> shift;
    }
}

inline uint8_t clamp_u8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// de-quantised coefficients (row-major) -> 8x8 samples.  Column pass: 10 of the 12 fraction bits dropped (2 guard bits stay); row
// pass: 12 + 2 + 3 bits dropped with the +128 level shift folded into the rounding bias.
CB_SIMD_CLONES void idct_block_generic(const int16_t* d, uint8_t* out, int stride) {
    uint64_t ac = 0;                                            // DC-only blocks (most of a smooth image) are one flat value
    for (int i = 0; i < 16; i++) { uint64_t v; memcpy(&v, d + 4 * i, 8); ac |= i ? v : (v >> 16 << 16); }
    // (little-endian: d[0] is the low 16 bits of the first word)
    if (ac == 0) {
        const u32 col = (u32)((int32_t)d[0] * 4);               // what the column pass leaves in row 0, column 0
        const uint8_t flat = clamp_u8((int32_t)(col * 4096u + 65536u + (128u << 17)) >> 17);
        for (int r = 0; r < 8; r++) memset(out + r * stride, flat, 8);
        return;
    }
    int32_t a[8][8], b[8][8];
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) a[r][c] = d[8 * r + c];
    llm_pass8(a, b, 512u, 10);                                  // lanes = columns
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) a[c][r] = b[r][c];          // transpose: the row pass runs with lanes = rows
    llm_pass8(a, b, 65536u + (128u << 17), 17);
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) out[r * stride + c] = clamp_u8(b[c][r]);
}

#if defined(__x86_64__) && defined(__GNUC__)
#define CB_HAVE_AVX2_PATH 1
#define CB_AVX2 __attribute__((target("avx2")))
// The same two passes with the eight lanes in one 256-bit register per row and both transposes done with shuffles (the generic
// version above spends more time transposing through memory than computing).  32-bit lanes wrap like the u32 arithmetic above.
CB_AVX2 inline void llm_pass_avx2(__m256i v[8], int bias, int shift) {
    const __m256i s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3], s4 = v[4], s5 = v[5], s6 = v[6], s7 = v[7];
#define CB_MUL(x, c) _mm256_mullo_epi32((x), _mm256_set1_epi32(c))
    const __m256i z = CB_MUL(_mm256_add_epi32(s2, s6), fx12(0.5411961f));
    const __m256i c2 = _mm256_add_epi32(z, CB_MUL(s6, fx12(-1.847759065f)));
    const __m256i c3 = _mm256_add_epi32(z, CB_MUL(s2, fx12(0.765366865f)));
    const __m256i vb = _mm256_set1_epi32(bias);
    const __m256i a = _mm256_add_epi32(_mm256_slli_epi32(_mm256_add_epi32(s0, s4), 12), vb);
    const __m256i b = _mm256_add_epi32(_mm256_slli_epi32(_mm256_sub_epi32(s0, s4), 12), vb);
    const __m256i e0 = _mm256_add_epi32(a, c3), e3 = _mm256_sub_epi32(a, c3), e1 = _mm256_add_epi32(b, c2), e2 = _mm256_sub_epi32(b, c2);
    __m256i p3 = _mm256_add_epi32(s7, s3), p4 = _mm256_add_epi32(s5, s1), p1 = _mm256_add_epi32(s7, s1), p2 = _mm256_add_epi32(s5, s3);
    const __m256i p5 = CB_MUL(_mm256_add_epi32(p3, p4), fx12(1.175875602f));
    const __m256i t0 = CB_MUL(s7, fx12(0.298631336f)), t1 = CB_MUL(s5, fx12(2.053119869f));
    const __m256i t2 = CB_MUL(s3, fx12(3.072711026f)), t3 = CB_MUL(s1, fx12(1.501321110f));
    p1 = _mm256_add_epi32(p5, CB_MUL(p1, fx12(-0.899976223f)));
    p2 = _mm256_add_epi32(p5, CB_MUL(p2, fx12(-2.562915447f)));
    p3 = CB_MUL(p3, fx12(-1.961570560f));
    p4 = CB_MUL(p4, fx12(-0.390180644f));
#undef CB_MUL
    const __m256i o0 = _mm256_add_epi32(t3, _mm256_add_epi32(p1, p4)), o1 = _mm256_add_epi32(t2, _mm256_add_epi32(p2, p3));
    const __m256i o2 = _mm256_add_epi32(t1, _mm256_add_epi32(p2, p4)), o3 = _mm256_add_epi32(t0, _mm256_add_epi32(p1, p3));
    const __m128i sh = _mm_cvtsi32_si128(shift);
    v[0] = _mm256_sra_epi32(_mm256_add_epi32(e0, o0), sh); v[7] = _mm256_sra_epi32(_mm256_sub_epi32(e0, o0), sh);
    v[1] = _mm256_sra_epi32(_mm256_add_epi32(e1, o1), sh); v[6] = _mm256_sra_epi32(_mm256_sub_epi32(e1, o1), sh);
    v[2] = _mm256_sra_epi32(_mm256_add_epi32(e2, o2), sh); v[5] = _mm256_sra_epi32(_mm256_sub_epi32(e2, o2), sh);
    v[3] = _mm256_sra_epi32(_mm256_add_epi32(e3, o3), sh); v[4] = _mm256_sra_epi32(_mm256_sub_epi32(e3, o3), sh);
}

CB_AVX2 inline void transpose8_avx2(__m256i v[8]) {
    const __m256i t0 = _mm256_unpacklo_epi32(v[0], v[1]), t1 = _mm256_unpackhi_epi32(v[0], v[1]);
    const __m256i t2 = _mm256_unpacklo_epi32(v[2], v[3]), t3 = _mm256_unpackhi_epi32(v[2], v[3]);
    const __m256i t4 = _mm256_unpacklo_epi32(v[4], v[5]), t5 = _mm256_unpackhi_epi32(v[4], v[5]);
    const __m256i t6 = _mm256_unpacklo_epi32(v[6], v[7]), t7 = _mm256_unpackhi_epi32(v[6], v[7]);
    const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2), u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
    const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6), u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
    v[0] = _mm256_permute2x128_si256(u0, u4, 0x20); v[1] = _mm256_permute2x128_si256(u1, u5, 0x20);
    v[2] = _mm256_permute2x128_si256(u2, u6, 0x20); v[3] = _mm256_permute2x128_si256(u3, u7, 0x20);
    v[4] = _mm256_permute2x128_si256(u0, u4, 0x31); v[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
    v[6] = _mm256_permute2x128_si256(u2, u6, 0x31); v[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

// eight int32 -> eight bytes, clamped to 0..255 (signed saturation to 16 bits, then unsigned saturation to 8)
CB_AVX2 inline __m128i clamp8_avx2(__m256i x) {
    const __m128i w = _mm_packs_epi32(_mm256_castsi256_si128(x), _mm256_extracti128_si256(x, 1));
    return _mm_packus_epi16(w, w);
}

CB_AVX2 void idct_block_avx2(const int16_t* d, uint8_t* out, int stride) {
    __m256i v[8];
    __m128i any = _mm_setzero_si128();
    for (int r = 0; r < 8; r++) {
        const __m128i row = _mm_loadu_si128((const __m128i*)(d + 8 * r));
        any = _mm_or_si128(any, r ? row : _mm_srli_si128(row, 2));   // every coefficient but d[0]
        v[r] = _mm256_cvtepi16_epi32(row);
    }
    if (_mm_testz_si128(any, any)) {                                 // DC only: one flat value
        const u32 col = (u32)((int32_t)d[0] * 4);
        const int flat = (int32_t)(col * 4096u + 65536u + (128u << 17)) >> 17;
        const __m128i f = _mm_set1_epi8((char)(flat < 0 ? 0 : (flat > 255 ? 255 : flat)));
        for (int r = 0; r < 8; r++) _mm_storel_epi64((__m128i*)(out + r * stride), f);
        return;
    }
    llm_pass_avx2(v, 512, 10);                                       // lanes = columns
    transpose8_avx2(v);
    llm_pass_avx2(v, 65536 + (128 << 17), 17);                       // lanes = rows
    transpose8_avx2(v);
    for (int r = 0; r < 8; r++) _mm_storel_epi64((__m128i*)(out + r * stride), clamp8_avx2(v[r]));
}
#endif

// CLIP_B200_JPEG_SIMD=0 keeps the portable loops (the tests run both and expect the same bytes)
bool use_avx2() {
#ifdef CB_HAVE_AVX2_PATH
    static const bool on = [] { const char* e = getenv("CLIP_B200_JPEG_SIMD"); return !(e && *e == '0') && __builtin_cpu_supports("avx2"); }();
    return on;
#else
    return false;
#endif
}

inline void idct_block(const int16_t* d, uint8_t* out, int stride) {
#ifdef CB_HAVE_AVX2_PATH
    if (use_avx2()) { idct_block_avx2(d, out, stride); return; }
#endif
    idct_block_generic(d, out, stride);
}

// ---------------------------------------------------------------------------------------------------
// up-sampling of one output row.  near / far are the two low-resolution rows that straddle it (equal at the image's top and
// bottom edge).  Taps and rounding follow the reference's decoder; ratios other than 1 and 2 replicate.
// ---------------------------------------------------------------------------------------------------
CB_SIMD_CLONES void upsample_row(uint8_t* out, const uint8_t* near, const uint8_t* far, int w, int hs, int vs, int16_t* tmp) {
    if (hs == 1 && vs == 1) { memcpy(out, near, (size_t)w); return; }
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        return;
    }
    if (hs == 2 && vs == 1) {
        if (w == 1) { out[0] = out[1] = near[0]; return; }
        out[0] = near[0];
        out[2 * w - 1] = near[w - 1];
        for (int i = 0; i + 1 < w; i++) out[2 * i + 1] = (uint8_t)((3 * near[i] + near[i + 1] + 2) >> 2);
        for (int i = 1; i + 1 < w; i++) out[2 * i] = (uint8_t)((3 * near[i] + near[i - 1] + 2) >> 2);
        // the reference's decoder weighs the last interior sample 3:1 towards near[w-2], not near[w-1]; kept, since the goal is the same pixels
        out[2 * w - 2] = (uint8_t)((3 * near[w - 2] + near[w - 1] + 2) >> 2);
        return;
    }
    if (hs == 2 && vs == 2) {
        for (int i = 0; i < w; i++) tmp[i] = (int16_t)(3 * near[i] + far[i]);      // vertical pass first, kept at 4x scale
        out[0] = (uint8_t)((tmp[0] + 2) >> 2);
        out[2 * w - 1] = (uint8_t)((tmp[w - 1] + 2) >> 2);
        for (int i = 1; i < w; i++) {
            out[2 * i - 1] = (uint8_t)((3 * tmp[i - 1] + tmp[i] + 8) >> 4);
            out[2 * i] = (uint8_t)((3 * tmp[i] + tmp[i - 1] + 8) >> 4);
        }
        return;
    }
    for (int i = 0; i < w; i++)
        for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
}

// YCbCr -> RGB, one pixel: 20 fraction bits, coefficients rounded to 12 bits then shifted up by 8; green's Cb product loses its
// low 16 bits before the sum (this is what keeps the reference's SIMD and scalar paths equal, and it shows in the last bit)
constexpr int fx20(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
inline void ycc_to_rgb(int y, int cb, int cr, uint8_t* out) {
    const int yf = (y << 20) + (1 << 19);
    cb -= 128; cr -= 128;
    const int r = yf + cr * fx20(1.40200f);
    const int g = yf + cr * -fx20(0.71414f) + (int)((uint32_t)(cb * -fx20(0.34414f)) & 0xffff0000u);
    const int b = yf + cb * fx20(1.77200f);
    out[0] = clamp_u8(r >> 20); out[1] = clamp_u8(g >> 20); out[2] = clamp_u8(b >> 20);
}

// x*y/255 rounded, for the K channel of CMYK / YCCK files
inline uint8_t mul255(uint8_t x, uint8_t y) {
    const uint32_t t = (uint32_t)x * y + 128;
    return (uint8_t)((t + (t >> 8)) >> 8);
}

CB_SIMD_CLONES void ycc_row_generic(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint8_t* out, int n) {
    for (int x = 0; x < n; x++) ycc_to_rgb(y[x], cb[x], cr[x], out + 3 * x);
}

#ifdef CB_HAVE_AVX2_PATH
// eight pixels per step in 32-bit lanes (same arithmetic as ycc_to_rgb), bytes interleaved to R,G,B triples with two byte shuffles
CB_AVX2 void ycc_row_avx2(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint8_t* out, int n) {
    const __m256i k128 = _mm256_set1_epi32(128), half = _mm256_set1_epi32(1 << 19), hi16 = _mm256_set1_epi32((int)0xffff0000u);
    const __m256i c_r = _mm256_set1_epi32(fx20(1.40200f)), c_g1 = _mm256_set1_epi32(-fx20(0.71414f)), c_g2 = _mm256_set1_epi32(-fx20(0.34414f)),
                  c_b = _mm256_set1_epi32(fx20(1.77200f));
    // from rg = r0 g0 r1 g1 .. r7 g7 and b = b0..b7: bytes 0..15 and 16..23 of the output
    const __m128i m_rg0 = _mm_setr_epi8(0, 1, -1, 2, 3, -1, 4, 5, -1, 6, 7, -1, 8, 9, -1, 10), m_b0 = _mm_setr_epi8(-1, -1, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1);
    const __m128i m_rg1 = _mm_setr_epi8(11, -1, 12, 13, -1, 14, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1), m_b1 = _mm_setr_epi8(-1, 5, -1, -1, 6, -1, -1, 7, -1, -1, -1, -1, -1, -1, -1, -1);
    int x = 0;
    for (; x + 8 <= n; x += 8) {
        const __m256i yy = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(y + x)));
        const __m256i b_ = _mm256_sub_epi32(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(cb + x))), k128);
        const __m256i r_ = _mm256_sub_epi32(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(cr + x))), k128);
        const __m256i yf = _mm256_add_epi32(_mm256_slli_epi32(yy, 20), half);
        const __m256i r = _mm256_srai_epi32(_mm256_add_epi32(yf, _mm256_mullo_epi32(r_, c_r)), 20);
        const __m256i g = _mm256_srai_epi32(_mm256_add_epi32(_mm256_add_epi32(yf, _mm256_mullo_epi32(r_, c_g1)), _mm256_and_si256(_mm256_mullo_epi32(b_, c_g2), hi16)), 20);
        const __m256i b = _mm256_srai_epi32(_mm256_add_epi32(yf, _mm256_mullo_epi32(b_, c_b)), 20);
        const __m128i rg = _mm_unpacklo_epi8(clamp8_avx2(r), clamp8_avx2(g)), bb = clamp8_avx2(b);
        _mm_storeu_si128((__m128i*)(out + 3 * x), _mm_or_si128(_mm_shuffle_epi8(rg, m_rg0), _mm_shuffle_epi8(bb, m_b0)));
        _mm_storel_epi64((__m128i*)(out + 3 * x + 16), _mm_or_si128(_mm_shuffle_epi8(rg, m_rg1), _mm_shuffle_epi8(bb, m_b1)));
    }
    for (; x < n; x++) ycc_to_rgb(y[x], cb[x], cr[x], out + 3 * x);
}
#endif

inline void ycc_row(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint8_t* out, int n) {
#ifdef CB_HAVE_AVX2_PATH
    if (use_avx2()) { ycc_row_avx2(y, cb, cr, out, n); return; }
#endif
    ycc_row_generic(y, cb, cr, out, n);
}

// [0, n) split into contiguous ranges over `threads` workers (the calling thread takes the first)
template <class F>
void parallel_ranges(int n, int threads, F&& body) {
    if (threads > n) threads = n;
    if (threads <= 1) { body(0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++)
        pool.emplace_back([&body, n, threads, t] { body((int)((int64_t)n * t / threads), (int)((int64_t)n * (t + 1) / threads)); });
    body(0, n / threads);
    for (std::thread& th : pool) th.join();
}

// Entropy decoding is serial; the inverse DCT and the colour rows are not.  Images of a megapixel and more fan those two phases out over
// up to 8 host threads (CLIP_B200_DECODE_THREADS overrides; 1 = never).
int decode_threads(int64_t pixels) {
    const char* e = getenv("CLIP_B200_DECODE_THREADS");
    if (e && *e) { const int v = atoi(e); return v < 1 ? 1 : (v > 64 ? 64 : v); }
    if (pixels < (1 << 20)) return 1;
    const unsigned hw = std::thread::hardware_concurrency();
    return hw >= 8 ? 8 : (hw > 1 ? (int)hw : 1);
}

// ---------------------------------------------------------------------------------------------------
struct Decoder {
    const uint8_t* base;
    size_t size;
    size_t pos = 0;

    int width = 0, height = 0, ncomp = 0;
    bool progressive = false, have_frame = false;
    bool jfif = false;
    int adobe_transform = -1;
    int restart_interval = 0;
    int hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    uint16_t quant[4][64];                 // natural order
    Huffman dc_tab[4], ac_tab[4];
    Component comp[4];

    // current scan
    int scan_n = 0, order[4] = {0, 0, 0, 0};
    int ss = 0, se = 63, ah = 0, al = 0;
    int eob_run = 0;
    BitReader br;

    Decoder(const uint8_t* b, size_t n) : base(b), size(n) { memset(quant, 0, sizeof quant); }

    bool need(size_t k) const { return pos + k <= size; }
    int u8() { return pos < size ? base[pos++] : 0; }
    int u16() { const int hi = u8(); return (hi << 8) | u8(); }

    // next marker at or after pos: skips anything that is not FF xx (xx != 00, FF).  Returns 0 at the end of the data.
    int next_marker() {
        while (pos + 1 < size) {
            if (base[pos] != 0xFF) { pos++; continue; }
            const uint8_t c = base[pos + 1];
            if (c == 0xFF) { pos++; continue; }
            pos += 2;
            if (c != 0) return c;
        }
        pos = size;
        return 0;
    }

    bool read_dqt() {
        int len = u16() - 2;
        while (len > 0) {
            const int q = u8(), wide = q >> 4, t = q & 15;
            if (wide > 1 || t > 3) return false;
            if (!need(wide ? 128 : 64)) return false;
            for (int k = 0; k < 64; k++) quant[t][kNatural[k]] = (uint16_t)(wide ? u16() : u8());
            len -= wide ? 129 : 65;
        }
        return len == 0;
    }

    bool read_dht() {
        int len = u16() - 2;
        while (len > 0) {
            const int q = u8(), cls = q >> 4, t = q & 15;
            if (cls > 1 || t > 3 || !need(16)) return false;
            uint8_t counts[16], symbols[256];
            int total = 0;
            for (int i = 0; i < 16; i++) { counts[i] = (uint8_t)u8(); total += counts[i]; }
            if (total > 256 || !need((size_t)total)) return false;
            for (int i = 0; i < total && i < 256; i++) symbols[i] = (uint8_t)u8();
            if (!(cls ? ac_tab[t] : dc_tab[t]).build(counts, symbols)) return false;
            len -= 17 + total;
        }
        return len == 0;
    }

    bool read_app(int m) {
        int len = u16();
        if (len < 2) return false;
        len -= 2;
        const size_t next = pos + (size_t)len;
        if (m == 0xE0 && len >= 5 && need(5) && memcmp(base + pos, "JFIF\0", 5) == 0) jfif = true;
        if (m == 0xEE && len >= 12 && need(12) && memcmp(base + pos, "Adobe\0", 6) == 0) adobe_transform = base[pos + 11];
        pos = next < size ? next : size;
        return true;
    }

    bool read_frame(int m) {
        if (have_frame) return false;
        const int len = u16();
        if (u8() != 8) return false;                             // sample precision
        height = u16(); width = u16(); ncomp = u8();
        if (width <= 0 || height <= 0) return false;             // height 0 = "defined later by DNL": not handled, as in the reference
        if (!(ncomp == 1 || ncomp == 3 || ncomp == 4) || len != 8 + 3 * ncomp) return false;
        if ((uint64_t)width * (uint64_t)height > (1ull << 28)) return false;
        progressive = (m == 0xC2);
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.id = u8();
            const int q = u8();
            c.h = q >> 4; c.v = q & 15; c.tq = u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return false;
            if (c.h > hmax) hmax = c.h;
            if (c.v > vmax) vmax = c.v;
        }
        for (int i = 0; i < ncomp; i++)
            if (hmax % comp[i].h || vmax % comp[i].v) return false;      // fractional ratios: refused by the reference too
        mcus_x = (width + 8 * hmax - 1) / (8 * hmax);
        mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.px = (width * c.h + hmax - 1) / hmax;
            c.py = (height * c.v + vmax - 1) / vmax;
            c.bw = mcus_x * c.h;
            c.bh = mcus_y * c.v;
            c.coef.assign((size_t)c.bw * c.bh * 64, 0);
            c.plane.assign((size_t)c.bw * c.bh * 64, 0);
        }
        have_frame = true;
        return true;
    }

    bool read_scan_header() {
        const int len = u16();
        scan_n = u8();
        if (scan_n < 1 || scan_n > ncomp || len != 6 + 2 * scan_n) return false;
        for (int i = 0; i < scan_n; i++) {
            const int id = u8(), q = u8();
            int which = -1;
            for (int k = 0; k < ncomp; k++)
                if (comp[k].id == id) { which = k; break; }
            if (which < 0 || (q >> 4) > 3 || (q & 15) > 3) return false;
            comp[which].td = q >> 4;
            comp[which].ta = q & 15;
            order[i] = which;
        }
        ss = u8(); se = u8();
        const int a = u8();
        ah = a >> 4; al = a & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return false;
            if (ss == 0 && se != 0) return false;                // DC and AC never share a scan
            if (ss != 0 && scan_n != 1) return false;            // AC scans are single-component
        } else {
            if (ss != 0 || ah != 0 || al != 0) return false;
            se = 63;
        }
        return true;
    }

    // ---- block decoders -------------------------------------------------------------------------------
    // sequential: DC difference + run/size coded AC, coefficients de-quantised as they are stored (16-bit wrap like the reference)
    bool block_sequential(Component& c, int16_t* blk) {
        const Huffman& hd = dc_tab[c.td];
        const Huffman& ha = ac_tab[c.ta];
        const uint16_t* q = quant[c.tq];
        if (!hd.valid || !ha.valid) return false;
        memset(blk, 0, 64 * sizeof(int16_t));
        const int t = hd.decode(br);
        if (t < 0 || t > 15) return false;
        c.dc_pred = (int)((uint32_t)c.dc_pred + (uint32_t)(t ? br.receive_extend(t) : 0));
        blk[0] = (int16_t)((uint32_t)c.dc_pred * q[0]);
        for (int k = 1; k < 64;) {
            const int32_t fast = ha.ac_quick[br.peek16() >> 7];
            if (fast) {                                          // short code + small value: one lookup
                br.skip(fast & 255);
                k += (fast >> 8) & 15;
                if (k > 63) return false;
                const int nat = kNatural[k++];
                blk[nat] = (int16_t)((fast >> 16) * q[nat]);
                continue;
            }
            const int rs = ha.decode(br);
            if (rs < 0) return false;
            const int run = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (run != 15) break;                            // end of block
                k += 16;
                continue;
            }
            k += run;
            if (k > 63) return false;
            const int nat = kNatural[k++];
            blk[nat] = (int16_t)(br.receive_extend(sz) * q[nat]);
        }
        return true;
    }

    // progressive DC (T.81 G.1.2.1): first pass stores the prediction-coded value shifted by Al, later passes add one bit
    bool block_dc_progressive(Component& c, int16_t* blk) {
        if (ah == 0) {
            const Huffman& hd = dc_tab[c.td];
            if (!hd.valid) return false;
            memset(blk, 0, 64 * sizeof(int16_t));
            const int t = hd.decode(br);
            if (t < 0 || t > 15) return false;
            c.dc_pred = (int)((uint32_t)c.dc_pred + (uint32_t)(t ? br.receive_extend(t) : 0));
            blk[0] = (int16_t)((uint32_t)c.dc_pred << al);
        } else if (br.bit()) {
            blk[0] = (int16_t)(blk[0] + (1 << al));
        }
        return true;
    }

    // one correction bit for an already non-zero coefficient (G.1.2.3): move it away from zero by 1 << Al
    inline void refine(int16_t& v, int step) {
        if (br.bit() && (v & step) == 0) v = (int16_t)(v > 0 ? v + step : v - step);
    }

    // progressive AC, band [ss, se] (G.1.2.2 first pass, G.1.2.3 refinement), with end-of-band runs spanning blocks
    bool block_ac_progressive(Component& c, int16_t* blk) {
        const Huffman& ha = ac_tab[c.ta];
        if (!ha.valid) return false;
        const int step = 1 << al;
        if (ah == 0) {
            if (eob_run) { eob_run--; return true; }
            for (int k = ss; k <= se;) {
                const int32_t fast = ha.ac_quick[br.peek16() >> 7];
                if (fast) {
                    br.skip(fast & 255);
                    k += (fast >> 8) & 15;
                    if (k > 63) return false;
                    blk[kNatural[k++]] = (int16_t)((fast >> 16) * step);
                    continue;
                }
                const int rs = ha.decode(br);
                if (rs < 0) return false;
                const int run = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (run < 15) {                              // EOBn: this block and (2^run + extra bits - 1) more end here
                        eob_run = (1 << run) - 1;
                        if (run) eob_run += (int)br.bits(run);
                        break;
                    }
                    k += 16;
                    continue;
                }
                k += run;
                if (k > 63) return false;
                blk[kNatural[k++]] = (int16_t)(br.receive_extend(sz) * step);
            }
            return true;
        }
        if (eob_run) {
            eob_run--;
            for (int k = ss; k <= se; k++) {
                int16_t& v = blk[kNatural[k]];
                if (v) refine(v, step);
            }
            return true;
        }
        for (int k = ss; k <= se;) {
            const int rs = ha.decode(br);
            if (rs < 0) return false;
            int run = rs >> 4;
            const int sz = rs & 15;
            int fresh = 0;                                       // value of the newly non-zero coefficient, if this code carries one
            if (sz == 0) {
                if (run < 15) {
                    eob_run = (1 << run) - 1;
                    if (run) eob_run += (int)br.bits(run);
                    run = 64;                                    // only correction bits remain in this block
                }
            } else {
                if (sz != 1) return false;
                fresh = br.bit() ? step : -step;
            }
            // skip `run` zero-history coefficients (correcting the non-zero ones passed on the way), then place `fresh`
            while (k <= se) {
                int16_t& v = blk[kNatural[k++]];
                if (v) refine(v, step);
                else if (run == 0) { v = (int16_t)fresh; break; }
                else run--;
            }
        }
        return true;
    }

    // ---- one scan -------------------------------------------------------------------------------------
    void reset_entropy() {
        br.restart();
        for (int i = 0; i < 4; i++) comp[i].dc_pred = 0;
        eob_run = 0;
    }

    // after `restart_interval` MCUs: realign on the RSTn marker.  Returns false when the stream does not continue with one
    // (the scan then ends early with what has been decoded, like the reference).
    bool take_restart() {
        br.acc = 0; br.n = 0;
        br.fill();
        if (br.marker < 0xD0 || br.marker > 0xD7) return false;
        br.p += 2;                                               // step over FF Dn
        reset_entropy();
        return true;
    }

    bool decode_block(Component& c, int bx, int by) {
        int16_t* blk = &c.coef[((size_t)by * c.bw + bx) * 64];
        if (!progressive) return block_sequential(c, blk);
        return ss == 0 ? block_dc_progressive(c, blk) : block_ac_progressive(c, blk);
    }

    bool decode_scan() {
        br.p = base + pos;
        br.end = base + size;
        reset_entropy();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        bool ok = true, more = true;
        if (scan_n == 1) {                                       // non-interleaved: the component's own block grid, one block per MCU
            Component& c = comp[order[0]];
            const int w = (c.px + 7) >> 3, h = (c.py + 7) >> 3;
            for (int by = 0; by < h && ok && more; by++)
                for (int bx = 0; bx < w && ok && more; bx++) {
                    ok = decode_block(c, bx, by);
                    if (ok && --todo <= 0) { more = take_restart(); todo = restart_interval; }
                }
        } else {
            for (int my = 0; my < mcus_y && ok && more; my++)
                for (int mx = 0; mx < mcus_x && ok && more; mx++) {
                    for (int k = 0; k < scan_n && ok; k++) {
                        Component& c = comp[order[k]];
                        for (int y = 0; y < c.v && ok; y++)
                            for (int x = 0; x < c.h && ok; x++) ok = decode_block(c, mx * c.h + x, my * c.v + y);
                    }
                    if (ok && --todo <= 0) { more = take_restart(); todo = restart_interval; }
                }
        }
        pos = (size_t)(br.p - base);
        return ok;
    }

    // ---- samples --------------------------------------------------------------------------------------
    void reconstruct(int threads) {
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            const int w = (c.px + 7) >> 3, h = (c.py + 7) >> 3, stride = c.bw * 8;
            const uint16_t* q = quant[c.tq];
            parallel_ranges(h, threads, [&](int by0, int by1) {
                for (int by = by0; by < by1; by++)
                    for (int bx = 0; bx < w; bx++) {
                        int16_t* blk = &c.coef[((size_t)by * c.bw + bx) * 64];
                        if (progressive)
                            for (int k = 0; k < 64; k++) blk[k] = (int16_t)(blk[k] * q[k]);
                        idct_block(blk, &c.plane[(size_t)by * 8 * stride + (size_t)bx * 8], stride);
                    }
            });
        }
    }

    void rows_to_rgb(uint8_t* rgb, int j0, int j1) const {
        std::vector<uint8_t> line[4];
        for (int i = 0; i < ncomp; i++) line[i].resize((size_t)width + 8);
        std::vector<int16_t> tmp((size_t)width + 8);
        const bool named_rgb = ncomp == 3 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B';
        const bool is_rgb = ncomp == 3 && (named_rgb || (adobe_transform == 0 && !jfif));
        for (int j = j0; j < j1; j++) {
            for (int i = 0; i < ncomp; i++) {
                const Component& c = comp[i];
                const int hs = hmax / c.h, vs = vmax / c.v, stride = c.bw * 8;
                // output row j sits in the lower half of low-res row `k` or the upper half of the next one: the nearer row gets the
                // 3/4 tap.  Rows are clamped to the component's real height, so the MCU padding below the image is never read.
                const int t = j + (vs >> 1), k = t / vs;
                const bool lower = (t % vs) >= (vs >> 1);
                const int last = c.py - 1;
                const int r1 = k < last ? k : last, r0 = k == 0 ? 0 : (k - 1 < last ? k - 1 : last);
                const uint8_t* near = &c.plane[(size_t)(lower ? r1 : r0) * stride];
                const uint8_t* far = &c.plane[(size_t)(lower ? r0 : r1) * stride];
                upsample_row(line[i].data(), near, far, (width + hs - 1) / hs, hs, vs, tmp.data());
            }
            uint8_t* out = rgb + (size_t)j * width * 3;
            if (ncomp == 1) {
                for (int x = 0; x < width; x++) out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = line[0][x];
            } else if (is_rgb) {
                for (int x = 0; x < width; x++) { out[3 * x] = line[0][x]; out[3 * x + 1] = line[1][x]; out[3 * x + 2] = line[2][x]; }
            } else if (ncomp == 4 && adobe_transform == 0) {     // CMYK stored inverted (Adobe): colour * K / 255
                for (int x = 0; x < width; x++)
                    for (int ch = 0; ch < 3; ch++) out[3 * x + ch] = mul255(line[ch][x], line[3][x]);
            } else {
                ycc_row(line[0].data(), line[1].data(), line[2].data(), out, width);
                if (ncomp == 4 && adobe_transform == 2)          // YCCK
                    for (int x = 0; x < 3 * width; x++) out[x] = mul255((uint8_t)(255 - out[x]), line[3][x / 3]);
            }
        }
    }

    void to_rgb(std::vector<uint8_t>& rgb, int threads) {
        rgb.resize((size_t)width * height * 3);
        uint8_t* dst = rgb.data();
        parallel_ranges(height, threads, [&](int j0, int j1) { rows_to_rgb(dst, j0, j1); });
    }

    bool run(std::vector<uint8_t>& rgb, int& nx, int& ny) {
        if (size < 4 || base[0] != 0xFF || base[1] != 0xD8) return false;
        pos = 2;
        bool scanned = false;
        for (;;) {
            int m = next_marker();
            if (m >= 0xD0 && m <= 0xD7) continue;               // stray restart marker between segments
            if (m == 0 || m == 0xD9) break;                      // end of image (or of the data: keep what was decoded)
            bool ok;
            if (m == 0xDB) ok = read_dqt();
            else if (m == 0xC4) ok = read_dht();
            else if (m == 0xDD) { ok = u16() == 4; restart_interval = u16(); }
            else if (m == 0xC0 || m == 0xC1 || m == 0xC2) ok = read_frame(m);
            else if (m == 0xDC) { ok = u16() == 4 && u16() == height; }
            else if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) ok = read_app(m);
            else if (m == 0xDA) {
                if (!have_frame) return false;
                ok = read_scan_header() && decode_scan();
                scanned = scanned || ok;
            } else return false;                                 // SOF3/5..15 (lossless, hierarchical, arithmetic), DAC, ...
            if (!ok) return false;
        }
        if (!have_frame || !scanned) return false;
        const int threads = decode_threads((int64_t)width * height);
        reconstruct(threads);
        to_rgb(rgb, threads);
        nx = width; ny = height;
        return true;
    }
};

}  // namespace

bool decode_jpeg(const uint8_t* data, size_t size, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    Decoder d(data, size);
    return d.run(rgb, nx, ny);
}

}  // namespace cb
