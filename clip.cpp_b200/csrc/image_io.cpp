// image_io.cpp -- clip_image_load_from_file's decoders other than JPEG (jpeg.cpp): PNG, BMP, binary PGM / PPM.  Host-side only.
// The reference hands files to stb_image and asks for 3 channels (clip.cpp:709-726); every decoder here returns those same pixels.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <zlib.h>

#include <new>
#include <vector>

#include "host_ops.h"

namespace cb {

namespace {
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
}  // namespace

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

// GIF, first frame only, as stb_image composes it for a 3-channel request: the canvas starts black; every pixel of the first image
// whose colour is not the transparent index is painted; canvas pixels the image does not cover take the background colour when the
// background index is non-zero; transparent pixels stay black.  GIF87a / GIF89a, global or local colour table, interlaced rows.
static bool decode_gif(const std::vector<uint8_t>& buf, std::vector<uint8_t>& rgb, int& nx, int& ny) {
    size_t p = 0;
    bool eof = false;
    auto u8 = [&]() -> int { if (p < buf.size()) return buf[p++]; eof = true; return 0; };
    auto u16 = [&]() -> int { const int lo = u8(); return lo | (u8() << 8); };
    if (buf.size() < 13 || memcmp(buf.data(), "GIF8", 4) != 0 || (buf[4] != '7' && buf[4] != '9') || buf[5] != 'a') return false;
    p = 6;
    const int W = u16(), H = u16(), flags = u8(), bgindex = u8();
    u8();                                                              // pixel aspect ratio
    if (W <= 0 || H <= 0 || (int64_t)W * H > (1ll << 28)) return false;
    uint8_t gpal[256][3], lpal[256][3];
    memset(gpal, 0, sizeof gpal);
    if (flags & 0x80)
        for (int i = 0, n = 2 << (flags & 7); i < n; i++) { gpal[i][0] = (uint8_t)u8(); gpal[i][1] = (uint8_t)u8(); gpal[i][2] = (uint8_t)u8(); }
    int transparent = -1;                                              // index made transparent by the last graphic control extension
    for (;;) {
        if (eof) return false;
        const int tag = u8();
        if (tag == 0x21) {                                             // extension
            const int ext = u8();
            int len;
            if (ext == 0xF9) {
                len = u8();
                if (len == 4) {
                    const int eflags = u8();
                    u16();                                             // delay
                    const int t = u8();
                    transparent = (eflags & 1) ? t : -1;
                } else { p += (size_t)len; continue; }
            }
            while ((len = u8()) != 0 && !eof) p += (size_t)len;       // sub-blocks
            continue;
        }
        if (tag != 0x2C) return false;                                 // 0x3B (trailer) before any image, or garbage
        const int x0 = u16(), y0 = u16(), w = u16(), h = u16(), lflags = u8();
        if (x0 + w > W || y0 + h > H) return false;
        const uint8_t(*pal)[3];
        if (lflags & 0x80) {
            memset(lpal, 0, sizeof lpal);
            for (int i = 0, n = 2 << (lflags & 7); i < n; i++) { lpal[i][0] = (uint8_t)u8(); lpal[i][1] = (uint8_t)u8(); lpal[i][2] = (uint8_t)u8(); }
            pal = lpal;
        } else if (flags & 0x80) pal = gpal;
        else return false;
        rgb.assign((size_t)W * H * 3, 0);
        std::vector<uint8_t> visited((size_t)W * H, 0);
        // pixel cursor over the image rectangle; interlaced files deliver rows 0, 8, 16.. then 4, 12.. then 2, 6.. then 1, 3..
        int cx = 0, cy = 0, pass = (lflags & 0x40) ? 3 : 0, step = (lflags & 0x40) ? 8 : 1;
        bool full = (w == 0 || h == 0);
        auto put = [&](int idx) {
            if (full) return;
            const size_t at = (size_t)(y0 + cy) * W + (x0 + cx);
            visited[at] = 1;
            if (idx != transparent) { rgb[3 * at] = pal[idx][0]; rgb[3 * at + 1] = pal[idx][1]; rgb[3 * at + 2] = pal[idx][2]; }
            if (++cx < w) return;
            cx = 0;
            cy += step;
            while (cy >= h && pass > 0) { step = 1 << pass; cy = step >> 1; pass--; }
            if (cy >= h) full = true;
        };
        // LZW (variable code width, LSB first, data in sub-blocks)
        const int min_bits = u8();
        if (min_bits > 12) return false;
        const int clear = 1 << min_bits;
        struct Code { int16_t prefix; uint8_t first, suffix; };
        std::vector<Code> codes(8192);
        for (int i = 0; i < clear; i++) codes[i] = {-1, (uint8_t)i, (uint8_t)i};
        std::vector<uint8_t> stack(8192 + 1);
        int width = min_bits + 1, mask = (1 << width) - 1, avail = clear + 2, old = -1, block = 0, nbits = 0;
        uint32_t acc = 0;
        bool seen_clear = false;
        for (;;) {
            if (nbits < width) {
                if (block == 0) { block = u8(); if (block == 0) break; }
                block--;
                acc |= (uint32_t)u8() << nbits;
                nbits += 8;
                if (eof) break;
                continue;
            }
            const int code = (int)(acc & (uint32_t)mask);
            acc >>= width;
            nbits -= width;
            if (code == clear) { width = min_bits + 1; mask = (1 << width) - 1; avail = clear + 2; old = -1; seen_clear = true; continue; }
            if (code == clear + 1) break;                              // end of information
            if (code > avail || !seen_clear) return false;
            if (old >= 0) {
                if (avail >= 8192) return false;
                Code& c = codes[avail++];
                c.prefix = (int16_t)old;
                c.first = codes[old].first;
                c.suffix = code == avail - 1 ? c.first : codes[code].first;
            } else if (code == avail) return false;
            int n = 0;
            for (int c = code; c >= 0 && n < 8192; c = codes[c].prefix) stack[n++] = codes[c].suffix;
            while (n > 0) put(stack[--n]);
            if ((avail & mask) == 0 && avail <= 0x0FFF) { width++; mask = (1 << width) - 1; }
            old = code;
        }
        // (the reference copies its B,G,R-ordered table entry straight into the R,G,B canvas here: background red and blue arrive swapped)
        if (bgindex > 0)
            for (size_t i = 0; i < visited.size(); i++)
                if (!visited[i]) { rgb[3 * i] = gpal[bgindex][2]; rgb[3 * i + 1] = gpal[bgindex][1]; rgb[3 * i + 2] = gpal[bgindex][0]; }
        nx = W; ny = H;
        return true;
    }
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
    if (buf.size() >= 13 && buf[0] == 'G' && buf[1] == 'I' && buf[2] == 'F') return decode_gif(buf, rgb, nx, ny);
    return false;
}

}  // namespace cb
