// jpeg_decode.cpp — baseline (SOF0/SOF1, Huffman) JPEG decoder for texture and sky images.
//
// Host-side scene staging, not part of the hot path: it stands in for the reference's `jpeg-decoder` crate
// (materials.rs:213-219, config.rs:36-47), which is not vendored; the exact IDCT/upsampling of that crate is
// therefore a documented "parity unpinned" boundary. Written from the JPEG standard (ITU T.81):
// 8-bit precision, 1 or 3 components, sampling factors 1..2, restart intervals, Adobe APP14 colour transform flag.
// Output: RGB8, row-major, 3 bytes per pixel.
#include "jpeg_decode.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace rthost {
namespace {

struct Huff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0, bw = 0, bh = 0; std::vector<uint8_t> plane; };

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int nbits = 0; bool hit_marker = false;
    void fill() {
        while (nbits <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    int b2 = p < end ? *p : 0xD9;
                    if (b2 == 0) ++p;                       // stuffed zero
                    else { hit_marker = true; --p; b = 0; } // a real marker: feed zeros, leave p at 0xFF
                }
            }
            acc |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int get(int n) {
        if (n == 0) return 0;
        if (nbits < n) fill();
        int v = (int)(acc >> (32 - n));
        acc <<= n; nbits -= n;
        return v;
    }
    void reset() { acc = 0; nbits = 0; hit_marker = false; }
};

int decode_symbol(BitReader& br, const Huff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | br.get(1);
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

const int kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// separable double-precision inverse DCT (T.81 A.3.3), rounded and level-shifted
void idct8x8(const int* coef, const uint16_t* q, uint8_t* out, int stride) {
    static double c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; ++x)
            for (int u = 0; u < 8; ++u) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0);
        init = true;
    }
    double tmp[64], blk[64];
    for (int i = 0; i < 64; ++i) blk[i] = (double)coef[i] * q[i];
    for (int y = 0; y < 8; ++y)          // rows
        for (int x = 0; x < 8; ++x) {
            double s = 0;
            for (int u = 0; u < 8; ++u) s += c[x][u] * blk[y * 8 + u];
            tmp[y * 8 + x] = s;
        }
    for (int x = 0; x < 8; ++x)          // columns
        for (int y = 0; y < 8; ++y) {
            double s = 0;
            for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
            int val = (int)std::lround(s + 128.0);
            out[y * stride + x] = (uint8_t)(val < 0 ? 0 : val > 255 ? 255 : val);
        }
}

}  // namespace

bool decode_jpeg(const uint8_t* data, size_t size, Image* out, std::string* err) {
    auto fail = [&](const char* m) { if (err) *err = m; return false; };
    if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG (no SOI)");
    uint16_t qt[4][64] = {{0}};
    Huff hdc[4], hac[4];
    Comp comps[3];
    int ncomp = 0, width = 0, height = 0, restart = 0, adobe_transform = -1;
    size_t i = 2;
    bool got_sof = false;
    while (i + 4 <= size) {
        if (data[i] != 0xFF) return fail("marker expected");
        int m = data[i + 1];
        if (m == 0xFF) { ++i; continue; }
        if (m == 0xD9) break;
        size_t L = ((size_t)data[i + 2] << 8) | data[i + 3];
        if (L < 2) return fail("segment length below 2");
        if (i + 2 + L > size) return fail("truncated segment");
        const uint8_t* s = data + i + 4;
        size_t n = L - 2;                                  // payload bytes: every read below is checked against it
        if (m == 0xDB) {                                   // DQT
            size_t k = 0;
            while (k < n) {
                int pq = s[k] >> 4, tq = s[k] & 15; ++k;
                if (tq > 3 || pq > 1) return fail("bad DQT id / precision");
                if (k + (size_t)64 * (pq ? 2 : 1) > n) return fail("truncated DQT");
                for (int j = 0; j < 64; ++j) {
                    uint16_t v = pq ? (uint16_t)((s[k] << 8) | s[k + 1]) : s[k];
                    k += pq ? 2 : 1;
                    qt[tq][kZigzag[j]] = v;
                }
            }
        } else if (m == 0xC4) {                            // DHT
            size_t k = 0;
            while (k < n) {
                int tc = s[k] >> 4, th = s[k] & 15; ++k;
                if (th > 3 || tc > 1) return fail("bad DHT id");
                Huff& h = tc ? hac[th] : hdc[th];
                int total = 0;
                if (k + 16 > n) return fail("truncated DHT");
                for (int l = 1; l <= 16; ++l) { h.bits[l] = s[k++]; total += h.bits[l]; }
                if (total > 256) return fail("bad DHT");
                if (k + (size_t)total > n) return fail("truncated DHT");
                for (int j = 0; j < total; ++j) h.vals[j] = s[k++];
                h.build();
            }
        } else if (m == 0xC0 || m == 0xC1) {               // SOF0 / SOF1
            if (n < 6) return fail("truncated SOF");
            if (s[0] != 8) return fail("only 8-bit JPEG is supported");
            height = (s[1] << 8) | s[2]; width = (s[3] << 8) | s[4]; ncomp = s[5];
            if (width == 0 || height == 0) return fail("zero image dimension");
            if ((size_t)width * (size_t)height > ((size_t)1 << 28)) return fail("image larger than 2^28 pixels");   // textures of this path are a few megapixels
            if (ncomp != 1 && ncomp != 3) return fail("only 1 or 3 components are supported");
            if (n < (size_t)(6 + 3 * ncomp)) return fail("truncated SOF");
            for (int c = 0; c < ncomp; ++c) {
                comps[c].id = s[6 + 3 * c]; comps[c].h = s[7 + 3 * c] >> 4; comps[c].v = s[7 + 3 * c] & 15; comps[c].tq = s[8 + 3 * c];
                if (comps[c].tq > 3) return fail("bad quantisation table id");
                comps[c].td = 0; comps[c].ta = 0;
            }
            got_sof = true;
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return fail("progressive / arithmetic / lossless JPEG is not supported");
        } else if (m == 0xDD) {
            if (n < 2) return fail("truncated DRI");
            restart = (s[0] << 8) | s[1];
        } else if (m == 0xEE && n >= 12 && std::memcmp(s, "Adobe", 5) == 0) {
            adobe_transform = s[11];
        } else if (m == 0xDA) {                            // SOS: the entropy-coded data follows
            if (!got_sof) return fail("SOS before SOF");
            if (n < 1) return fail("truncated SOS");
            int ns = s[0];
            if (ns != ncomp) return fail("non-interleaved scans are not supported");
            if (n < (size_t)(1 + 2 * ns)) return fail("truncated SOS");
            for (int k = 0; k < ns; ++k) {
                int cid = s[1 + 2 * k];
                const int td = s[2 + 2 * k] >> 4, ta = s[2 + 2 * k] & 15;
                if (td > 3 || ta > 3) return fail("bad Huffman table id");
                for (int c = 0; c < ncomp; ++c)
                    if (comps[c].id == cid) { comps[c].td = td; comps[c].ta = ta; }
            }
            i += 2 + L;
            int hmax = 1, vmax = 1;
            for (int c = 0; c < ncomp; ++c) { if (comps[c].h > hmax) hmax = comps[c].h; if (comps[c].v > vmax) vmax = comps[c].v; }
            int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (int c = 0; c < ncomp; ++c) {
                if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2) return fail("sampling factors above 2 are not supported");
                comps[c].bw = mcux * comps[c].h * 8; comps[c].bh = mcuy * comps[c].v * 8;
                comps[c].plane.assign((size_t)comps[c].bw * comps[c].bh, 0);
                comps[c].pred = 0;
                if (!hdc[comps[c].td].present || !hac[comps[c].ta].present) return fail("missing Huffman table");
            }
            BitReader br{data + i, data + size};
            int coef[64];
            int mcus_left = restart;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart && mcus_left == 0) {       // RSTn: byte-align, skip the marker, reset predictors
                        const uint8_t* q = br.p;
                        while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                        if (q + 1 >= br.end) return fail("restart marker missing");
                        br.p = q + 2; br.reset();
                        for (int c = 0; c < ncomp; ++c) comps[c].pred = 0;
                        mcus_left = restart;
                    }
                    for (int c = 0; c < ncomp; ++c) {
                        Comp& cp = comps[c];
                        for (int by = 0; by < cp.v; ++by)
                            for (int bx = 0; bx < cp.h; ++bx) {
                                std::memset(coef, 0, sizeof coef);
                                int t = decode_symbol(br, hdc[cp.td]);
                                if (t < 0 || t > 11) return fail("bad DC code");
                                int diff = t ? extend(br.get(t), t) : 0;
                                cp.pred += diff;
                                coef[0] = cp.pred;
                                for (int k = 1; k < 64;) {
                                    int rs = decode_symbol(br, hac[cp.ta]);
                                    if (rs < 0) return fail("bad AC code");
                                    int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                    k += r;
                                    if (k > 63) return fail("AC run past block end");
                                    coef[kZigzag[k]] = extend(br.get(sz), sz);
                                    ++k;
                                }
                                uint8_t* dst = cp.plane.data() + (size_t)((my * cp.v + by) * 8) * cp.bw + (size_t)(mx * cp.h + bx) * 8;
                                idct8x8(coef, qt[cp.tq], dst, cp.bw);
                            }
                    }
                    if (restart) --mcus_left;
                }
            // upsample (triangle filter for 2x, like libjpeg's "fancy" upsampling) + colour conversion
            out->width = width; out->height = height;
            out->rgb.assign((size_t)width * height * 3, 0);
            auto sample = [&](const Comp& cp, int x, int y) -> double {
                if (cp.h == hmax && cp.v == vmax) return cp.plane[(size_t)y * cp.bw + x];
                double fx = cp.h == hmax ? x : (x + 0.5) * cp.h / hmax - 0.5, fy = cp.v == vmax ? y : (y + 0.5) * cp.v / vmax - 0.5;
                int cw = (width * cp.h + hmax - 1) / hmax, ch = (height * cp.v + vmax - 1) / vmax;
                if (fx < 0) fx = 0; if (fy < 0) fy = 0;
                if (fx > cw - 1) fx = cw - 1; if (fy > ch - 1) fy = ch - 1;
                int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1 < cw ? x0 + 1 : x0, y1 = y0 + 1 < ch ? y0 + 1 : y0;
                double ax = fx - x0, ay = fy - y0;
                const uint8_t* P = cp.plane.data();
                return (1 - ay) * ((1 - ax) * P[(size_t)y0 * cp.bw + x0] + ax * P[(size_t)y0 * cp.bw + x1]) +
                       ay * ((1 - ax) * P[(size_t)y1 * cp.bw + x0] + ax * P[(size_t)y1 * cp.bw + x1]);
            };
            auto clamp8 = [](double v) { int r = (int)std::lround(v); return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r); };
            const bool ycc = ncomp == 3 && adobe_transform != 0;
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) {
                    uint8_t* px = out->rgb.data() + ((size_t)y * width + x) * 3;
                    if (ncomp == 1) { uint8_t g = clamp8(sample(comps[0], x, y)); px[0] = px[1] = px[2] = g; continue; }
                    double Y = sample(comps[0], x, y), Cb = sample(comps[1], x, y), Cr = sample(comps[2], x, y);
                    if (ycc) {
                        px[0] = clamp8(Y + 1.402 * (Cr - 128.0));
                        px[1] = clamp8(Y - 0.344136 * (Cb - 128.0) - 0.714136 * (Cr - 128.0));
                        px[2] = clamp8(Y + 1.772 * (Cb - 128.0));
                    } else { px[0] = clamp8(Y); px[1] = clamp8(Cb); px[2] = clamp8(Cr); }
                }
            return true;
        }
        i += 2 + L;
    }
    return fail("no scan found");
}

bool decode_jpeg_file(const std::string& path, Image* out, std::string* err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return decode_jpeg(buf.data(), buf.size(), out, err);
}

}  // namespace rthost
