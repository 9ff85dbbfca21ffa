#include "png_writer.hpp"

#include <zlib.h>

#include <cstdio>
#include <vector>

namespace rthost {
namespace {
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
    put32(out, (uint32_t)data.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}
}  // namespace

bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h, std::string* err) {
    std::vector<uint8_t> raw((size_t)h * (w * 3 + 1));
    for (uint32_t y = 0; y < h; ++y) {
        raw[(size_t)y * (w * 3 + 1)] = 0;   // filter type None
        std::copy(rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3, raw.begin() + (size_t)y * (w * 3 + 1) + 1);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { if (err) *err = "zlib compress failed"; return false; }
    comp.resize(clen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w); put32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit, colour type 2 (RGB)
    chunk(out, "IHDR", ihdr); chunk(out, "IDAT", comp); chunk(out, "IEND", {});
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { if (err) *err = "cannot create " + path; return false; }
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = fclose(f) == 0 && ok;
    if (!ok && err) *err = "error writing " + path;
    return ok;
}
}  // namespace rthost
