// Baseline JPEG decoder used by the host-side scene staging (stands in for the reference's jpeg-decoder crate).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
namespace rthost {
struct Image { int width = 0, height = 0; std::vector<uint8_t> rgb; };
bool decode_jpeg(const uint8_t* data, size_t size, Image* out, std::string* err);
bool decode_jpeg_file(const std::string& path, Image* out, std::string* err);
}  // namespace rthost
