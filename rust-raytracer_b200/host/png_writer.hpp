// write_image (reference raytracer.rs:33-42): RGB8 PNG, zlib-compressed, no interlace.
#pragma once
#include <cstdint>
#include <string>
namespace rthost { bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height, std::string* err); }
