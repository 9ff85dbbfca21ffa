// C entry points of the host-side helpers that ship inside librtb200.so (texture decode for non-C++ hosts).
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/rtb200.h"
#include "jpeg_decode.hpp"

extern "C" {
// load_texture_image (reference materials.rs:213-219 / config.rs:36-47): decode a JPEG file to RGB8.
// *out_rgb8 is malloc'ed by the callee and released with rtb200_free(). Returns 0 on success.
int rtb200_decode_jpeg_file(const char* path, uint8_t** out_rgb8, uint64_t* width, uint64_t* height) {
    if (!path || !out_rgb8 || !width || !height) return RT_ERR_INVALID;
    *out_rgb8 = nullptr;
    try {   // nothing may unwind through the C boundary (a header can claim 65535 x 65535 pixels: std::bad_alloc)
        rthost::Image img; std::string err;
        if (!rthost::decode_jpeg_file(path, &img, &err)) return RT_ERR_INVALID;
        *out_rgb8 = (uint8_t*)malloc(img.rgb.size() ? img.rgb.size() : 1);
        if (!*out_rgb8) return RT_ERR_OOM;
        memcpy(*out_rgb8, img.rgb.data(), img.rgb.size());
        *width = (uint64_t)img.width; *height = (uint64_t)img.height;
        return RT_OK;
    } catch (const std::bad_alloc&) {
        return RT_ERR_OOM;
    } catch (...) {
        return RT_ERR_INVALID;
    }
}
void rtb200_free(void* p) { free(p); }
}
