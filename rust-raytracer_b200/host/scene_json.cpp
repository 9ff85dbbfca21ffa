#include "scene_json.hpp"

#include <fstream>
#include <stdexcept>

#include "json.hpp"

namespace rthost {
namespace {
rt_vec3 vec3(const Json& j) { return rt_vec3{j.at("x").number(), j.at("y").number(), j.at("z").number()}; }   // point3d.rs:10-15
void albedo(const Json& j, float out[3]) {                                                                  // SrgbAsArray, materials.rs:18-25
    if (j.kind != Json::Arr || j.arr.size() != 3) throw std::runtime_error("albedo: array of 3 numbers expected");
    for (int i = 0; i < 3; ++i) out[i] = (float)j.arr[i].number();
}
// serde parses width/height/max_depth as usize and samples_per_pixel as u32 (config.rs:68-71): negative, fractional or
// out-of-range numbers are parse errors there; casting them blindly would be undefined behaviour here.
uint64_t uint_field(const Json& j, const char* name, double max_value) {
    const double v = j.number();
    if (!(v >= 0.0) || v > max_value || v != (double)(uint64_t)v) throw std::runtime_error(std::string(name) + ": non-negative integer expected");
    return (uint64_t)v;
}
bool load_image(const std::string& path, const std::string& base_dir, Image* img) {
    std::string err;
    if (decode_jpeg_file(path, img, &err)) return true;
    if (!base_dir.empty() && decode_jpeg_file(base_dir + "/" + path, img, &err)) return true;
    throw std::runtime_error(path + ": " + err);   // the reference panics with the path (materials.rs:214)
}
}  // namespace

void load_scene_json(const std::string& text, const std::string& base_dir, SceneHolder* out) {
    Json root = JsonParser::parse(text);
    if (root.kind != Json::Obj) throw std::runtime_error("Unable to parse config json: object expected");
    rt_scene& s = out->scene;
    s.width = (uint32_t)uint_field(root.at("width"), "width", 4294967295.0); s.height = (uint32_t)uint_field(root.at("height"), "height", 4294967295.0);
    s.samples_per_pixel = (uint32_t)uint_field(root.at("samples_per_pixel"), "samples_per_pixel", 4294967295.0);
    s.max_depth = (uint32_t)uint_field(root.at("max_depth"), "max_depth", 4294967295.0);
    s.seed = 0x5EED;
    // camera: CameraParams -> Camera::new (camera.rs:29-42)
    const Json& cam = root.at("camera");
    rt_camera_params cp{vec3(cam.at("look_from")), vec3(cam.at("look_at")), vec3(cam.at("vup")), cam.at("vfov").number(), cam.at("aspect").number()};
    if (rtb200_camera_from_params(&cp, &s.camera) != 0) throw std::runtime_error(rtb200_last_error());
    // sky: missing or null -> None (black); {"texture": ""} -> gradient; path -> texture (config.rs:49-64)
    s.sky.mode = RT_SKY_NONE;
    if (const Json* sky = root.find("sky")) {
        if (sky->kind == Json::Obj) {
            const std::string& t = sky->at("texture").string();
            if (t.empty()) s.sky.mode = RT_SKY_GRADIENT;
            else {
                load_image(t, base_dir, &out->sky_image);
                s.sky.mode = RT_SKY_TEXTURE;
                s.sky.tex = rt_image{out->sky_image.rgb.data(), (uint64_t)out->sky_image.width, (uint64_t)out->sky_image.height, (uint64_t)out->sky_image.rgb.size()};
            }
        } else if (sky->kind != Json::Null) throw std::runtime_error("sky: object or null expected");
    }
    const Json& objs = root.at("objects");
    if (objs.kind != Json::Arr) throw std::runtime_error("objects: array expected");
    out->spheres.resize(objs.arr.size());
    out->images.reserve(objs.arr.size());
    std::vector<std::pair<uint64_t, uint64_t>> dims;
    for (size_t i = 0; i < objs.arr.size(); ++i) {
        const Json& o = objs.arr[i];
        rt_sphere& sp = out->spheres[i];
        sp = rt_sphere{};
        sp.center = vec3(o.at("center")); sp.radius = o.at("radius").number(); sp.texture = -1;
        const Json& m = o.at("material");
        if (m.kind != Json::Obj || m.obj.size() != 1) throw std::runtime_error("material: externally tagged enum expected (materials.rs:35-42)");
        const std::string& tag = m.obj[0].first; const Json& b = m.obj[0].second;
        if (tag == "Lambertian") { sp.kind = RT_LAMBERTIAN; albedo(b.at("albedo"), sp.albedo); }
        else if (tag == "Metal") { sp.kind = RT_METAL; albedo(b.at("albedo"), sp.albedo); sp.param = b.at("fuzz").number(); }
        else if (tag == "Glass") { sp.kind = RT_GLASS; sp.param = b.at("index_of_refraction").number(); }
        else if (tag == "Light") { sp.kind = RT_LIGHT; }
        else if (tag == "Texture") {
            sp.kind = RT_TEXTURE; albedo(b.at("albedo"), sp.albedo); sp.param = b.at("h_offset").number();
            out->images.emplace_back();
            load_image(b.at("pixels").string(), base_dir, &out->images.back());
            const uint64_t w = uint_field(b.at("width"), "texture width", 9007199254740992.0), h = uint_field(b.at("height"), "texture height", 9007199254740992.0);   // JSON dims (materials.rs:208-209)
            const uint64_t texels = out->images.back().rgb.size() / 3;
            if (w == 0 || h == 0 || w > texels || h > texels / w) throw std::runtime_error("texture: JSON width/height exceed the decoded image");   // overflow-safe w*h*3 <= size
            dims.emplace_back(w, h);
            sp.texture = (int32_t)out->images.size() - 1;
        } else throw std::runtime_error("unknown variant `" + tag + "`, expected one of `Lambertian`, `Metal`, `Glass`, `Texture`, `Light`");
    }
    out->textures.resize(out->images.size());
    for (size_t t = 0; t < out->images.size(); ++t) out->textures[t] = rt_image{out->images[t].rgb.data(), dims[t].first, dims[t].second, (uint64_t)out->images[t].rgb.size()};
    s.spheres = out->spheres.data(); s.n_spheres = out->spheres.size();
    s.textures = out->textures.data(); s.n_textures = out->textures.size();
}
}  // namespace rthost
