// raytracer <config_file> <output_file> — the reference CLI (main.rs:7-20) on top of librtb200.so.
// Same argument contract, same two stdout lines ("\nRendering <file>", "Frame time: <ms>ms"); errors that make the
// reference panic print a message to stderr and exit with status 101 (Rust's panic exit code).
// Extra knobs, so the CLI stays identical: RTB200_SEED, RTB200_DEVICE, RTB200_GPUS=<n|0=all> (row bands dealt over n GPUs of
// this process, rtb200_render_rgb8_multi), RTB200_STATS=1 (prints rays / Mrays/s to stderr).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "../../include/rtb200.h"
#include "png_writer.hpp"
#include "scene_json.hpp"

int main(int argc, char** argv) {
    if (argc != 3) {                                                       // main.rs:9-12
        printf("Usage: %s <config_file> <output_file>\n", argc > 0 ? argv[0] : "raytracer");
        return 0;
    }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { fprintf(stderr, "Unable to read config file.: %s\n", argv[1]); return 101; }              // main.rs:14
    std::stringstream ss; ss << f.rdbuf();
    rthost::SceneHolder holder;
    try {
        std::string path = argv[1];
        size_t slash = path.find_last_of('/');
        rthost::load_scene_json(ss.str(), slash == std::string::npos ? std::string(".") : path.substr(0, slash), &holder);
    } catch (const std::exception& e) { fprintf(stderr, "Unable to parse config json: %s\n", e.what()); return 101; }   // main.rs:15
    if (const char* sd = getenv("RTB200_SEED")) holder.scene.seed = strtoull(sd, nullptr, 0);
    printf("\nRendering %s\n", argv[2]);                                  // main.rs:18
    fflush(stdout);
    const rt_scene& s = holder.scene;
    std::vector<uint8_t> pixels((size_t)s.width * s.height * 3);          // raytracer.rs:254
    rt_options opts{};
    opts.device = getenv("RTB200_DEVICE") ? atoi(getenv("RTB200_DEVICE")) : -1; opts.rank = 0; opts.world = 1; opts.band_rows = 1;
    rt_stats st{};
    auto t0 = std::chrono::steady_clock::now();                           // raytracer.rs:259
    const char* gpus = getenv("RTB200_GPUS");
    int rc = gpus ? rtb200_render_rgb8_multi(&s, &opts, atoi(gpus), pixels.data(), &st)   // replaces raytracer.rs:260-262
                  : rtb200_render_rgb8(&s, &opts, pixels.data(), &st);
    if (rc != 0) { fprintf(stderr, "render failed (%d): %s\n", rc, rtb200_last_error()); return 101; }
    long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    printf("Frame time: %lldms\n", ms);                                   // raytracer.rs:263
    if (getenv("RTB200_STATS"))
        fprintf(stderr, "rays=%llu samples=%llu device_ms=%.3f Mrays/s=%.1f gpus=%d\n", (unsigned long long)st.rays, (unsigned long long)st.samples, st.device_ms,
                st.device_ms > 0 ? st.rays / st.device_ms / 1e3 : 0.0, (int)st.gpus_used);
    std::string err;
    if (!rthost::write_png_rgb8(argv[2], pixels.data(), s.width, s.height, &err)) { fprintf(stderr, "error writing image: %s\n", err.c_str()); return 101; }   // raytracer.rs:265
    return 0;
}
