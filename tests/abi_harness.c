/* abi_harness.c — the call sequence of integration/rust/render_replacement.rs in plain C11: proves that a host with no
 * Python, no torch and no C++ can drive the drop-in boundary (include/rtb200.h) exactly as the Rust shim would.
 *   gcc -std=c11 -O1 -I include tests/abi_harness.c -L rust-raytracer_b200 -lrtb200 -Wl,-rpath,... -o abi_harness
 *   abi_harness <scene.bin> <out.rgb> [n_gpus]
 * scene.bin (written by the test): u32 width,height,spp,max_depth; rt_camera (96 B); u32 sky_mode, u32 n_spheres;
 * n_spheres x rt_sphere (64 B each, the header's layout). Textures are not used by this harness (cover scene).
 * The frame it writes is compared byte for byte with the Python host's frame in tests/test_gpu_parity.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtb200.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <scene.bin> <out.rgb> [n_gpus]\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("scene.bin"); return 2; }
    uint32_t hdr[4], tail[2];
    rt_scene scene;
    memset(&scene, 0, sizeof scene);
    if (fread(hdr, 4, 4, f) != 4 || fread(&scene.camera, sizeof scene.camera, 1, f) != 1 || fread(tail, 4, 2, f) != 2) { fprintf(stderr, "short scene file\n"); return 2; }
    scene.width = hdr[0]; scene.height = hdr[1]; scene.samples_per_pixel = hdr[2]; scene.max_depth = hdr[3];
    scene.sky.mode = tail[0];
    rt_sphere* spheres = (rt_sphere*)calloc(tail[1] ? tail[1] : 1, sizeof(rt_sphere));
    if (fread(spheres, sizeof(rt_sphere), tail[1], f) != tail[1]) { fprintf(stderr, "short sphere list\n"); return 2; }
    fclose(f);
    scene.spheres = spheres; scene.n_spheres = tail[1];
    scene.textures = NULL; scene.n_textures = 0;
    scene.seed = 0x5EED;                                               /* like render_replacement.rs */
    uint8_t* pixels = (uint8_t*)malloc((size_t)scene.width * scene.height * 3);    /* raytracer.rs:254 */
    rt_stats stats;
    int n_gpus = argc > 3 ? atoi(argv[3]) : 0;
    int rc = rtb200_render_rgb8_multi(&scene, NULL, n_gpus, pixels, &stats);       /* replaces raytracer.rs:260-262 */
    if (rc != 0) { fprintf(stderr, "rtb200: %s\n", rtb200_last_error()); return 101; }   /* the Rust shim panics here */
    printf("Frame time: %dms\n", (int)stats.wall_ms);                  /* raytracer.rs:263 */
    printf("rays=%llu gpus=%d\n", (unsigned long long)stats.rays, (int)stats.gpus_used);
    f = fopen(argv[2], "wb");
    if (!f || fwrite(pixels, 1, (size_t)scene.width * scene.height * 3, f) != (size_t)scene.width * scene.height * 3) { perror("out.rgb"); return 2; }
    fclose(f);
    free(pixels); free(spheres);
    return 0;
}
