// Part of the Rust-side binding described in INTEGRATION.md (N3). Not compiled in this repository: the build image has no
// cargo/rustc. Replaces `pub fn render` in dps/rust-raytracer's `raytracer/src/raytracer.rs` (lines 250-266).
// Needs `pub(crate)` on materials::Texture::{width, height, h_offset} (private today, materials.rs:208-210).
// tests/abi_harness.c performs exactly this call sequence from plain C and is diffed against the Python host's frame.
use crate::rtb200_sys::*;

fn image(pixels: &[u8], width: u64, height: u64) -> rt_image {
    // the library reads width*height*3 bytes; `bytes` lets it reject a JSON that claims more than the decoded file holds
    rt_image { rgb8: pixels.as_ptr(), width, height, bytes: pixels.len() as u64 }
}

fn flatten(scene: &Config) -> (Vec<rt_sphere>, Vec<rt_image>) {
    let mut textures = Vec::new();
    let spheres = scene.objects.iter().map(|s| {
        let v = |p: &Point3D| rt_vec3 { x: p.x(), y: p.y(), z: p.z() };
        let (kind, albedo, param, texture) = match &s.material {
            Material::Lambertian(l) => (0, [l.albedo.red, l.albedo.green, l.albedo.blue], 0.0, -1),
            Material::Metal(m)      => (1, [m.albedo.red, m.albedo.green, m.albedo.blue], m.fuzz, -1),
            Material::Glass(g)      => (2, [1.0; 3], g.index_of_refraction, -1),
            Material::Texture(t)    => { textures.push(image(&t.pixels, t.width, t.height));
                                         (3, [t.albedo.red, t.albedo.green, t.albedo.blue], t.h_offset, textures.len() as i32 - 1) }
            Material::Light(_)      => (4, [1.0; 3], 0.0, -1),
        };
        rt_sphere { center: v(&s.center), radius: s.radius, kind, albedo, param, texture, reserved: 0 }
    }).collect();
    (spheres, textures)
}

pub fn render(filename: &str, scene: Config) {
    let (w, h) = (scene.width, scene.height);
    let mut pixels = vec![0u8; w * h * 3];                                   // raytracer.rs:254
    let (spheres, textures) = flatten(&scene);
    let v = |p: &Point3D| rt_vec3 { x: p.x(), y: p.y(), z: p.z() };
    let none = rt_image { rgb8: std::ptr::null(), width: 0, height: 0, bytes: 0 };
    let sky = match &scene.sky {                                             // raytracer.rs:137-161
        None => rt_sky { mode: 0, reserved: 0, tex: none },
        Some(s) => match &s.texture {
            None => rt_sky { mode: 1, reserved: 0, tex: none },
            Some((px, tw, th, _)) => rt_sky { mode: 2, reserved: 0, tex: image(px, *tw as u64, *th as u64) },
        },
    };
    let c = &scene.camera;
    let rs = rt_scene {
        width: w as u32, height: h as u32, samples_per_pixel: scene.samples_per_pixel, max_depth: scene.max_depth as u32,
        camera: rt_camera { origin: v(&c.origin), lower_left_corner: v(&c.lower_left_corner), horizontal: v(&c.horizontal), vertical: v(&c.vertical) },
        sky, spheres: spheres.as_ptr(), n_spheres: spheres.len() as u64,
        textures: textures.as_ptr(), n_textures: textures.len() as u64,
        seed: 0x5EED,                                                        // the reference is unseeded (thread_rng)
    };
    let mut stats = rt_stats::default();
    let start = Instant::now();
    // replaces the rayon loop, raytracer.rs:260-262: every GPU of the box (0 = all), row bands dealt round-robin like rayon's rows
    let rc = unsafe { rtb200_render_rgb8_multi(&rs, std::ptr::null(), 0, pixels.as_mut_ptr(), &mut stats) };
    if rc != 0 { panic!("rtb200: {}", unsafe { std::ffi::CStr::from_ptr(rtb200_last_error()) }.to_string_lossy()); }
    println!("Frame time: {}ms", start.elapsed().as_millis());              // raytracer.rs:263
    write_image(filename, &pixels, (w, h)).expect("error writing image");    // raytracer.rs:265
}
