// Part of the Rust-side binding described in INTEGRATION.md (N3). Not compiled in this repository: the build image has no
// cargo/rustc. Drop into dps/rust-raytracer's `raytracer/` crate as `src/rtb200_sys.rs`.
// Mirrors include/rtb200.h (ABI 2) one to one; tests/test_rust_shim_layout.py parses THIS file and checks every field
// offset, size and alignment against the header's structs, and both extern signatures against the header's prototypes.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)] #[derive(Clone, Copy, Default)] pub struct rt_vec3 { pub x: f64, pub y: f64, pub z: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct rt_camera { pub origin: rt_vec3, pub lower_left_corner: rt_vec3, pub horizontal: rt_vec3, pub vertical: rt_vec3 }
#[repr(C)] #[derive(Clone, Copy)] pub struct rt_sphere {
    pub center: rt_vec3, pub radius: f64,
    pub kind: u32, pub albedo: [f32; 3], pub param: f64, pub texture: i32, pub reserved: i32,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct rt_image { pub rgb8: *const u8, pub width: u64, pub height: u64, pub bytes: u64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct rt_sky { pub mode: u32, pub reserved: u32, pub tex: rt_image }
#[repr(C)] pub struct rt_scene {
    pub width: u32, pub height: u32, pub samples_per_pixel: u32, pub max_depth: u32,
    pub camera: rt_camera, pub sky: rt_sky,
    pub spheres: *const rt_sphere, pub n_spheres: u64,
    pub textures: *const rt_image, pub n_textures: u64,
    pub seed: u64,
}
#[repr(C)] #[derive(Default)] pub struct rt_stats {
    pub rays: u64, pub samples: u64, pub candidates: u64,
    pub device_ms: f64, pub trace_ms: f64, pub wall_ms: f64,
    pub kernel_launches: u32, pub batches: u32, pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub clusters: u64, pub frames: u64, pub nodes: u64, pub gpus_used: i32, pub reserved: i32,
}
#[repr(C)] pub struct rt_options { pub device: i32, pub rank: i32, pub world: i32, pub band_rows: u32, pub variant: u32, pub flags: u32, pub sample_buffer_bytes: u64 }

extern "C" {
    pub fn rtb200_render_rgb8(scene: *const rt_scene, opts: *const rt_options, out_rgb8: *mut u8, stats: *mut rt_stats) -> c_int;
    pub fn rtb200_render_rgb8_multi(scene: *const rt_scene, opts: *const rt_options, n_gpus: i32, out_rgb8: *mut u8, stats: *mut rt_stats) -> c_int;
    pub fn rtb200_last_error() -> *const c_char;
}
