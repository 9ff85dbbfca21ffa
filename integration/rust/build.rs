// Part of the Rust-side binding described in INTEGRATION.md (N3). Not compiled in this repository: the build image has no
// cargo/rustc. Drop into dps/rust-raytracer's `raytracer/` crate as the file name says.
fn main() {
    println!("cargo:rustc-link-search=native={}", std::env::var("RTB200_LIB_DIR").unwrap());
    println!("cargo:rustc-link-lib=dylib=rtb200");
}
