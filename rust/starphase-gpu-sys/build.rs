// Link against the prebuilt libstarphase_gpu.so; STARPHASE_GPU_LIB_DIR points at pb_starphase_b200/.
fn main() {
    if let Ok(dir) = std::env::var("STARPHASE_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=starphase_gpu");
    println!("cargo:rerun-if-env-changed=STARPHASE_GPU_LIB_DIR");
}
