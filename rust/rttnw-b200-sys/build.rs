// Links librttnw_b200.so (built by `make all` at the repository root).
fn main() {
    let dir = std::env::var("RTTNW_B200_LIB_DIR").unwrap_or_else(|_| "../../rttnw_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rttnw_b200");
    println!("cargo:rerun-if-env-changed=RTTNW_B200_LIB_DIR");
    // --features reference-scenes: where the reference's src/ lives (its scenes.rs is include!d, not copied)
    if let Ok(dir) = std::env::var("RTTNW_REFERENCE_DIR") {
        println!("cargo:rustc-env=RTTNW_REFERENCE_SRC={}/src", dir);
    }
    println!("cargo:rerun-if-env-changed=RTTNW_REFERENCE_DIR");
}
