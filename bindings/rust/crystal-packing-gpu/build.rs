// Points the linker at lib/libcpgpu.so (built by `make -C crystal-packing_b200`).
fn main() {
    let dir = std::env::var("CPGPU_LIB_DIR").unwrap_or_else(|_| "../../../crystal-packing_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=cpgpu");
    println!("cargo:rerun-if-env-changed=CPGPU_LIB_DIR");
}
