// Links libg753.so.  G753_LIB_DIR points at the directory holding it (ginger-lib_b200/ in the CUDA repo).
fn main() {
    let dir = std::env::var("G753_LIB_DIR").expect("set G753_LIB_DIR to the directory containing libg753.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=g753");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=G753_LIB_DIR");
}
