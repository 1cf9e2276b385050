// Links libtotsu_b200.so.  TOTSU_B200_LIB_DIR points at the directory holding it (default: the in-tree build output
// totsu_b200/ two levels up, produced by `make -C totsu_b200/csrc`).
fn main() {
    let dir = std::env::var("TOTSU_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../totsu_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=totsu_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=TOTSU_B200_LIB_DIR");
}
