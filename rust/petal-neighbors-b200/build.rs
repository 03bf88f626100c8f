// Links libpetal_b200.so (built by `make -C petal-neighbors_b200/csrc`).
// PETAL_B200_LIB_DIR overrides the default in-tree location.
fn main() {
    let dir = std::env::var("PETAL_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../../petal-neighbors_b200/lib")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=petal_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PETAL_B200_LIB_DIR");
}
