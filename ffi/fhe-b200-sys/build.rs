// Links libfhe_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` into learn-fhe_b200/).
fn main() {
    let dir = std::env::var("FHE_B200_LIB_DIR").expect("set FHE_B200_LIB_DIR to the directory that holds libfhe_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fhe_b200");
    println!("cargo:rerun-if-env-changed=FHE_B200_LIB_DIR");
}
