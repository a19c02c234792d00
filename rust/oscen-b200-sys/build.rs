// Links the in-tree shared library built by `python -m oscen_b200.build` (oscen_b200/lib/liboscen_b200.so).
fn main() {
    let dir = std::env::var("OSCEN_B200_LIB_DIR").unwrap_or_else(|_| "../../oscen_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=oscen_b200");
    println!("cargo:rerun-if-env-changed=OSCEN_B200_LIB_DIR");
}
