// links libs2c_b200.so; S2C_B200_LIB_DIR = directory that holds it (default: the in-tree build of this repository)
fn main() {
    let dir = std::env::var("S2C_B200_LIB_DIR").unwrap_or_else(|_| "../../zk_symmetric_crypto_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=s2c_b200");
    println!("cargo:rerun-if-env-changed=S2C_B200_LIB_DIR");
}
