// Links the C-ABI library built by `python -c "import __graft_entry__ as g; g.build()"` (alphazero-chess_b200/libaz_b200.so).
fn main() {
    let dir = std::env::var("AZ_B200_LIB_DIR").unwrap_or_else(|_| "../../../alphazero-chess_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=az_b200");
    println!("cargo:rerun-if-env-changed=AZ_B200_LIB_DIR");
}
