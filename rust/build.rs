// build.rs — only acts when the `gpu` feature is on: links libibu_b200.so (built by
// `make -C ibu_b200/csrc` of the ibu_b200 repository; sm_100a only, no CPU fallback inside).
fn main() {
    println!("cargo:rerun-if-env-changed=IBU_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_GPU").is_ok() {
        let dir = std::env::var("IBU_B200_LIB_DIR")
            .expect("the `gpu` feature needs IBU_B200_LIB_DIR = directory that holds libibu_b200.so");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=ibu_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
}
