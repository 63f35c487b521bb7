// Links libce_gpu.so.  CE_GPU_LIB_DIR points at the directory holding it (codec_eval_b200/ after
// `python -m codec_eval_b200.build`).  The library links the CUDA runtime statically; only libcuda.so.1
// (the driver) is needed at run time.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=CE_GPU_LIB_DIR");
    if let Ok(dir) = env::var("CE_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ce_gpu");
}
