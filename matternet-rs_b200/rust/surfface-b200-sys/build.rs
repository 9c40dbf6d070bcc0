// Compiles the CUDA sources of the graph-wiring path for sm_100a (no multi-arch fat binary, no CPU fallback).
// SURFFACE_B200_CSRC points at matternet-rs_b200/csrc of the B200 repository (default: ../../csrc).
fn main() {
    let csrc = std::env::var("SURFFACE_B200_CSRC").unwrap_or_else(|_| "../../csrc".to_string());
    let include = std::env::var("SURFFACE_B200_INCLUDE").unwrap_or_else(|_| "../../../include".to_string());
    let files = ["api.cu", "knn.cu", "knn_exact.cu", "knn_screen.cu", "laplacian.cu", "lambda.cu", "pipeline.cu", "comm.cu", "bc.cu", "project.cu"];
    let mut b = cc::Build::new();
    b.cuda(true)
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-std=c++17").flag("--fmad=false").flag("-lineinfo")
        .include(&include);
    for f in files {
        b.file(format!("{csrc}/{f}"));
        println!("cargo:rerun-if-changed={csrc}/{f}");
    }
    b.compile("surfface_b200");
    println!("cargo:rustc-link-lib=static=cudart_static");   // as csrc/Makefile links it
    println!("cargo:rustc-link-lib=rt");
    println!("cargo:rustc-link-lib=pthread");
    println!("cargo:rustc-link-lib=cuda");
    println!("cargo:rustc-link-lib=dl");
}
