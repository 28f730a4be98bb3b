// Link stanza for libode_b200.so (built by `make -C ode_solvers_b200/csrc`).
// ODE_B200_LIB_DIR overrides the search directory.
fn main() {
    let dir = std::env::var("ODE_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../ode_solvers_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=ode_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=ODE_B200_LIB_DIR");
}
