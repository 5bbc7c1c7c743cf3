//! Bodies a maintainer pastes into han0110/learn-fhe's `util` crate (file:line given per function) so that the scheme
//! crates' call sites stay unchanged while the ring arithmetic runs on the B200.  Written against the reference's types:
//! `Zq { q: u64, v: u64 }` (util/src/zq.rs:21-26, read through `q()` / `Into<u64>` / `Zq::from_u64`) and `T64(u64)`
//! (util/src/torus.rs:12).  NOT compiled in the build image (no cargo / rustc there): shipped as source, kept mechanical.
//! Batched scheme-level replacements (Fhew::op over a Vec of gates, tfhe::Bootstrapping::bootstrap, Ckks::mul) are
//! listed in INTEGRATION.md §3; they call fhe_fhew_bootstrap_batch_host / fhe_tfhe_pbs_batch_host /
//! fhe_ckks_mul_relin_rescale_batch_host the same way.
use fhe_b200_sys as sys;
use std::{ffi::CStr, sync::OnceLock};

struct Ctx(*mut sys::fhe_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

/// One process-wide context replaces the reference's global twiddle `Mutex<HashMap>` (util/src/ring/fft/zq.rs:38-56).
fn ctx() -> *mut sys::fhe_ctx {
    static CTX: OnceLock<Ctx> = OnceLock::new();
    CTX.get_or_init(|| {
        let mut p = std::ptr::null_mut();
        let st = unsafe { sys::fhe_ctx_create(0, &mut p) };
        assert_eq!(st, sys::FHE_OK, "fhe_ctx_create failed: no sm_100-class GPU (there is no CPU fallback)");
        Ctx(p)
    })
    .0
}

/// The reference panics on violated preconditions (assert!/unwrap); so does the shim, with the library's message.
fn check(st: sys::fhe_status) {
    if st != sys::FHE_OK {
        let msg = unsafe { CStr::from_ptr(sys::fhe_last_error(ctx())) }.to_string_lossy().into_owned();
        panic!("fhe_b200 error {st}: {msg}");
    }
}

/// util/src/ring/fft/zq.rs:27-30 — `a` as (modulus, residues); the caller gathers `Zq::into::<u64>()` and scatters back
/// with `Zq::from_u64(q, v)` because `Zq` is a 16-byte `{q, v}` pair.
pub fn nega_cyclic_ntt_in_place(q: u64, a: &mut [u64]) {
    check(unsafe { sys::fhe_ntt_fwd_host(ctx(), q, a.as_mut_ptr(), a.len(), 1) });
}
/// util/src/ring/fft/zq.rs:32-36
pub fn nega_cyclic_intt_in_place(q: u64, a: &mut [u64]) {
    check(unsafe { sys::fhe_ntt_inv_host(ctx(), q, a.as_mut_ptr(), a.len(), 1) });
}
/// util/src/ring/fft/zq.rs:14-25 (`Rq *= &Rq`, util/src/ring.rs:256-264)
pub fn nega_cyclic_ntt_mul_assign(q: u64, a: &mut [u64], b: &[u64]) {
    assert_eq!(a.len(), b.len());
    check(unsafe { sys::fhe_negacyclic_mul_host(ctx(), q, a.as_mut_ptr(), b.as_ptr(), a.len(), 1) });
}
/// util/src/ring/fft/c64.rs:11 (`Rt *= &Rt`, util/src/ring.rs:315-320); `T64` is a transparent u64 newtype
pub fn nega_cyclic_fft64_mul_assign_rt(a: &mut [u64], b: &[u64]) {
    assert_eq!(a.len(), b.len());
    check(unsafe { sys::fhe_fft64_negacyclic_mul_host(ctx(), a.as_mut_ptr(), b.as_ptr(), a.len(), 1) });
}
