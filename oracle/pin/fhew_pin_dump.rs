//! TEST INFRASTRUCTURE of learn-fhe_b200 (not part of han0110/learn-fhe): known-answer dump of the FHEW crate.
//!
//! Installed by oracle/pin/apply.sh as `scheme/fhew/src/bootstrapping/pin_dump.rs` with the line
//! `#[cfg(test)] mod pin_dump;` appended to `scheme/fhew/src/bootstrapping.rs` (a child module sees the private key
//! accessors of `BootstrappingKey`).  Run
//!     FHE_PIN_OUT=/path/to/learn-fhe_b200/tests/golden/ref cargo test --release -p fhew pin_dump -- --nocapture
//! Keys, ciphertexts and noise come from `StdRng::seed_from_u64`; every output is computed by the reference's own
//! `Lwe::key_switch`, `Rgsw::external_product`, `Rlwe::automorphism`, `Bootstrapping::bootstrap`.  The file follows
//! the `fhew_tiny` schema of learn-fhe_b200/tests/golden/util_fhew.json (plus `steps` for single external products /
//! automorphisms) so that tests/test_cpu_refpin.py can replay it.
use super::*;
use crate::{
    lwe::{Lwe, LweCiphertext, LweParam},
    rgsw::{Rgsw, RgswParam},
    rlwe::{Rlwe, RlweCiphertext, RlweParam},
};
use rand::{rngs::StdRng, SeedableRng};
use std::{env, fs, path::PathBuf};
use util::{two_adic_primes, Rq, Zq};

fn arr<T: ToString>(v: impl IntoIterator<Item = T>) -> String {
    format!("[{}]", v.into_iter().map(|x| x.to_string()).collect::<Vec<_>>().join(","))
}
fn poly(p: &Rq) -> String {
    arr(p.iter().map(|z| z.to_u64()))
}
fn lwe(ct: &LweCiphertext) -> String {
    arr(ct.a().iter().chain([ct.b()]).map(|z| z.to_u64()))
}

fn dump(name: &str, log_q: usize, log_n: usize, log_b: usize, d: usize, n_s: usize, log_q_ks: usize, ks: (usize, usize), w: usize, seed: u64) -> String {
    let mut rng = StdRng::seed_from_u64(seed);
    let p = 4;
    let big_q = two_adic_primes(log_q, log_n + 1).next().unwrap();
    let rlwe = RlweParam::new(big_q, p, log_n).with_decomposor(log_b, d);
    let rgsw = RgswParam::new(rlwe, log_b, d);
    let lwe_s = LweParam::new(1 << log_q_ks, p, n_s).with_decomposor(ks.0, ks.1);
    let param = BootstrappingParam::new(rgsw, lwe_s, w);
    let z = Rlwe::sk_gen(param.rlwe(), &mut rng);
    let bk = Bootstrapping::key_gen(&param, &z, &mut rng);
    let n = param.n();

    // keys in the layout of fhe_fhew_key_upload: ksk_a [N ks_d][n_s], ksk_b [N ks_d], brk [n_s][2d][2 (a, b)][N], ak [w+1][d][2][N]
    let ksk_a = arr(bk.ksk().a().map(|r| arr(r.iter().map(|z| z.to_u64()))));
    let ksk_b = arr(bk.ksk().b().map(|z| z.to_u64()));
    let brk = arr(bk.brk().iter().map(|g| arr(g.a().zip(g.b()).map(|(a, b)| format!("[{},{}]", poly(a), poly(b))))));
    let ak = arr(bk.ak().iter().map(|k| arr(k.a().zip(k.b()).map(|(a, b)| format!("[{},{}]", poly(a), poly(b))))));
    let ak_t = arr(bk.ak().iter().map(|k| k.t()));

    // NAND gate polynomial and epilogue constant exactly as Fhew::op builds them (fhew.rs:31-39)
    let table = [1usize, 1, 1, 0];
    let map = [-bk.big_q_by_8(), bk.big_q_by_8()];
    let f: Rq = table
        .into_iter()
        .flat_map(|out| core::iter::repeat(map[out]).take(bk.q_by_8()))
        .collect::<Vec<_>>()
        .into();

    let mut cases = Vec::new();
    for (m0, m1) in [(0u64, 0u64), (0, 1), (1, 0), (1, 1)] {
        let enc = |m: u64, rng: &mut StdRng| {
            let pt = Lwe::encode(param.lwe_z(), Zq::from_u64(p, m));
            Lwe::sk_encrypt(param.lwe_z(), &z, pt, rng)
        };
        let ct = enc(m0, &mut rng) + enc(m1, &mut rng);
        let pro = {
            let c = ct.mod_switch(bk.big_q_ks());
            let c = Lwe::key_switch(bk.lwe_s(), bk.ksk(), c);
            c.mod_switch_odd(bk.q())
        };
        let LweCiphertext(a, b) = Bootstrapping::bootstrap(&bk, &f, ct.clone());
        let out = LweCiphertext(a, b + bk.big_q_by_8());
        let bit = Lwe::decode(param.lwe_z(), Lwe::decrypt(param.lwe_z(), &z, out.clone())).to_u64();
        assert_eq!(bit, 1 - (m0 & m1));
        cases.push(format!("{{\"ct\":{},\"prologue\":{},\"out\":{},\"bit\":{}}}", lwe(&ct), lwe(&pro), lwe(&out), bit));
    }

    // single steps: Rgsw::external_product(brk[idx], acc) and Rlwe::automorphism(ak[idx], acc) on uniform accumulators
    let mut steps = Vec::new();
    for idx in [0usize, n_s / 2, n_s - 1] {
        let acc = RlweCiphertext(Rq::sample_uniform(big_q, n, &mut rng), Rq::sample_uniform(big_q, n, &mut rng));
        let o = Rgsw::external_product(param.rgsw(), &bk.brk()[idx], &acc);
        steps.push(format!("{{\"kind\":\"external_product\",\"idx\":{},\"acc\":[{},{}],\"out\":[{},{}]}}", idx, poly(acc.a()), poly(acc.b()), poly(o.a()), poly(o.b())));
    }
    for idx in [0usize, 1, w] {
        let acc = RlweCiphertext(Rq::sample_uniform(big_q, n, &mut rng), Rq::sample_uniform(big_q, n, &mut rng));
        let o = Rlwe::automorphism(param.rlwe(), &bk.ak()[idx], acc.clone());
        steps.push(format!("{{\"kind\":\"automorphism\",\"idx\":{},\"acc\":[{},{}],\"out\":[{},{}]}}", idx, poly(acc.a()), poly(acc.b()), poly(o.a()), poly(o.b())));
    }

    format!(
        "\"{}\":{{\"param\":{{\"n\":{},\"log_n\":{},\"big_q\":{},\"p\":{},\"rlwe_log_b\":{},\"rlwe_d\":{},\"rgsw_log_b\":{},\"rgsw_d\":{},\"n_s\":{},\"q_ks\":{},\"ks_log_b\":{},\"ks_d\":{},\"w\":{}}},\
         \"keys\":{{\"ksk_a\":{},\"ksk_b\":{},\"brk\":{},\"ak\":{},\"ak_t\":{}}},\"table\":[1,1,1,0],\"f\":{},\"post_add\":{},\"cases\":[{}],\"steps\":[{}]}}",
        name, n, log_n, big_q, p, log_b, d, log_b, d, n_s, 1u64 << log_q_ks, ks.0, ks.1, w,
        ksk_a, ksk_b, brk, ak, ak_t, poly(&f), bk.big_q_by_8().to_u64(), cases.join(","), steps.join(",")
    )
}

#[test]
fn pin_dump_fhew() {
    // a tiny set (N = 16, 20-bit Q: the shape of tests/golden/util_fhew.json) and one at N = 64, 28-bit Q, decomposor (7, 4),
    // which takes the same kernels as the reference's own single-key test parameters (fhew/boolean.rs:225-239)
    let sections = [
        dump("fhew_tiny", 20, 4, 5, 4, 6, 10, (2, 5), 3, 0x5EED_0201),
        dump("fhew_n64", 28, 6, 7, 4, 12, 16, (4, 4), 10, 0x5EED_0202),
    ];
    let dir = PathBuf::from(env::var("FHE_PIN_OUT").unwrap_or_else(|_| ".".into()));
    fs::create_dir_all(&dir).unwrap();
    fs::write(dir.join("ref_fhew.json"), format!("{{{}}}\n", sections.join(",\n"))).unwrap();
    println!("wrote {}", dir.join("ref_fhew.json").display());
}
