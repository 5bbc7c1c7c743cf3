//! TEST INFRASTRUCTURE of learn-fhe_b200 (not part of han0110/learn-fhe): known-answer dump of the TFHE crate.
//!
//! Installed by oracle/pin/apply.sh as `scheme/tfhe/src/bootstrapping/pin_dump.rs` with `#[cfg(test)] mod pin_dump;`
//! appended to `scheme/tfhe/src/bootstrapping.rs`.  Run
//!     FHE_PIN_OUT=/path/to/learn-fhe_b200/tests/golden/ref cargo test --release -p tfhe pin_dump -- --nocapture
//! Everything random comes from `StdRng::seed_from_u64`; outputs are the reference's own `Tggsw::external_product`,
//! `Tggsw::cmux`, `Tlwe::key_switch` and `Bootstrapping::bootstrap` (f64 FFT products of util/src/ring/fft/c64.rs).
//! Schema: the `tggsw`, `tlwe_key_switch` and `tfhe_pbs` sections of learn-fhe_b200/tests/golden/tfhe_ckks.json
//! (lists of such objects here); tests/test_cpu_refpin.py replays them bit for bit against the oracle.
use super::*;
use crate::{
    tggsw::{Tggsw, TggswParam},
    tglwe::TglweCiphertext,
    tlwe::{Tlwe, TlweCiphertext, TlweParam},
};
use rand::{rngs::StdRng, SeedableRng};
use std::{env, fs, path::PathBuf};
use util::{AVec, Rq, Rt, Zq, T64};

fn arr<T: ToString>(v: impl IntoIterator<Item = T>) -> String {
    format!("[{}]", v.into_iter().map(|x| x.to_string()).collect::<Vec<_>>().join(","))
}
fn words<'a>(v: impl IntoIterator<Item = &'a T64>) -> String {
    arr(v.into_iter().map(|t| t.to_u64()))
}
fn glwe(ct: &TglweCiphertext) -> String {
    arr(ct.a().iter().chain([ct.b()]).map(|p| words(p.iter())))
}
fn tlwe(ct: &TlweCiphertext) -> String {
    words(ct.a().iter().chain([ct.b()]))
}
fn uniform_glwe(k: usize, big_n: usize, rng: &mut StdRng) -> TglweCiphertext {
    TglweCiphertext((0..k).map(|_| Rt::sample_uniform(big_n, rng)).collect(), Rt::sample_uniform(big_n, rng))
}

#[allow(clippy::too_many_arguments)]
fn dump(n: usize, big_n: usize, k: usize, bs: (usize, usize), ks: (usize, usize), sd: (f64, f64), seed: u64, pbs: &mut Vec<String>, ext: &mut Vec<String>, ksw: &mut Vec<String>) {
    let mut rng = StdRng::seed_from_u64(seed);
    let (log_p, padding) = (4usize, 1usize);
    let param = {
        let tlwe = TlweParam::new(log_p, padding, n, sd.0).with_decomposor(ks.0, ks.1);
        let tggsw = TggswParam::new(log_p, padding, big_n, k, sd.1, bs.0, bs.1);
        BootstrappingParam::new(tlwe, tggsw)
    };
    let z = Tlwe::sk_gen(&param, &mut rng);
    let bk = Bootstrapping::key_gen(&param, &z, &mut rng);
    // brk [n][(k+1) d rows][(k+1) components a_0.., b][N]; ksk_a [(k N) d_ks][n], ksk_b [(k N) d_ks]
    let row = |ct: &crate::tggsw::TggswCiphertext| arr(ct.a().zip(ct.b()).map(|(a, b)| arr(a.iter().chain([b]).map(|p| words(p.iter())))));
    let brk = arr(bk.brk().iter().map(row));
    let ksk_a = arr(bk.ksk().a().map(|r| words(r.iter())));
    let ksk_b = words(bk.ksk().b());

    // the reference's test look-up table layout (bootstrapping.rs:118-128) for f(m) = 3 m + 1 mod p
    let p = 1u64 << log_p;
    let m_per = big_n >> log_p;
    let table: Vec<Zq> = (0..p).map(|v| Zq::from_u64(p, (3 * v + 1) % p)).collect();
    let v: Rq = core::iter::repeat(table[0])
        .take(m_per / 2)
        .chain(table[1..].iter().flat_map(|t| core::iter::repeat(*t).take(m_per)))
        .chain(core::iter::repeat(-table[0]).take(m_per / 2))
        .collect();
    let (mut cts, mut outs) = (Vec::new(), Vec::new());
    for m in [0u64, 1, 7, 15] {
        let ct = Tlwe::sk_encrypt(&param, &z, Tlwe::encode(&param, Zq::from_u64(p, m)), &mut rng);
        let out = Bootstrapping::bootstrap(&bk, &v, ct.clone());
        assert_eq!(Tlwe::decode(&param, Tlwe::decrypt(&param, &z, out.clone())).to_u64(), (3 * m + 1) % p);
        cts.push(tlwe(&ct));
        outs.push(tlwe(&out));
    }
    pbs.push(format!(
        "{{\"log_p\":{},\"padding\":{},\"n\":{},\"big_n\":{},\"k\":{},\"bs_log_b\":{},\"bs_d\":{},\"ks_log_b\":{},\"ks_d\":{},\"brk\":{},\"ksk_a\":{},\"ksk_b\":{},\"v\":{},\"cts\":[{}],\"out\":[{}]}}",
        log_p, padding, n, big_n, k, bs.0, bs.1, ks.0, ks.1, brk, ksk_a, ksk_b, arr(v.iter().map(|z| z.to_u64())), cts.join(","), outs.join(",")
    ));

    // Tggsw::external_product / cmux (tggsw.rs:100-121) of brk[0] on uniform TGLWE ciphertexts
    let (ct0, ct1) = (uniform_glwe(k, big_n, &mut rng), uniform_glwe(k, big_n, &mut rng));
    let e = Tggsw::external_product(param.tggsw(), &bk.brk()[0], ct0.clone());
    let c = Tggsw::cmux(param.tggsw(), &bk.brk()[0], ct0.clone(), ct1.clone());
    ext.push(format!(
        "{{\"k\":{},\"d\":{},\"log_b\":{},\"n\":{},\"rows\":{},\"ct0\":{},\"ct1\":{},\"external_product\":{},\"cmux\":{}}}",
        k, bs.1, bs.0, big_n, row(&bk.brk()[0]), glwe(&ct0), glwe(&ct1), glwe(&e), glwe(&c)
    ));

    // Tlwe::key_switch (tlwe.rs:144-153) of a uniform extracted ciphertext of dimension k N
    let big = TlweCiphertext(AVec::<T64>::sample_uniform(k * big_n, &mut rng), T64::sample_uniform(&mut rng));
    let o = Tlwe::key_switch(&bk, bk.ksk(), big.clone());
    ksw.push(format!(
        "{{\"log_b\":{},\"d\":{},\"ksk_a\":{},\"ksk_b\":{},\"a\":{},\"b\":{},\"out\":{}}}",
        ks.0, ks.1, ksk_a, ksk_b, words(big.a().iter()), big.b().to_u64(), tlwe(&o)
    ));
}

#[test]
fn pin_dump_tfhe() {
    let (mut pbs, mut ext, mut ksw) = (Vec::new(), Vec::new(), Vec::new());
    // tiny (the shape of tests/golden/tfhe_ckks.json), the reference's tggsw test shape (N = 256, k = 2, d = 8), and a reduced
    // TFHE-T (N = 2048, log_b 23, d 1 as in bootstrapping.rs:141-152, n cut to 8 so that the key stays small)
    dump(4, 16, 1, (8, 2), (4, 5), (1.0e-9, 1.0e-15), 0x5EED_0301, &mut pbs, &mut ext, &mut ksw);
    dump(6, 256, 2, (8, 8), (4, 5), (1.339775301998614e-7, 2.845267479601915e-15), 0x5EED_0302, &mut pbs, &mut ext, &mut ksw);
    dump(8, 2048, 1, (23, 1), (4, 5), (1.339775301998614e-7, 2.845267479601915e-15), 0x5EED_0303, &mut pbs, &mut ext, &mut ksw);
    let dir = PathBuf::from(env::var("FHE_PIN_OUT").unwrap_or_else(|_| ".".into()));
    fs::create_dir_all(&dir).unwrap();
    let body = format!("{{\"tfhe_pbs\":[{}],\n\"tggsw\":[{}],\n\"tlwe_key_switch\":[{}]}}\n", pbs.join(","), ext.join(","), ksw.join(","));
    fs::write(dir.join("ref_tfhe.json"), body).unwrap();
    println!("wrote {}", dir.join("ref_tfhe.json").display());
}
