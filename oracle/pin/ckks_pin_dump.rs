//! TEST INFRASTRUCTURE of learn-fhe_b200 (not part of han0110/learn-fhe): known-answer dump of the CKKS crate.
//!
//! Installed by oracle/pin/apply.sh as `scheme/ckks/src/ckks/pin_dump.rs` with `#[cfg(test)] mod pin_dump;` appended to
//! `scheme/ckks/src/ckks.rs`.  Run
//!     FHE_PIN_OUT=/path/to/learn-fhe_b200/tests/golden/ref cargo test --release -p ckks pin_dump -- --nocapture
//! Keys and ciphertexts come from `StdRng::seed_from_u64`; outputs are the reference's own `Ckks::mul`, `Ckks::key_switch`
//! and `Ckks::rotate`.  Limbs are read back as (CRT value mod q_i) because `RnsRq` exposes no per-limb accessor.
//! Schema: a list of objects shaped like the `ckks` section of learn-fhe_b200/tests/golden/tfhe_ckks.json.
use super::*;
use rand::{rngs::StdRng, SeedableRng};
use std::{env, fs, path::PathBuf};

fn arr<T: ToString>(v: impl IntoIterator<Item = T>) -> String {
    format!("[{}]", v.into_iter().map(|x| x.to_string()).collect::<Vec<_>>().join(","))
}
fn rem(v: &BigInt, q: u64) -> u64 {
    let q = BigInt::from(q);
    (((v % &q) + &q) % &q).to_string().parse().unwrap()
}
// [limb][coefficient] residues of an RnsRq
fn limbs(x: &RnsRq) -> String {
    let qs = x.qs();
    let v = x.clone().into_bigint();
    arr(qs.iter().map(|q| arr(v.iter().map(|c| rem(c, *q)))))
}
// ciphertext tuple order (b, a) as in ckks.rs:112-121
fn ct(c: &CkksCiphertext) -> String {
    format!("[{},{}]", limbs(c.b()), limbs(c.a()))
}

fn dump(log_n: usize, log_qi: usize, big_l: usize, seed: u64) -> String {
    let mut rng = StdRng::seed_from_u64(seed);
    let param = CkksParam::new(log_n, log_qi, big_l);
    let sk = Ckks::sk_gen(&param, &mut rng);
    let rlk = Ckks::rlk_gen(&param, &sk, &mut rng);
    let rtk = Ckks::rtk_gen(&param, &sk, 1, &mut rng);
    let fresh = |rng: &mut StdRng| {
        let pt = CkksPlaintext(RnsRq::sample_i64(param.qs(), param.n(), dg(3.2, 6), rng));
        Ckks::sk_encrypt(&param, &sk, pt, rng)
    };
    let (ct0, ct1) = (fresh(&mut rng), fresh(&mut rng));
    let mul = Ckks::mul(&param, &rlk, ct0.clone(), ct1.clone());
    let mul2 = Ckks::mul(&param, &rlk, mul.clone(), mul.clone()); // one level down: the modulus-intersection product of rns.rs:143-158
    let ksw = Ckks::key_switch(&param, &rlk, ct0.clone());
    let rot = Ckks::rotate(&param, &rtk, ct0.clone());
    format!(
        "{{\"log_n\":{},\"qs\":{},\"ps\":{},\"sk\":{},\"ksk\":{},\"ct0\":{},\"ct1\":{},\"mul\":{},\"mul_again\":{},\"key_switch_ct0\":{},\
         \"rot_keys\":[{{\"j\":{},\"t\":{},\"ksk\":{}}}],\"rotate1_ct0\":{}}}",
        log_n,
        arr(param.qs().iter()),
        arr(param.ps().iter()),
        arr(sk.0.iter()),
        ct(&rlk),
        ct(&ct0),
        ct(&ct1),
        ct(&mul),
        ct(&mul2),
        ct(&ksw),
        rtk.j(),
        param.pow5(rtk.j()),
        ct(&rtk),
        ct(&rot)
    )
}

#[test]
fn pin_dump_ckks() {
    let cases = [dump(3, 55, 3, 0x5EED_0401), dump(6, 55, 4, 0x5EED_0402), dump(9, 45, 3, 0x5EED_0403)];
    let dir = PathBuf::from(env::var("FHE_PIN_OUT").unwrap_or_else(|_| ".".into()));
    fs::create_dir_all(&dir).unwrap();
    fs::write(dir.join("ref_ckks.json"), format!("{{\"ckks\":[{}]}}\n", cases.join(",\n"))).unwrap();
    println!("wrote {}", dir.join("ref_ckks.json").display());
}
