//! TEST INFRASTRUCTURE of learn-fhe_b200 (not part of han0110/learn-fhe): known-answer dump of the `util` crate.
//!
//! Drop this file into the reference checkout as `util/tests/pin_dump.rs` (oracle/pin/apply.sh does it) and run
//!     FHE_PIN_OUT=/path/to/learn-fhe_b200/tests/golden/ref cargo test --release -p util --test pin_dump -- --nocapture
//! It draws every input from `StdRng::seed_from_u64`, evaluates it with the reference's own public API and writes
//! `ref_util.json` in the schema of learn-fhe_b200/tests/golden/util_fhew.json / tfhe_ckks.json, which
//! tests/test_cpu_refpin.py replays against the C++ oracle (and tests/test_gpu_refpin.py against the CUDA library).
use rand::{rngs::StdRng, Rng, SeedableRng};
use std::{env, fs, path::PathBuf};
use util::{two_adic_primes, Base2Decomposor, BigInt, RnsRq, Rq, Rt, Zq, T64, X};

fn arr<T: ToString>(v: impl IntoIterator<Item = T>) -> String {
    format!("[{}]", v.into_iter().map(|x| x.to_string()).collect::<Vec<_>>().join(","))
}
fn zqs(q: u64, v: &[u64]) -> Vec<Zq> {
    v.iter().map(|x| Zq::from_u64(q, *x)).collect()
}
fn u64s<'a>(v: impl IntoIterator<Item = &'a Zq>) -> Vec<u64> {
    v.into_iter().map(|z| z.to_u64()).collect()
}
fn t64s(v: &[u64]) -> Vec<T64> {
    v.iter().map(|x| T64::from(*x)).collect()
}
fn rem(v: &BigInt, q: u64) -> u64 {
    let q = BigInt::from(q);
    let r = ((v % &q) + &q) % &q;
    r.to_string().parse().unwrap()
}

#[test]
fn pin_dump_util() {
    let mut rng = StdRng::seed_from_u64(0x5EED_0100);
    let mut out: Vec<String> = Vec::new();

    // ntt: util/src/ring/fft/zq.rs:27-36 through Rq::to_evaluation (ring.rs:140-144)
    let mut ntt = Vec::new();
    for (bits, log_n) in [(28usize, 3usize), (28, 9), (45, 5), (55, 7), (55, 12), (61, 4)] {
        let q = two_adic_primes(bits, log_n + 1).next().unwrap();
        let a: Vec<u64> = (0..1usize << log_n).map(|_| rng.gen_range(0..q)).collect();
        let ev = Rq::from(zqs(q, &a)).to_evaluation();
        assert_eq!(u64s(ev.to_coefficient().iter()), a);
        ntt.push(format!(
            "{{\"q\":{},\"a\":{},\"fwd\":{},\"generator\":{}}}",
            q,
            arr(a.iter()),
            arr(u64s(ev.iter())),
            Zq::generator(q).to_u64()
        ));
    }
    out.push(format!("\"ntt\":[{}]", ntt.join(",")));

    // coefficient-form negacyclic product (ring.rs:256-264)
    {
        let q = two_adic_primes(45, 5).next().unwrap();
        let a: Vec<u64> = (0..16).map(|_| rng.gen_range(0..q)).collect();
        let b: Vec<u64> = (0..16).map(|_| rng.gen_range(0..q)).collect();
        let mut p = Rq::from(zqs(q, &a));
        p *= &Rq::from(zqs(q, &b));
        out.push(format!(
            "\"negacyclic_mul\":{{\"q\":{},\"a\":{},\"b\":{},\"out\":{}}}",
            q,
            arr(a.iter()),
            arr(b.iter()),
            arr(u64s(p.iter()))
        ));
    }

    // decomposers (misc/decompose.rs): digits[i] = all d digits of v[i], least significant first
    let mut dz = Vec::new();
    for (q, log_b, d) in [
        (268409857u64, 7usize, 4usize),
        (1 << 16, 4, 4),
        (two_adic_primes(55, 12).next().unwrap(), 11, 5),
        (268409857, 5, 4),
        (97, 2, 3),
    ] {
        let mut v: Vec<u64> = (0..24).map(|_| rng.gen_range(0..q)).collect();
        v.extend([0, 1, q - 1, q / 2, q / 2 + 1, q / 2 - 1]);
        let dec = Base2Decomposor::<Zq>::new(q, log_b, d);
        let digits = v.iter().map(|x| arr(u64s(dec.decompose(&Zq::from_u64(q, *x)).collect::<Vec<_>>().iter())));
        dz.push(format!(
            "{{\"q\":{},\"log_b\":{},\"d\":{},\"v\":{},\"digits\":[{}]}}",
            q,
            log_b,
            d,
            arr(v.iter()),
            digits.collect::<Vec<_>>().join(",")
        ));
    }
    out.push(format!("\"decompose_zq\":[{}]", dz.join(",")));
    let mut dt = Vec::new();
    for (log_b, d) in [(23usize, 1usize), (4, 5), (8, 8), (2, 8), (1, 3), (7, 3)] {
        let mut v: Vec<u64> = (0..24).map(|_| rng.gen()).collect();
        v.extend([0, 1, u64::MAX, 1 << 63, (1 << 63) - 1]);
        let dec = Base2Decomposor::<T64>::new(log_b, d);
        let digits = v
            .iter()
            .map(|x| arr(dec.decompose(&T64::from(*x)).map(|t| t.to_u64()).collect::<Vec<_>>()));
        dt.push(format!(
            "{{\"log_b\":{},\"d\":{},\"v\":{},\"digits\":[{}]}}",
            log_b,
            d,
            arr(v.iter()),
            digits.collect::<Vec<_>>().join(",")
        ));
    }
    out.push(format!("\"decompose_t64\":[{}]", dt.join(",")));

    // mod switches (zq.rs:128-140)
    let mut ms = Vec::new();
    for (q, qp) in [(268409857u64, 1u64 << 16), (1 << 16, 1024), (1024, 268409857)] {
        let mut v: Vec<u64> = (0..40).map(|_| rng.gen_range(0..q)).collect();
        v.extend([0, 1, q - 1, q / 2]);
        let a = v.iter().map(|x| Zq::from_u64(q, *x).mod_switch(qp).to_u64());
        let b = v.iter().map(|x| Zq::from_u64(q, *x).mod_switch_odd(qp).to_u64());
        ms.push(format!(
            "{{\"q\":{},\"qp\":{},\"v\":{},\"mod_switch\":{},\"mod_switch_odd\":{}}}",
            q,
            qp,
            arr(v.iter()),
            arr(a),
            arr(b)
        ));
    }
    out.push(format!("\"mod_switch\":[{}]", ms.join(",")));

    // automorphism (avec.rs:34-50) and monomial product (ring.rs:299-313)
    let (mut au, mut mo) = (Vec::new(), Vec::new());
    let q = 268409857u64;
    for t in [5i64, -5, 25, 3, 31] {
        let a: Vec<u64> = (0..16).map(|_| rng.gen_range(0..q)).collect();
        let o = Rq::from(zqs(q, &a)).automorphism(t);
        au.push(format!("{{\"q\":{},\"a\":{},\"t\":{},\"out\":{}}}", q, arr(a.iter()), t, arr(u64s(o.iter()))));
    }
    for k in [0i64, 1, 15, 16, 17, 31, -1, -16, 40] {
        let a: Vec<u64> = (0..16).map(|_| rng.gen_range(0..q)).collect();
        let o = Rq::from(zqs(q, &a)) * (X ^ k);
        mo.push(format!("{{\"q\":{},\"a\":{},\"k\":{},\"out\":{}}}", q, arr(a.iter()), k, arr(u64s(o.iter()))));
    }
    out.push(format!("\"automorphism\":[{}]", au.join(",")));
    out.push(format!("\"monomial_mul\":[{}]", mo.join(",")));

    // f64 FFT torus product (ring/fft/c64.rs:11-56): full-range torus words times signed digits
    let mut ff = Vec::new();
    for (log_n, log_b) in [(0usize, 8u32), (1, 8), (4, 12), (6, 17), (8, 23), (11, 23)] {
        let n = 1usize << log_n;
        let a: Vec<u64> = (0..n).map(|_| rng.gen()).collect();
        let b: Vec<u64> = (0..n)
            .map(|_| (rng.gen_range(0..1i64 << log_b) - (1i64 << (log_b - 1))) as u64)
            .collect();
        let mut p = Rt::from(t64s(&a));
        p *= &Rt::from(t64s(&b));
        ff.push(format!(
            "{{\"a\":{},\"b\":{},\"out\":{}}}",
            arr(a.iter()),
            arr(b.iter()),
            arr(p.iter().map(|t| t.to_u64()))
        ));
    }
    out.push(format!("\"fft64_mul\":[{}]", ff.join(",")));

    // RNS base extension and rescaling (ring/rns.rs:83-132); limbs are read back as value mod modulus
    {
        let mut pr = two_adic_primes(55, 8);
        let qs: Vec<u64> = pr.by_ref().take(3).collect();
        let ps: Vec<u64> = pr.by_ref().take(3).collect();
        let big_q: BigInt = qs.iter().map(|q| BigInt::from(*q)).product();
        let mut cases = Vec::new();
        for i in 0..12 {
            let v: BigInt = match i {
                0 => BigInt::from(0),
                1 => BigInt::from(1),
                2 => &big_q - 1,
                3 => &big_q / 2,
                4 => &big_q / 2 + 1,
                _ => (0..3).fold(BigInt::from(0), |acc, _| (acc << 60) + BigInt::from(rng.gen::<u64>() >> 4)) % &big_q,
            };
            let x: Vec<u64> = qs.iter().map(|q| rem(&v, *q)).collect();
            let y = RnsRq::from_bigint(qs.clone(), &[v.clone()]).extend_bases(&ps).into_bigint();
            cases.push(format!("{{\"x\":{},\"out\":{}}}", arr(x.iter()), arr(ps.iter().map(|p| rem(&y[0], *p)))));
        }
        out.push(format!(
            "\"rns_extend_bases\":{{\"qs\":{},\"ps\":{},\"cases\":[{}]}}",
            arr(qs.iter()),
            arr(ps.iter()),
            cases.join(",")
        ));
        let mut rk = Vec::new();
        let all: Vec<u64> = qs.iter().chain(ps.iter()).copied().collect();
        let big_all: BigInt = all.iter().map(|q| BigInt::from(*q)).product();
        for (nq, k) in [(3usize, 1usize), (4, 2), (6, 3)] {
            let m = &all[..nq];
            let vals: Vec<BigInt> = (0..8)
                .map(|_| (0..6).fold(BigInt::from(0), |acc, _| (acc << 60) + BigInt::from(rng.gen::<u64>() >> 4)) % &big_all)
                .collect();
            let x = vals.iter().map(|v| arr(m.iter().map(|q| rem(v, *q))));
            let y = RnsRq::from_bigint(m.to_vec(), &vals).rescale_k(k).into_bigint();
            let o = y.iter().map(|v| arr(m[..nq - k].iter().map(|q| rem(v, *q))));
            rk.push(format!(
                "{{\"qs\":{},\"k\":{},\"x\":[{}],\"out\":[{}]}}",
                arr(m.iter()),
                k,
                x.collect::<Vec<_>>().join(","),
                o.collect::<Vec<_>>().join(",")
            ));
        }
        out.push(format!("\"rns_rescale_k\":[{}]", rk.join(",")));
    }

    let dir = PathBuf::from(env::var("FHE_PIN_OUT").unwrap_or_else(|_| ".".into()));
    fs::create_dir_all(&dir).unwrap();
    fs::write(dir.join("ref_util.json"), format!("{{{}}}\n", out.join(",\n"))).unwrap();
    println!("wrote {}", dir.join("ref_util.json").display());
}
