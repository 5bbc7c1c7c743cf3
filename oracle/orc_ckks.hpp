// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// Restatement of the RNS ciphertext arithmetic of scheme/ckks/src/ckks.rs (mul / relinearize /
// key_switch / rescale / rotate / conjugate) on top of orc_rns.hpp.  Encode/decode (256-bit floats,
// sfft.rs) are out of scope; tests encrypt integer plaintext polynomials directly.
#pragma once
#include "orc_fhew.hpp"  // Rng, DiscreteGaussian
#include "orc_rns.hpp"

namespace orc {

struct CkksParam {  // ckks.rs:12-35
    unsigned log_n;
    Vec qs, ps;
    size_t n() const { return (size_t)1 << log_n; }
    Vec qps() const {
        Vec r = qs;
        r.insert(r.end(), ps.begin(), ps.end());
        return r;
    }
};
static inline CkksParam ckks_param_new(unsigned log_n, unsigned log_qi, unsigned big_l) {
    CkksParam P;
    P.log_n = log_n;
    Vec primes = two_adic_primes(log_qi, log_n + 1, 2 * (size_t)big_l);
    if (primes.size() < 2 * (size_t)big_l) throw std::runtime_error("not enough primes");
    P.qs.assign(primes.begin(), primes.begin() + big_l);
    P.ps.assign(primes.begin() + big_l, primes.end());
    return P;
}
struct CkksCt {  // ckks.rs:108-121: tuple order is (b, a)
    RnsPoly b, a;
};
struct CkksKey {
    CkksParam param;
    std::vector<i64> sk;
    CkksCt rlk;                           // ckks.rs:164-167
    std::vector<std::pair<i64, CkksCt>> autk;  // (t, key) for rotations / conjugation
};

// RnsRq * AVec<i64> (rns.rs:160-167 -> ring.rs:272-276)
static inline RnsPoly rns_mul_i64(const RnsPoly& a, const std::vector<i64>& s) { return rns_mul(a, rns_from_i64(a.qs, s)); }

// ckks.rs:215-225 sk_encrypt: a uniform over pt.qs, e = dg(3.2, 6), b = -(a*sk) + e + pt
static inline CkksCt ckks_sk_encrypt(const std::vector<i64>& sk, const RnsPoly& pt, Rng& rng) {
    DiscreteGaussian dg(3.2, 6);
    size_t n = pt.n();
    CkksCt ct;
    ct.a = rns_zero(pt.qs, n);
    for (size_t i = 0; i < pt.qs.size(); ++i)
        for (size_t c = 0; c < n; ++c) ct.a.limbs[i][c] = rng.below(pt.qs[i]);
    std::vector<i64> e(n);
    for (auto& x : e) x = dg.sample(rng);
    ct.b = rns_add(rns_add(rns_neg(rns_mul_i64(ct.a, sk)), rns_from_i64(pt.qs, e)), pt);
    return ct;
}
// ckks.rs:241-248 decrypt: pt = b + a*sk
static inline RnsPoly ckks_decrypt(const std::vector<i64>& sk, const CkksCt& ct) { return rns_add(ct.b, rns_mul_i64(ct.a, sk)); }

// ckks.rs:154-162 ksk_gen: pt = from_i64(qps, sk') * P ; encrypt under sk over Q ∪ P
static inline CkksCt ckks_ksk_gen(const CkksParam& P, const std::vector<i64>& sk, const std::vector<i64>& sk_prime, Rng& rng) {
    Vec qps = P.qps();
    RnsPoly pt = rns_from_i64(qps, sk_prime);
    for (size_t i = 0; i < qps.size(); ++i) {
        u64 pm = prod_mod(P.ps, qps[i]);
        for (auto& v : pt.limbs[i]) v = zq_mul(qps[i], v, pm);
    }
    return ckks_sk_encrypt(sk, pt, rng);
}
// exact negacyclic product of two small-integer polynomials via one NTT prime (|coeff| << q/2).
// The reference uses Karatsuba over i64 (ckks.rs:79-81 -> karatsuba.rs) — same integers.
static inline std::vector<i64> small_negacyclic_mul(u64 q, const std::vector<i64>& a, const std::vector<i64>& b) {
    size_t n = a.size();
    Vec x(n), y(n);
    for (size_t i = 0; i < n; ++i) {
        x[i] = zq_from_i64(q, a[i]);
        y[i] = zq_from_i64(q, b[i]);
    }
    nega_cyclic_ntt_mul_assign(q, x.data(), y.data(), n);
    std::vector<i64> r(n);
    for (size_t i = 0; i < n; ++i) r[i] = zq_to_i64(q, x[i]);
    return r;
}
static inline std::vector<i64> automorphism_i64(const std::vector<i64>& z, i64 t) {
    size_t n = z.size();
    i64 m = 2 * (i64)n;
    size_t tt = (size_t)(((t % m) + m) % m);
    std::vector<i64> za = z;
    for (size_t i = 0; i < n; ++i) {
        size_t it = (i * tt) % (2 * n);
        if (it < n)
            za[it] = z[i];
        else
            za[it - n] = -z[i];
    }
    return za;
}
// ckks.rs:139-141 sk_gen (zo(0.5)), 164-184 rlk / conj / rot keys
static inline CkksKey ckks_key_gen(const CkksParam& P, u64 seed, const std::vector<i64>& auto_ts) {
    CkksKey K;
    K.param = P;
    Rng rng(seed);
    K.sk.resize(P.n());
    for (auto& v : K.sk) {  // distribution.rs:10-21 zo(0.5)
        double u = rng.unif();
        v = u <= 0.25 ? -1 : (u <= 0.5 ? 1 : 0);
    }
    K.rlk = ckks_ksk_gen(P, K.sk, small_negacyclic_mul(P.qs[0], K.sk, K.sk), rng);
    for (i64 t : auto_ts) K.autk.push_back({t, ckks_ksk_gen(P, K.sk, automorphism_i64(K.sk, t), rng)});
    return K;
}
// ckks.rs:123-125 rescale
static inline CkksCt ckks_rescale(const CkksCt& ct) { return CkksCt{rns_rescale_k(ct.b, 1), rns_rescale_k(ct.a, 1)}; }
// ckks.rs:284-293 key_switch
static inline CkksCt ckks_key_switch(const CkksParam& P, const CkksCt& ksk, const CkksCt& ct) {
    RnsPoly ct_a = rns_extend_bases(ct.a, P.ps);
    CkksCt o;
    o.b = rns_add(rns_rescale_k(rns_mul(ksk.b, ct_a), P.ps.size()), ct.b);
    o.a = rns_rescale_k(rns_mul(ksk.a, ct_a), P.ps.size());
    return o;
}
// ckks.rs:255-272 mul (+ relinearize)
static inline CkksCt ckks_mul(const CkksParam& P, const CkksCt& rlk, const CkksCt& c0, const CkksCt& c1) {
    RnsPoly d0 = rns_mul(c0.b, c1.b);
    RnsPoly d1 = rns_add(rns_mul(c0.b, c1.a), rns_mul(c0.a, c1.b));
    RnsPoly d2 = rns_mul(c0.a, c1.a);
    CkksCt quad{rns_zero(d2.qs, d2.n()), d2};
    CkksCt rl = ckks_key_switch(P, rlk, quad);
    CkksCt sum{rns_add(d0, rl.b), rns_add(d1, rl.a)};
    return ckks_rescale(sum);
}
// ckks.rs:274-282 conjugate / rotate = automorphism + key_switch (t = -1 or 5^j mod 2N)
static inline CkksCt ckks_automorphism_ks(const CkksParam& P, const CkksCt& ksk, i64 t, const CkksCt& ct) {
    CkksCt au{rns_automorphism(ct.b, t), rns_automorphism(ct.a, t)};
    return ckks_key_switch(P, ksk, au);
}

}  // namespace orc
