// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// Restatement of scheme/tfhe/src/{tlwe,tglwe,tggsw,bootstrapping}.rs.  Polynomial products over T64 go
// through the f64 FFT of util/src/ring/fft/c64.rs exactly as the reference does (ring.rs:315-320), each
// product rounded back to u64 before the wrapping sum of `Dot` (misc.rs:59-61).
#pragma once
#include "orc_fhew.hpp"  // Rng
#include "orc_util.hpp"

namespace orc {

struct TfheParam {
    unsigned log_p, padding;
    unsigned n;          // TLWE dimension
    double tlwe_std;     // TLWE noise
    unsigned ks_log_b, ks_d;
    unsigned big_n;      // ring degree N
    unsigned k;          // GLWE dimension ("n" of TglweParam, tglwe.rs:13-18)
    double tglwe_std;
    unsigned bs_log_b, bs_d;
    u64 p() const { return 1ull << log_p; }
    unsigned log_delta() const { return 64 - (log_p + padding); }  // tlwe.rs:47-49
    DecomposorT64 ks_dec() const { return DecomposorT64(ks_log_b, ks_d); }
    DecomposorT64 bs_dec() const { return DecomposorT64(bs_log_b, bs_d); }
};
// tfhe/bootstrapping.rs:141-152 test parameter set
static inline TfheParam tfhe_testing_param() {
    TfheParam P;
    P.log_p = 4;
    P.padding = 1;
    P.n = 1024;
    P.tlwe_std = 1.339775301998614e-7;
    P.ks_log_b = 4;
    P.ks_d = 5;
    P.big_n = 2048;
    P.k = 1;
    P.tglwe_std = 2.845267479601915e-15;
    P.bs_log_b = 23;
    P.bs_d = 1;
    return P;
}

struct TlweCt {
    Vec a;
    u64 b;
};
struct TglweCt {
    std::vector<Vec> a;  // k polynomials
    Vec b;
};
struct TfheKey {
    TfheParam param;
    std::vector<i64> z;  // TLWE secret (n bits)
    std::vector<i64> s;  // TGLWE secret (k*N bits)
    std::vector<std::vector<TglweCt>> brk;  // n TGGSW ciphertexts of (k+1)*d rows
    std::vector<TlweCt> ksk;                // (k*N)*d_ks TLWE ciphertexts, index = digit*(kN) + coefficient
};

// distribution.rs:51-54 tdg: Normal(0, sd) -> fractional part scaled by 2^64
static inline u64 tdg_sample(double sd, Rng& rng) {
    double u1 = rng.unif(), u2 = rng.unif();
    if (u1 < 1e-300) u1 = 1e-300;
    double v = sd * std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    double frac = v - std::round(v);
    return t64_from_f64(frac * 18446744073709551616.0);
}
static inline Vec rt_mul(const Vec& a, const Vec& b) {
    Vec r = a;
    nega_cyclic_fft64_mul_assign_rt(r.data(), b.data(), r.size());
    return r;
}
static inline Vec rt_from_i64(const i64* v, size_t n) {
    Vec r(n);
    for (size_t i = 0; i < n; ++i) r[i] = (u64)v[i];
    return r;
}
// tlwe.rs:122-132 sk_encrypt
static inline TlweCt tlwe_sk_encrypt(size_t n, double sd, const std::vector<i64>& sk, u64 pt, Rng& rng) {
    TlweCt ct;
    ct.a.resize(n);
    u64 dot = 0;
    for (size_t i = 0; i < n; ++i) {
        ct.a[i] = rng.next();
        dot += ct.a[i] * (u64)sk[i];
    }
    ct.b = dot + tdg_sample(sd, rng) + pt;
    return ct;
}
// tlwe.rs:134-142 decrypt (mu_star rounded at log_delta) ; 118-120 decode
static inline u64 tlwe_decrypt_raw(const std::vector<i64>& sk, const TlweCt& ct) {
    u64 dot = 0;
    for (size_t i = 0; i < ct.a.size(); ++i) dot += ct.a[i] * (u64)sk[i];
    return ct.b - dot;
}
static inline u64 tlwe_decode(const TfheParam& P, u64 mu_star) {
    unsigned ld = P.log_delta();
    u64 mu = DecomposorT64::rounding_shr_bits(mu_star, ld) << ld;  // decompose.rs:120-122 round
    return zq_from_u64(P.p(), mu >> ld);
}
// tglwe.rs:92-103 sk_encrypt
static inline TglweCt tglwe_sk_encrypt(const TfheParam& P, const std::vector<i64>& s, const Vec& pt, Rng& rng) {
    size_t N = P.big_n;
    TglweCt ct;
    ct.b.assign(N, 0);
    for (unsigned j = 0; j < P.k; ++j) {
        Vec a(N);
        for (auto& x : a) x = rng.next();
        Vec as = rt_mul(a, rt_from_i64(s.data() + (size_t)j * N, N));
        for (size_t i = 0; i < N; ++i) ct.b[i] += as[i];
        ct.a.push_back(std::move(a));
    }
    for (size_t i = 0; i < N; ++i) ct.b[i] += tdg_sample(P.tglwe_std, rng) + pt[i];
    return ct;
}
// tggsw.rs:73-89 sk_encrypt
static inline std::vector<TglweCt> tggsw_sk_encrypt(const TfheParam& P, const std::vector<i64>& s, const Vec& pt, Rng& rng) {
    DecomposorT64 dec = P.bs_dec();
    size_t N = P.big_n;
    std::vector<TglweCt> rows;
    Vec zero(N, 0);
    for (unsigned r = 0; r < (P.k + 1) * dec.d; ++r) rows.push_back(tglwe_sk_encrypt(P, s, zero, rng));
    for (unsigned j = 0; j < P.k; ++j)
        for (unsigned i = 0; i < dec.d; ++i)
            for (size_t c = 0; c < N; ++c) rows[j * dec.d + i].a[j][c] += pt[c] * dec.base(i);
    for (unsigned i = 0; i < dec.d; ++i)
        for (size_t c = 0; c < N; ++c) rows[P.k * dec.d + i].b[c] += pt[c] * dec.base(i);
    return rows;
}
// tfhe/bootstrapping.rs:59-76 key_gen (z given by tlwe.rs:96-98 sk_gen: binary)
static inline TfheKey tfhe_key_gen(const TfheParam& P, u64 seed) {
    TfheKey K;
    K.param = P;
    Rng rng(seed);
    K.z.resize(P.n);
    for (auto& v : K.z) v = rng.unif() <= 0.5 ? 0 : 1;  // distribution.rs:6-8
    K.s.resize((size_t)P.k * P.big_n);
    for (auto& v : K.s) v = rng.unif() <= 0.5 ? 0 : 1;
    for (unsigned i = 0; i < P.n; ++i) {
        Vec pt(P.big_n, 0);
        pt[0] = (u64)K.z[i];
        K.brk.push_back(tggsw_sk_encrypt(P, K.s, pt, rng));
    }
    // tlwe.rs:100-111 ksk_gen(param, sk0 = z, sk1 = s): pt = power_up(-s).flatten()
    DecomposorT64 kd = P.ks_dec();
    for (unsigned d = 0; d < kd.d; ++d)
        for (size_t i = 0; i < K.s.size(); ++i) {
            u64 pt = (u64)(-K.s[i]) * kd.base(d);
            K.ksk.push_back(tlwe_sk_encrypt(P.n, P.tlwe_std, K.z, pt, rng));
        }
    return K;
}
// tglwe.rs:61-66 rotate
static inline TglweCt tglwe_rotate(const TglweCt& ct, i64 i) {
    TglweCt o = ct;
    for (auto& a : o.a) monomial_mul_t64(a.data(), a.size(), i);
    monomial_mul_t64(o.b.data(), o.b.size(), i);
    return o;
}
// tggsw.rs:100-112 external_product
static inline TglweCt tggsw_external_product(const TfheParam& P, const std::vector<TglweCt>& ct0, const TglweCt& ct1) {
    DecomposorT64 dec = P.bs_dec();
    size_t N = P.big_n;
    std::vector<Vec> limbs;  // [a_0 digits .., a_{k-1} digits .., b digits]
    auto push = [&](const Vec& v) {
        Vec flat(dec.d * N);
        dec.decompose_vec(v.data(), N, flat.data());
        for (unsigned i = 0; i < dec.d; ++i) limbs.emplace_back(flat.begin() + (size_t)i * N, flat.begin() + (size_t)(i + 1) * N);
    };
    for (auto& a : ct1.a) push(a);
    push(ct1.b);
    TglweCt o;
    o.a.assign(P.k, Vec(N, 0));
    o.b.assign(N, 0);
    for (size_t r = 0; r < limbs.size(); ++r) {
        for (unsigned j = 0; j < P.k; ++j) {
            Vec pr = rt_mul(ct0[r].a[j], limbs[r]);
            for (size_t c = 0; c < N; ++c) o.a[j][c] += pr[c];
        }
        Vec pr = rt_mul(ct0[r].b, limbs[r]);
        for (size_t c = 0; c < N; ++c) o.b[c] += pr[c];
    }
    return o;
}
// tggsw.rs:114-121 cmux = ct0 + ext(b, ct1 - ct0)
static inline TglweCt tggsw_cmux(const TfheParam& P, const std::vector<TglweCt>& b, const TglweCt& ct0, const TglweCt& ct1) {
    TglweCt diff = ct1;
    for (unsigned j = 0; j < P.k; ++j)
        for (size_t c = 0; c < P.big_n; ++c) diff.a[j][c] -= ct0.a[j][c];
    for (size_t c = 0; c < P.big_n; ++c) diff.b[c] -= ct0.b[c];
    TglweCt e = tggsw_external_product(P, b, diff);
    TglweCt o = ct0;
    for (unsigned j = 0; j < P.k; ++j)
        for (size_t c = 0; c < P.big_n; ++c) o.a[j][c] += e.a[j][c];
    for (size_t c = 0; c < P.big_n; ++c) o.b[c] += e.b[c];
    return o;
}
// tfhe/bootstrapping.rs:99-104 mod_switch
static inline void tfhe_mod_switch(const TfheParam& P, const TlweCt& ct, std::vector<i64>& a, i64& b) {
    unsigned lg = 0;
    while ((1u << lg) < 2 * P.big_n) ++lg;
    unsigned rb = 64 - lg;
    a.clear();
    for (u64 v : ct.a) a.push_back((i64)DecomposorT64::rounding_shr_bits(v, rb));
    b = (i64)DecomposorT64::rounding_shr_bits(ct.b, rb);
}
// tfhe/bootstrapping.rs:84-96 blind_rotate; v is the LUT polynomial over Z_p (tglwe.rs:80-84 encode)
static inline TglweCt tfhe_blind_rotate(const TfheKey& K, const Vec& v, const TlweCt& ct) {
    const TfheParam& P = K.param;
    TglweCt acc;
    acc.a.assign(P.k, Vec(P.big_n, 0));
    acc.b.resize(P.big_n);
    for (size_t i = 0; i < P.big_n; ++i) acc.b[i] = v[i] << P.log_delta();
    std::vector<i64> a;
    i64 b;
    tfhe_mod_switch(P, ct, a, b);
    acc = tglwe_rotate(acc, -b);
    for (size_t i = 0; i < a.size(); ++i) acc = tggsw_cmux(P, K.brk[i], acc, tglwe_rotate(acc, a[i]));
    return acc;
}
// tglwe.rs:115-127 sample_extract
static inline TlweCt tglwe_sample_extract(const TglweCt& ct, size_t i) {
    TlweCt o;
    for (auto& a : ct.a) {
        size_t N = a.size();
        for (size_t k = i + 1; k-- > 0;) o.a.push_back(a[k]);
        for (size_t k = N; k-- > i + 1;) o.a.push_back((u64)(0 - a[k]));
    }
    o.b = ct.b[i];
    return o;
}
// tlwe.rs:144-153 key_switch
static inline TlweCt tlwe_key_switch(const TfheParam& P, const std::vector<TlweCt>& ksk, const TlweCt& ct) {
    DecomposorT64 dec = P.ks_dec();
    size_t m = ct.a.size();
    Vec limbs(dec.d * m);
    dec.decompose_vec(ct.a.data(), m, limbs.data());
    TlweCt o;
    o.a.assign(P.n, 0);
    o.b = 0;
    for (size_t idx = 0; idx < limbs.size(); ++idx) {
        u64 l = limbs[idx];
        const TlweCt& kk = ksk[idx];
        for (size_t j = 0; j < P.n; ++j) o.a[j] += kk.a[j] * l;
        o.b += kk.b * l;
    }
    o.b += ct.b;
    return o;
}
// tfhe/bootstrapping.rs:78-82 bootstrap
static inline TlweCt tfhe_bootstrap(const TfheKey& K, const Vec& v, const TlweCt& ct) {
    TglweCt acc = tfhe_blind_rotate(K, v, ct);
    TlweCt ex = tglwe_sample_extract(acc, 0);
    return tlwe_key_switch(K.param, K.ksk, ex);
}
// tfhe/bootstrapping.rs:118-127 test LUT layout from a table over Z_p
static inline Vec tfhe_lut_poly(const TfheParam& P, const Vec& table /* p entries mod p */) {
    size_t m = P.big_n >> P.log_p;
    Vec v;
    for (size_t r = 0; r < m / 2; ++r) v.push_back(table[0]);
    for (size_t t = 1; t < table.size(); ++t)
        for (size_t r = 0; r < m; ++r) v.push_back(table[t]);
    for (size_t r = 0; r < m / 2; ++r) v.push_back(zq_neg(P.p(), table[0]));
    return v;
}

}  // namespace orc
