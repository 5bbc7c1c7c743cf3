// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// Restatement of scheme/fhew/src/{lwe,rlwe,rgsw,bootstrapping,fhew}.rs with the reference's dataflow:
// coefficient-form keys, every polynomial product = 3 transforms (ring.rs:256-264 -> fft/zq.rs:14-25).
// Key generation follows the reference structure; the noise/uniform sampler is a seeded splitmix64
// stream (util/src/misc/distribution.rs is NOT on the parity path — any small-noise source works).
#pragma once
#include "orc_util.hpp"

namespace orc {

struct Rng {
    u64 s;
    explicit Rng(u64 seed) : s(seed) {}
    u64 next() {
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double unif() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    u64 below(u64 q) { return (u64)(((u128)next() * (u128)q) >> 64); }
};

// distribution.rs:25-49 dg(std_dev, n): discrete Gaussian on [-floor(n*sd), floor(n*sd)] by CDF weights
struct DiscreteGaussian {
    std::vector<double> cum;
    i64 mx;
    DiscreteGaussian(double sd, unsigned n) {
        auto erf_as = [](double x) {
            double p = 0.3275911, a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429;
            double t = 1.0 / (1.0 + p * std::fabs(x));
            double pos = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * std::exp(-x * x);
            return std::signbit(x) ? -pos : pos;
        };
        auto cdf = [&](double x) { return (1.0 + erf_as(x / (sd * std::sqrt(2.0)))) / 2.0; };
        mx = (i64)std::floor((double)n * sd);
        double tot = 0;
        for (i64 i = -mx; i <= mx; ++i) {
            tot += cdf((double)i + 0.5) - cdf((double)i - 0.5);
            cum.push_back(tot);
        }
        for (auto& c : cum) c /= tot;
    }
    i64 sample(Rng& r) const {
        double u = r.unif();
        size_t lo = 0, hi = cum.size() - 1;
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (cum[mid] > u)
                hi = mid;
            else
                lo = mid + 1;
        }
        return (i64)lo - mx;
    }
};

// ---------------------------------------------------------------------------------------------
// Parameters: bootstrapping.rs:21-90, rgsw.rs:18-27, rlwe.rs:13-20, lwe.rs:17-53
// ---------------------------------------------------------------------------------------------
struct FhewParam {
    unsigned log_n;
    u64 big_q;        // RLWE/RGSW modulus Q (prime, NTT friendly)
    u64 p;            // plaintext modulus (4 for gates)
    unsigned rlwe_log_b, rlwe_d;  // RLWE key-switch decomposor (automorphism keys)
    unsigned rgsw_log_b, rgsw_d;  // RGSW decomposor (external product)
    unsigned n_s;     // LWE_s dimension
    u64 q_ks;         // LWE_s modulus
    unsigned ks_log_b, ks_d;      // LWE key-switch decomposor
    unsigned w;       // LMKCDEY window
    size_t n() const { return (size_t)1 << log_n; }
    u64 q() const { return 2 * (u64)n(); }                                               // bootstrapping.rs:74-76
    u64 big_q_by_8() const { return zq_from_f64(big_q, (double)big_q / 8.0); }            // :62-64
    u64 big_q_by_4() const { return zq_from_f64(big_q, (double)big_q / 4.0); }            // :66-68
    DecomposorZq rlwe_dec() const { return DecomposorZq(big_q, rlwe_log_b, rlwe_d); }
    DecomposorZq rgsw_dec() const { return DecomposorZq(big_q, rgsw_log_b, rgsw_d); }
    DecomposorZq ks_dec() const { return DecomposorZq(q_ks, ks_log_b, ks_d); }
    // bootstrapping.rs:86-89 ak_t: [-g, g^1 .. g^w] mod 2N, as centred i64
    std::vector<i64> ak_t() const {
        u64 m = q();
        u64 g = zq_from_i64(m, 5);
        std::vector<i64> t;
        t.push_back(zq_to_i64(m, zq_neg(m, g)));
        u64 pw = g;
        for (unsigned i = 0; i < w; ++i) {
            t.push_back(zq_to_i64(m, pw));
            pw = zq_mul(m, pw, g);
        }
        return t;
    }
};
// fhew/boolean.rs:225-239 single_key_testing_param
static inline FhewParam fhew_single_key_testing_param() {
    FhewParam p;
    p.log_n = 9;
    p.big_q = two_adic_primes(28, 10, 1)[0];
    p.p = 4;
    p.rlwe_log_b = 7;
    p.rlwe_d = 4;
    p.rgsw_log_b = 7;
    p.rgsw_d = 4;
    p.n_s = 100;
    p.q_ks = 1ull << 16;
    p.ks_log_b = 4;
    p.ks_d = 4;
    p.w = 10;
    return p;
}

struct LweCt {  // lwe.rs:77 LweCiphertext(a, b)
    Vec a;
    u64 b;
};
struct RlweCt {  // rlwe.rs:69 RlweCiphertext(a, b)
    Vec a, b;
};

struct FhewKey {
    FhewParam param;
    std::vector<i64> z;  // RLWE secret (dim N) == LWE_z secret
    std::vector<i64> s;  // LWE_s secret (dim n_s)
    // lwe.rs:108-119 ksk: N*d_ks LWE_s ciphertexts, index = digit*N + coefficient
    std::vector<LweCt> ksk;
    // bootstrapping.rs:126-133 brk[j]: RGSW(X^{s_j}) = 2d RLWE rows (rgsw.rs:84-105)
    std::vector<std::vector<RlweCt>> brk;
    // bootstrapping.rs:134-137 ak[v]: d RLWE rows for automorphism t = ak_t[v]
    std::vector<std::vector<RlweCt>> ak;
    std::vector<i64> ak_t;
};

// Rq * AVec<i64> (ring.rs:272-276 -> from_i64 then coefficient product)
static inline Vec rq_mul_i64(u64 q, const Vec& a, const std::vector<i64>& s) {
    Vec r = a;
    Vec sb(s.size());
    for (size_t i = 0; i < s.size(); ++i) sb[i] = zq_from_i64(q, s[i]);
    nega_cyclic_ntt_mul_assign(q, r.data(), sb.data(), r.size());
    return r;
}
// rlwe.rs:146-156 sk_encrypt
static inline RlweCt rlwe_sk_encrypt(u64 q, size_t n, const std::vector<i64>& sk, const Vec& pt, Rng& rng, const DiscreteGaussian& dg) {
    RlweCt ct;
    ct.a.resize(n);
    for (auto& x : ct.a) x = rng.below(q);
    Vec as = rq_mul_i64(q, ct.a, sk);
    ct.b.resize(n);
    for (size_t i = 0; i < n; ++i) {
        u64 e = zq_from_i64(q, dg.sample(rng));
        ct.b[i] = zq_add(q, zq_add(q, as[i], e), pt[i]);
    }
    return ct;
}
// rlwe.rs:171-175 decrypt
static inline Vec rlwe_decrypt(u64 q, const std::vector<i64>& sk, const RlweCt& ct) {
    Vec as = rq_mul_i64(q, ct.a, sk);
    Vec pt(ct.b.size());
    for (size_t i = 0; i < pt.size(); ++i) pt[i] = zq_sub(q, ct.b[i], as[i]);
    return pt;
}
// lwe.rs:130-140 sk_encrypt
static inline LweCt lwe_sk_encrypt(u64 q, size_t n, const std::vector<i64>& sk, u64 pt, Rng& rng, const DiscreteGaussian& dg) {
    LweCt ct;
    ct.a.resize(n);
    for (auto& x : ct.a) x = rng.below(q);
    u64 e = zq_from_i64(q, dg.sample(rng));
    u64 dot = zq_mul(q, ct.a[0], zq_from_i64(q, sk[0]));
    for (size_t i = 1; i < n; ++i) dot = zq_add(q, dot, zq_mul(q, ct.a[i], zq_from_i64(q, sk[i])));
    ct.b = zq_add(q, zq_add(q, dot, pt), e);
    return ct;
}
// lwe.rs:142-149 decrypt
static inline u64 lwe_decrypt(u64 q, const std::vector<i64>& sk, const LweCt& ct) {
    u64 dot = zq_mul(q, ct.a[0], zq_from_i64(q, sk[0]));
    for (size_t i = 1; i < ct.a.size(); ++i) dot = zq_add(q, dot, zq_mul(q, ct.a[i], zq_from_i64(q, sk[i])));
    return zq_sub(q, ct.b, dot);
}

// bootstrapping.rs:122-146 key_gen (+ rlwe.rs:109-132, rgsw.rs:84-105, lwe.rs:108-119)
static inline FhewKey fhew_key_gen(const FhewParam& P, u64 seed) {
    FhewKey K;
    K.param = P;
    Rng rng(seed);
    DiscreteGaussian dg(3.2, 6);
    size_t n = P.n();
    u64 Q = P.big_q;
    K.z.resize(n);
    for (auto& v : K.z) v = dg.sample(rng);  // rlwe.rs:94-96
    K.s.resize(P.n_s);
    for (auto& v : K.s) v = dg.sample(rng);  // lwe.rs:103-106
    // ksk: pt = power_up(-z).flatten()  => pt[k*N + i] = base_k * (-z_i) mod q_ks
    DecomposorZq ksd = P.ks_dec();
    for (unsigned k = 0; k < P.ks_d; ++k)
        for (size_t i = 0; i < n; ++i) {
            u64 pt = zq_mul(P.q_ks, ksd.base(k), zq_from_i64(P.q_ks, -K.z[i]));
            K.ksk.push_back(lwe_sk_encrypt(P.q_ks, P.n_s, K.s, pt, rng, dg));
        }
    // brk[j] = RGSW_z(X^{s_j})
    DecomposorZq gd = P.rgsw_dec();
    for (size_t j = 0; j < P.n_s; ++j) {
        Vec pt(n, 0);
        pt[0] = 1 % Q;
        monomial_mul_zq(Q, pt.data(), n, K.s[j]);
        std::vector<RlweCt> rows;
        Vec zero(n, 0);
        for (unsigned r = 0; r < 2 * P.rgsw_d; ++r) rows.push_back(rlwe_sk_encrypt(Q, n, K.z, zero, rng, dg));
        for (unsigned k = 0; k < P.rgsw_d; ++k)
            for (size_t i = 0; i < n; ++i) {
                u64 v = zq_mul(Q, pt[i], gd.base(k));
                rows[k].a[i] = zq_add(Q, rows[k].a[i], v);                          // rgsw.rs:102
                rows[P.rgsw_d + k].b[i] = zq_add(Q, rows[P.rgsw_d + k].b[i], v);    // rgsw.rs:103
            }
        K.brk.push_back(std::move(rows));
    }
    // ak[v] = ksk_gen(z, z.automorphism(t))
    DecomposorZq rd = P.rlwe_dec();
    K.ak_t = P.ak_t();
    for (i64 t : K.ak_t) {
        // LweSecretKey automorphism on i64 (avec.rs:34-50 with i64 negation)
        std::vector<i64> za(n);
        {
            i64 m = 2 * (i64)n;
            size_t tt = (size_t)(((t % m) + m) % m);
            za = K.z;
            for (size_t i = 0; i < n; ++i) {
                size_t it = (i * tt) % (2 * n);
                if (it < n)
                    za[it] = K.z[i];
                else
                    za[it - n] = -K.z[i];
            }
        }
        std::vector<RlweCt> rows;
        for (unsigned k = 0; k < P.rlwe_d; ++k) {
            Vec pt(n);
            for (size_t i = 0; i < n; ++i) pt[i] = zq_mul(Q, rd.base(k), zq_from_i64(Q, -za[i]));  // power_up(-sk1)
            rows.push_back(rlwe_sk_encrypt(Q, n, K.z, pt, rng, dg));
        }
        K.ak.push_back(std::move(rows));
    }
    return K;
}

// lwe.rs:90-99
static inline LweCt lwe_mod_switch(u64 q, const LweCt& ct, u64 qp) {
    LweCt o;
    for (u64 v : ct.a) o.a.push_back(zq_mod_switch(q, v, qp));
    o.b = zq_mod_switch(q, ct.b, qp);
    return o;
}
static inline LweCt lwe_mod_switch_odd(u64 q, const LweCt& ct, u64 qp) {
    LweCt o;
    for (u64 v : ct.a) o.a.push_back(zq_mod_switch_odd(q, v, qp));
    o.b = zq_mod_switch_odd(q, ct.b, qp);
    return o;
}
// lwe.rs:151-160 key_switch
static inline LweCt lwe_key_switch(const FhewParam& P, const std::vector<LweCt>& ksk, const LweCt& ct) {
    u64 q = P.q_ks;
    DecomposorZq dec = P.ks_dec();
    size_t n = ct.a.size();
    Vec limbs(dec.d * n);
    dec.decompose_vec(ct.a.data(), n, limbs.data());  // limb-major flatten
    LweCt o;
    o.a.assign(P.n_s, 0);
    o.b = 0;
    for (size_t idx = 0; idx < limbs.size(); ++idx) {
        u64 l = limbs[idx];
        if (idx == 0) {
            for (size_t j = 0; j < P.n_s; ++j) o.a[j] = zq_mul(q, ksk[idx].a[j], l);
            o.b = zq_mul(q, ksk[idx].b, l);
        } else {
            for (size_t j = 0; j < P.n_s; ++j) o.a[j] = zq_add(q, o.a[j], zq_mul(q, ksk[idx].a[j], l));
            o.b = zq_add(q, o.b, zq_mul(q, ksk[idx].b, l));
        }
    }
    o.b = zq_add(q, o.b, ct.b);
    return o;
}

// Dot of key polynomials with decomposed limbs (misc.rs:50-62): Σ_k row_k * limb_k, each `*` is a
// coefficient-form product (3 transforms)
static inline Vec dot_rows(u64 q, size_t n, const std::vector<const Vec*>& rows, const Vec& limbs) {
    Vec acc;
    for (size_t k = 0; k < rows.size(); ++k) {
        Vec prod = *rows[k];
        nega_cyclic_ntt_mul_assign(q, prod.data(), limbs.data() + k * n, n);
        if (k == 0)
            acc = prod;
        else
            for (size_t i = 0; i < n; ++i) acc[i] = zq_add(q, acc[i], prod[i]);
    }
    return acc;
}
// rgsw.rs:116-128 external_product
static inline RlweCt rgsw_external_product(const FhewParam& P, const std::vector<RlweCt>& ct0, const RlweCt& ct1) {
    size_t n = P.n();
    DecomposorZq dec = P.rgsw_dec();
    Vec limbs(2 * dec.d * n);
    dec.decompose_vec(ct1.a.data(), n, limbs.data());
    dec.decompose_vec(ct1.b.data(), n, limbs.data() + dec.d * n);
    std::vector<const Vec*> ra, rb;
    for (auto& r : ct0) {
        ra.push_back(&r.a);
        rb.push_back(&r.b);
    }
    RlweCt o;
    o.a = dot_rows(P.big_q, n, ra, limbs);
    o.b = dot_rows(P.big_q, n, rb, limbs);
    return o;
}
// rgsw.rs:130-150 internal_product: ct0's rows are moved to the evaluation domain once; every row of ct1 is decomposed, its limbs
// transformed, dotted with ct0's a / b columns in the evaluation domain and transformed back (the reference's own dataflow)
static inline std::vector<RlweCt> rgsw_internal_product(u64 q, size_t n, const DecomposorZq& dec, const std::vector<RlweCt>& ct0,
                                                        const std::vector<RlweCt>& ct1) {
    std::vector<Vec> e0a, e0b;
    for (const RlweCt& r : ct0) {
        e0a.push_back(r.a);
        e0b.push_back(r.b);
        nega_cyclic_ntt_in_place(q, e0a.back().data(), n);
        nega_cyclic_ntt_in_place(q, e0b.back().data(), n);
    }
    std::vector<RlweCt> out;
    for (const RlweCt& r : ct1) {
        Vec limbs(2 * dec.d * n);
        dec.decompose_vec(r.a.data(), n, limbs.data());
        dec.decompose_vec(r.b.data(), n, limbs.data() + dec.d * n);
        for (size_t k = 0; k < 2 * dec.d; ++k) nega_cyclic_ntt_in_place(q, limbs.data() + k * n, n);
        RlweCt o;
        o.a.assign(n, 0);
        o.b.assign(n, 0);
        for (size_t k = 0; k < 2 * dec.d; ++k)
            for (size_t i = 0; i < n; ++i) {
                o.a[i] = zq_add(q, o.a[i], zq_mul(q, e0a[k][i], limbs[k * n + i]));
                o.b[i] = zq_add(q, o.b[i], zq_mul(q, e0b[k][i], limbs[k * n + i]));
            }
        nega_cyclic_intt_in_place(q, o.a.data(), n);
        nega_cyclic_intt_in_place(q, o.b.data(), n);
        out.push_back(o);
    }
    return out;
}
// rlwe.rs:177-186 key_switch
static inline RlweCt rlwe_key_switch(const FhewParam& P, const std::vector<RlweCt>& ksk, const RlweCt& ct) {
    size_t n = P.n();
    DecomposorZq dec = P.rlwe_dec();
    Vec limbs(dec.d * n);
    dec.decompose_vec(ct.a.data(), n, limbs.data());
    std::vector<const Vec*> ra, rb;
    for (auto& r : ksk) {
        ra.push_back(&r.a);
        rb.push_back(&r.b);
    }
    RlweCt o;
    o.a = dot_rows(P.big_q, n, ra, limbs);
    o.b = dot_rows(P.big_q, n, rb, limbs);
    for (size_t i = 0; i < n; ++i) o.b[i] = zq_add(P.big_q, o.b[i], ct.b[i]);
    return o;
}
// rlwe.rs:188-191 automorphism (+ :80-82)
static inline RlweCt rlwe_automorphism(const FhewParam& P, const std::vector<RlweCt>& ak, i64 t, const RlweCt& ct) {
    RlweCt au;
    au.a = automorphism_zq(P.big_q, ct.a.data(), P.n(), t);
    au.b = automorphism_zq(P.big_q, ct.b.data(), P.n(), t);
    return rlwe_key_switch(P, ak, au);
}
// rlwe.rs:193-202 sample_extract
static inline LweCt rlwe_sample_extract(u64 q, const RlweCt& ct, size_t i) {
    size_t n = ct.a.size();
    LweCt o;
    for (size_t k = i + 1; k-- > 0;) o.a.push_back(ct.a[k]);
    for (size_t k = n; k-- > i + 1;) o.a.push_back(zq_neg(q, ct.a[k]));
    o.b = ct.b[i];
    return o;
}

// bootstrapping.rs:212-231 i_minus_i_plus / log_g_map
struct ISets {
    std::vector<std::vector<size_t>> minus, plus;
};
static inline ISets i_minus_i_plus(size_t n, const Vec& a) {
    u64 m = 2 * (u64)n;
    std::map<u64, size_t> lm, lp;
    u64 g = zq_from_i64(m, 5), pw = 1 % m;
    for (size_t l = 0; l < n / 2; ++l) {
        lm[zq_from_i64(m, -(i64)pw)] = l;  // g^l * (-1)
        lp[pw] = l;
        pw = zq_mul(m, pw, g);
    }
    ISets S;
    S.minus.assign(n / 2, {});
    S.plus.assign(n / 2, {});
    for (size_t i = 0; i < a.size(); ++i) {
        auto im = lm.find(a[i]), ip = lp.find(a[i]);
        if (im != lm.end() && ip == lp.end())
            S.minus[im->second].push_back(i);
        else if (im == lm.end() && ip != lp.end())
            S.plus[ip->second].push_back(i);
        else if (a[i] == 0) {
        } else
            throw std::runtime_error("i_minus_i_plus: unreachable (even non-zero exponent)");
    }
    return S;
}

// One step of the blind-rotation schedule: kind 0 = external product with brk[idx], kind 1 = automorphism with ak[idx]
struct BrStep {
    int kind;
    size_t idx;
};
// bootstrapping.rs:172-209 blind_rotate_core, control flow only
static inline std::vector<BrStep> blind_rotate_schedule(const FhewParam& P, const Vec& a) {
    ISets S = i_minus_i_plus(P.n(), a);
    std::vector<BrStep> steps;
    size_t v = 0;
    auto sweep = [&](const std::vector<std::vector<size_t>>& I) {
        for (size_t l = I.size() - 1; l >= 1; --l) {
            for (size_t j : I[l]) steps.push_back({0, j});
            v += 1;
            if (!I[l - 1].empty() || v == P.w || l == 1) {
                steps.push_back({1, v});
                v = 0;
            }
        }
        for (size_t j : I[0]) steps.push_back({0, j});
    };
    sweep(S.minus);
    steps.push_back({1, 0});
    sweep(S.plus);
    return steps;
}
// bootstrapping.rs:158-209 blind_rotate
static inline RlweCt fhew_blind_rotate(const FhewKey& K, const Vec& f, const LweCt& ct /* mod 2N */) {
    const FhewParam& P = K.param;
    size_t n = P.n();
    u64 m = P.q();
    // f' = f.automorphism(-g) * X^(b*g)   (centred exponent: ring.rs:400-406)
    Vec fp = automorphism_zq(P.big_q, f.data(), n, -5);
    i64 e = zq_to_i64(m, zq_mul(m, ct.b, zq_from_i64(m, 5)));
    monomial_mul_zq(P.big_q, fp.data(), n, e);
    RlweCt acc;
    acc.a.assign(n, 0);
    acc.b = fp;
    for (const BrStep& st : blind_rotate_schedule(P, ct.a)) {
        if (st.kind == 0)
            acc = rgsw_external_product(P, K.brk[st.idx], acc);
        else
            acc = rlwe_automorphism(P, K.ak[st.idx], K.ak_t[st.idx], acc);
    }
    return acc;
}
// bootstrapping.rs:149-155 bootstrap, first three steps (mod_switch, key_switch, mod_switch_odd)
static inline LweCt fhew_bootstrap_prologue(const FhewKey& K, const LweCt& ct) {
    const FhewParam& P = K.param;
    LweCt c1 = lwe_mod_switch(P.big_q, ct, P.q_ks);
    LweCt c2 = lwe_key_switch(P, K.ksk, c1);
    return lwe_mod_switch_odd(P.q_ks, c2, P.q());
}
static inline LweCt fhew_bootstrap(const FhewKey& K, const Vec& f, const LweCt& ct) {
    LweCt c3 = fhew_bootstrap_prologue(K, ct);
    RlweCt acc = fhew_blind_rotate(K, f, c3);
    return rlwe_sample_extract(K.param.big_q, acc, 0);
}
// fhew.rs:31-39 Fhew::op
static inline Vec fhew_gate_poly(const FhewParam& P, const int table[4]) {
    u64 q8 = P.big_q_by_8();
    u64 map[2] = {zq_neg(P.big_q, q8), q8};
    Vec f;
    for (int t = 0; t < 4; ++t)
        for (size_t r = 0; r < P.q() / 8; ++r) f.push_back(map[table[t]]);
    return f;
}
static inline LweCt fhew_op(const FhewKey& K, const int table[4], const LweCt& ct) {
    Vec f = fhew_gate_poly(K.param, table);
    LweCt o = fhew_bootstrap(K, f, ct);
    o.b = zq_add(K.param.big_q, o.b, K.param.big_q_by_8());
    return o;
}
static inline LweCt lwe_add(u64 q, const LweCt& x, const LweCt& y) {
    LweCt o = x;
    for (size_t i = 0; i < o.a.size(); ++i) o.a[i] = zq_add(q, x.a[i], y.a[i]);
    o.b = zq_add(q, x.b, y.b);
    return o;
}
// fhew/boolean.rs:18-25 sk_encrypt of a bit; lwe.rs:121-124 encode
static inline LweCt fhew_encrypt_bit(const FhewKey& K, bool m, Rng& rng) {
    const FhewParam& P = K.param;
    DiscreteGaussian dg(3.2, 6);
    double delta = (double)P.big_q / (double)P.p;
    u64 pt = zq_from_f64(P.big_q, (double)zq_to_i64(P.p, m ? 1 : 0) * delta);
    return lwe_sk_encrypt(P.big_q, P.n(), K.z, pt, rng, dg);
}
// fhew/boolean.rs:37-41 decrypt; fhew.rs:20-25 decode; lwe.rs:126-128
static inline int fhew_decrypt_bit(const FhewKey& K, const LweCt& ct) {
    const FhewParam& P = K.param;
    u64 pt = lwe_decrypt(P.big_q, K.z, ct);
    double delta = (double)P.big_q / (double)P.p;
    u64 m = zq_from_f64(P.p, (double)zq_to_i64(P.big_q, pt) / delta);
    return (int)m;  // caller asserts m in {0,1}
}

}  // namespace orc
