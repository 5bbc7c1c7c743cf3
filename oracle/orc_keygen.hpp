// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// Key generation of scheme/fhew/src/bootstrapping.rs:122-146 (with rlwe.rs:109-132, rgsw.rs:84-105, lwe.rs:108-119) exactly as
// fhew_key_gen (orc_fhew.hpp) states it, but drawing every random word from the COUNTER-BASED stream the device key generation
// uses (learn-fhe_b200/csrc/keygen_stream.cuh) instead of a sequential generator: the reference takes its randomness from the
// caller's RngCore, so which generator feeds the formulas is not part of the reference's semantics.  This is the checker of
// fhe_fhew_keygen: same seed -> the same key, word for word.
#pragma once
#include "../learn-fhe_b200/csrc/keygen_stream.cuh"
#include "orc_ckks.hpp"
#include "orc_fhew.hpp"
#include "orc_tfhe.hpp"

namespace orc {

static inline FhewKey fhew_key_gen_ctr(const FhewParam& P, u64 seed) {
    using namespace fhe;
    FhewKey K;
    K.param = P;
    const size_t n = P.n();
    const u64 Q = P.big_q;
    K.z.resize(n);
    for (size_t i = 0; i < n; ++i) K.z[i] = ks_gauss(seed, KS_FHEW_Z, i);          // rlwe.rs:94-96
    K.s.resize(P.n_s);
    for (size_t j = 0; j < P.n_s; ++j) K.s[j] = ks_gauss(seed, KS_FHEW_S, j);      // lwe.rs:103-106
    // ksk (lwe.rs:108-119, 130-140): row idx = digit * N + coefficient, b = <a, s> + pt + e, pt = base_k * (-z_i)
    DecomposorZq ksd = P.ks_dec();
    for (unsigned k = 0; k < P.ks_d; ++k)
        for (size_t i = 0; i < n; ++i) {
            const u64 idx = (u64)k * n + i;
            LweCt ct;
            ct.a.resize(P.n_s);
            u64 dot = 0;
            for (size_t j = 0; j < P.n_s; ++j) {
                ct.a[j] = ks_uniform(seed, KS_FHEW_KSK_A, idx * P.n_s + j, P.q_ks);
                dot = zq_add(P.q_ks, dot, zq_mul(P.q_ks, ct.a[j], zq_from_i64(P.q_ks, K.s[j])));
            }
            const u64 pt = zq_mul(P.q_ks, ksd.base(k), zq_from_i64(P.q_ks, -K.z[i]));
            ct.b = zq_add(P.q_ks, zq_add(P.q_ks, dot, pt), zq_from_i64(P.q_ks, ks_gauss(seed, KS_FHEW_KSK_E, idx)));
            K.ksk.push_back(ct);
        }
    // RLWE encryption of `pt` in row R of domain (da, de): a uniform, b = a * z + e + pt   (rlwe.rs:146-156)
    auto rlwe_row = [&](uint32_t da, uint32_t de, u64 R, const Vec& pt) {
        RlweCt ct;
        ct.a.resize(n);
        for (size_t c = 0; c < n; ++c) ct.a[c] = ks_uniform(seed, da, R * n + c, Q);
        Vec as = rq_mul_i64(Q, ct.a, K.z);
        ct.b.resize(n);
        for (size_t c = 0; c < n; ++c) ct.b[c] = zq_add(Q, zq_add(Q, as[c], zq_from_i64(Q, ks_gauss(seed, de, R * n + c))), pt[c]);
        return ct;
    };
    DecomposorZq gd = P.rgsw_dec();
    const Vec zero(n, 0);
    for (size_t j = 0; j < P.n_s; ++j) {  // brk[j] = RGSW_z(X^{s_j})   (rgsw.rs:84-105)
        Vec m(n, 0);
        m[0] = 1 % Q;
        monomial_mul_zq(Q, m.data(), n, K.s[j]);
        std::vector<RlweCt> rows;
        for (unsigned r = 0; r < 2 * P.rgsw_d; ++r) rows.push_back(rlwe_row(KS_FHEW_BRK_A, KS_FHEW_BRK_E, (u64)j * 2 * P.rgsw_d + r, zero));
        for (unsigned k = 0; k < P.rgsw_d; ++k)
            for (size_t i = 0; i < n; ++i) {
                const u64 v = zq_mul(Q, m[i], gd.base(k));
                rows[k].a[i] = zq_add(Q, rows[k].a[i], v);
                rows[P.rgsw_d + k].b[i] = zq_add(Q, rows[P.rgsw_d + k].b[i], v);
            }
        K.brk.push_back(std::move(rows));
    }
    DecomposorZq rd = P.rlwe_dec();
    K.ak_t = P.ak_t();
    for (size_t v = 0; v < K.ak_t.size(); ++v) {  // ak[v] = ksk_gen(z, z(X^t))   (rlwe.rs:109-132)
        const i64 t = K.ak_t[v], m2 = 2 * (i64)n;
        const size_t tt = (size_t)(((t % m2) + m2) % m2);
        std::vector<i64> za(n);
        for (size_t i = 0; i < n; ++i) {
            const size_t it = (i * tt) % (2 * n);
            if (it < n)
                za[it] = K.z[i];
            else
                za[it - n] = -K.z[i];
        }
        std::vector<RlweCt> rows;
        for (unsigned k = 0; k < P.rlwe_d; ++k) {
            Vec pt(n);
            for (size_t i = 0; i < n; ++i) pt[i] = zq_mul(Q, rd.base(k), zq_from_i64(Q, -za[i]));
            rows.push_back(rlwe_row(KS_FHEW_AK_A, KS_FHEW_AK_E, (u64)v * P.rlwe_d + k, pt));
        }
        K.ak.push_back(std::move(rows));
    }
    return K;
}

// Ckks::sk_gen / rlk_gen / cjk_gen / rtk_gen (scheme/ckks/src/ckks.rs:139-184, 215-225) as ckks_key_gen states them, fed by the counter
// stream of the device key generation (fhe_ckks_keygen): key 0 = relinearisation key, key 1 + i = automorphism key of auto_ts[i]
static inline CkksCt ckks_ksk_gen_ctr(const CkksParam& P, const std::vector<i64>& sk, const std::vector<i64>& sk_prime, u64 seed, uint32_t key) {
    using namespace fhe;
    Vec qps = P.qps();
    const size_t n = P.n();
    RnsPoly pt = rns_from_i64(qps, sk_prime);
    for (size_t i = 0; i < qps.size(); ++i) {
        u64 pm = prod_mod(P.ps, qps[i]);
        for (auto& v : pt.limbs[i]) v = zq_mul(qps[i], v, pm);
    }
    CkksCt ct;
    ct.a = rns_zero(qps, n);
    for (size_t i = 0; i < qps.size(); ++i)
        for (size_t c = 0; c < n; ++c) ct.a.limbs[i][c] = ks_uniform(seed, ks_ckks_a(key), (u64)i * n + c, qps[i]);
    std::vector<i64> e(n);
    for (size_t c = 0; c < n; ++c) e[c] = ks_gauss(seed, ks_ckks_e(key), c);
    ct.b = rns_add(rns_add(rns_neg(rns_mul_i64(ct.a, sk)), rns_from_i64(qps, e)), pt);
    return ct;
}
static inline CkksKey ckks_key_gen_ctr(const CkksParam& P, u64 seed, const std::vector<i64>& auto_ts) {
    using namespace fhe;
    CkksKey K;
    K.param = P;
    K.sk.resize(P.n());
    for (size_t c = 0; c < P.n(); ++c) K.sk[c] = ks_ternary(seed, KS_CKKS_SK, c);
    K.rlk = ckks_ksk_gen_ctr(P, K.sk, small_negacyclic_mul(P.qs[0], K.sk, K.sk), seed, 0);
    for (size_t i = 0; i < auto_ts.size(); ++i)
        K.autk.push_back({auto_ts[i], ckks_ksk_gen_ctr(P, K.sk, automorphism_i64(K.sk, auto_ts[i]), seed, (uint32_t)(1 + i))});
    return K;
}

// tfhe/bootstrapping.rs:59-76 key_gen (tlwe.rs:96-111, 122-132; tglwe.rs:92-103; tggsw.rs:73-89) as tfhe_key_gen states it, fed by
// the counter stream of fhe_tfhe_keygen; the torus noise is ks_tgauss (integer-only, see keygen_stream.cuh) with
// sigma_q = round(sigma * 2^64)
static inline TfheKey tfhe_key_gen_ctr(const TfheParam& P, u64 seed) {
    using namespace fhe;
    TfheKey K;
    K.param = P;
    const size_t N = P.big_n, kn = (size_t)P.k * N;
    const u64 sq_tlwe = (u64)std::llround(P.tlwe_std * 18446744073709551616.0), sq_tglwe = (u64)std::llround(P.tglwe_std * 18446744073709551616.0);
    K.z.resize(P.n);
    for (size_t i = 0; i < P.n; ++i) K.z[i] = ks_binary(seed, KS_TFHE_Z, i);
    K.s.resize(kn);
    for (size_t i = 0; i < kn; ++i) K.s[i] = ks_binary(seed, KS_TFHE_S, i);
    DecomposorT64 dec = P.bs_dec();
    const size_t rows_per = (size_t)(P.k + 1) * dec.d;
    for (unsigned i = 0; i < P.n; ++i) {
        std::vector<TglweCt> rows;
        for (size_t r = 0; r < rows_per; ++r) {  // TGLWE encryption of zero (tglwe.rs:92-103)
            const u64 R = (u64)i * rows_per + r;
            TglweCt ct;
            ct.b.assign(N, 0);
            for (unsigned j = 0; j < P.k; ++j) {
                Vec a(N);
                for (size_t c = 0; c < N; ++c) a[c] = ks_u64(seed, KS_TFHE_BRK_A, (R * P.k + j) * N + c);
                Vec as = rt_mul(a, rt_from_i64(K.s.data() + (size_t)j * N, N));
                for (size_t c = 0; c < N; ++c) ct.b[c] += as[c];
                ct.a.push_back(std::move(a));
            }
            for (size_t c = 0; c < N; ++c) ct.b[c] += ks_tgauss(seed, KS_TFHE_BRK_E, R * N + c, sq_tglwe);
            rows.push_back(std::move(ct));
        }
        const u64 zi = (u64)K.z[i];  // plaintext: the constant polynomial z_i (tggsw.rs:84-87)
        for (unsigned j = 0; j < P.k; ++j)
            for (unsigned d = 0; d < dec.d; ++d) rows[j * dec.d + d].a[j][0] += zi * dec.base(d);
        for (unsigned d = 0; d < dec.d; ++d) rows[P.k * dec.d + d].b[0] += zi * dec.base(d);
        K.brk.push_back(std::move(rows));
    }
    DecomposorT64 kd = P.ks_dec();
    for (unsigned d = 0; d < kd.d; ++d)
        for (size_t i = 0; i < kn; ++i) {  // tlwe.rs:100-111: pt = power_up(-s).flatten(), encrypted under z
            const u64 idx = (u64)d * kn + i;
            TlweCt ct;
            ct.a.resize(P.n);
            u64 dot = 0;
            for (size_t j = 0; j < P.n; ++j) {
                ct.a[j] = ks_u64(seed, KS_TFHE_KSK_A, idx * P.n + j);
                dot += ct.a[j] * (u64)K.z[j];
            }
            ct.b = dot + ks_tgauss(seed, KS_TFHE_KSK_E, idx, sq_tlwe) + (u64)(-K.s[i]) * kd.base(d);
            K.ksk.push_back(std::move(ct));
        }
    return K;
}

}  // namespace orc
