// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// extern "C" surface of the CPU restatement, loaded with ctypes by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.  Never linked into libfhe_b200.so.
#include <atomic>
#include <thread>

#include "orc_ckks.hpp"
#include "orc_fhew.hpp"
#include "orc_keygen.hpp"
#include "orc_rns.hpp"
#include "orc_tfhe.hpp"
#include "orc_util.hpp"

using namespace orc;

#define ORC_TRY(...)                       \
    try {                                  \
        __VA_ARGS__;                       \
        return 0;                          \
    } catch (const std::exception& e) {    \
        g_err = e.what();                  \
        return -1;                         \
    }

static thread_local std::string g_err;

template <typename F>
static void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t)
        ts.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& t : ts) t.join();
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---------------- Zq / primes ----------------
int orc_two_adic_primes(unsigned bits, unsigned log_n, size_t count, u64* out) {
    ORC_TRY({
        Vec p = two_adic_primes(bits, log_n, count);
        if (p.size() < count) throw std::runtime_error("fewer primes than requested");
        std::memcpy(out, p.data(), count * 8);
    })
}
int orc_is_prime(u64 q) { return is_prime(q) ? 1 : 0; }
u64 orc_zq_generator(u64 q) { return zq_generator(q); }
u64 orc_zq_two_adic_generator(u64 q, unsigned log_n) { return zq_two_adic_generator(q, log_n); }
u64 orc_zq_pow(u64 q, u64 v, u64 e) { return zq_pow(q, v, e); }
u64 orc_zq_inv(u64 q, u64 v) { return zq_inv(q, v); }
u64 orc_zq_mul(u64 q, u64 a, u64 b) { return zq_mul(q, a, b); }
u64 orc_zq_add(u64 q, u64 a, u64 b) { return zq_add(q, a, b); }
u64 orc_zq_sub(u64 q, u64 a, u64 b) { return zq_sub(q, a, b); }
u64 orc_zq_neg(u64 q, u64 a) { return zq_neg(q, a); }
int64_t orc_zq_to_i64(u64 q, u64 a) { return zq_to_i64(q, a); }
// element-wise vector ops (avec.rs:166-291)
void orc_vec_add(u64 q, const u64* a, const u64* b, u64* o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = zq_add(q, a[i], b[i]);
}
void orc_vec_sub(u64 q, const u64* a, const u64* b, u64* o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = zq_sub(q, a[i], b[i]);
}
void orc_vec_mul(u64 q, const u64* a, const u64* b, u64* o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = zq_mul(q, a[i], b[i]);
}
void orc_vec_neg(u64 q, const u64* a, u64* o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = zq_neg(q, a[i]);
}
void orc_mod_switch(u64 q, u64 qp, const u64* in, u64* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = zq_mod_switch(q, in[i], qp);
}
void orc_mod_switch_odd(u64 q, u64 qp, const u64* in, u64* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = zq_mod_switch_odd(q, in[i], qp);
}

// ---------------- NTT ----------------
// returns the table length (2^(s-1)); fills up to cap entries of the bit-reversed tables
long orc_twiddles(u64 q, u64* fwd, u64* inv, size_t cap) {
    try {
        const Twiddle& t = twiddle(q);
        size_t n = t.fwd.size() < cap ? t.fwd.size() : cap;
        if (fwd) std::memcpy(fwd, t.fwd.data(), n * 8);
        if (inv) std::memcpy(inv, t.inv.data(), n * 8);
        return (long)t.fwd.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
int orc_ntt_fwd(u64 q, u64* a, size_t n, size_t batch, int threads) {
    ORC_TRY({
        twiddle(q);
        parallel_for(batch, threads, [&](size_t b) { nega_cyclic_ntt_in_place(q, a + b * n, n); });
    })
}
int orc_ntt_inv(u64 q, u64* a, size_t n, size_t batch, int threads) {
    ORC_TRY({
        twiddle(q);
        parallel_for(batch, threads, [&](size_t b) { nega_cyclic_intt_in_place(q, a + b * n, n); });
    })
}
int orc_ntt_mul(u64 q, u64* a, const u64* b, size_t n, size_t batch, int threads) {
    ORC_TRY({
        twiddle(q);
        parallel_for(batch, threads, [&](size_t i) { nega_cyclic_ntt_mul_assign(q, a + i * n, b + i * n, n); });
    })
}
int orc_schoolbook_zq(u64 q, const u64* a, const u64* b, u64* out, size_t n) {
    ORC_TRY({
        Vec c = schoolbook_zq(q, a, b, n);
        std::memcpy(out, c.data(), n * 8);
    })
}
int orc_schoolbook_t64(const u64* a, const u64* b, u64* out, size_t n) {
    ORC_TRY({
        Vec c = schoolbook_t64(a, b, n);
        std::memcpy(out, c.data(), n * 8);
    })
}
int orc_fft64_mul(u64* a, const u64* b, size_t n, size_t batch, int threads) {
    ORC_TRY({
        twiddle64(n);
        if (n > 1) twiddle64(n / 2);
        parallel_for(batch, threads, [&](size_t i) { nega_cyclic_fft64_mul_assign_rt(a + i * n, b + i * n, n); });
    })
}
u64 orc_f64_mod_u64(double v) { return f64_mod_u64(v); }

// ---------------- ring ops ----------------
int orc_automorphism_zq(u64 q, const u64* in, u64* out, size_t n, int64_t t) {
    ORC_TRY({
        Vec v = automorphism_zq(q, in, n, t);
        std::memcpy(out, v.data(), n * 8);
    })
}
int orc_automorphism_t64(const u64* in, u64* out, size_t n, int64_t t) {
    ORC_TRY({
        Vec v = automorphism_t64(in, n, t);
        std::memcpy(out, v.data(), n * 8);
    })
}
int orc_monomial_mul_zq(u64 q, u64* a, size_t n, int64_t k) { ORC_TRY({ monomial_mul_zq(q, a, n, k); }) }
int orc_monomial_mul_t64(u64* a, size_t n, int64_t k) { ORC_TRY({ monomial_mul_t64(a, n, k); }) }

// ---------------- decomposition ----------------
int orc_decompose_zq(u64 q, unsigned log_b, unsigned d, const u64* in, size_t n, u64* out) {
    ORC_TRY({ DecomposorZq(q, log_b, d).decompose_vec(in, n, out); })
}
int orc_decompose_t64(unsigned log_b, unsigned d, const u64* in, size_t n, u64* out) {
    ORC_TRY({ DecomposorT64(log_b, d).decompose_vec(in, n, out); })
}
int orc_decomposor_zq_info(u64 q, unsigned log_b, unsigned d, unsigned* log_q, unsigned* rounding_bits, u64* bases) {
    ORC_TRY({
        DecomposorZq dc(q, log_b, d);
        *log_q = dc.log_q;
        *rounding_bits = dc.rounding_bits;
        for (unsigned i = 0; i < d; ++i) bases[i] = dc.base(i);
    })
}
int orc_rounding_shr_t64(const u64* in, u64* out, size_t n, unsigned bits) {
    ORC_TRY({
        for (size_t i = 0; i < n; ++i) out[i] = DecomposorT64::rounding_shr_bits(in[i], bits);
    })
}

// ---------------- RNS ----------------
static RnsPoly rns_wrap(const u64* qs, size_t nq, const u64* data, size_t n) {
    RnsPoly x;
    x.qs.assign(qs, qs + nq);
    for (size_t i = 0; i < nq; ++i) x.limbs.emplace_back(data + i * n, data + (i + 1) * n);
    return x;
}
static void rns_unwrap(const RnsPoly& x, u64* out) {
    for (size_t i = 0; i < x.limbs.size(); ++i) std::memcpy(out + i * x.n(), x.limbs[i].data(), x.n() * 8);
}
// out: (nq + np) limbs
int orc_rns_extend_bases(const u64* qs, size_t nq, const u64* ps, size_t np, const u64* in, u64* out, size_t n) {
    ORC_TRY({ rns_unwrap(rns_extend_bases(rns_wrap(qs, nq, in, n), Vec(ps, ps + np)), out); })
}
// out: np limbs
int orc_rns_switch_bases(const u64* qs, size_t nq, const u64* ps, size_t np, const u64* in, u64* out, size_t n) {
    ORC_TRY({ rns_unwrap(rns_switch_bases(rns_wrap(qs, nq, in, n), Vec(ps, ps + np)), out); })
}
// out: nq - k limbs
int orc_rns_rescale_k(const u64* qs, size_t nq, size_t k, const u64* in, u64* out, size_t n) {
    ORC_TRY({ rns_unwrap(rns_rescale_k(rns_wrap(qs, nq, in, n), k), out); })
}

// ---------------- FHEW ----------------
struct orc_fhew_param_c {
    unsigned log_n;
    u64 big_q, p;
    unsigned rlwe_log_b, rlwe_d, rgsw_log_b, rgsw_d, n_s;
    u64 q_ks;
    unsigned ks_log_b, ks_d, w;
};
static FhewParam to_param(const orc_fhew_param_c& c) {
    FhewParam p;
    p.log_n = c.log_n;
    p.big_q = c.big_q;
    p.p = c.p;
    p.rlwe_log_b = c.rlwe_log_b;
    p.rlwe_d = c.rlwe_d;
    p.rgsw_log_b = c.rgsw_log_b;
    p.rgsw_d = c.rgsw_d;
    p.n_s = c.n_s;
    p.q_ks = c.q_ks;
    p.ks_log_b = c.ks_log_b;
    p.ks_d = c.ks_d;
    p.w = c.w;
    return p;
}
void orc_fhew_testing_param(orc_fhew_param_c* o) {
    FhewParam p = fhew_single_key_testing_param();
    o->log_n = p.log_n;
    o->big_q = p.big_q;
    o->p = p.p;
    o->rlwe_log_b = p.rlwe_log_b;
    o->rlwe_d = p.rlwe_d;
    o->rgsw_log_b = p.rgsw_log_b;
    o->rgsw_d = p.rgsw_d;
    o->n_s = p.n_s;
    o->q_ks = p.q_ks;
    o->ks_log_b = p.ks_log_b;
    o->ks_d = p.ks_d;
    o->w = p.w;
}
void* orc_fhew_keygen(const orc_fhew_param_c* c, u64 seed) {
    try {
        return new FhewKey(fhew_key_gen(to_param(*c), seed));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// the same key generation fed by the counter-based stream of the device keygen (oracle/orc_keygen.hpp)
void* orc_fhew_keygen_ctr(const orc_fhew_param_c* c, u64 seed) {
    try {
        return new FhewKey(fhew_key_gen_ctr(to_param(*c), seed));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_fhew_key_free(void* h) { delete (FhewKey*)h; }
// key export in reference layout (u64 words):
//   ksk_a [N*d_ks][n_s], ksk_b [N*d_ks]            (index = digit*N + coefficient)
//   brk   [n_s][2*d_rgsw][2 (a,b)][N]
//   ak    [w+1][d_rlwe][2 (a,b)][N], ak_t [w+1]
//   z [N] (i64), s [n_s] (i64)
int orc_fhew_key_export(void* h, u64* ksk_a, u64* ksk_b, u64* brk, u64* ak, int64_t* ak_t, int64_t* z, int64_t* s) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        const FhewParam& P = K.param;
        size_t n = P.n();
        for (size_t i = 0; i < K.ksk.size(); ++i) {
            if (ksk_a) std::memcpy(ksk_a + i * P.n_s, K.ksk[i].a.data(), P.n_s * 8);
            if (ksk_b) ksk_b[i] = K.ksk[i].b;
        }
        if (brk)
            for (size_t j = 0; j < K.brk.size(); ++j)
                for (size_t r = 0; r < K.brk[j].size(); ++r) {
                    u64* base = brk + ((j * K.brk[j].size() + r) * 2) * n;
                    std::memcpy(base, K.brk[j][r].a.data(), n * 8);
                    std::memcpy(base + n, K.brk[j][r].b.data(), n * 8);
                }
        if (ak)
            for (size_t v = 0; v < K.ak.size(); ++v)
                for (size_t r = 0; r < K.ak[v].size(); ++r) {
                    u64* base = ak + ((v * K.ak[v].size() + r) * 2) * n;
                    std::memcpy(base, K.ak[v][r].a.data(), n * 8);
                    std::memcpy(base + n, K.ak[v][r].b.data(), n * 8);
                }
        if (ak_t)
            for (size_t v = 0; v < K.ak_t.size(); ++v) ak_t[v] = K.ak_t[v];
        if (z) std::memcpy(z, K.z.data(), n * 8);
        if (s) std::memcpy(s, K.s.data(), P.n_s * 8);
    })
}
// inverse of orc_fhew_key_export for the PUBLIC part (evaluation keys only; z and s stay empty): lets tests and
// bench.py's checker run the oracle on arbitrary (e.g. synthetic uniformly random) key material.
void* orc_fhew_key_import(const orc_fhew_param_c* c, const u64* ksk_a, const u64* ksk_b, const u64* brk, const u64* ak, const int64_t* ak_t) {
    try {
        FhewKey* K = new FhewKey();
        K->param = to_param(*c);
        const FhewParam& P = K->param;
        const size_t n = P.n();
        K->ksk.resize(n * P.ks_d);
        for (size_t i = 0; i < K->ksk.size(); ++i) {
            K->ksk[i].a.assign(ksk_a + i * P.n_s, ksk_a + (i + 1) * P.n_s);
            K->ksk[i].b = ksk_b[i];
        }
        const size_t rg = 2 * P.rgsw_d;
        K->brk.resize(P.n_s);
        for (size_t j = 0; j < P.n_s; ++j) {
            K->brk[j].resize(rg);
            for (size_t r = 0; r < rg; ++r) {
                const u64* base = brk + ((j * rg + r) * 2) * n;
                K->brk[j][r].a.assign(base, base + n);
                K->brk[j][r].b.assign(base + n, base + 2 * n);
            }
        }
        K->ak.resize(P.w + 1);
        K->ak_t.resize(P.w + 1);
        for (size_t v = 0; v <= P.w; ++v) {
            K->ak[v].resize(P.rlwe_d);
            for (size_t r = 0; r < P.rlwe_d; ++r) {
                const u64* base = ak + ((v * P.rlwe_d + r) * 2) * n;
                K->ak[v][r].a.assign(base, base + n);
                K->ak[v][r].b.assign(base + n, base + 2 * n);
            }
            K->ak_t[v] = ak_t[v];
        }
        return K;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// ct layout: [a_0 .. a_{N-1}, b]
static LweCt lwe_wrap(const u64* ct, size_t n) { return LweCt{Vec(ct, ct + n), ct[n]}; }
static void lwe_unwrap(const LweCt& c, u64* out) {
    std::memcpy(out, c.a.data(), c.a.size() * 8);
    out[c.a.size()] = c.b;
}
int orc_fhew_encrypt(void* h, const int* bits, size_t count, u64 seed, u64* cts) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        Rng rng(seed);
        for (size_t i = 0; i < count; ++i) lwe_unwrap(fhew_encrypt_bit(K, bits[i] != 0, rng), cts + i * (n + 1));
    })
}
int orc_fhew_decrypt(void* h, const u64* cts, size_t count, int* out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        for (size_t i = 0; i < count; ++i) out[i] = fhew_decrypt_bit(K, lwe_wrap(cts + i * (n + 1), n));
    })
}
// raw phase b - <a,z> mod Q (for noise inspection)
int orc_fhew_phase(void* h, const u64* cts, size_t count, u64* out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        for (size_t i = 0; i < count; ++i) out[i] = lwe_decrypt(K.param.big_q, K.z, lwe_wrap(cts + i * (n + 1), n));
    })
}
int orc_fhew_gate_poly(const orc_fhew_param_c* c, const int* table, u64* f) {
    ORC_TRY({
        Vec v = fhew_gate_poly(to_param(*c), table);
        std::memcpy(f, v.data(), v.size() * 8);
    })
}
// Fhew::op on `count` already-linearly-combined ciphertexts (fhew.rs:31-39)
int orc_fhew_op(void* h, const int* table, const u64* cts_in, size_t count, u64* cts_out, int threads) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        twiddle(K.param.big_q);
        parallel_for(count, threads, [&](size_t i) { lwe_unwrap(fhew_op(K, table, lwe_wrap(cts_in + i * (n + 1), n)), cts_out + i * (n + 1)); });
    })
}
// Bootstrapping::bootstrap with an arbitrary test polynomial f (bootstrapping.rs:149-155)
int orc_fhew_bootstrap(void* h, const u64* f, const u64* cts_in, size_t count, u64* cts_out, int threads) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        Vec fv(f, f + n);
        twiddle(K.param.big_q);
        parallel_for(count, threads, [&](size_t i) { lwe_unwrap(fhew_bootstrap(K, fv, lwe_wrap(cts_in + i * (n + 1), n)), cts_out + i * (n + 1)); });
    })
}
// first three steps of bootstrap: out is [a (n_s), b] mod 2N
int orc_fhew_prologue(void* h, const u64* cts_in, size_t count, u64* out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        for (size_t i = 0; i < count; ++i) lwe_unwrap(fhew_bootstrap_prologue(K, lwe_wrap(cts_in + i * (n + 1), n)), out + i * (K.param.n_s + 1));
    })
}
int orc_lwe_key_switch(void* h, const u64* cts_in, size_t count, u64* out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        for (size_t i = 0; i < count; ++i) lwe_unwrap(lwe_key_switch(K.param, K.ksk, lwe_wrap(cts_in + i * (n + 1), n)), out + i * (K.param.n_s + 1));
    })
}
// schedule of blind_rotate_core for an LWE mask a (mod 2N): writes pairs (kind, idx); returns count or -1
long orc_fhew_schedule(const orc_fhew_param_c* c, const u64* a, int* steps, size_t cap) {
    try {
        FhewParam P = to_param(*c);
        auto st = blind_rotate_schedule(P, Vec(a, a + P.n_s));
        if (st.size() > cap) throw std::runtime_error("schedule cap too small");
        for (size_t i = 0; i < st.size(); ++i) {
            steps[2 * i] = st[i].kind;
            steps[2 * i + 1] = (int)st[i].idx;
        }
        return (long)st.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
// acc layout [a (N), b (N)]
int orc_fhew_external_product(void* h, size_t j, const u64* acc_in, u64* acc_out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        RlweCt c{Vec(acc_in, acc_in + n), Vec(acc_in + n, acc_in + 2 * n)};
        RlweCt o = rgsw_external_product(K.param, K.brk.at(j), c);
        std::memcpy(acc_out, o.a.data(), n * 8);
        std::memcpy(acc_out + n, o.b.data(), n * 8);
    })
}
// Rgsw::internal_product (rgsw.rs:130-150): ct0, ct1, out are RGSW ciphertexts [2d rows][2 (a, b)][n] over Z_q
int orc_rgsw_internal_product(u64 q, unsigned log_n, unsigned log_b, unsigned d, const u64* ct0, const u64* ct1, u64* out) {
    ORC_TRY({
        const size_t n = (size_t)1 << log_n, rows = 2 * (size_t)d;
        auto wrap = [&](const u64* p) {
            std::vector<RlweCt> v(rows);
            for (size_t r = 0; r < rows; ++r) {
                v[r].a.assign(p + (2 * r) * n, p + (2 * r + 1) * n);
                v[r].b.assign(p + (2 * r + 1) * n, p + (2 * r + 2) * n);
            }
            return v;
        };
        const std::vector<RlweCt> o = rgsw_internal_product(q, n, DecomposorZq(q, log_b, d), wrap(ct0), wrap(ct1));
        for (size_t r = 0; r < rows; ++r) {
            std::memcpy(out + (2 * r) * n, o[r].a.data(), n * 8);
            std::memcpy(out + (2 * r + 1) * n, o[r].b.data(), n * 8);
        }
    })
}
int orc_fhew_automorphism(void* h, size_t v, const u64* acc_in, u64* acc_out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        RlweCt c{Vec(acc_in, acc_in + n), Vec(acc_in + n, acc_in + 2 * n)};
        RlweCt o = rlwe_automorphism(K.param, K.ak.at(v), K.ak_t.at(v), c);
        std::memcpy(acc_out, o.a.data(), n * 8);
        std::memcpy(acc_out + n, o.b.data(), n * 8);
    })
}
// blind rotation only: input [a (n_s), b] mod 2N, output acc [a (N), b (N)]
int orc_fhew_blind_rotate(void* h, const u64* f, const u64* ct2n, u64* acc_out) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        RlweCt o = fhew_blind_rotate(K, Vec(f, f + n), lwe_wrap(ct2n, K.param.n_s));
        std::memcpy(acc_out, o.a.data(), n * 8);
        std::memcpy(acc_out + n, o.b.data(), n * 8);
    })
}
int orc_rlwe_decrypt(void* h, const u64* acc, u64* pt) {
    ORC_TRY({
        FhewKey& K = *(FhewKey*)h;
        size_t n = K.param.n();
        RlweCt c{Vec(acc, acc + n), Vec(acc + n, acc + 2 * n)};
        Vec p = rlwe_decrypt(K.param.big_q, K.z, c);
        std::memcpy(pt, p.data(), n * 8);
    })
}

// ---------------- TFHE ----------------
struct orc_tfhe_param_c {
    unsigned log_p, padding, n;
    double tlwe_std;
    unsigned ks_log_b, ks_d, big_n, k;
    double tglwe_std;
    unsigned bs_log_b, bs_d;
};
static TfheParam to_tparam(const orc_tfhe_param_c& c) {
    TfheParam P;
    P.log_p = c.log_p;
    P.padding = c.padding;
    P.n = c.n;
    P.tlwe_std = c.tlwe_std;
    P.ks_log_b = c.ks_log_b;
    P.ks_d = c.ks_d;
    P.big_n = c.big_n;
    P.k = c.k;
    P.tglwe_std = c.tglwe_std;
    P.bs_log_b = c.bs_log_b;
    P.bs_d = c.bs_d;
    return P;
}
void orc_tfhe_testing_param(orc_tfhe_param_c* o) {
    TfheParam P = tfhe_testing_param();
    o->log_p = P.log_p;
    o->padding = P.padding;
    o->n = P.n;
    o->tlwe_std = P.tlwe_std;
    o->ks_log_b = P.ks_log_b;
    o->ks_d = P.ks_d;
    o->big_n = P.big_n;
    o->k = P.k;
    o->tglwe_std = P.tglwe_std;
    o->bs_log_b = P.bs_log_b;
    o->bs_d = P.bs_d;
}
void* orc_tfhe_keygen(const orc_tfhe_param_c* c, u64 seed) {
    try {
        return new TfheKey(tfhe_key_gen(to_tparam(*c), seed));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// public evaluation keys only, in the layout of orc_tfhe_key_export (known-answer replays: tests/test_cpu_refpin.py)
void* orc_tfhe_key_import(const orc_tfhe_param_c* c, const u64* brk, const u64* ksk_a, const u64* ksk_b) {
    try {
        TfheKey* K = new TfheKey();
        K->param = to_tparam(*c);
        const TfheParam& P = K->param;
        const size_t N = P.big_n, rows = (P.k + 1) * P.bs_d, polys = P.k + 1, kn = (size_t)P.k * N;
        K->brk.resize(P.n);
        for (size_t i = 0; i < P.n; ++i) {
            K->brk[i].resize(rows);
            for (size_t r = 0; r < rows; ++r) {
                const u64* base = brk + ((i * rows + r) * polys) * N;
                K->brk[i][r].a.resize(P.k);
                for (unsigned j = 0; j < P.k; ++j) K->brk[i][r].a[j].assign(base + j * N, base + (j + 1) * N);
                K->brk[i][r].b.assign(base + P.k * N, base + (P.k + 1) * N);
            }
        }
        K->ksk.resize(kn * P.ks_d);
        for (size_t i = 0; i < K->ksk.size(); ++i) {
            K->ksk[i].a.assign(ksk_a + i * P.n, ksk_a + (i + 1) * P.n);
            K->ksk[i].b = ksk_b[i];
        }
        return K;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void* orc_tfhe_keygen_ctr(const orc_tfhe_param_c* c, u64 seed) {
    try {
        return new TfheKey(tfhe_key_gen_ctr(to_tparam(*c), seed));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_tfhe_key_free(void* h) { delete (TfheKey*)h; }
// export: brk [n][(k+1)*d][(k+1) polys: a_0..a_{k-1}, b][N];  ksk_a [(kN)*d_ks][n], ksk_b [(kN)*d_ks]; z [n]; s [kN]
int orc_tfhe_key_export(void* h, u64* brk, u64* ksk_a, u64* ksk_b, int64_t* z, int64_t* s) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        size_t N = P.big_n, rows = (P.k + 1) * P.bs_d, polys = P.k + 1;
        if (brk)
            for (size_t i = 0; i < K.brk.size(); ++i)
                for (size_t r = 0; r < rows; ++r) {
                    u64* base = brk + ((i * rows + r) * polys) * N;
                    for (unsigned j = 0; j < P.k; ++j) std::memcpy(base + j * N, K.brk[i][r].a[j].data(), N * 8);
                    std::memcpy(base + P.k * N, K.brk[i][r].b.data(), N * 8);
                }
        for (size_t i = 0; i < K.ksk.size(); ++i) {
            if (ksk_a) std::memcpy(ksk_a + i * P.n, K.ksk[i].a.data(), P.n * 8);
            if (ksk_b) ksk_b[i] = K.ksk[i].b;
        }
        if (z) std::memcpy(z, K.z.data(), P.n * 8);
        if (s) std::memcpy(s, K.s.data(), K.s.size() * 8);
    })
}
// ct layout [a (n), b]
int orc_tfhe_encrypt(void* h, const u64* msgs, size_t count, u64 seed, u64* cts) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        Rng rng(seed);
        for (size_t i = 0; i < count; ++i) {
            TlweCt c = tlwe_sk_encrypt(P.n, P.tlwe_std, K.z, msgs[i] << P.log_delta(), rng);
            std::memcpy(cts + i * (P.n + 1), c.a.data(), P.n * 8);
            cts[i * (P.n + 1) + P.n] = c.b;
        }
    })
}
int orc_tfhe_decrypt(void* h, const u64* cts, size_t count, u64* msgs, u64* raw_phase) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        for (size_t i = 0; i < count; ++i) {
            TlweCt c{Vec(cts + i * (P.n + 1), cts + i * (P.n + 1) + P.n), cts[i * (P.n + 1) + P.n]};
            u64 mu = tlwe_decrypt_raw(K.z, c);
            if (raw_phase) raw_phase[i] = mu;
            if (msgs) msgs[i] = tlwe_decode(P, mu);
        }
    })
}
int orc_tfhe_lut_poly(const orc_tfhe_param_c* c, const u64* table, u64* v) {
    ORC_TRY({
        TfheParam P = to_tparam(*c);
        Vec r = tfhe_lut_poly(P, Vec(table, table + P.p()));
        std::memcpy(v, r.data(), r.size() * 8);
    })
}
int orc_tfhe_bootstrap(void* h, const u64* v, const u64* cts_in, size_t count, u64* cts_out, int threads) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        Vec vv(v, v + P.big_n);
        twiddle64(P.big_n);
        twiddle64(P.big_n / 2);
        parallel_for(count, threads, [&](size_t i) {
            TlweCt c{Vec(cts_in + i * (P.n + 1), cts_in + i * (P.n + 1) + P.n), cts_in[i * (P.n + 1) + P.n]};
            TlweCt o = tfhe_bootstrap(K, vv, c);
            std::memcpy(cts_out + i * (P.n + 1), o.a.data(), P.n * 8);
            cts_out[i * (P.n + 1) + P.n] = o.b;
        });
    })
}
// blind rotation + sample extract only (before key switch): out [a (kN), b]
int orc_tfhe_blind_rotate_extract(void* h, const u64* v, const u64* ct_in, u64* out) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        TlweCt c{Vec(ct_in, ct_in + P.n), ct_in[P.n]};
        TlweCt o = tglwe_sample_extract(tfhe_blind_rotate(K, Vec(v, v + P.big_n), c), 0);
        std::memcpy(out, o.a.data(), o.a.size() * 8);
        out[o.a.size()] = o.b;
    })
}
// glwe layout [(k+1) polys][N]; row i of the bootstrapping key
int orc_tfhe_external_product(void* h, size_t i, const u64* glwe_in, u64* glwe_out) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        size_t N = P.big_n;
        TglweCt c;
        for (unsigned j = 0; j < P.k; ++j) c.a.emplace_back(glwe_in + j * N, glwe_in + (j + 1) * N);
        c.b.assign(glwe_in + P.k * N, glwe_in + (P.k + 1) * N);
        TglweCt o = tggsw_external_product(P, K.brk.at(i), c);
        for (unsigned j = 0; j < P.k; ++j) std::memcpy(glwe_out + j * N, o.a[j].data(), N * 8);
        std::memcpy(glwe_out + P.k * N, o.b.data(), N * 8);
    })
}
int orc_tfhe_key_switch(void* h, const u64* ct_in /* [kN + 1] */, u64* ct_out /* [n + 1] */) {
    ORC_TRY({
        TfheKey& K = *(TfheKey*)h;
        const TfheParam& P = K.param;
        size_t m = (size_t)P.k * P.big_n;
        TlweCt c{Vec(ct_in, ct_in + m), ct_in[m]};
        TlweCt o = tlwe_key_switch(P, K.ksk, c);
        std::memcpy(ct_out, o.a.data(), P.n * 8);
        ct_out[P.n] = o.b;
    })
}

// ---------------- CKKS ----------------
void* orc_ckks_keygen(unsigned log_n, unsigned log_qi, unsigned big_l, u64 seed, const int64_t* auto_ts, size_t n_ts) {
    try {
        CkksParam P = ckks_param_new(log_n, log_qi, big_l);
        for (u64 q : P.qps()) twiddle(q);
        return new CkksKey(ckks_key_gen(P, seed, std::vector<i64>(auto_ts, auto_ts + n_ts)));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// public evaluation keys only: explicit moduli, rlk and n_auto automorphism keys [2 (b, a)][2L][N] each with exponents ts
void* orc_ckks_key_import(unsigned log_n, const u64* qs, const u64* ps, size_t big_l, const u64* rlk, const int64_t* ts, const u64* autk,
                          size_t n_auto) {
    try {
        CkksKey* K = new CkksKey();
        K->param.log_n = log_n;
        K->param.qs.assign(qs, qs + big_l);
        K->param.ps.assign(ps, ps + big_l);
        const Vec qps = K->param.qps();
        for (u64 q : qps) twiddle(q);
        const size_t n = K->param.n(), words = 2 * big_l * n;
        auto wrap = [&](const u64* p) {
            CkksCt c;
            c.b = rns_wrap(qps.data(), 2 * big_l, p, n);
            c.a = rns_wrap(qps.data(), 2 * big_l, p + words, n);
            return c;
        };
        K->rlk = wrap(rlk);
        for (size_t i = 0; i < n_auto; ++i) K->autk.emplace_back((i64)ts[i], wrap(autk + i * 2 * words));
        return K;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void* orc_ckks_keygen_ctr(unsigned log_n, unsigned log_qi, unsigned big_l, u64 seed, const int64_t* auto_ts, size_t n_ts) {
    try {
        CkksParam P = ckks_param_new(log_n, log_qi, big_l);
        for (u64 q : P.qps()) twiddle(q);
        return new CkksKey(ckks_key_gen_ctr(P, seed, std::vector<i64>(auto_ts, auto_ts + n_ts)));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_ckks_key_free(void* h) { delete (CkksKey*)h; }
int orc_ckks_moduli(void* h, u64* qs, u64* ps) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        std::memcpy(qs, K.param.qs.data(), K.param.qs.size() * 8);
        std::memcpy(ps, K.param.ps.data(), K.param.ps.size() * 8);
    })
}
// key-switching key export: [2 (b, a)][2L limbs (qs then ps)][N]; which = -1 -> rlk, else autk[which]
int orc_ckks_ksk_export(void* h, int which, u64* out, int64_t* sk) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        const CkksCt& k = which < 0 ? K.rlk : K.autk.at(which).second;
        size_t n = K.param.n(), l2 = k.b.limbs.size();
        if (out) {
            rns_unwrap(k.b, out);
            rns_unwrap(k.a, out + l2 * n);
        }
        if (sk) std::memcpy(sk, K.sk.data(), n * 8);
    })
}
// ciphertext layout: [2 (b, a)][l limbs][N] over qs[0..l)
static CkksCt ct_wrap(const CkksParam& P, const u64* ct, size_t l) {
    size_t n = P.n();
    CkksCt c;
    c.b = rns_wrap(P.qs.data(), l, ct, n);
    c.a = rns_wrap(P.qs.data(), l, ct + l * n, n);
    return c;
}
static void ct_unwrap(const CkksCt& c, u64* out) {
    rns_unwrap(c.b, out);
    rns_unwrap(c.a, out + c.b.limbs.size() * c.b.n());
}
// encrypt an integer plaintext polynomial (i64 coefficients) at level l
int orc_ckks_encrypt(void* h, const int64_t* pt, size_t l, u64 seed, u64* ct) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        Rng rng(seed);
        Vec qs(K.param.qs.begin(), K.param.qs.begin() + l);
        RnsPoly p = rns_from_i64(qs, std::vector<i64>(pt, pt + K.param.n()));
        ct_unwrap(ckks_sk_encrypt(K.sk, p, rng), ct);
    })
}
// decrypt to RNS limbs [l][N]
int orc_ckks_decrypt(void* h, const u64* ct, size_t l, u64* pt) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        rns_unwrap(ckks_decrypt(K.sk, ct_wrap(K.param, ct, l)), pt);
    })
}
// Ckks::mul on `count` pairs at level l; output at level l-1
int orc_ckks_mul(void* h, const u64* ct0, const u64* ct1, size_t l, size_t count, u64* out, int threads) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        size_t n = K.param.n();
        size_t in_sz = 2 * l * n, out_sz = 2 * (l - 1) * n;
        parallel_for(count, threads, [&](size_t i) {
            ct_unwrap(ckks_mul(K.param, K.rlk, ct_wrap(K.param, ct0 + i * in_sz, l), ct_wrap(K.param, ct1 + i * in_sz, l)), out + i * out_sz);
        });
    })
}
// Ckks::key_switch with rlk (which=-1) or autk[which] (after applying its automorphism): level preserved
int orc_ckks_key_switch(void* h, int which, int apply_auto, const u64* ct, size_t l, u64* out) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        CkksCt c = ct_wrap(K.param, ct, l);
        CkksCt o;
        if (which < 0)
            o = ckks_key_switch(K.param, K.rlk, c);
        else if (apply_auto)
            o = ckks_automorphism_ks(K.param, K.autk.at(which).second, K.autk.at(which).first, c);
        else
            o = ckks_key_switch(K.param, K.autk.at(which).second, c);
        ct_unwrap(o, out);
    })
}
int orc_ckks_rescale(void* h, const u64* ct, size_t l, u64* out) {
    ORC_TRY({
        CkksKey& K = *(CkksKey*)h;
        ct_unwrap(ckks_rescale(ct_wrap(K.param, ct, l)), out);
    })
}

}  // extern "C"
